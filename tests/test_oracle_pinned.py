"""CPU, build container only: the oracle against the LIVE reference modules in /root/reference
(skipped where the reference tree is absent, e.g. on the GPU box — the golden fixtures cover that)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import metrics_ref, step_ref
from oracle.chunked_ref import chunked_forward_backward
from oracle.data import structured_batch, uniform_batch
from oracle.unet_ref import UNetRef, clone_sd, make_state_dict, param_names

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        import metrics as ref_metrics
        from models.unet import UNet
    finally:
        sys.path.remove(REF)
    return UNet, ref_metrics


def test_state_dict_keys_and_shapes(ref):
    UNet, _ = ref
    m = UNet(21)
    sd = make_state_dict(0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and v.dtype == sd[k].dtype, k


def test_forward_backward_bit_close_to_reference(ref):
    UNet, _ = ref
    sd = make_state_dict(11)
    x, y = uniform_batch(12, 2, 32, 48)
    m = UNet(21)
    m.load_state_dict(sd)
    m.train()
    out = m(x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    l2, logits, grads, _ = step_ref.forward_backward(clone_sd(sd), x, y)
    assert torch.equal(out.detach(), logits)  # same ops in the same order on the same machine
    assert l2 == float(loss)
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, grads[k], rtol=1e-4, atol=1e-7), k


def test_chunked_oracle_equals_dataparallel_semantics(ref):
    """per-replica BN statistics + global-mean CE == average of per-chunk gradients (trainer.py:120-122)."""
    UNet, _ = ref
    sd = make_state_dict(5)
    x, y = structured_batch(6, 4, 32, 32)
    loss, grads = chunked_forward_backward(clone_sd(sd), x, y, 2)
    m = UNet(21)
    m.load_state_dict(sd)
    m.train()
    tot = 0.0
    for xc, yc in zip(x.chunk(2), y.chunk(2)):
        l = torch.nn.CrossEntropyLoss()(m(xc), yc) / 2
        l.backward()
        tot += float(l)
    assert abs(tot - loss) < 1e-6
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, grads[k], rtol=1e-4, atol=1e-7), k


def test_metrics_against_reference(ref):
    _, rm = ref
    rng = np.random.Generator(np.random.PCG64(9))
    t = torch.from_numpy(rng.integers(0, 21, size=(3, 16, 16), dtype=np.int64))
    p = torch.from_numpy(rng.integers(0, 21, size=(3, 16, 16), dtype=np.int64))
    a = rm.eval_metrics(t, p, 22)
    b = metrics_ref.eval_metrics(t, p, 22)
    assert [float(v) for v in a] == [float(v) for v in b]
    assert float(rm.mean_IU_(t.numpy(), p.numpy())) == float(metrics_ref.mean_iu_binary(t.numpy(), p.numpy()))
    # a prediction >= nc on a kept last-row target breaks the reference's reshape
    with pytest.raises(RuntimeError):
        rm._fast_conf_matrix(torch.tensor([2]), torch.tensor([5]), 3)
