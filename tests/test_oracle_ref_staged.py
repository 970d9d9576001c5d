"""CPU: the oracle restatement against the reference's OWN modules staged under oracle/_ref (byte-compiled from
/root/reference by oracle/build_ref.py; skipped when they were never built).  Unlike tests/test_oracle_pinned.py
this also runs where /root/reference does not exist (the GPU box): the staged files travel with the snapshot."""
import numpy as np
import pytest
import torch

from continual_learning_b200.synthetic import structured_batch, uniform_batch
from oracle import build_ref, metrics_ref
from oracle.unet_ref import UNetRef, clone_sd, make_state_dict

build_ref.build_ref()
pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not built (no /root/reference here)")


def test_staged_reference_unet_equals_the_restatement_bit_for_bit():
    UNet, _ = build_ref.load()
    sd = make_state_dict(3)
    x, y = structured_batch(4, 2, 32, 32)
    m = UNet(21)
    m.load_state_dict(sd)
    m.train()
    out = m(x)
    torch.nn.CrossEntropyLoss()(out, y).backward()
    work = clone_sd(sd, requires_grad=True)
    ref = UNetRef(work, 21, training=True)(x)
    torch.nn.functional.cross_entropy(ref, y).backward()
    assert torch.equal(out, ref)
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, work[k].grad, rtol=1e-4, atol=1e-7), k


def test_staged_reference_metrics_equal_the_restatement():
    _, mt = build_ref.load()
    t, _ = None, None
    _, t = uniform_batch(5, 2, 48, 40, 21)
    _, p = uniform_batch(6, 2, 48, 40, 21)
    got = mt.eval_metrics(t, p, 22)
    want = metrics_ref.eval_metrics(t, p, 22)
    assert [float(a) for a in got] == [float(b) for b in want]
    conf = mt._fast_conf_matrix(t.flatten(), p.flatten(), 22)
    assert np.array_equal(conf.numpy().astype(np.int64), metrics_ref.conf_matrix_int(t.numpy(), p.numpy(), 22))
