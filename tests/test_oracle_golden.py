"""CPU: the oracle restatement reproduces the golden vectors minted from the unmodified reference
(tests/golden/make_golden.py).  fp32 on CPU against fp32 on CPU: tolerances are round-off only."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref, step_ref
from oracle.data import structured_batch
from oracle.unet_ref import UNetRef, clone_sd, make_state_dict, param_names

CASES = [("unet_b1_32x32", 0, 1, 1, 32, 32, 21), ("unet_b2_48x32_c7", 2, 4, 2, 48, 32, 7),
         ("unet_b2_64x64", 0, 1, 2, 64, 64, 21)]


@pytest.mark.parametrize("name,seed_w,seed_x,b,h,w,nc", CASES)
def test_forward_backward_matches_reference_golden(golden_dir, name, seed_w, seed_x, b, h, w, nc):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = make_state_dict(seed_w, nc)
    x, y = structured_batch(seed_x, b, h, w, nc)
    loss, logits, grads, _ = step_ref.forward_backward(sd, x, y, num_classes=nc)
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(logits.numpy(), g["logits_train"], rtol=1e-5, atol=1e-5)
    names = list(g["grad_names"])
    assert names == param_names(sd)
    for i, k in enumerate(names):
        gr = grads[k]
        assert abs(float(gr.double().norm()) - g["grad_norms"][i]) <= 1e-4 * g["grad_norms"][i] + 1e-9, k
        head = gr.flatten()[:16].numpy()
        np.testing.assert_allclose(head, g["grad_heads"][i][:head.size], rtol=2e-3, atol=1e-6 + 1e-4 * g["grad_norms"][i])
    for k in ("enc1.2.running_mean", "enc1.2.running_var", "dec1.block.5.running_var", "last.5.running_mean"):
        np.testing.assert_allclose(sd[k].numpy(), g["buf_" + k], rtol=1e-5, atol=1e-6)
    assert int(sd["enc1.2.num_batches_tracked"]) == int(g["nbt"])
    with torch.no_grad():
        ev = UNetRef(clone_sd(sd), nc, training=False)(x)
    np.testing.assert_allclose(ev.numpy(), g["logits_eval"], rtol=1e-5, atol=1e-5)


def test_train_trajectory_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_traj_b2_32x32.npz"))
    sd = make_state_dict(3)
    batches = [structured_batch(100 + i, 2, 32, 32) for i in range(3)]
    losses = step_ref.train_steps(sd, batches, lr=1e-4, betas=(0.5, 0.99))
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=2e-5)
    np.testing.assert_allclose(sd["last.6.weight"].numpy(), g["w_head"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["enc1.0.bias"].numpy(), g["b_enc1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["dec1.block.3.weight"].flatten()[:256].numpy(), g["w_dec1_slice"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["last.5.running_var"].numpy(), g["rv_last5"], rtol=1e-4)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_metrics_match_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    t, p, nc = torch.from_numpy(g[f"{tag}_target"]), torch.from_numpy(g[f"{tag}_pred"]), int(g[f"{tag}_nc"])
    conf = sum(metrics_ref.conf_matrix_int(a.numpy(), b.numpy(), nc) for a, b in zip(t, p))
    assert np.array_equal(conf.astype(np.float32), g[f"{tag}_conf"])  # bit-exact counts
    out = metrics_ref.eval_metrics(t, p, nc)
    got = np.array([float(v) for v in out], dtype=np.float32)
    assert np.array_equal(got, g[f"{tag}_out"])  # same float32 formulas on the same matrix
    assert metrics_ref.mean_iu_binary(t.numpy(), p.numpy()) == float(g[f"{tag}_miu"])
    acc, tot, cor = metrics_ref.pixel_acc(t, p, 100.0, float((t == p).sum()))
    assert [acc, tot, cor] == list(g[f"{tag}_pixacc"])


def test_conf_matrix_rejects_out_of_range_prediction():
    t = np.array([0, 1, 2]); p = np.array([0, 5, 1])
    with pytest.raises(RuntimeError):
        metrics_ref.conf_matrix_int(t, p, 3)
    # an out-of-range TARGET is masked, not an error (metrics.py:33)
    assert metrics_ref.conf_matrix_int(np.array([0, 7, -1]), np.array([0, 1, 1]), 3).sum() == 1
