"""Generate the golden fixtures in this directory from the UNMODIFIED reference modules.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The reference ships no tests or golden vectors; these fixtures are outputs of
`models/unet.py` and `metrics.py` themselves on seeded inputs (weights from oracle.unet_ref.make_state_dict,
data from oracle.data — both numpy-PCG64, torch-RNG independent).  They pin the oracle on machines where
/root/reference is absent (the GPU box).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np
import torch
import torch.nn as nn

import metrics as ref_metrics  # /root/reference/metrics.py
from models.unet import UNet as RefUNet  # /root/reference/models/unet.py
from oracle.data import structured_batch, uniform_batch
from oracle.unet_ref import make_state_dict

torch.set_num_threads(8)


def unet_case(name, seed_w, seed_x, b, h, w, num_classes=21):
    sd = make_state_dict(seed_w, num_classes)
    x, y = structured_batch(seed_x, b, h, w, num_classes)
    m = RefUNet(num_classes)
    m.load_state_dict(sd)
    m.train()
    out = m(x)                                   # trainer.py:172
    loss = nn.CrossEntropyLoss()(out, y)         # trainer.py:113,174
    loss.backward()                              # trainer.py:175
    rec = {"logits_train": out.detach().numpy(), "loss": np.float64(loss.item())}
    names, norms, heads = [], [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()))
        heads.append(p.grad.flatten()[:16].numpy().copy() if p.numel() >= 16 else
                     np.pad(p.grad.flatten().numpy(), (0, 16 - p.numel())))
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms)
    rec["grad_heads"] = np.stack(heads)
    st = m.state_dict()
    for k in ("enc1.2.running_mean", "enc1.2.running_var", "dec1.block.5.running_var", "last.5.running_mean"):
        rec["buf_" + k] = st[k].numpy().copy()
    rec["nbt"] = np.int64(st["enc1.2.num_batches_tracked"].item())
    m.eval()                                     # trainer.py:271
    with torch.no_grad():
        rec["logits_eval"] = m(x).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "loss", rec["loss"])


def traj_case(name, steps=3, b=2, h=32, w=32):
    sd = make_state_dict(3)
    m = RefUNet(21)
    m.load_state_dict(sd)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, betas=[0.5, 0.99])   # trainer.py:108-110
    c_loss = nn.CrossEntropyLoss()
    losses = []
    for i in range(steps):
        x, y = structured_batch(100 + i, b, h, w)
        out = m(x)
        opt.zero_grad()
        loss = c_loss(out, y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    st = m.state_dict()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), losses=np.array(losses),
                        w_head=st["last.6.weight"].numpy(), b_enc1=st["enc1.0.bias"].numpy(),
                        w_dec1_slice=st["dec1.block.3.weight"].flatten()[:256].numpy(),
                        rv_last5=st["last.5.running_var"].numpy())
    print(name, losses)


def metrics_case(name):
    rng = np.random.Generator(np.random.PCG64(5))
    rec = {}
    for tag, nc_data, nc_call in (("a", 21, 22), ("b", 21, 21), ("c", 5, 22)):
        t = torch.from_numpy(rng.integers(0, nc_data, size=(2, 32, 32), dtype=np.int64))
        p = torch.from_numpy(rng.integers(0, nc_data, size=(2, 32, 32), dtype=np.int64))
        if tag == "b":
            t[0, :4] = 255  # void-like labels outside [0, nc): masked by the reference
        o = ref_metrics.eval_metrics(t, p, nc_call)                      # metrics.py:55-63 (trainer.py:188)
        rec[f"{tag}_target"], rec[f"{tag}_pred"], rec[f"{tag}_nc"] = t.numpy(), p.numpy(), np.int64(nc_call)
        rec[f"{tag}_out"] = np.array([float(v) for v in o], dtype=np.float32)
        rec[f"{tag}_conf"] = sum(ref_metrics._fast_conf_matrix(a.flatten(), b_.flatten(), nc_call)
                                 for a, b_ in zip(t, p)).numpy()
        rec[f"{tag}_miu"] = np.float64(ref_metrics.mean_IU_(t.numpy(), p.numpy()))      # metrics.py:67-71
        acc, tot, cor = ref_metrics.pixel_acc(t, p, 100.0, float((t == p).sum()))        # metrics.py:6-9
        rec[f"{tag}_pixacc"] = np.array([acc, tot, cor], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, rec["a_out"])


if __name__ == "__main__":
    unet_case("unet_b1_32x32", 0, 1, 1, 32, 32)
    unet_case("unet_b2_48x32_c7", 2, 4, 2, 48, 32, num_classes=7)
    unet_case("unet_b2_64x64", 0, 1, 2, 64, 64)
    traj_case("train_traj_b2_32x32")
    metrics_case("metrics")
