"""Functional CPU restatement of the reference U-Net (models/unet.py) on a plain state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Layer order per block is Conv3x3(+bias) -> ReLU -> BatchNorm2d (models/unet.py:13-18, 28-33, 50-55,
66-71), max-pool 2x2 in front of the encoder blocks (models/unet.py:12) and after enc4
(models/unet.py:80), ConvTranspose2d(k=2, s=2) at the end of each decoder block (models/unet.py:34),
skip concat with the encoder tensor FIRST (models/unet.py:83-87), 1x1 head (models/unet.py:72).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default (models/unet.py:15)
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


def layer_table(num_classes, in_dim=3, conv_dim=64):
    """[(state_dict prefix, kind, Cin, Cout)] in forward order — the module tree of models/unet.py:48-72."""
    c = conv_dim
    t = [("enc1.0", "conv3", in_dim, c), ("enc1.2", "bn", c, c), ("enc1.3", "conv3", c, c), ("enc1.5", "bn", c, c)]
    for name, ci, co in (("enc2", c, 2 * c), ("enc3", 2 * c, 4 * c), ("enc4", 4 * c, 8 * c)):
        t += [(f"{name}.block.1", "conv3", ci, co), (f"{name}.block.3", "bn", co, co),
              (f"{name}.block.4", "conv3", co, co), (f"{name}.block.6", "bn", co, co)]
    for name, ci, cm, co in (("dec1", 8 * c, 16 * c, 8 * c), ("dec2", 16 * c, 8 * c, 4 * c),
                             ("dec3", 8 * c, 4 * c, 2 * c), ("dec4", 4 * c, 2 * c, c)):
        t += [(f"{name}.block.0", "conv3", ci, cm), (f"{name}.block.2", "bn", cm, cm),
              (f"{name}.block.3", "conv3", cm, cm), (f"{name}.block.5", "bn", cm, cm),
              (f"{name}.block.6", "convT", cm, co)]
    t += [("last.0", "conv3", 2 * c, c), ("last.2", "bn", c, c), ("last.3", "conv3", c, c), ("last.5", "bn", c, c),
          ("last.6", "conv1", c, num_classes)]
    return t


def make_state_dict(seed, num_classes=21, in_dim=3, conv_dim=64, dtype=torch.float32):
    """Deterministic (numpy PCG64, torch-version independent) weights with the reference's keys/shapes.

    Scales follow PyTorch's default init magnitudes (kaiming-uniform(a=sqrt 5) => U(-1/sqrt(fan_in), +)),
    BN affine is perturbed away from (1, 0) so that gamma/beta paths are exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def uni(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32)).to(dtype)

    for prefix, kind, ci, co in layer_table(num_classes, in_dim, conv_dim):
        if kind == "conv3":
            b = 1.0 / np.sqrt(ci * 9)
            sd[prefix + ".weight"] = uni((co, ci, 3, 3), b)
            sd[prefix + ".bias"] = uni((co,), b)
        elif kind == "conv1":
            b = 1.0 / np.sqrt(ci)
            sd[prefix + ".weight"] = uni((co, ci, 1, 1), b)
            sd[prefix + ".bias"] = uni((co,), b)
        elif kind == "convT":
            b = 1.0 / np.sqrt(co * 4)  # fan_in of ConvTranspose2d weight (Cin, Cout, 2, 2) is Cout*4
            sd[prefix + ".weight"] = uni((ci, co, 2, 2), b)
            sd[prefix + ".bias"] = uni((co,), b)
        else:
            sd[prefix + ".weight"] = 1.0 + uni((co,), 0.2)
            sd[prefix + ".bias"] = uni((co,), 0.2)
            sd[prefix + ".running_mean"] = torch.zeros(co, dtype=dtype)
            sd[prefix + ".running_var"] = torch.ones(co, dtype=dtype)
            sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or
                                  k.endswith("num_batches_tracked"))]


def _round_bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class UNetRef:
    """Forward pass over a state_dict of tensors (which may require grad).

    matched_rounding=True rounds every conv/convT weight and every conv/convT INPUT to bf16 (fp32
    accumulate, fp32 everywhere else) — the arithmetic the CUDA path performs, used for the tight
    e2e tolerance of SURVEY.md §8(c).
    """

    def __init__(self, sd, num_classes=21, in_dim=3, conv_dim=64, training=True, matched_rounding=False,
                 update_running_stats=True):
        self.sd = sd
        self.table = layer_table(num_classes, in_dim, conv_dim)
        self.training = training
        self.mr = matched_rounding
        self.update = update_running_stats
        self.captured = OrderedDict()

    # -- single layers ---------------------------------------------------------------------
    def _q(self, t):
        return _round_bf16(t) if self.mr else t

    def _qa(self, t):
        # activations: straight-through rounding so autograd still flows
        if not self.mr:
            return t
        return t + (_round_bf16(t.detach()) - t.detach())

    def conv3(self, x, p):  # nn.Conv2d(k=3, s=1, p=1)  models/unet.py:13
        return F.conv2d(self._qa(x), self._q(self.sd[p + ".weight"]), self.sd[p + ".bias"], stride=1, padding=1)

    def conv1(self, x, p):  # nn.Conv2d(k=1)  models/unet.py:72
        return F.conv2d(self._qa(x), self._q(self.sd[p + ".weight"]), self.sd[p + ".bias"])

    def convT(self, x, p):  # nn.ConvTranspose2d(k=2, s=2)  models/unet.py:34
        return F.conv_transpose2d(self._qa(x), self._q(self.sd[p + ".weight"]), self.sd[p + ".bias"], stride=2)

    def bn(self, x, p):  # nn.BatchNorm2d  models/unet.py:15
        rm, rv = self.sd[p + ".running_mean"], self.sd[p + ".running_var"]
        if self.training:
            out = F.batch_norm(x, rm if self.update else None, rv if self.update else None, self.sd[p + ".weight"],
                               self.sd[p + ".bias"], True, BN_MOMENTUM, BN_EPS)
            if self.update:
                self.sd[p + ".num_batches_tracked"] += 1
            return out
        return F.batch_norm(x, rm, rv, self.sd[p + ".weight"], self.sd[p + ".bias"], False, BN_MOMENTUM, BN_EPS)

    def crb(self, x, pc, pb):  # conv -> ReLU -> BN  (models/unet.py:13-15)
        y = F.relu(self.conv3(x, pc))
        if self.mr:
            y = self._qa(y)  # the CUDA path stores relu(conv) as bf16 before normalising
        z = self.bn(y, pb)
        self.captured[pc] = z
        return z

    # -- blocks (models/unet.py:8-38, 48-72) ---------------------------------------------------
    def down(self, x, name):
        x = F.max_pool2d(x, kernel_size=2, stride=2)
        x = self.crb(x, f"{name}.block.1", f"{name}.block.3")
        return self.crb(x, f"{name}.block.4", f"{name}.block.6")

    def up(self, x, name):
        x = self.crb(x, f"{name}.block.0", f"{name}.block.2")
        x = self.crb(x, f"{name}.block.3", f"{name}.block.5")
        return self.convT(x, f"{name}.block.6")

    def forward(self, x):  # models/unet.py:74-92
        enc1 = self.crb(self.crb(x, "enc1.0", "enc1.2"), "enc1.3", "enc1.5")
        enc2 = self.down(enc1, "enc2")
        enc3 = self.down(enc2, "enc3")
        enc4 = self.down(enc3, "enc4")
        center = F.max_pool2d(enc4, kernel_size=2, stride=2)
        dec1 = self.up(center, "dec1")
        dec2 = self.up(torch.cat([enc4, dec1], dim=1), "dec2")
        dec3 = self.up(torch.cat([enc3, dec2], dim=1), "dec3")
        dec4 = self.up(torch.cat([enc2, dec3], dim=1), "dec4")
        x = self.crb(torch.cat([enc1, dec4], dim=1), "last.0", "last.2")
        x = self.crb(x, "last.3", "last.5")
        return self.conv1(x, "last.6")

    __call__ = forward


def clone_sd(sd, requires_grad=False):
    out = OrderedDict()
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
            t.requires_grad_(True)
        out[k] = t
    return out
