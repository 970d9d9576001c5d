"""Continual-learning regulariser — PARITY UNPINNED (the reference contains no such code).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Specification adopted from BASELINE.json:north_star / SURVEY.md §8(c), LwF-style as cited by the
reference README (README.md:1,70-71):
    L = CE(z, y) + lam * T^2 * (1/P) * sum_pixels KL( softmax(z_old/T) || softmax(z[:, :C_old]/T) )
with z_old the logits of the frozen previous-task UNet(C_old) in eval mode.
"""
import torch.nn.functional as F


def continual_loss(logits, labels, old_logits, T=2.0, lam=1.0):
    c_old = old_logits.shape[1]
    p = logits.shape[0] * logits.shape[2] * logits.shape[3]
    ce = F.cross_entropy(logits, labels)
    kd = F.kl_div(F.log_softmax(logits[:, :c_old] / T, dim=1), F.softmax(old_logits / T, dim=1), reduction="sum") / p
    return ce + lam * T * T * kd


def expand_head(old_sd, new_sd, c_old):
    """new head = old head rows in the first C_old channels, all other layers copied."""
    for k, v in old_sd.items():
        if k.startswith("last.6."):
            new_sd[k][:c_old] = v
        else:
            new_sd[k] = v.clone()
    return new_sd
