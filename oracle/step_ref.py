"""CPU restatement of the reference training step (trainer.py:172-176) and of torch.optim.Adam.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F

from .continual_ref import continual_loss
from .unet_ref import UNetRef, clone_sd, param_names


def cross_entropy(logits, labels):
    """nn.CrossEntropyLoss() with defaults: mean over B*H*W, no weights (trainer.py:113,174)."""
    return F.cross_entropy(logits, labels)


class AdamRef:
    """optim.Adam(lr, betas) with eps=1e-8, weight_decay=0, amsgrad=False (trainer.py:108-110), restated
    elementwise: m=b1 m+(1-b1) g; v=b2 v+(1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""

    def __init__(self, names, lr=1e-4, betas=(0.5, 0.99), eps=1e-8):
        self.names, self.lr, self.b1, self.b2, self.eps = list(names), lr, betas[0], betas[1], eps
        self.t = 0
        self.m, self.v = {}, {}

    @torch.no_grad()
    def step(self, sd, grads):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2s = math.sqrt(1.0 - self.b2 ** self.t)
        for k in self.names:
            g = grads[k]
            if k not in self.m:
                self.m[k] = torch.zeros_like(g)
                self.v[k] = torch.zeros_like(g)
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            denom = self.v[k].sqrt() / bc2s + self.eps
            sd[k].sub_((self.lr / bc1) * self.m[k] / denom)


def forward_backward(sd, x, labels, num_classes=21, in_dim=3, conv_dim=64, training=True, matched_rounding=False,
                     old=None, T=2.0, lam=1.0, loss_scale=1.0, update_running_stats=True):
    """outputs = model(inputs); loss = c_loss(outputs, labels); loss.backward()  (trainer.py:172-175).

    `old` = (old_state_dict, num_old_classes) adds the distillation term (continual_ref.continual_loss).
    Returns (loss float, logits, {name: grad}); BN running stats in `sd` are updated in place like the
    reference module does in train mode.
    """
    work = clone_sd(sd, requires_grad=True)
    for k in sd:  # running stats must alias the caller's tensors (in-place EMA)
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            work[k] = sd[k]
    net = UNetRef(work, num_classes, in_dim, conv_dim, training=training, matched_rounding=matched_rounding,
                  update_running_stats=update_running_stats)
    logits = net(x)
    if old is None:
        loss = cross_entropy(logits, labels)
    else:
        old_sd, c_old = old
        with torch.no_grad():
            old_logits = UNetRef(old_sd, c_old, in_dim, conv_dim, training=False,
                                 matched_rounding=matched_rounding)(x)
        loss = continual_loss(logits, labels, old_logits, T, lam)
    (loss * loss_scale).backward()
    grads = {k: work[k].grad for k in param_names(sd)}
    return float(loss.detach()), logits.detach(), grads, net.captured


def train_steps(sd, batches, lr=1e-4, betas=(0.5, 0.99), **kw):
    """the hot loop of trainer.py:165-176 over a list of (x, labels); returns the loss trajectory."""
    opt = AdamRef(param_names(sd), lr, betas)
    losses = []
    for x, y in batches:
        loss, _, grads, _ = forward_backward(sd, x, y, **kw)
        opt.step(sd, grads)
        losses.append(loss)
    return losses
