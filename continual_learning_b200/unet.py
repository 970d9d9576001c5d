"""Drop-in `UNet` (reference models/unet.py:40-92) whose forward/backward run on libclk kernels.

Same constructor `UNet(num_classes, in_dim=3, conv_dim=64)`, same attribute tree (so `state_dict()`
has the reference's 136 keys, shapes and dtypes and checkpoints interchange, trainer.py:68-102), same
`forward(x[B,in_dim,H,W]) -> [B,num_classes,H,W]`.  The stock nn layers below only OWN the fp32
parameters and buffers; they are never called.  CUDA (sm_100a) only — no CPU fallback.
"""
import torch
import torch.nn as nn

from . import ops
from .engine import UNetEngine


def _crb(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.BatchNorm2d(cout)]


class DownBlock(nn.Module):
    """pool -> (conv, relu, bn) x2 under `.block` (reference models/unet.py:8-22)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.block = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=2), *_crb(in_dim, out_dim), *_crb(out_dim, out_dim))


class UpBlock(nn.Module):
    """(conv, relu, bn) x2 -> ConvTranspose2d 2x2/s2 under `.block` (reference models/unet.py:24-38)."""

    def __init__(self, in_dim, mid_dim, out_dim):
        super().__init__()
        self.block = nn.Sequential(*_crb(in_dim, mid_dim), *_crb(mid_dim, mid_dim),
                                   nn.ConvTranspose2d(mid_dim, out_dim, kernel_size=2, stride=2))


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, training, *params):
        logits = engine.forward(x, training=training)
        ctx.engine = engine
        ctx.generation = engine.generation  # the engine keeps the activations of its LATEST forward only
        out = logits.permute(0, 3, 1, 2)  # [B, C, H, W] view over NHWC memory
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.engine
        if ctx.generation != eng.generation:
            raise RuntimeError("UNet.backward: another forward ran on this module since the one being differentiated; "
                               "the engine keeps the activations of the latest forward only (forward -> backward, as "
                               "in trainer.py:172-175)")
        pending = getattr(eng, "_pending_dlogits", None)
        eng._pending_dlogits = None
        if pending is not None and grad_out.data_ptr() == pending[0].data_ptr():
            dl = pending[1]  # fused loss already produced bf16 [B,H,W,64] dlogits
        elif pending is not None:
            # `outputs` had a second differentiable consumer: autograd summed its gradient with the (all-zero) token
            # of the fused loss, so grad_out holds only the OTHER consumers' part -> add the fused loss's dlogits
            g = grad_out.permute(0, 2, 3, 1)
            dl = pending[1].clone()
            dl[..., :g.shape[3]] += g.to(dl.dtype)
        else:
            g = grad_out.permute(0, 2, 3, 1)
            dl = torch.zeros((*g.shape[:3], 64), device=g.device, dtype=torch.bfloat16)
            dl[..., :g.shape[3]] = g
        comm = getattr(eng, "comm", None)
        views = eng.backward(dl, after_group=(lambda g: comm.launch_group(eng.G, g)) if comm is not None else None)
        if comm is not None:
            comm.finish(eng.G)
        grads = []
        for p, v in zip(eng.params, views):
            # autograd accumulates `p.grad += g` when a grad already exists; never alias then
            grads.append(v.clone() if (p.grad is not None and p.grad.data_ptr() == v.data_ptr()) else v)
        eng.release()
        return (None, None, None, *grads)


class UNet(nn.Module):
    def __init__(self, num_classes, in_dim=3, conv_dim=64):
        super().__init__()
        self.num_classes = num_classes
        self.in_dim = in_dim
        self.conv_dim = conv_dim
        c = conv_dim
        self.enc1 = nn.Sequential(*_crb(in_dim, c), *_crb(c, c))
        self.enc2 = DownBlock(c, c * 2)
        self.enc3 = DownBlock(c * 2, c * 4)
        self.enc4 = DownBlock(c * 4, c * 8)
        self.dec1 = UpBlock(c * 8, c * 16, c * 8)
        self.dec2 = UpBlock(c * 16, c * 8, c * 4)
        self.dec3 = UpBlock(c * 8, c * 4, c * 2)
        self.dec4 = UpBlock(c * 4, c * 2, c)
        self.last = nn.Sequential(*_crb(c * 2, c), *_crb(c, c), nn.Conv2d(c, num_classes, kernel_size=1, stride=1))
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", UNetEngine(self))
        return self._engine

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("continual_learning_b200.UNet.forward needs a CUDA tensor on an sm_100a device; "
                               "there is no CPU fallback (use the reference model for CPU runs)")
        eng = self.engine
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _UNetFn.apply(x, eng, self.training, *self.parameters())
        logits = eng.forward(x, training=self.training, save_for_backward=False)
        out = logits.permute(0, 3, 1, 2)
        return out

    @torch.no_grad()
    def evaluate_batch(self, x, labels, nc=None, conf=None, correct=None, want_pred=False):
        """statistics / validation step in inference mode (trainer.py:183-189, 271-280): forward with BatchNorm
        folded into the conv epilogues, then the 1x1 head, argmax, the correct-pixel count and the confusion matrix
        (rows = target, `nc` classes) in ONE kernel — the logits are never written.  Accumulates into `conf`
        (int64 [nc*nc]) and `correct` (int64 [1]) when given.  Returns (pred int64 [B,H,W] or None, conf, correct).
        Uses the current mode's statistics: call `.eval()` first for the reference's test() semantics."""
        if not x.is_cuda:
            raise RuntimeError("continual_learning_b200.UNet needs CUDA tensors (sm_100a); there is no CPU fallback")
        eng = self.engine
        if self.conv_dim != 64:  # the one-kernel head needs the reference's 64 head input channels
            logits = eng.forward(x, training=self.training, save_for_backward=False)
            out = ops.argmax_confusion(logits, labels.contiguous(), nc, want_pred=want_pred, conf=conf, correct=correct)
            eng.release()
            return out
        z = eng.forward(x, training=self.training, save_for_backward=False, head=False)
        out = ops.head_argmax_confusion(z, eng.hwf, eng.head.bias.detach(), labels.contiguous(), self.num_classes, nc=nc,
                                        want_pred=want_pred, conf=conf, correct=correct)
        eng.release()
        return out

    def logits_nhwc(self):
        """fp32 [B, H, W, num_classes] logits of the last forward (the memory behind forward()'s view)."""
        return self.engine.logits
