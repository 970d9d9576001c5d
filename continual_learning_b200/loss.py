"""CrossEntropyDistillLoss — drop-in for `nn.CrossEntropyLoss()` (reference trainer.py:113,174) with the
continual-learning distillation term fused in (not in the reference).

    c_loss = CrossEntropyDistillLoss()                      # plain CE, mean over B*H*W
    c_loss = CrossEntropyDistillLoss(old_model, T=2, lam=1) # + lam*T^2*KL(softmax(z_old/T) || softmax(z[:, :C_old]/T))
    loss = c_loss(outputs, labels); loss.backward()

Forward and backward of the loss are ONE pass over the logits (clk_ce_kd_loss); the bf16 dlogits go
straight to the U-Net backward when `outputs` came from continual_learning_b200.UNet.
"""
import torch
import torch.nn as nn

from . import ops


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels, old_logits, T, lam, owner):
        z = outputs.permute(0, 2, 3, 1)  # NHWC view
        if not z.is_contiguous():
            z = z.contiguous()
        z = z.float()
        c = z.shape[-1]
        p = z.numel() // c
        err = torch.zeros(1, device=z.device, dtype=torch.int32)
        acc, dl = ops.ce_kd_loss(z, labels.contiguous(), old_logits, T=T, lam=lam, err_flag=err)
        owner.last_error_flag = err
        loss = acc[0] / p
        if old_logits is not None:
            loss = loss + (lam * T * T / p) * acc[1]
        ctx.dl = dl
        ctx.shape = outputs.shape
        ctx.engine = getattr(owner, "_engine_hint", None)
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        if g.numel() == 1:
            dl.mul_(g.to(dl.dtype))  # in place: the loss owns this buffer
        eng = ctx.engine
        if eng is not None:
            # the real gradient travels next to autograd as bf16 [B,H,W,64]; the returned token is all zeros so that
            # a second consumer of `outputs` (auxiliary loss, regulariser) still sums to the right thing
            token = torch.zeros((1,), device=dl.device, dtype=torch.float32).expand(ctx.shape)
            eng._pending_dlogits = (token, dl)
            return token, None, None, None, None, None
        c = ctx.shape[1]
        return dl[..., :c].float().permute(0, 3, 1, 2), None, None, None, None, None


class CrossEntropyDistillLoss(nn.Module):
    def __init__(self, old_model=None, T=2.0, lam=1.0):
        super().__init__()
        object.__setattr__(self, "old_model", old_model)  # frozen: not registered, not trained, not saved
        self.T, self.lam = float(T), float(lam)
        self.last_error_flag = None
        self._engine_hint = None
        self._old_logits = None

    def check_labels(self):
        """nn.CrossEntropyLoss raises on a label outside [0, C) (trainer.py:174); the fused kernel records it in a
        device flag instead of stalling the stream.  Call where a host sync happens anyway (Trainer does, in its
        every-10th-iteration statistics block): raises IndexError like the reference's CPU loss."""
        flag = self.last_error_flag
        if flag is not None and int(flag.item()) != 0:
            self.last_error_flag = None
            raise IndexError("Target out of bounds: a label outside [0, num_classes) reached the loss")

    def observe(self, inputs):
        """run the frozen previous-task network (eval mode, no grad) on this batch's inputs."""
        if self.old_model is None:
            return
        with torch.no_grad():
            was = self.old_model.training
            self.old_model.eval()
            self.old_model(inputs)
            self._old_logits = self.old_model.logits_nhwc()
            self.old_model.engine.release()
            self.old_model.train(was)

    def forward(self, outputs, labels):
        if not outputs.is_cuda:
            raise RuntimeError("CrossEntropyDistillLoss runs on CUDA (sm_100a) only; no CPU fallback")
        old = None
        if self.old_model is not None:
            if self._old_logits is None:
                raise RuntimeError("call c_loss.observe(inputs) before the loss when distilling from an old model")
            old, self._old_logits = self._old_logits, None
        fn = outputs.grad_fn
        self._engine_hint = getattr(fn, "engine", None) if fn is not None else None
        return _LossFn.apply(outputs, labels, old, self.T, self.lam, self)
