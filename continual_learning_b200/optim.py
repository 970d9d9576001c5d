"""FusedAdam — torch.optim.Adam semantics (reference trainer.py:108-110,176) in one multi-tensor launch.

State keys (`step`, `exp_avg`, `exp_avg_sq`) match torch.optim.Adam so `optimizer_state` in the
reference checkpoints (trainer.py:76,98) round-trips between the two.
"""
import math

import torch

from . import _lib

_CHUNK = 65536


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if weight_decay != 0.0:
            raise ValueError("FusedAdam implements the reference configuration only (weight_decay=0)")
        betas = (float(betas[0]), float(betas[1]))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0))
        self._tables = {}
        self.state_epoch = 0  # bumped whenever the state is replaced from outside (load_state_dict)

    def load_state_dict(self, state_dict):
        """torch.optim.Adam-compatible (trainer.py:98): `TrainStep` re-reads the step count afterwards so that a
        resumed run continues the bias correction where the checkpoint stopped."""
        super().load_state_dict(state_dict)
        self.state_epoch += 1
        self._tables = {}

    def group_step(self, plist):
        """the common optimiser step count of `plist` (0 before the first step); raises if they disagree."""
        seen, steps = set(), set()
        for p in plist:
            t = self.state.get(p, {}).get("step")
            if t is None:
                steps.add(0.0)
            elif id(t) not in seen:  # the step tensor is shared between the parameters: one read in the common case
                seen.add(id(t))
                steps.add(float(t))
        if len(steps) > 1:
            raise RuntimeError("FusedAdam expects all parameters of a group to share the step count")
        return steps.pop() if steps else 0.0

    def set_group_step(self, plist, step):
        t = torch.tensor(float(step))
        for p in plist:
            self.state[p]["step"] = t

    def _table(self, gi, plist):
        key = (gi, tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                          self.state[p]["exp_avg_sq"].data_ptr()) for p in plist))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2], hit[3]
        rows, blocks = [], []
        for ti, p in enumerate(plist):
            st = self.state[p]
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         p.numel()])
            for ch in range((p.numel() + _CHUNK - 1) // _CHUNK):
                blocks.append([ti, ch])
        dev = plist[0].device
        t = torch.tensor(rows, dtype=torch.int64).to(dev)
        b = torch.tensor(blocks, dtype=torch.int32).to(dev)
        self._tables[gi] = (key, t, b, len(blocks))
        return t, b, len(blocks)

    default_grad_scale = 1.0  # set to 1/world by parallel.attach()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None, hyper_dev=None):
        """hyper_dev: optional device float[4] {lr, bc1, bc2_sqrt, gscale} (see `hyper_values`) read by the
        kernel instead of the host scalars — lets a captured CUDA graph be replayed."""
        if grad_scale is None:
            grad_scale = self.default_grad_scale
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            _lib.ensure_device(plist[0].device.index)
            for p in plist:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            step = self.group_step(plist) + 1.0
            self.set_group_step(plist, step)
            b1, b2 = group["betas"]
            bc1 = 1.0 - b1 ** step
            bc2_sqrt = math.sqrt(1.0 - b2 ** step)
            t, b, nblocks = self._table(gi, plist)
            _lib.call("clk_adam_multi_tensor", t, b, nblocks, _CHUNK, float(group["lr"]), b1, b2,
                      float(group["eps"]), bc1, bc2_sqrt, float(grad_scale), hyper_dev)
        _lib.param_epoch += 1  # the kernel wrote the parameters through raw pointers: invalidate packed copies
        return loss

    def hyper_values(self, step, grad_scale=1.0, group=0):
        """[lr, 1-b1^t, sqrt(1-b2^t), grad_scale] for optimiser step number `step` (1-based)."""
        g = self.param_groups[group]
        b1, b2 = g["betas"]
        return [float(g["lr"]), 1.0 - b1 ** step, math.sqrt(1.0 - b2 ** step), float(grad_scale)]
