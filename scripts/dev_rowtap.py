"""paired-tap kernel for 64-output-channel conv3x3 layers (igemm_conv3r_kernel) vs the tap-by-tap pair kernel:
correctness on odd shapes + timing at the benchmark shapes.  usage: python scripts/dev_rowtap.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from continual_learning_b200 import _lib, ops

_lib.ensure_device(0)
torch.backends.cudnn.allow_tf32 = False
bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


def check(n, h, w, c0, c1, co, kind):
    g = torch.Generator(device="cuda").manual_seed(n + h + w + c0)
    if kind == "fprop":
        ci = c0 + c1
        x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(bf16)
        wt = (torch.randn(co, ci, 3, 3, device="cuda", generator=g) * 0.05).to(bf16).float()
        b = torch.randn(co, device="cuda", generator=g) * 0.1
        wf, _ = ops.pack_conv3x3(wt)
        x0 = x[..., :c0].contiguous()
        x1 = x[..., c0:].contiguous() if c1 else None
        s1 = torch.zeros(co, device="cuda", dtype=torch.float64)
        s2 = torch.zeros_like(s1)
        y = ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(s1, s2))
        ref = torch.relu(F.conv2d(nchw(x), wt, b, padding=1))
        yd = y.double().reshape(-1, co)
        e = rel(nchw(y), ref)
        es = max(rel(s1, yd.sum(0)), rel(s2, (yd * yd).sum(0)))
        return e, es
    else:  # dgrad: dy has co channels, dx has c0 (=64) channels
        dy = torch.randn(n, h, w, co, device="cuda", generator=g).to(bf16)
        wt = (torch.randn(co, c0, 3, 3, device="cuda", generator=g) * 0.05).to(bf16).float()
        _, wd = ops.pack_conv3x3(wt)
        dx0, _ = ops.conv3x3_dgrad(dy, wd, c0, 0)
        ref = F.conv_transpose2d(nchw(dy), wt, padding=1)
        return rel(nchw(dx0), ref), 0.0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


shapes = [(2, 16, 16, 64, 0, 64), (1, 1, 1, 64, 0, 64), (2, 3, 5, 64, 64, 64), (1, 64, 64, 64, 0, 64), (2, 40, 72, 64, 64, 64),
          (1, 33, 61, 64, 0, 64), (3, 8, 30, 64, 0, 64), (1, 9, 31, 64, 64, 64), (16, 256, 256, 64, 0, 64),
          (16, 256, 256, 64, 64, 64)]
for rt in (1, 0):
    _lib.set_tuning("conv3_rowtap", rt)
    for sh in shapes:
        print(f"rowtap={rt} fprop {sh}: rel %.2e stats %.2e" % check(*sh, "fprop"), flush=True)
    for sh in [(2, 16, 16, 64, 0, 64), (2, 6, 10, 64, 0, 128), (1, 64, 48, 64, 0, 64), (16, 128, 128, 64, 0, 128),
               (16, 256, 256, 64, 0, 64)]:
        print(f"rowtap={rt} dgrad {sh}: rel %.2e" % check(*sh, "dgrad")[0], flush=True)

# timing at the benchmark shapes
for (n, h, w, c0, c1, co, kind) in [(16, 256, 256, 64, 0, 64, "fprop"), (16, 256, 256, 64, 64, 64, "fprop"),
                                    (16, 256, 256, 64, 0, 64, "dgrad"), (16, 128, 128, 64, 0, 128, "dgrad")]:
    g = torch.Generator(device="cuda").manual_seed(1)
    if kind == "fprop":
        ci = c0 + c1
        x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(bf16)
        wt = torch.randn(co, ci, 3, 3, device="cuda", generator=g) * 0.05
        wf, _ = ops.pack_conv3x3(wt)
        x0 = x[..., :c0].contiguous()
        x1 = x[..., c0:].contiguous() if c1 else None
        b = torch.zeros(co, device="cuda")
        s1 = torch.zeros(co, device="cuda", dtype=torch.float64)
        s2 = torch.zeros_like(s1)
        y = torch.empty(n, h, w, co, device="cuda", dtype=bf16)
        fn = lambda: ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(s1, s2), out=y)
        fl = 2.0 * n * h * w * co * ci * 9
    else:
        dy = torch.randn(n, h, w, co, device="cuda", generator=g).to(bf16)
        wt = torch.randn(co, c0, 3, 3, device="cuda", generator=g) * 0.05
        _, wd = ops.pack_conv3x3(wt)
        dx = torch.empty(n, h, w, c0, device="cuda", dtype=bf16)
        fn = lambda: ops.conv3x3_dgrad(dy, wd, c0, 0, out0=dx)
        fl = 2.0 * n * h * w * co * c0 * 9
    for rt in (0, 1):
        _lib.set_tuning("conv3_rowtap", rt)
        us = timeit(fn)
        print(f"{kind} {(n, h, w, c0, c1, co)} rowtap={rt}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
