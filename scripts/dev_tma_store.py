"""generic GEMM kernel: TMA-store epilogue on/off — correctness + timing for ConvTranspose2d forward / dgrad and the stem GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from continual_learning_b200 import _lib, ops
_lib.ensure_device(0)
bf16 = torch.bfloat16
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


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


g = torch.Generator(device="cuda").manual_seed(0)
for (n, h, w, ci, co) in [(2, 8, 8, 128, 64), (2, 4, 4, 256, 128), (3, 2, 6, 64, 64), (1, 16, 16, 1024, 512), (2, 5, 16, 64, 64),
                          (16, 16, 16, 1024, 512), (16, 32, 32, 512, 256), (16, 64, 64, 256, 128), (16, 128, 128, 128, 64)]:
    x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(bf16)
    dy = torch.randn(n, 2 * h, 2 * w, co, device="cuda", generator=g).to(bf16)
    wt = (torch.randn(ci, co, 2, 2, device="cuda", generator=g) * 0.05).to(bf16).float()
    b = torch.randn(co, device="cuda", generator=g)
    wf, wd = ops.pack_convT(wt)
    ref_y = F.conv_transpose2d(nchw(x), wt, b, stride=2)
    ref_dx = F.conv2d(nchw(dy), wt, stride=2)
    line = f"convT {(n, h, w, ci, co)}:"
    for ts in (1, 0):
        _lib.set_tuning("tma_store", ts)
        y = ops.convT_fprop(x, wf, b)
        dx = ops.convT_dgrad(dy, wd)
        line += f"  tma_store={ts} fprop rel {rel(nchw(y), ref_y):.2e} dgrad rel {rel(nchw(dx), ref_dx):.2e}"
        if n == 16:
            yo, dxo = torch.empty_like(y), torch.empty_like(dx)
            line += f" fprop {timeit(lambda: ops.convT_fprop(x, wf, b, out=yo)):.1f} us dgrad {timeit(lambda: ops.convT_dgrad(dy, wd, out=dxo)):.1f} us"
    print(line, flush=True)
for P, K, N in [(1000, 64, 64), (4096, 256, 256), (300, 64, 512), (1, 64, 64), (16 * 256 * 256, 64, 64)]:
    a = torch.randn(P, K, device="cuda", generator=g).to(bf16)
    wm = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(bf16)
    b = torch.randn(N, device="cuda", generator=g)
    ref = torch.relu(a.float() @ wm.float().t() + b)
    line = f"gemm {(P, K, N)}:"
    for ts in (1, 0):
        _lib.set_tuning("tma_store", ts)
        s1 = torch.zeros(N, device="cuda", dtype=torch.float64)
        s2 = torch.zeros_like(s1)
        out = ops.gemm_fprop(a, wm, b, N, relu=True, stats=(s1, s2))
        line += f"  tma_store={ts} rel {rel(out, ref):.2e} stats {rel(s1, out.double().sum(0)):.1e}"
        if P > 100000:
            o2 = torch.empty_like(out)
            line += f" {timeit(lambda: ops.gemm_fprop(a, wm, b, N, relu=True, stats=(s1, s2), out=o2)):.1f} us"
    print(line, flush=True)
