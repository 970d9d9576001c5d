"""direct stem kernel vs im2col + GEMM: timing at 16x3x256x256"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from continual_learning_b200 import _lib, ops
_lib.ensure_device(0)
x = torch.randn(16, 3, 256, 256, device="cuda")
wf = ops.pack_stem(torch.randn(64, 3, 3, 3, device="cuda") * 0.2)
b = torch.zeros(64, device="cuda")
s1, s2 = (torch.zeros(64, device="cuda", dtype=torch.float64) for _ in range(2))
y = torch.empty(16, 256, 256, 64, device="cuda", dtype=torch.bfloat16)
def timeit(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
print(f"direct stem kernel: {timeit(lambda: ops.stem_conv(x, wf, b, relu=True, stats=(s1, s2), out=y)):.1f} us")
print(f"im2col + GEMM:      {timeit(lambda: ops.gemm_fprop(ops.im2col_stem(x), wf, b, 64, relu=True, stats=(s1, s2), out=y)):.1f} us")
sc, sh = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
print(f"direct, no statistics (inference affine): {timeit(lambda: ops.stem_conv(x, wf, b, relu=True, scale=sc, shift=sh, out=y)):.1f} us")
print(f"direct, no statistics, no affine:         {timeit(lambda: ops.stem_conv(x, wf, b, relu=True, out=y)):.1f} us")
