"""What the BatchNorm statistics cost in the conv epilogue: forward conv with / without the column sums at the
benchmark layer shapes.  Developer tool (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from continual_learning_b200 import _lib, ops
_lib.ensure_device(0)
bf16 = torch.bfloat16
def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for (c0, co, d) in ((64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4), (512, 512, 8)):
    h = 256 // d
    x = torch.randn(16, h, h, c0, device="cuda").to(bf16)
    wt = torch.randn(co, c0, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_conv3x3(wt)
    b = torch.zeros(co, device="cuda")
    s, q = torch.zeros(co, device="cuda", dtype=torch.float64), torch.zeros(co, device="cuda", dtype=torch.float64)
    y = torch.empty(16, h, h, co, device="cuda", dtype=bf16)
    t1 = timeit(lambda: ops.conv3x3_fprop(x, None, wf, b, relu=True, stats=(s, q), out=y))
    t2 = timeit(lambda: ops.conv3x3_fprop(x, None, wf, b, relu=True, stats=None, out=y))
    t3 = timeit(lambda: ops.conv3x3_fprop(x, None, wf, None, relu=False, stats=None, out=y))
    print(f"{c0}->{co} @{h}: stats {t1*1e3:.1f} us, no stats {t2*1e3:.1f} us, no bias/relu/stats {t3*1e3:.1f} us")
