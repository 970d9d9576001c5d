import sys, os
sys.path.insert(0, '/root/repo')
import torch
from continual_learning_b200 import _lib, ops
_lib.ensure_device(0)
bf16=torch.bfloat16
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
B=16
for (c0,co,d) in [(64,64,1),(128,128,2),(256,256,4)]:
    h=w=256//d
    x=torch.randn(B,h,w,c0,device='cuda').to(bf16)
    wt=torch.randn(co,c0,3,3,device='cuda')*0.05
    wf,wd=ops.pack_conv3x3(wt)
    bias=torch.zeros(co,device='cuda')
    y=torch.empty(B,h,w,co,device='cuda',dtype=bf16)
    ss=torch.zeros(co,device='cuda',dtype=torch.float64); sq=ss.clone()
    gf=2.0*B*h*w*co*c0*9/1e9
    for name,relu,stats in (("full",1,(ss,sq)),("no stats",1,None),("mainloop only",77,None)):
        ms=timeit(lambda: _lib.call("clk_conv3x3_fprop", x, c0, None, 0, wf, bias, y, stats[0] if stats else None, stats[1] if stats else None, B,h,w,co, relu))
        print(f"{c0}->{co}@{h}: {name:14s} {ms:.3f} ms {gf/ms:.0f} TF/s")
