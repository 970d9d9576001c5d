"""first contact with the CTA-pair conv kernel: correctness on a few shapes (each in its own process), then speed."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one(n, h, w, c0, c1, co):
    import torch, torch.nn.functional as F
    from continual_learning_b200 import _lib, ops
    _lib.ensure_device(0)
    _lib.set_tuning("conv3_v2", 4)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(n, c0 + c1, h, w, generator=g)).to(torch.bfloat16).float()
    wt = torch.randn(co, c0 + c1, 3, 3, generator=g) * 0.05
    b = torch.randn(co, generator=g)
    xh = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    x0 = xh[..., :c0].contiguous(); x1 = xh[..., c0:].contiguous() if c1 else None
    wf, wd = ops.pack_conv3x3(wt.cuda())
    ss = torch.zeros(co, device="cuda", dtype=torch.float64); sq = ss.clone()
    y = ops.conv3x3_fprop(x0, x1, wf, b.cuda(), relu=True, stats=(ss, sq))
    torch.cuda.synchronize()
    ref = torch.relu(F.conv2d(x, wt.to(torch.bfloat16).float(), b, padding=1))
    got = y.float().cpu().permute(0, 3, 1, 2)
    err = float((got - ref).norm() / ref.norm())
    q = y.float().cpu().double().reshape(-1, co)
    es = float((ss.cpu() - q.sum(0)).norm() / q.sum(0).norm())
    dy = torch.randn(n, co, h, w, generator=g).to(torch.bfloat16).float()
    dx0, dx1 = ops.conv3x3_dgrad(dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda(), wd, c0, c1)
    torch.cuda.synchronize()
    refd = F.conv_transpose2d(dy, wt.to(torch.bfloat16).float(), padding=1)
    gd = dx0.float().cpu().permute(0, 3, 1, 2) if dx1 is None else torch.cat([dx0.float().cpu().permute(0, 3, 1, 2), dx1.float().cpu().permute(0, 3, 1, 2)], 1)
    ed = float((gd - refd).norm() / refd.norm())
    print(f"pair kernel n={n} {h}x{w} {c0}+{c1}->{co}: fprop rel {err:.2e} stats {es:.1e} dgrad rel {ed:.2e}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(v) for v in sys.argv[1:]])
        sys.exit(0)
    for case in [(2, 16, 16, 64, 0, 64), (2, 32, 32, 64, 0, 128), (1, 64, 64, 128, 0, 256), (2, 40, 24, 64, 64, 128), (3, 16, 16, 256, 0, 512)]:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(v) for v in case], timeout=120)
        if r.returncode != 0:
            print("FAILED case", case, "rc", r.returncode, flush=True)
