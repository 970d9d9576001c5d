"""first contact with the CTA-pair wgrad kernel (each case in its own process)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one(n, h, w, c0, c1, co):
    import torch, torch.nn.functional as F
    from continual_learning_b200 import _lib, ops
    _lib.ensure_device(0)
    _lib.set_tuning("wgrad_v2", 2)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, c0 + c1, h, w, generator=g).to(torch.bfloat16).float()
    dy = (torch.randn(n, co, h, w, generator=g) * 0.1).to(torch.bfloat16).float()
    xh = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    x0 = xh[..., :c0].contiguous(); x1 = xh[..., c0:].contiguous() if c1 else None
    dw = ops.conv3x3_wgrad(dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda(), x0, x1)
    torch.cuda.synchronize()
    wref = torch.zeros(co, c0 + c1, 3, 3, requires_grad=True)
    F.conv2d(x, wref, padding=1).backward(dy)
    got = dw.reshape(3, 3, c0 + c1, co).permute(3, 2, 0, 1).cpu()
    err = float((got - wref.grad).norm() / wref.grad.norm())
    per_tap = [(float((got[:, :, r, s] - wref.grad[:, :, r, s]).norm() / wref.grad[:, :, r, s].norm())) for r in range(3) for s in range(3)]
    print(f"pair wgrad n={n} {h}x{w} {c0}+{c1}->{co}: rel {err:.2e} per-tap {['%.0e' % e for e in per_tap]}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(v) for v in sys.argv[1:]])
        sys.exit(0)
    for case in [(2, 16, 16, 64, 0, 128), (2, 32, 32, 128, 0, 128), (1, 64, 64, 128, 0, 256), (2, 40, 24, 64, 64, 128), (3, 16, 16, 256, 0, 512)]:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(v) for v in case], timeout=120)
        if r.returncode != 0:
            print("FAILED case", case, "rc", r.returncode, flush=True)
