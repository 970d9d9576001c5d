"""Developer smoke check of every libclk kernel against stock PyTorch on the GPU box (not a pytest).

Each group runs in its own subprocess so a faulting kernel cannot poison the others:
    python scripts/dev_check.py            # all groups
    python scripts/dev_check.py gemm conv  # selected groups
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["gemm", "conv", "dgrad", "wgrad", "convT", "head", "bn", "loss", "metrics", "adam", "pack"]


def rel(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, err, tol):
    print(f"  [{'ok' if err <= tol else 'FAIL'}] {name}: rel-L2 {err:.3e} (tol {tol:g})", flush=True)


def run_group(g):
    import torch
    import torch.nn.functional as F

    from continual_learning_b200 import ops

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(1234)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=gen) * scale).to(dev)

    def bf(x):
        return x.to(torch.bfloat16)

    def nhwc(x):  # NCHW fp32 -> NHWC bf16
        return bf(x.permute(0, 2, 3, 1).contiguous())

    def nchw(x):  # NHWC -> NCHW fp32
        return x.float().permute(0, 3, 1, 2).contiguous()

    if g == "gemm":
        for (P, K, N) in [(128, 64, 64), (1000, 128, 128), (4096, 256, 256), (300, 64, 512)]:
            a = bf(rnd(P, K))
            w = bf(rnd(N, K, scale=0.1))
            b = rnd(N)
            s_sum = torch.zeros(N, device=dev, dtype=torch.float64)
            s_sq = torch.zeros(N, device=dev, dtype=torch.float64)
            out = ops.gemm_fprop(a, w, b, N, relu=True, stats=(s_sum, s_sq))
            torch.cuda.synchronize()
            ref = torch.relu(a.float() @ w.float().t() + b)
            report(f"gemm_fprop P={P} K={K} N={N}", rel(out, ref), 5e-3)
            refq = out.float()
            report("  stats sum", rel(s_sum, refq.sum(0)), 1e-5)
            report("  stats sq", rel(s_sq, (refq * refq).sum(0)), 1e-5)
    elif g == "conv":
        for (n, h, w_, c0, c1, co) in [(2, 16, 16, 64, 0, 64), (3, 8, 24, 128, 0, 256), (2, 16, 16, 64, 64, 128),
                                       (5, 4, 4, 128, 128, 64), (1, 32, 32, 64, 0, 64), (33, 2, 2, 64, 0, 128)]:
            x = rnd(n, c0 + c1, h, w_)
            wt = rnd(co, c0 + c1, 3, 3, scale=0.05)
            b = rnd(co)
            xh = nhwc(x)
            x0 = xh[..., :c0].contiguous()
            x1 = xh[..., c0:].contiguous() if c1 else None
            wf, wd = ops.pack_conv3x3(wt)
            s_sum = torch.zeros(co, device=dev, dtype=torch.float64)
            s_sq = torch.zeros(co, device=dev, dtype=torch.float64)
            y = ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(s_sum, s_sq))
            torch.cuda.synchronize()
            ref = torch.relu(F.conv2d(xh.float().permute(0, 3, 1, 2), bf(wt).float(), b, padding=1))
            report(f"conv3x3_fprop n={n} {h}x{w_} {c0}+{c1}->{co}", rel(nchw(y), ref), 5e-3)
            yq = y.float().reshape(-1, co)
            report("  stats sum", rel(s_sum, yq.sum(0)), 1e-5)
            report("  stats sq", rel(s_sq, (yq * yq).sum(0)), 1e-5)
    elif g == "dgrad":
        for (n, h, w_, c0, c1, co) in [(2, 16, 16, 64, 0, 64), (2, 8, 8, 128, 128, 128), (1, 16, 16, 256, 0, 64),
                                       (2, 16, 16, 64, 64, 64)]:
            dy = rnd(n, co, h, w_)
            wt = rnd(co, c0 + c1, 3, 3, scale=0.05)
            wf, wd = ops.pack_conv3x3(wt)
            dyh = nhwc(dy)
            dx0, dx1 = ops.conv3x3_dgrad(dyh, wd, c0, c1)
            torch.cuda.synchronize()
            ref = F.conv_transpose2d(dyh.float().permute(0, 3, 1, 2), bf(wt).float(), padding=1)
            got = nchw(dx0) if dx1 is None else torch.cat([nchw(dx0), nchw(dx1)], 1)
            report(f"conv3x3_dgrad n={n} {h}x{w_} {co}->{c0}+{c1}", rel(got, ref), 5e-3)
    elif g == "wgrad":
        for (n, h, w_, c0, c1, co) in [(2, 16, 16, 64, 0, 64), (2, 8, 8, 128, 0, 256), (4, 16, 16, 64, 64, 128),
                                       (16, 32, 32, 64, 0, 64), (3, 4, 4, 128, 0, 128)]:
            x = rnd(n, c0 + c1, h, w_)
            dy = rnd(n, co, h, w_, scale=0.1)
            xh, dyh = nhwc(x), nhwc(dy)
            x0 = xh[..., :c0].contiguous()
            x1 = xh[..., c0:].contiguous() if c1 else None
            dw = ops.conv3x3_wgrad(dyh, x0, x1)
            torch.cuda.synchronize()
            xr = xh.float().permute(0, 3, 1, 2).requires_grad_(False)
            wref = torch.zeros(co, c0 + c1, 3, 3, device=dev, requires_grad=True)
            F.conv2d(xr, wref, padding=1).backward(dyh.float().permute(0, 3, 1, 2))
            got = dw.reshape(3, 3, c0 + c1, co).permute(3, 2, 0, 1)
            report(f"conv3x3_wgrad n={n} {h}x{w_} {c0}+{c1}->{co}", rel(got, wref.grad), 5e-3)
            g2 = torch.empty(co, c0 + c1, 3, 3, device=dev)
            ops.unpack_wgrad(dw, g2, co, c0 + c1, 9, co, c0 + c1, transposed=True)
            report("  unpack_wgrad", rel(g2, got), 1e-7)
        for (P, cu, ct) in [(1000, 64, 64), (5000, 128, 64), (777, 64, 128)]:
            u, t = bf(rnd(P, cu)), bf(rnd(P, ct))
            out = ops.gemm_wgrad(u, t)
            torch.cuda.synchronize()
            report(f"gemm_wgrad P={P} {cu}x{ct}", rel(out, u.float().t() @ t.float()), 5e-3)
    elif g == "convT":
        for (n, h, w_, ci, co) in [(2, 8, 8, 128, 64), (2, 4, 4, 256, 128), (3, 2, 6, 64, 64), (1, 16, 16, 1024, 512)]:
            x = rnd(n, ci, h, w_)
            wt = rnd(ci, co, 2, 2, scale=0.05)
            b = rnd(co)
            wf, wd = ops.pack_convT(wt)
            xh = nhwc(x)
            y = ops.convT_fprop(xh, wf, b)
            torch.cuda.synchronize()
            ref = F.conv_transpose2d(xh.float().permute(0, 3, 1, 2), bf(wt).float(), b, stride=2)
            report(f"convT_fprop n={n} {h}x{w_} {ci}->{co}", rel(nchw(y), ref), 5e-3)
            dy = rnd(n, co, 2 * h, 2 * w_)
            dyh = nhwc(dy)
            dx = ops.convT_dgrad(dyh, wd)
            torch.cuda.synchronize()
            refdx = F.conv2d(dyh.float().permute(0, 3, 1, 2), bf(wt).float(), stride=2)
            report("  convT_dgrad", rel(nchw(dx), refdx), 5e-3)
            dw = ops.convT_wgrad(xh, dyh)
            torch.cuda.synchronize()
            wref = torch.zeros(ci, co, 2, 2, device=dev, requires_grad=True)
            F.conv_transpose2d(xh.float().permute(0, 3, 1, 2), wref, stride=2).backward(dyh.float().permute(0, 3, 1, 2))
            got = dw.reshape(2, 2, ci, co).permute(2, 3, 0, 1)
            report("  convT_wgrad", rel(got, wref.grad), 5e-3)
    elif g == "head":
        for (n, h, w_, ci, nc) in [(2, 16, 16, 64, 21), (1, 8, 8, 64, 2), (3, 4, 4, 128, 16)]:
            x = nhwc(rnd(n, ci, h, w_))
            wt = rnd(nc, ci, 1, 1, scale=0.1)
            b = rnd(nc)
            wf, wd = ops.pack_head(wt)
            lg = ops.gemm_fprop(x, wf, b, nc, out_f32=True)
            torch.cuda.synchronize()
            ref = x.float() @ bf(wt).float().reshape(nc, ci).t() + b
            report(f"head fprop {ci}->{nc}", rel(lg, ref), 2e-3)
            dl = torch.zeros(n, h, w_, 64, device=dev, dtype=torch.bfloat16)
            dl[..., :nc] = bf(rnd(n, h, w_, nc))
            dz = ops.gemm_fprop(dl, wd, None, ci)
            torch.cuda.synchronize()
            report("  head dgrad", rel(dz, dl[..., :nc].float() @ bf(wt).float().reshape(nc, ci)), 5e-3)
            dw = ops.gemm_wgrad(dl, x)
            torch.cuda.synchronize()
            report("  head wgrad", rel(dw[:nc], dl[..., :nc].float().reshape(-1, nc).t() @ x.float().reshape(-1, ci)), 5e-3)
    elif g == "bn":
        for (n, h, w_, c) in [(2, 16, 16, 64), (3, 8, 8, 256), (2, 4, 4, 1024)]:
            y = bf(torch.relu(rnd(n, h, w_, c)))
            P = n * h * w_
            s_sum = torch.zeros(c, device=dev, dtype=torch.float64)
            s_sq = torch.zeros(c, device=dev, dtype=torch.float64)
            ops.bn_stats(y, s_sum, s_sq)
            gamma, beta = rnd(c) + 1.0, rnd(c)
            rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
            mean, invstd, scale, shift = (torch.empty(c, device=dev) for _ in range(4))
            ops.bn_finalize(s_sum, s_sq, gamma, beta, rm, rv, mean, invstd, scale, shift, P)
            z = ops.bn_apply(y, scale, shift)
            torch.cuda.synchronize()
            bn = torch.nn.BatchNorm2d(c).to(dev)
            with torch.no_grad():
                bn.weight.copy_(gamma)
                bn.bias.copy_(beta)
            yin = y.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            zr = bn(yin)
            report(f"bn fwd c={c}", rel(nchw(z), zr), 5e-3)
            report("  running_mean", rel(rm, bn.running_mean), 1e-5)
            report("  running_var", rel(rv, bn.running_var), 1e-5)
            z2, pooled, idx = ops.bn_apply_pool(y, scale, shift)
            torch.cuda.synchronize()
            pr, ir = F.max_pool2d(z2.float().permute(0, 3, 1, 2), 2, 2, return_indices=True)
            report("  pool value", rel(nchw(pooled), pr), 1e-7)
            hh, ww = ir // w_ - 2 * torch.arange(h // 2, device=dev).view(1, 1, -1, 1), ir % w_ - 2 * torch.arange(w_ // 2, device=dev).view(1, 1, 1, -1)
            report("  pool index", float((idx.permute(0, 3, 1, 2).long() != hh * 2 + ww).float().mean()), 0.0)
            dz = bf(rnd(n, h, w_, c))
            s1 = torch.zeros(c, device=dev, dtype=torch.float64)
            s2 = torch.zeros(c, device=dev, dtype=torch.float64)
            ops.bn_bwd_reduce(dz, y, s1, s2)
            dgamma, dbeta, ka, kb, kc = (torch.empty(c, device=dev) for _ in range(5))
            ops.bn_bwd_finalize(s1, s2, gamma, mean, invstd, dgamma, dbeta, ka, kb, kc, P)
            dbias = torch.zeros(c, device=dev, dtype=torch.float64)
            dpre = ops.bn_relu_bwd_apply(dz, y, ka, kb, kc, dbias)
            torch.cuda.synchronize()
            zr.backward(dz.float().permute(0, 3, 1, 2))
            refd = yin.grad * (yin > 0)
            report("  bn bwd dx*relu", rel(nchw(dpre), refd), 1e-2)
            report("  dgamma", rel(dgamma, bn.weight.grad), 1e-3)
            report("  dbeta", rel(dbeta, bn.bias.grad), 1e-3)
            report("  dbias", rel(dbias, refd.sum((0, 2, 3))), 1e-2)
            dp = bf(rnd(n, h // 2, w_ // 2, c))
            skip = bf(rnd(n, h, w_, c))
            din = ops.maxpool_bwd_add(dp, idx, skip)
            torch.cuda.synchronize()
            refp = F.max_unpool2d(dp.float().permute(0, 3, 1, 2), ir, 2, 2) + skip.float().permute(0, 3, 1, 2)
            report("  maxpool_bwd_add", rel(nchw(din), refp), 5e-3)
    elif g == "loss":
        for (P, c, cold) in [(1000, 21, 0), (4096, 21, 16), (513, 2, 0)]:
            z = rnd(P, c, scale=2.0).requires_grad_(True)
            lab = torch.randint(0, c, (P,), generator=gen).to(dev)
            zo = rnd(P, cold, scale=2.0) if cold else None
            T, lam = 2.0, 1.0
            acc, dl = ops.ce_kd_loss(z.detach(), lab, zo, T=T, lam=lam)
            torch.cuda.synchronize()
            loss = F.cross_entropy(z, lab)
            if cold:
                loss = loss + lam * T * T * F.kl_div(F.log_softmax(z[:, :cold] / T, 1), F.softmax(zo / T, 1), reduction="sum") / P
            loss.backward()
            got = (acc[0] + (lam * T * T * acc[1] if cold else 0)) / P
            report(f"loss P={P} C={c} Cold={cold}", abs(float(got) - float(loss)) / abs(float(loss)), 1e-5)
            report("  dlogits", rel(dl[:, :c], z.grad), 5e-3)
            report("  dlogits pad", float(dl[:, c:].float().abs().max()), 0.0)
    elif g == "metrics":
        for (n, nc) in [(100001, 22), (65536, 21), (7, 3)]:
            t = torch.randint(-1, nc + 1, (n,), generator=gen).to(dev)
            p = torch.randint(0, nc, (n,), generator=gen).to(dev)
            conf = ops.confusion_matrix(t, p, nc)
            m = (t >= 0) & (t < nc)
            ref = torch.bincount(nc * t[m] + p[m], minlength=nc * nc)
            torch.cuda.synchronize()
            report(f"confusion n={n} nc={nc}", float((conf != ref).sum()), 0.0)
        lg = rnd(5000, 21)
        lab = torch.randint(0, 21, (5000,), generator=gen).to(dev)
        pred, conf, correct = ops.argmax_confusion(lg, lab, nc=22, want_pred=True)
        torch.cuda.synchronize()
        rp = lg.argmax(1)
        report("argmax pred", float((pred != rp).sum()), 0.0)
        report("argmax conf", float((conf != torch.bincount(22 * lab + rp, minlength=484)).sum()), 0.0)
        report("argmax correct", abs(int(correct) - int((rp == lab).sum())), 0.0)
    elif g == "adam":
        from continual_learning_b200.optim import FusedAdam
        ps = [torch.nn.Parameter(rnd(*s)) for s in [(64, 3, 3, 3), (64,), (1024, 1024, 3, 3), (21, 64, 1, 1), (7,)]]
        qs = [torch.nn.Parameter(p_.detach().clone()) for p_ in ps]
        o1 = FusedAdam(ps, lr=1e-3, betas=(0.5, 0.99))
        o2 = torch.optim.Adam(qs, lr=1e-3, betas=(0.5, 0.99))
        for it in range(3):
            for a, b in zip(ps, qs):
                gr = rnd(*a.shape)
                a.grad = gr.clone()
                b.grad = gr.clone()
            o1.step()
            o2.step()
        torch.cuda.synchronize()
        for a, b in zip(ps, qs):
            report(f"adam {tuple(a.shape)}", rel(a, b), 1e-6)
    elif g == "pack":
        x = rnd(2, 3, 16, 16)
        a = ops.im2col_stem(x)
        ref = F.unfold(x, 3, padding=1).permute(0, 2, 1).reshape(2, 16, 16, 27)
        torch.cuda.synchronize()
        report("im2col stem", rel(a[..., :27], bf(ref).float()), 1e-7)
        report("im2col pad", float(a[..., 27:].float().abs().max()), 0.0)
        wt = rnd(128, 64, 3, 3)
        wf, wd = ops.pack_conv3x3(wt)
        report("pack wf", rel(wf, bf(wt.permute(2, 3, 0, 1).reshape(9, 128, 64)).float()), 1e-7)
        report("pack wd", rel(wd, bf(wt.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 64, 128)).float()), 1e-7)
        xx = nhwc(rnd(2, 64, 8, 8))
        report("nhwc->nchw", rel(ops.nhwc_to_nchw_f32(xx), xx.float().permute(0, 3, 1, 2)), 1e-7)
        report("nchw->nhwc", rel(ops.nchw_to_nhwc_bf16(x, 8)[..., :3], bf(x.permute(0, 2, 3, 1)).float()), 1e-7)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--group":
        run_group(sys.argv[2])
        sys.exit(0)
    groups = sys.argv[1:] or GROUPS
    bad = 0
    for g in groups:
        print(f"== {g}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=300)
            if r.returncode != 0:
                print(f"  [FAIL] group {g} exited with {r.returncode}", flush=True)
                bad += 1
        except subprocess.TimeoutExpired:
            print(f"  [FAIL] group {g} timed out", flush=True)
            bad += 1
    sys.exit(1 if bad else 0)
