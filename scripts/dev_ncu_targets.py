"""Launch a fixed list of representative kernels at the benchmark shapes (each twice: warm-up + profiled) so that
`ncu --set full -k regex:... ` can capture them in one short run.  Developer tool (run under gpurun).

    python scripts/dev_ncu_targets.py [names...]     # default: all
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from continual_learning_b200 import _lib, ops

bf16 = torch.bfloat16
dev = "cuda"
B, S = 16, 256


def t(*shape):
    return torch.randn(*shape, device=dev).to(bf16)


def conv_case(c0, c1, co, d):
    h = w = S // d
    x0, x1 = t(B, h, w, c0), (t(B, h, w, c1) if c1 else None)
    wt = torch.randn(co, c0 + c1, 3, 3, device=dev) * 0.05
    wf, wd = ops.pack_conv3x3(wt)
    dy = t(B, h, w, co)
    return x0, x1, wf, wd, dy, torch.zeros(co, device=dev), h, w


def main():
    _lib.ensure_device(0)
    want = set(sys.argv[1:])
    cases = {}

    def case(name):
        def deco(fn):
            cases[name] = fn
            return fn
        return deco

    @case("fprop64")
    def _():
        x0, x1, wf, wd, dy, b, h, w = conv_case(64, 0, 64, 1)
        ss = torch.zeros(64, device=dev, dtype=torch.float64)
        return lambda: ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(ss, ss.clone()))

    @case("fprop128")
    def _():
        x0, x1, wf, wd, dy, b, h, w = conv_case(128, 0, 128, 2)
        ss = torch.zeros(128, device=dev, dtype=torch.float64)
        return lambda: ops.conv3x3_fprop(x0, x1, wf, b, relu=True, stats=(ss, ss.clone()))

    @case("dgrad64")
    def _():
        x0, x1, wf, wd, dy, b, h, w = conv_case(64, 0, 64, 1)
        return lambda: ops.conv3x3_dgrad(dy, wd, 64, 0)

    @case("wgrad64")
    def _():
        x0, x1, wf, wd, dy, b, h, w = conv_case(64, 0, 64, 1)
        dw = torch.zeros(9, 64, 64, device=dev)
        return lambda: ops.conv3x3_wgrad(dy, x0, None, out=dw)

    @case("wgrad512")
    def _():
        x0, x1, wf, wd, dy, b, h, w = conv_case(512, 0, 512, 8)
        dw = torch.zeros(9, 512, 512, device=dev)
        return lambda: ops.conv3x3_wgrad(dy, x0, None, out=dw)

    @case("stem")
    def _():
        a = t(B, S, S, 64)
        wf = t(64, 64)
        ss = torch.zeros(64, device=dev, dtype=torch.float64)
        b = torch.zeros(64, device=dev)
        return lambda: ops.gemm_fprop(a, wf, b, 64, relu=True, stats=(ss, ss.clone()))

    @case("stem_direct")
    def _():
        x = torch.randn(B, 3, S, S, device=dev)
        wf = ops.pack_stem(torch.randn(64, 3, 3, 3, device=dev) * 0.2)
        ss = torch.zeros(64, device=dev, dtype=torch.float64)
        b = torch.zeros(64, device=dev)
        return lambda: ops.stem_conv(x, wf, b, relu=True, stats=(ss, ss.clone()))

    @case("head")
    def _():
        a = t(B, S, S, 64)
        wf = t(32, 64)
        b = torch.zeros(21, device=dev)
        return lambda: ops.gemm_fprop(a, wf, b, 21, out_f32=True)

    @case("convT_wgrad")
    def _():
        x = t(B, S // 2, S // 2, 128)
        dy = t(B, S, S, 64)
        dw = torch.zeros(4, 128, 64, device=dev)
        return lambda: ops.convT_wgrad(x, dy, out=dw)

    @case("convT_fprop")
    def _():
        x = t(B, S // 2, S // 2, 128)
        wf = t(256, 128)
        b = torch.zeros(64, device=dev)
        return lambda: ops.convT_fprop(x, wf, b)

    @case("bn_bwd")
    def _():
        y, dz = t(B, S, S, 64), t(B, S, S, 64)
        k = torch.ones(64, device=dev)
        db = torch.zeros(64, device=dev, dtype=torch.float64)
        out = torch.empty_like(y)
        return lambda: ops.bn_relu_bwd_apply(dz, y, k, k, k, db, out=out)

    @case("bn_reduce")
    def _():
        y, dz = t(B, S, S, 64), t(B, S, S, 64)
        s1 = torch.zeros(64, device=dev, dtype=torch.float64)
        return lambda: ops.bn_bwd_reduce(dz, y, s1, s1.clone())

    @case("bn_apply")
    def _():
        y = t(B, S, S, 64)
        k = torch.ones(64, device=dev)
        out = torch.empty_like(y)
        return lambda: ops.bn_apply(y, k, k, out=out)

    @case("bn_apply_pool")
    def _():
        y = t(B, S, S, 64)
        k = torch.ones(64, device=dev)
        return lambda: ops.bn_apply_pool(y, k, k)

    @case("head_loss")
    def _():
        z = t(B, S, S, 64)
        wf, wd = t(32, 64), t(64, 64)
        b = torch.zeros(21, device=dev)
        y = torch.randint(0, 21, (B, S, S), device=dev)
        dz = torch.empty_like(z)
        dw = torch.zeros(64, 64, device=dev)
        db = torch.zeros(64, device=dev, dtype=torch.float64)
        la = torch.zeros(2, device=dev, dtype=torch.float64)
        return lambda: ops.head_loss_bwd(z, wf, wd, b, y, 21, dz=dz, dw=dw, dbias=db, loss_acc=la)

    # ---- HBM-bound kernels of the statistics / loss / input paths at the config-2 / config-5 sizes (16 x 256 x 256)
    @case("confusion")
    def _():
        tgt = torch.randint(0, 21, (B, S, S), device=dev)
        prd = torch.randint(0, 21, (B, S, S), device=dev)
        conf = torch.zeros(21 * 21, device=dev, dtype=torch.int64)
        return lambda: ops.confusion_matrix(tgt, prd, 21, conf=conf)

    @case("argmax_confusion")
    def _():
        lg = torch.randn(B, S, S, 21, device=dev)
        y = torch.randint(0, 21, (B, S, S), device=dev)
        conf = torch.zeros(22 * 22, device=dev, dtype=torch.int64)
        ok = torch.zeros(1, device=dev, dtype=torch.int64)
        return lambda: ops.argmax_confusion(lg, y, nc=22, want_pred=True, conf=conf, correct=ok)

    @case("head_argmax")
    def _():
        z = t(B, S, S, 64)
        wf = t(32, 64)
        b = torch.zeros(21, device=dev)
        y = torch.randint(0, 21, (B, S, S), device=dev)
        conf = torch.zeros(21 * 21, device=dev, dtype=torch.int64)
        ok = torch.zeros(1, device=dev, dtype=torch.int64)
        return lambda: ops.head_argmax_confusion(z, wf, b, y, 21, nc=21, conf=conf, correct=ok)

    @case("ce_kd_loss")
    def _():
        lg = torch.randn(B, S, S, 21, device=dev)
        old = torch.randn(B, S, S, 16, device=dev)
        y = torch.randint(0, 21, (B, S, S), device=dev)
        dl = torch.zeros(B, S, S, 64, device=dev, dtype=bf16)
        la = torch.zeros(2, device=dev, dtype=torch.float64)
        return lambda: ops.ce_kd_loss(lg, y, old, T=2.0, lam=1.0, dlogits=dl, loss_acc=la)

    @case("maxpool_bwd")
    def _():
        dp = t(B, S // 2, S // 2, 64)
        skip, yy = t(B, S, S, 64), t(B, S, S, 64)
        idx = torch.randint(0, 4, (B, S // 2, S // 2, 64), device=dev, dtype=torch.uint8)
        s1 = torch.zeros(64, device=dev, dtype=torch.float64)
        out = torch.empty_like(skip)
        return lambda: ops.maxpool_bwd_add_reduce(dp, idx, skip, yy, s1, s1.clone(), out=out)

    @case("im2col_stem")
    def _():
        x = torch.randn(B, 3, S, S, device=dev)
        return lambda: ops.im2col_stem(x)

    @case("adam")
    def _():
        import continual_learning_b200 as clk
        ps = [torch.nn.Parameter(torch.randn(31_044_821, device=dev))]
        ps[0].grad = torch.randn_like(ps[0])
        opt = clk.FusedAdam(ps, lr=1e-4, betas=(0.5, 0.99))
        return lambda: opt.step()

    for name, make in cases.items():
        if want and name not in want:
            continue
        fn = make()
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"ran {name}: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
