"""Per-layer timing of the igemm kernels at the benchmark shape (B=16, 256x256): TFLOP/s per layer for
fprop / dgrad / wgrad, plus GB/s of the HBM-bound kernels.  Developer tool (run under gpurun).

    python scripts/dev_perf.py [--batch 16] [--size 256] [--reps 5] [--tune key=value ...] [--only fprop,dgrad,wgrad,mem]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from continual_learning_b200 import _lib, ops

bf16 = torch.bfloat16


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    ts = []
    for _ in range(reps):
        flush.zero_()  # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tune", nargs="*", default=[])
    ap.add_argument("--only", default="fprop,dgrad,wgrad,convT,mem")
    args = ap.parse_args()
    _lib.ensure_device(0)
    for kv in args.tune:
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))
    only = set(args.only.split(","))
    B, S = args.batch, args.size
    dev = "cuda"
    c = 64
    layers = [("enc1.3", c, 0, c, 1)]
    for nm, ci, co, d in (("enc2", c, 2 * c, 2), ("enc3", 2 * c, 4 * c, 4), ("enc4", 4 * c, 8 * c, 8)):
        layers += [(nm + ".1", ci, 0, co, d), (nm + ".4", co, 0, co, d)]
    layers += [("dec1.0", 8 * c, 0, 16 * c, 16), ("dec1.3", 16 * c, 0, 16 * c, 16)]
    for nm, ch, cm, d in (("dec2", 8 * c, 8 * c, 8), ("dec3", 4 * c, 4 * c, 4), ("dec4", 2 * c, 2 * c, 2)):
        layers += [(nm + ".0", ch, ch, cm, d), (nm + ".3", cm, 0, cm, d)]
    layers += [("last.0", c, c, c, 1), ("last.3", c, 0, c, 1)]
    tot = {"fprop": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    print(f"{'layer':10s} {'shape':28s} {'GFLOP':>8s} | {'fprop ms':>9s} {'TF/s':>7s} | {'dgrad ms':>9s} {'TF/s':>7s} | {'wgrad ms':>9s} {'TF/s':>7s}")
    for nm, c0, c1, co, d in layers:
        h = w = S // d
        x0 = torch.randn(B, h, w, c0, device=dev).to(bf16)
        x1 = torch.randn(B, h, w, c1, device=dev).to(bf16) if c1 else None
        wt = torch.randn(co, c0 + c1, 3, 3, device=dev) * 0.05
        wf, wd = ops.pack_conv3x3(wt)
        bias = torch.zeros(co, device=dev)
        ss = torch.zeros(co, device=dev, dtype=torch.float64)
        sq = torch.zeros(co, device=dev, dtype=torch.float64)
        y = torch.empty(B, h, w, co, device=dev, dtype=bf16)
        dy = torch.randn(B, h, w, co, device=dev).to(bf16)
        dx0 = torch.empty_like(x0)
        dx1 = torch.empty_like(x1) if c1 else None
        dw = torch.zeros(9, c0 + c1, co, device=dev)
        gf = 2.0 * B * h * w * co * (c0 + c1) * 9 / 1e9
        row = f"{nm:10s} {f'{c0}+{c1}->{co} @{h}x{w}':28s} {gf:8.1f} |"
        for kind, fn in (("fprop", lambda: ops.conv3x3_fprop(x0, x1, wf, bias, relu=True, stats=(ss, sq), out=y)),
                         ("dgrad", lambda: ops.conv3x3_dgrad(dy, wd, c0, c1, out0=dx0, out1=dx1)),
                         ("wgrad", lambda: ops.conv3x3_wgrad(dy, x0, x1, out=dw))):
            if kind in only:
                ms = timeit(fn, args.reps)
                tot[kind][0] += ms
                tot[kind][1] += gf
                row += f" {ms:9.3f} {gf / ms:7.1f} |"
            else:
                row += f" {'-':>9s} {'-':>7s} |"
        print(row, flush=True)
    for kind, (ms, gf) in tot.items():
        if ms:
            print(f"total {kind}: {ms:.3f} ms, {gf / ms:.1f} TFLOP/s")
    if "convT" in only:
        for nm, cm, co, d in (("dec1.6", 16 * c, 8 * c, 16), ("dec2.6", 8 * c, 4 * c, 8), ("dec3.6", 4 * c, 2 * c, 4), ("dec4.6", 2 * c, c, 2)):
            h = w = S // d
            x = torch.randn(B, h, w, cm, device=dev).to(bf16)
            wt = torch.randn(cm, co, 2, 2, device=dev) * 0.05
            wf, wd = ops.pack_convT(wt)
            bias = torch.zeros(co, device=dev)
            y = torch.empty(B, 2 * h, 2 * w, co, device=dev, dtype=bf16)
            dy = torch.randn(B, 2 * h, 2 * w, co, device=dev).to(bf16)
            dx = torch.empty_like(x)
            dw = torch.zeros(4, cm, co, device=dev)
            gf = 2.0 * B * h * w * 4 * co * cm / 1e9
            a = timeit(lambda: ops.convT_fprop(x, wf, bias, out=y), args.reps)
            b = timeit(lambda: ops.convT_dgrad(dy, wd, out=dx), args.reps)
            cc = timeit(lambda: ops.convT_wgrad(x, dy, out=dw), args.reps)
            print(f"{nm:10s} {f'{cm}->{co} @{h}x{w}':28s} {gf:8.1f} | {a:9.3f} {gf / a:7.1f} | {b:9.3f} {gf / b:7.1f} | {cc:9.3f} {gf / cc:7.1f} |")
    if "mem" in only:
        P = B * S * S
        y = torch.randn(B, S, S, 64, device=dev).to(bf16)
        z = torch.empty_like(y)
        sc, sh = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        s1 = torch.zeros(64, device=dev, dtype=torch.float64)
        s2 = torch.zeros(64, device=dev, dtype=torch.float64)
        nbytes = y.numel() * 2
        for name, fn, traffic in (
                ("bn_apply", lambda: ops.bn_apply(y, sc, sh, out=z), 2 * nbytes),
                ("bn_apply_pool", lambda: ops.bn_apply_pool(y, sc, sh, z=z), 2 * nbytes + nbytes // 4 + nbytes // 8),
                ("bn_bwd_reduce", lambda: ops.bn_bwd_reduce(y, z, s1, s2), 2 * nbytes),
                ("bn_relu_bwd_apply", lambda: ops.bn_relu_bwd_apply(y, z, sc, sh, sh, s1, out=z), 3 * nbytes),
                ("bn_stats", lambda: ops.bn_stats(y, s1, s2), nbytes)):
            ms = timeit(fn, args.reps)
            print(f"{name:20s} {ms:8.3f} ms  {traffic / ms / 1e6:8.1f} GB/s")
        x = torch.randn(B, 3, S, S, device=dev)
        ms = timeit(lambda: ops.im2col_stem(x), args.reps)
        print(f"{'im2col_stem':20s} {ms:8.3f} ms  {(x.numel() * 4 + P * 128) / ms / 1e6:8.1f} GB/s")
        lg = torch.randn(P, 21, device=dev)
        lab = torch.randint(0, 21, (P,), device=dev)
        dl = torch.empty(P, 64, device=dev, dtype=bf16)
        acc = torch.zeros(2, device=dev, dtype=torch.float64)
        ms = timeit(lambda: ops.ce_kd_loss(lg, lab, dlogits=dl, loss_acc=acc), args.reps)
        print(f"{'ce_loss':20s} {ms:8.3f} ms  {(P * 21 * 4 + P * 8 + P * 128) / ms / 1e6:8.1f} GB/s")
        ms = timeit(lambda: ops.argmax_confusion(lg, lab, nc=22), args.reps)
        print(f"{'argmax_confusion':20s} {ms:8.3f} ms  {(P * 21 * 4 + P * 8) / ms / 1e6:8.1f} GB/s")
        pr = torch.randint(0, 21, (P,), device=dev)
        ms = timeit(lambda: ops.confusion_matrix(lab, pr, 22), args.reps)
        print(f"{'confusion_matrix':20s} {ms:8.3f} ms  {(P * 16) / ms / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()
