"""A/B timing of engine options on ONE box (CUDA-graph TrainStep at the benchmark shape). Developer tool."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import continual_learning_b200 as clk
from continual_learning_b200.synthetic import uniform_batch


def time_cfg(flags, tune, steps=40):
    from continual_learning_b200 import _lib
    for k, v in tune.items():
        _lib.set_tuning(k, v)
    torch.manual_seed(0)
    m = clk.UNet(21).cuda().train()
    for k, v in flags.items():
        setattr(m.engine, k, v)
    opt = clk.FusedAdam(m.parameters(), lr=1e-4, betas=(0.5, 0.99))
    ts = clk.TrainStep(m, opt, use_graph=True)
    bx, by = uniform_batch(1, 16, 256, 256)
    bx, by = bx.cuda(), by.cuda()
    for _ in range(5):
        ts.step(bx, by)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ts.step(bx, by)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if __name__ == "__main__":
    base = dict(use_side_stream=True, side_pack=True)
    nos = dict(use_side_stream=False, side_pack=False)
    nos = dict(use_side_stream=False, side_pack=False)
    nos = dict(use_side_stream=False, side_pack=False)
    variants = [("default", {}, {}), ("pdl off", {}, {"pdl": 0}), ("default", {}, {"pdl": 1}),
                ("one stream", nos, {}), ("default", {}, {})]
    if "--all" in sys.argv:
        sys.argv.remove("--all")
        variants += [("no side_pack", dict(side_pack=False), {}),
                     ("fprop_bn=64", {}, {"fprop_bn": 64}), ("conv3 generic only", {}, {"fprop_bn": 0, "conv3_v2": 0}),
                     ("conv3 halo 1-CTA everywhere", {}, {"conv3_v2": 2, "conv3_pair": 0}),
                     ("wgrad 1-CTA halo", {}, {"conv3_v2": 4, "conv3_pair": 1, "wgrad_v2": 1}),
                     ("default (again)", {}, {"wgrad_v2": 2})]
    for extra in sys.argv[1:]:
        k, v = extra.split("=")
        variants.append((extra, {}, {k: int(v)}))
    for name, fl, tune in variants:
        f = dict(base)
        f.update(fl)
        print(f"{name:36s} {time_cfg(f, tune):8.3f} ms/step", flush=True)
