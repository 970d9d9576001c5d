#!/usr/bin/env python
"""bench.py — U-Net 256x256 training throughput (images/sec) on N B200s, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # N=1 directly, N>1 under torchrun
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference arm: CPU oracle step

A "step" is one pass of the hot path over one synthetic VOC-shaped batch: U-Net forward, softmax
cross-entropy, backward, gradient all-reduce (N>1), Adam (trainer.py:168-176).  Workload at every N:
configs[1] of BASELINE.json — U-Net 21-class, 256x256, batch 16 per GPU (weak scaling).
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

METRIC = "unet256_train_images_per_sec"
UNIT = "images/s"
BATCH, H, W, NUM_CLASSES = 16, 256, 256, 21
GFLOP_PER_IMG_TRAIN = 289.28  # 3x fwd - dgrad(enc1.0), true (unpadded) dims, 256x256
WORKLOAD = "unet21_256x256_b16_train_single_task"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, mean over the conv3x3 forward +
# dgrad launches of one step (bytes).  ncu cannot run inside the timed region, so this is the figure of the committed
# `ncu --set full` capture of the same command (scripts/summarize_full.py prints it); the file is named in the line.
ROOFLINE_TRAFFIC = 87.1e6
ROOFLINE_TRAFFIC_SOURCE = "profiles/round1_pair_kernels_ncu_full.md"
_tr = os.path.join(ROOT, "profiles", "round2_conv_kernels_ncu_full.json")
if os.path.exists(_tr):
    try:
        _d = json.load(open(_tr))
        ROOFLINE_TRAFFIC, ROOFLINE_TRAFFIC_SOURCE = float(_d["mean_dram_bytes_per_launch"]), "profiles/round2_conv_kernels_ncu_full.json"
    except (ValueError, KeyError):
        pass


def igemm_flops_per_step(batch, h, w, conv_dim=64, num_classes=NUM_CLASSES, in_dim=3):
    """algorithmic FLOPs (2*M*N*K, true dims) of every tcgen05 igemm launch of one training step."""
    c = conv_dim
    convs = [(in_dim, c, 1, False), (c, c, 1, True)]
    for ci, co, d in ((c, 2 * c, 2), (2 * c, 4 * c, 4), (4 * c, 8 * c, 8)):
        convs += [(ci, co, d, True), (co, co, d, True)]
    convT = []
    for ci, cm, co, d in ((8 * c, 16 * c, 8 * c, 16), (16 * c, 8 * c, 4 * c, 8), (8 * c, 4 * c, 2 * c, 4), (4 * c, 2 * c, c, 2)):
        convs += [(ci, cm, d, True), (cm, cm, d, True)]
        convT.append((cm, co, d))
    convs += [(2 * c, c, 1, True), (c, c, 1, True)]
    total = 0.0
    conv_fd = 0.0  # conv3x3 forward + dgrad launches only (the dominant kernel, igemm_conv3x2_kernel)
    for ci, co, d, has_dgrad in convs:
        m = batch * (h // d) * (w // d)
        total += 2.0 * m * co * ci * 9 * (3 if has_dgrad else 2)
        if ci != in_dim:  # the 3-channel stem runs as an im2col GEMM on the generic kernel
            conv_fd += 2.0 * m * co * ci * 9 * 2
    for cm, co, d in convT:
        m = batch * (h // d) * (w // d)
        total += 2.0 * m * (4 * co) * cm * 3
    total += 2.0 * batch * h * w * num_classes * c * 3
    return total, conv_fd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [t.strip() for t in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts_, r in self.rows:
            if t_begin is not None and not (t_begin <= ts_ <= t_end):
                continue  # keep only the samples taken inside the timed region
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    """{burst, sustained} dense-bf16 TFLOP/s, HBM GB/s, the SM clock the sustained figure was measured at, source."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        burst = d.get("bf16_tflops")
        sus = d.get("bf16_tflops_sustained", burst)
        return {"burst": burst or sus, "sustained": sus, "hbm": d.get("hbm_gbs"),
                "sustained_mhz": (d.get("clocks_under_load") or {}).get("sm_mhz_median"),
                "source": "measured (MEASURED_PEAKS.json: cuBLAS bf16 8192^3 best of 10 = burst, back to back for 4 s = sustained)"}
    return {"burst": 1590.0, "sustained": 1590.0, "hbm": 6650.0, "sustained_mhz": None,
            "source": "fallback (B200_PROFILING.md)"}


def pick_regime(clocks, peaks):
    """which measured peak matches the run: the sustained figure was taken power-capped (MEASURED_PEAKS.json records
    its median SM clock); a run whose clocks stayed near the maximum is a burst measurement."""
    mhz, mx = (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz")
    if not mhz or not mx:
        return "sustained"
    lo = peaks.get("sustained_mhz") or 0.7 * mx
    return "burst" if mhz >= 0.5 * (lo + mx) else "sustained"


# ---------------------------------------------------------------------------------------------
def cpu_step_throughput(steps, warmup, batch, threads=None, budget_s=150.0):
    """the reference training step (trainer.py:172-176) on the host cores, fp32, all host threads (torchrun exports
    OMP_NUM_THREADS=1: overridden here, rank 0 is the only rank doing CPU work).  Runs the REFERENCE'S OWN U-Net
    module when `oracle/_ref` was staged (byte-compiled from /root/reference by `oracle/build_ref.py`; kind
    "reference"), else the oracle port (kind "port").  Stops early when the time budget is used up, so `--steps 100`
    still ends within a few minutes on a slow host.  Returns (img/s, s/step, threads, steps timed, kind)."""
    from continual_learning_b200.synthetic import uniform_batch
    from oracle import build_ref
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, threads))
    t_begin = time.perf_counter()
    x, y = uniform_batch(1, batch, H, W, NUM_CLASSES)
    if build_ref.available():
        kind = "reference"
        UNet, _ = build_ref.load()
        torch.manual_seed(0)
        model = UNet(num_classes=NUM_CLASSES, in_dim=3, conv_dim=64)                 # trainer.py:107
        optim = torch.optim.Adam(model.parameters(), lr=1e-4, betas=[0.5, 0.99])     # trainer.py:108-110
        c_loss = torch.nn.CrossEntropyLoss()                                         # trainer.py:113

        def one_step():                                                              # trainer.py:172-176
            outputs = model(x)
            optim.zero_grad()
            loss = c_loss(outputs, y)
            loss.backward()
            optim.step()
    else:
        kind = "port"
        from oracle import step_ref
        from oracle.unet_ref import make_state_dict, param_names
        sd = make_state_dict(0, NUM_CLASSES)
        opt = step_ref.AdamRef(param_names(sd), lr=1e-4, betas=(0.5, 0.99))

        def one_step():
            _, _, grads, _ = step_ref.forward_backward(sd, x, y)
            opt.step(sd, grads)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if time.perf_counter() - t_begin > budget_s:
                break
    return batch / statistics.median(times), statistics.median(times), torch.get_num_threads(), len(times), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 2
    val, sec, cores, ran, kind = cpu_step_throughput(max(1, args.steps), args.warmup, sample_batch)
    what = ("the reference's own models/unet.py UNet (oracle/_ref) + nn.CrossEntropyLoss + optim.Adam, loop of" if kind == "reference"
            else "oracle port of")
    sample = (f"{what} trainer.py:172-176 (fp32 CPU), batch {sample_batch} of the {BATCH}-image 256x256 step, "
              f"{ran} of {args.steps} steps timed after {args.warmup} warm-up (150 s budget), median {sec:.3f} s/step")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU reference arm: bounded sample (batch 2) per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    import continual_learning_b200 as clk
    from continual_learning_b200 import _lib, parallel
    from continual_learning_b200.synthetic import uniform_batch

    rank, local, world = parallel.init_from_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.ensure_device(local)
    warmup = max(3, args.warmup)

    torch.manual_seed(0)  # identical initial weights on every rank (reference default init, unet.py:41-72)
    model = clk.UNet(NUM_CLASSES).to(dev)
    model.train()
    opt = clk.FusedAdam(model.parameters(), lr=1e-4, betas=(0.5, 0.99))  # trainer.py:108-110, main.py:77-80
    comm = None
    if world > 1 or os.environ.get("CLK_FORCE_SEGMENTS"):  # (developer probe: the four-graph schedule on one GPU)
        names = [k for k, _ in model.named_parameters()]
        comm = parallel.GradAllReduce([p.numel() for p in model.parameters()], names)
    old_model = None
    if args.continual:  # BASELINE config 3: + frozen previous-task network UNet(16) in eval mode, T = 2, lambda = 1
        old_model = clk.UNet(16).to(dev)
        old_model.eval()
    ts = clk.TrainStep(model, opt, old_model=old_model, T=2.0, lam=1.0, use_graph=not args.no_graph,
                       comm=comm)

    # synthetic VOC-shaped data: a few distinct batches per rank, resident in HBM for `value`,
    # in pinned host memory for `e2e`
    nb = 4
    host = [uniform_batch(1000 * rank + i, BATCH, H, W, NUM_CLASSES) for i in range(nb)]
    pinned = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    devb = [(x.to(dev), y.to(dev)) for x, y in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start: launch it before the warm-up, filter by time later
    for i in range(warmup):
        ts.step(*devb[i % nb])
    barrier()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w_begin = time.time()
    e0.record()
    for i in range(args.steps):
        loss = ts.step(*devb[i % nb])
    e1.record()
    barrier()
    w_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(w_begin, w_end) if rank == 0 else None
    last_loss = float(loss)
    if ts.graph is not None:
        launches = ts.launches_per_step * args.steps
    else:
        launches = _lib.launch_count - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * BATCH * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST inputs (H2D + D2H inside the timed region)
    # warm-up through exactly the timed call pattern (copy stream, pinned loss buffers, allocator blocks of the
    # staged batches are created here, not inside the timed region)
    for i in range(warmup):
        ts.step_host(*pinned[i % nb], prefetch=pinned[(i + 1) % nb], defer_loss=True)
    ts.step_host(*pinned[warmup % nb], defer_loss=True)
    ts.flush_loss()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        # the H2D copy of batch i+1 is issued before step i's kernels (double-buffered input staging)
        # and the loss of step i is read back (async D2H into pinned memory) while step i+1 runs
        nxt = pinned[(i + 1) % nb] if i + 1 < args.steps else None
        ts.step_host(*pinned[i % nb], prefetch=nxt, defer_loss=True)
    e2e_last_loss = ts.flush_loss()
    e1.record()
    barrier()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
    d2h = 16  # {loss, out-of-range-label flag} as two float64, one read per step

    # ---- per-kernel timing of one eager step (CUDA events around every clk_* launch) -> roofline
    ts_prof = clk.TrainStep(model, opt, old_model=old_model, T=2.0, lam=1.0, use_graph=False, comm=comm)
    ts_prof.step_count = ts.step_count
    model.engine.use_side_stream = False  # serialise the weight-gradient stream: one kernel at a time under the events
    ts_prof.step(*devb[0])
    torch.cuda.synchronize()
    PROF_STEPS = 3
    prof = {}
    for i in range(PROF_STEPS):
        _lib.start_profile()
        ts_prof.step(*devb[(i + 1) % len(devb)])
        for k, v in _lib.stop_profile().items():
            e = prof.setdefault(k, [0, 0.0])
            e[0] += v[0] / PROF_STEPS
            e[1] += v[1] / PROF_STEPS
    model.engine.use_side_stream = True
    igemm_names = [k for k in prof if k.startswith(("clk_conv3x3_", "clk_gemm_", "clk_convT2x2_", "clk_head_loss_bwd"))]
    igemm_ms = sum(prof[k][1] for k in igemm_names)
    igemm_n = sum(prof[k][0] for k in igemm_names)
    all_ms = sum(v[1] for v in prof.values())
    flops, conv_fd_flops = igemm_flops_per_step(BATCH, H, W)
    peaks = measured_peaks()
    all_tensor_tf = flops / (igemm_ms / 1e3) / 1e12
    # dominant kernel: igemm_conv3x2_kernel = the 17 conv3x3 forward + 17 dgrad launches of a step
    dom_n = prof["clk_conv3x3_fprop"][0] + prof["clk_conv3x3_dgrad"][0]
    dom_ms = prof["clk_conv3x3_fprop"][1] + prof["clk_conv3x3_dgrad"][1]
    achieved_tf = conv_fd_flops / (dom_ms / 1e3) / 1e12
    breakdown = {k: {"launches": round(v[0]), "ms": round(v[1], 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    regime = pick_regime(clocks, peaks)
    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores, _, kind = cpu_step_throughput(steps=6, warmup=2, batch=2)
        what = "the reference's own U-Net module (oracle/_ref)" if kind == "reference" else "oracle port"
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{what}, loop of trainer.py:172-176, fp32, batch 2 x 256x256 (BASELINE config 1 shape), 6 steps after 2 warm-up, median {sec:.3f} s/step"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "per_gpu_batch": BATCH, "image": [H, W],
                   "num_classes": NUM_CLASSES, "parallelism": f"dp{world}", "optimizer": "adam(1e-4,0.5,0.99)",
                   "cuda_graph": ts.graph is not None,
                   "l2": "per-step working set (~2 GB of bf16 activations + 0.9 GB of fp32 parameter/optimizer "
                         "state) is >> the 126 MB L2, so nothing survives between timed steps; no explicit flush"},
        "model_tflops": value * GFLOP_PER_IMG_TRAIN / 1e3 / world,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e * 1e3 / args.steps,
                "api": "TrainStep.step_host(x_pinned, y_pinned, prefetch=next, defer_loss=True): H2D of every batch "
                       "(double-buffered) and an async D2H of every step's loss, consumed one step later",
                "loss_last_step": e2e_last_loss},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor",
                     "kernel": "conv3x3 forward + dgrad: igemm_conv3x2_kernel (CTA-pair tcgen05 halo kernel) and, for the "
                               "64-output-channel layers, igemm_conv3r_kernel (row-tap, N = 192)",
                     "achieved": achieved_tf, "peak": peaks[regime], "unit": "TFLOP/s", "frac": achieved_tf / peaks[regime],
                     "peak_regime": regime + " (median SM clock of the timed region vs the clock the sustained peak was "
                                    "measured at)",
                     "frac_of_burst_peak": achieved_tf / peaks["burst"],
                     "frac_of_sustained_peak": achieved_tf / peaks["sustained"],
                     "frac_of_spec_2250": achieved_tf / 2250.0,
                     "peaks": {"burst": peaks["burst"], "sustained": peaks["sustained"]},
                     "peak_source": peaks["source"], "traffic": ROOFLINE_TRAFFIC,
                     "traffic_source": ROOFLINE_TRAFFIC_SOURCE + " (ncu --set full capture of this command; not "
                                       "measurable inside an un-profiled run)",
                     "launches_per_step": round(dom_n), "avg_launch_us": dom_ms * 1e3 / dom_n,
                     "algorithmic_gflop_per_launch": conv_fd_flops / dom_n / 1e9,
                     "share_of_step": dom_ms / all_ms if all_ms else None,
                     "timing": f"CUDA events around every launch of {PROF_STEPS} eager steps run right after the "
                               "timed region (same process, same buffers, weight-gradient side stream serialised)",
                     "all_tensor_kernels": {"achieved": all_tensor_tf, "frac": all_tensor_tf / peaks[regime],
                                            "launches_per_step": round(igemm_n), "ms_per_step": igemm_ms,
                                            "share_of_step": igemm_ms / all_ms if all_ms else None,
                                            "algorithmic_gflop_per_step": flops / 1e9},
                     "all_kernels_ms_per_step": all_ms,
                     "note": "the per-kernel event profile serialises the weight-gradient side stream, so "
                             "all_kernels_ms_per_step exceeds ms_per_step (where wgrad overlaps dgrad / BatchNorm)"},
        "kernel_breakdown_ms": breakdown,
        "loss_last_step": last_loss,
        "cpu_baseline": cpu,
    }
    emit(line)



# ---------------------------------------------------------------------------------------------
def cpu_eval_throughput(budget_s=40.0, batch=2):
    """BASELINE config 5 on the host cores: the reference's eval-mode forward (trainer.py:271,278-279) + metrics
    (metrics.py:55-63) on batches of 2; returns (img/s, threads, batches timed, kind)."""
    from continual_learning_b200.synthetic import uniform_batch
    from oracle import build_ref
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, threads))
    x, y = uniform_batch(7, batch, 256, 256, NUM_CLASSES)
    if build_ref.available():
        kind = "reference"
        UNet, mt = build_ref.load()
        torch.manual_seed(0)
        model = UNet(num_classes=NUM_CLASSES, in_dim=3, conv_dim=64).eval()

        def one():
            with torch.no_grad():
                _, pred = torch.max(model(x).data, 1)
            mt.eval_metrics(y, pred, NUM_CLASSES)
    else:
        kind = "port"
        from oracle import metrics_ref
        from oracle.unet_ref import UNetRef, make_state_dict
        sd = make_state_dict(0, NUM_CLASSES)

        def one():
            with torch.no_grad():
                pred = UNetRef(sd, NUM_CLASSES, training=False)(x).argmax(1)
            metrics_ref.eval_metrics(y, pred, NUM_CLASSES)
    one()
    t0, times = time.perf_counter(), []
    while time.perf_counter() - t0 < budget_s and len(times) < 30:
        t1 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t1)
    return batch / statistics.median(times), torch.get_num_threads(), len(times), kind


def run_eval_sweep(args):
    """BASELINE config 5: validation sweep — U-Net inference (BatchNorm folded into the conv epilogues) + 1x1 head +
    argmax + correct count + 21x21 confusion matrix (ONE kernel, logits never written) over 10,000 synthetic 256x256
    images sharded across the ranks, one 442-element int64 all-reduce at the end, metrics on rank 0
    (trainer.py:270-284 + metrics.py:55-63).  A "step" is one batch of EVAL_BATCH images per GPU."""
    import torch.distributed as dist

    import continual_learning_b200 as clk
    from continual_learning_b200 import _lib, metrics as mt, parallel
    from continual_learning_b200.synthetic import uniform_batch

    rank, local, world = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.ensure_device(local)
    EVAL_BATCH, TOTAL = 64, 10000
    per_rank = (TOTAL + world - 1) // world
    sizes = [EVAL_BATCH] * (per_rank // EVAL_BATCH) + ([per_rank % EVAL_BATCH] if per_rank % EVAL_BATCH else [])
    torch.manual_seed(0)
    model = clk.UNet(NUM_CLASSES).to(dev).eval()
    parallel.broadcast_buffers(model)
    nb = 3
    host = [uniform_batch(5000 + 10 * rank + i, EVAL_BATCH, 256, 256, NUM_CLASSES) for i in range(nb)]
    pinned = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    devb = [(x.to(dev), y.to(dev)) for x, y in host]
    conf = torch.zeros(NUM_CLASSES * NUM_CLASSES, device=dev, dtype=torch.int64)
    correct = torch.zeros(1, device=dev, dtype=torch.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sweep(from_host):
        conf.zero_()
        correct.zero_()
        copy_stream = torch.cuda.Stream()
        staged = None
        for i, b in enumerate(sizes):
            if from_host:   # H2D of batch i+1 is issued before batch i's kernels (double-buffered)
                if staged is None:
                    staged = (pinned[i % nb][0][:b].to(dev, non_blocking=True), pinned[i % nb][1][:b].to(dev, non_blocking=True))
                x, y = staged
                if i + 1 < len(sizes):
                    with torch.cuda.stream(copy_stream):
                        nx = pinned[(i + 1) % nb][0][:sizes[i + 1]].to(dev, non_blocking=True)
                        ny = pinned[(i + 1) % nb][1][:sizes[i + 1]].to(dev, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                model.evaluate_batch(x, y, nc=NUM_CLASSES, conf=conf, correct=correct)
                if i + 1 < len(sizes):
                    torch.cuda.current_stream().wait_event(ev)
                    nx.record_stream(torch.cuda.current_stream())
                    ny.record_stream(torch.cuda.current_stream())
                    staged = (nx, ny)
            else:
                x, y = devb[i % nb]
                model.evaluate_batch(x[:b], y[:b], nc=NUM_CLASSES, conf=conf, correct=correct)
        c, k = parallel.all_reduce_confusion(conf, correct)   # ONE 442-element int64 all-reduce
        return c, k

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        model.evaluate_batch(*devb[0], nc=NUM_CLASSES, conf=conf, correct=correct)
    barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w_begin = time.time()
    e0.record()
    c_dev, k_dev = sweep(False)
    e1.record()
    barrier()
    w_end = time.time()
    launches = _lib.launch_count - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(w_begin, w_end) if rank == 0 else None
    e0.record()
    c_host, k_host = sweep(True)
    k_cpu = int(k_host.item())  # the sweep's result is read back: D2H inside the timed region
    c_cpu = c_host.cpu()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    n_img = world * per_rank
    # per-kernel profile of two batches
    prof = {}
    for i in range(2):
        _lib.start_profile()
        model.evaluate_batch(*devb[i % nb], nc=NUM_CLASSES, conf=conf, correct=correct)
        for k, v in _lib.stop_profile().items():
            e = prof.setdefault(k, [0, 0.0])
            e[0] += v[0] / 2
            e[1] += v[1] / 2
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = measured_peaks()
    regime = pick_regime(clocks, peaks)
    _, conv_fd = igemm_flops_per_step(EVAL_BATCH, 256, 256)
    conv_f = conv_fd / 2.0   # forward only
    dom_ms, dom_n = prof["clk_conv3x3_fprop_eval"][1], prof["clk_conv3x3_fprop_eval"][0]
    tf = conv_f / (dom_ms / 1e3) / 1e12
    P = EVAL_BATCH * 256 * 256
    head_ms = prof["clk_head_argmax_confusion"][1]
    head_gbs = P * (128 + 8) / (head_ms / 1e3) / 1e9
    res = mt.metrics_from_matrix(c_cpu.view(NUM_CLASSES, NUM_CLASSES))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, nbat, kind = cpu_eval_throughput()
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": kind,
               "sample": f"eval-mode forward + torch.max + metrics.eval_metrics on batches of 2 x 256x256, {nbat} batches"}
    h2d = EVAL_BATCH * (3 * 256 * 256 * 4 + 256 * 256 * 8)
    line = {
        "metric": "unet256_validation_sweep_images_per_sec", "value": n_img / (ms / 1e3), "unit": "images/s", "n_gpus": world,
        "steps": len(sizes), "warmup": max(3, args.warmup), "ms_per_step": ms / len(sizes), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "unet21_256x256_validation_sweep_10k", "images": n_img, "per_gpu_images": per_rank,
                   "per_gpu_batch": EVAL_BATCH, "parallelism": f"dp{world}", "collective": "one all-reduce of 442 int64",
                   "l2": "each batch reads/writes > 1 GB of activations: nothing survives in the 126 MB L2 between steps"},
        "clocks": clocks,
        "e2e": {"value": n_img / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 442 * 8,
                "api": "UNet.evaluate_batch on pinned host batches (H2D double-buffered), confusion matrix + correct "
                       "count read back at the end of the sweep"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "conv3x3 forward with BatchNorm folded into the epilogue (igemm_conv3x2_kernel / igemm_conv3r_kernel)",
                     "achieved": tf, "peak": peaks[regime], "unit": "TFLOP/s", "frac": tf / peaks[regime],
                     "frac_of_burst_peak": tf / peaks["burst"], "frac_of_sustained_peak": tf / peaks["sustained"],
                     "peak_regime": regime, "traffic": None, "launches_per_step": round(dom_n),
                     "hbm_kernels": {"head_argmax_kernel": {"us": head_ms * 1e3, "algorithmic_bytes": P * 136,
                                                            "achieved_gbs": head_gbs, "peak_gbs": peaks["hbm"],
                                                            "frac": head_gbs / peaks["hbm"]}}},
        "kernel_breakdown_ms": {k: {"launches": round(v[0]), "ms": round(v[1], 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        "result": {"pixel_acc": float(res[0]), "mean_class_acc": float(res[1]), "mean_iou": float(res[2]),
                   "correct_pixels": k_cpu, "confusion_sum": int(c_cpu.sum()),
                   "device_resident_sweep_equal": bool(torch.equal(c_dev.cpu(), c_cpu))},
        "cpu_baseline": cpu,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL's version banner, torchrun notices) write to fd 1 behind Python's back: route everything except
    # the JSON line to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--image", type=int, default=256, help="square image size (default = BASELINE config 2; 512 = "
                    "the per-GPU shape of config 4)")
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--continual", action="store_true", help="BASELINE config 3: the continual step (distillation "
                    "against a frozen UNet(16)); images/s of that step")
    ap.add_argument("--workload", default="train", choices=["train", "continual", "train512", "eval_sweep"],
                    help="train = BASELINE config 2 (the contract's default); continual = config 3; train512 = config 4's "
                         "per-GPU shape (16 x 512x512 per GPU); eval_sweep = config 5 (10,000-image validation sweep)")
    args = ap.parse_args()
    if args.workload == "continual":
        args.continual = True
    elif args.workload == "train512":
        args.image = 512
    elif args.workload == "eval_sweep" and args.impl != "reference":
        return run_eval_sweep(args)
    global BATCH, H, W, WORKLOAD, GFLOP_PER_IMG_TRAIN
    if args.image != 256 or args.batch != 16 or args.continual:
        GFLOP_PER_IMG_TRAIN *= (args.image * args.image) / float(H * W)
        BATCH, H, W = args.batch, args.image, args.image
        WORKLOAD = f"unet21_{H}x{W}_b{BATCH}_train_" + ("continual_kd16" if args.continual else "single_task")
        args.no_cpu_baseline = True  # the CPU sample is defined on the default workload only
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
