#!/usr/bin/env python
"""N-rank data-parallel parity (SURVEY.md §8e "Parity check"; reference semantics trainer.py:120-122).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/ddp_parity.py [--out gpurun_out/ddp_parity.json]

Every rank runs `TrainStep` (the path bench.py times at N > 1) on its shard of one global batch with the bucketed
all-reduce; rank 0 then compares against the chunked n-replica CPU oracle (`oracle/chunked_ref.py`):
  * the all-reduced gradient (sum over ranks) x 1/n   vs the oracle's averaged per-chunk gradients,
  * the weights after one Adam step                    vs `AdamRef` on the oracle gradient,
  * rank 0's BatchNorm running statistics              vs chunk 0's (DataParallel replica-0 semantics),
  * bit-identical weights on every rank after the step (replicas must not drift).
CLK_DIST_BACKEND=gloo + LOCAL_RANK=0 for every rank runs the same thing on ONE GPU (tests/test_gpu_ddp.py).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--per-rank", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--graph", type=int, default=-1, help="-1: TrainStep's default for this world size")
    args = ap.parse_args()

    import continual_learning_b200 as clk
    from continual_learning_b200 import parallel
    from continual_learning_b200.synthetic import structured_batch
    from oracle import step_ref
    from oracle.chunked_ref import chunked_forward_backward
    from oracle.unet_ref import clone_sd, make_state_dict, param_names

    rank, local, world = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sd = make_state_dict(5)
    x, y = structured_batch(6, world * args.per_rank, args.size, args.size)

    model = clk.UNet(21).to(dev)
    model.load_state_dict(sd)
    model.train()
    opt = clk.FusedAdam(model.parameters(), lr=1e-4, betas=(0.5, 0.99))
    names = [k for k, _ in model.named_parameters()]
    comm = parallel.GradAllReduce([p.numel() for p in model.parameters()], names) if world > 1 else None
    kw = {} if args.graph < 0 else {"use_graph": bool(args.graph)}
    ts = clk.TrainStep(model, opt, comm=comm, **kw)
    xs = parallel.shard_batch(x, rank, world).to(dev)
    ys = parallel.shard_batch(y, rank, world).to(dev)
    loss = float(ts.step(xs, ys))
    torch.cuda.synchronize()

    # replicas must hold bit-identical weights after the step
    w = torch.cat([p.detach().flatten() for p in model.parameters()])
    digest = torch.stack([w.double().sum(), w.double().abs().sum(), (w.double() * w.double()).sum()])
    if world > 1:
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        losses = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(losses, torch.tensor([loss], device=dev, dtype=torch.float64))
    else:
        gathered, losses = [digest], [torch.tensor([loss])]
    result = None
    if rank == 0:
        identical = all(torch.equal(g, gathered[0]) for g in gathered)
        # ---- chunked n-replica oracle on the host (fp32)
        sd_ref = clone_sd(sd)
        loss_ref, grads_ref = chunked_forward_backward(sd_ref, x, y, world)
        pn = param_names(sd)
        g = model.engine.G / world                       # all-reduced sum x 1/n (the scale Adam applies)
        g_ref = torch.cat([grads_ref[k].flatten() for k in pn])
        w0 = torch.cat([sd[k].flatten() for k in pn])
        step_ref.AdamRef(pn, lr=1e-4, betas=(0.5, 0.99)).step(sd_ref, grads_ref)
        w_ref = torch.cat([sd_ref[k].flatten() for k in pn])
        wc = w.cpu()
        big = g_ref.abs() > g_ref.abs().quantile(0.75) if g_ref.numel() < 2 ** 24 else g_ref.abs() > 2.0 * g_ref.abs().median()
        sign_agree = float(((wc - w0).sign() == (w_ref - w0).sign())[big].float().mean())
        st = model.state_dict()
        bn = {k: rel(st[k], sd_ref[k]) for k in ("enc1.2.running_mean", "enc1.2.running_var", "enc4.block.6.running_var",
                                                  "dec1.block.5.running_mean", "last.5.running_var")}
        result = {
            "world": world, "backend": dist.get_backend() if world > 1 else "none", "per_rank_batch": args.per_rank,
            "image": args.size, "cuda_graph": ts.graph is not None,
            "devices": sorted({int(os.environ.get("LOCAL_RANK", "0"))}) if world == 1 else "one per rank" if
            os.environ.get("CLK_DIST_BACKEND") != "gloo" else "shared cuda:0",
            "loss_mean_over_ranks": float(torch.stack(losses).mean()), "loss_oracle": loss_ref,
            "grad_rel_l2": rel(g, g_ref), "grad_cosine": cosine(g, g_ref),
            "adam_update_sign_agreement": sign_agree, "weights_rel_l2_after_step": rel(wc, w_ref),
            "bn_running_stats_rel": bn, "num_batches_tracked": int(st["enc1.2.num_batches_tracked"]),
            "replicas_bit_identical": identical,
        }
        print(json.dumps(result))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(result, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        ok = (abs(result["loss_mean_over_ranks"] - result["loss_oracle"]) <= 1e-3 * result["loss_oracle"]
              and result["grad_rel_l2"] <= 5e-2 and result["grad_cosine"] >= 0.998
              and result["adam_update_sign_agreement"] >= 0.97 and max(result["bn_running_stats_rel"].values()) <= 2e-2
              and result["replicas_bit_identical"] and result["num_batches_tracked"] == 1)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
