"""Data parallelism: one process per GPU, batch sharded by rank, gradients all-reduced (sum) over
NCCL/NVLink and scaled by 1/world in the Adam kernel.

Replaces the reference's single-process nn.DataParallel (trainer.py:120-122): replicas keep identical
fp32 master weights, BatchNorm statistics stay per replica (DataParallel semantics),
rank 0 owns prints and checkpoints.  The flat gradient buffer of UNetEngine is reduced in a few
large buckets ordered as the backward pass finishes them (decoder + head first, encoder last).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("CLK_FORCE_LOCAL_RANK") or os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            # CLK_DIST_BACKEND=gloo: several ranks sharing ONE GPU (NCCL refuses duplicate devices) — the test suite
            # uses it to run the real N-rank code path, kernels included, on a single-GPU box
            backend = os.environ.get("CLK_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_batch(x, rank, world, ragged=False):
    """rank r takes images [r*B/n, (r+1)*B/n) of the global batch.

    ragged=True (validation: the reference's val loader has no drop_last, main.py:39-41) tolerates a batch that is
    not a multiple of the world size: the first B % n ranks take one image more, a rank may get an empty shard."""
    b = x.shape[0]
    if b % world:
        if not ragged:
            raise ValueError(f"global batch {b} is not divisible by world size {world}")
        per, rem = divmod(b, world)
        start = rank * per + min(rank, rem)
        return x[start:start + per + (1 if rank < rem else 0)]
    per = b // world
    return x[rank * per:(rank + 1) * per]


def bucket_bounds(param_numels, names, n_buckets=4):
    """Split the flat gradient buffer (module.parameters() order) into contiguous buckets.

    Backward produces gradients in reverse parameter order except that the decoder runs before the
    encoder: [last, dec4..dec1] finish first, then [enc4..enc1].  Parameter order is
    enc1..enc4, dec1..dec4, last, so bucket 0 = encoder (reduced last), the rest split the decoder + head
    by size.  Returns [(start, end)] over the flat buffer, in the order they should be launched.
    """
    offs = [0]
    for n in param_numels:
        offs.append(offs[-1] + n)
    first_dec = next(i for i, nm in enumerate(names) if nm.startswith("dec"))
    enc = (0, offs[first_dec])
    dec_total = offs[-1] - offs[first_dec]
    target = dec_total / max(1, n_buckets - 1)
    bounds, start = [], first_dec
    acc = 0
    for i in range(first_dec, len(param_numels)):
        acc += param_numels[i]
        if acc >= target and len(bounds) < n_buckets - 2:
            bounds.append((offs[start], offs[i + 1]))
            start, acc = i + 1, 0
    if start < len(param_numels):
        bounds.append((offs[start], offs[-1]))
    # launch order: the tail of the parameter list (head, dec4 ...) is final first
    return list(reversed(bounds)) + [enc]


class GradAllReduce:
    """Sum-all-reduce of a flat fp32 gradient buffer in buckets (async ops, one wait at the end)."""

    def __init__(self, param_numels, names, group=None, n_buckets=4):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bounds = bucket_bounds(list(param_numels), list(names), n_buckets)

        self._works = []

    def start_decoder(self, flat):
        """launch the head + decoder buckets (final once the decoder backward is done); overlaps with the
        encoder backward that is still being enqueued on the compute stream."""
        if self.world_size == 1:
            return
        for a, b in self.bounds[:-1]:
            if b > a:
                self._works.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, flat):
        """launch the encoder bucket and wait for every bucket (stream-level wait for NCCL)."""
        if self.world_size == 1:
            return
        a, b = self.bounds[-1]
        if b > a:
            self._works.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in self._works:
            w.wait()
        self._works = []

    def all_reduce(self, flat):
        self.start_decoder(flat)
        self.finish(flat)


def all_reduce_confusion(conf, correct, group=None):
    """validation sweep: ONE small int64 all-reduce of the confusion matrix + the counters.
    `conf` may be empty and `correct` may hold any number of counters (correct pixels, total pixels, ...)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return conf, correct
    n = conf.numel()
    buf = torch.cat([conf.reshape(-1), correct.reshape(-1)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[:n].view_as(conf), buf[n:].view_as(correct)


def attach(model, optimizer, n_buckets=4):
    """module path (UNet + loss.backward() + optimizer.step()): all-reduce the flat gradient buffer inside the
    U-Net backward (decoder buckets overlap the encoder backward) and fold the 1/world scale into FusedAdam."""
    names = [k for k, _ in model.named_parameters()]
    comm = GradAllReduce([p.numel() for p in model.parameters()], names, n_buckets=n_buckets)
    model.engine.comm = comm
    optimizer.default_grad_scale = 1.0 / comm.world_size
    return comm
