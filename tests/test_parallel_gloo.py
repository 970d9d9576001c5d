"""CPU: the data-parallel host logic (batch sharding, gradient buckets, bucketed all-reduce, confusion-matrix
reduction) with world_size 2 over gloo.  The kernels are not involved; the N>1 GPU path runs the same code
over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from continual_learning_b200 import parallel
from oracle.unet_ref import make_state_dict, param_names


def _names_numels():
    sd = make_state_dict(0)
    names = param_names(sd)
    return names, [sd[k].numel() for k in names]


def test_bucket_bounds_cover_the_flat_buffer_in_backward_order():
    names, numels = _names_numels()
    groups = parallel.bucket_bounds(numels, names, n_buckets=4)
    bounds = [b for g in groups for b in g]
    total = sum(numels)
    # disjoint, complete
    spans = sorted(bounds)
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
        assert b0 == a1
    # three groups in the order the backward pass finishes them: head + decoder (3 buckets, tail of the parameter
    # list first), enc4, enc3..enc1
    enc_end = sum(n for k, n in zip(names, numels) if k.startswith("enc"))
    enc4_start = sum(n for k, n in zip(names, numels) if k.startswith(("enc1", "enc2", "enc3")))
    assert len(groups) == 3 and len(groups[0]) == 3
    assert groups[1] == [(enc4_start, enc_end)] and groups[2] == [(0, enc4_start)]
    assert groups[0][0][1] == total and groups[0][-1][0] == enc_end
    assert all(groups[0][i][0] >= groups[0][i + 1][0] for i in range(2))
    # the exposed tail (the group that can only start after the whole backward pass) is < 5 % of the bytes
    assert enc4_start / total < 0.05


def test_shard_batch():
    x = torch.arange(8).view(8, 1)
    assert parallel.shard_batch(x, 1, 4).flatten().tolist() == [2, 3]
    with pytest.raises(ValueError):
        parallel.shard_batch(x, 0, 3)
    # validation batches may be ragged (no drop_last on the reference's val loader, main.py:39-41)
    parts = [parallel.shard_batch(x, r, 3, ragged=True).flatten().tolist() for r in range(3)]
    assert parts == [[0, 1, 2], [3, 4, 5], [6, 7]]
    y = torch.arange(2).view(2, 1)
    assert [parallel.shard_batch(y, r, 4, ragged=True).shape[0] for r in range(4)] == [1, 1, 0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env(backend="gloo")
    names, numels = _names_numels()
    comm = parallel.GradAllReduce(numels, names, n_buckets=4)
    flat = torch.full((sum(numels),), float(rank + 1))
    flat[::1000] += torch.arange(flat[::1000].numel(), dtype=torch.float32) * (rank + 1)
    expect = torch.full_like(flat, 3.0)
    expect[::1000] += torch.arange(flat[::1000].numel(), dtype=torch.float32) * 3
    comm.launch_group(flat, 0)   # head + decoder buckets while "the encoder backward still runs"
    comm.launch_group(flat, 1)   # enc4
    comm.launch_group(flat, 1)   # launching a group twice in one step must not reduce it twice
    comm.finish(flat)            # launches enc3..enc1, waits for everything
    ok = torch.equal(flat, expect)
    conf = torch.full((21, 21), rank + 1, dtype=torch.int64)
    correct = torch.tensor([10 * (rank + 1)], dtype=torch.int64)
    conf2, correct2 = parallel.all_reduce_confusion(conf, correct)
    ok = ok and bool((conf2 == 3).all()) and int(correct2) == 30 and comm.world_size == 2 and (r, w) == (rank, world)
    # the form Trainer.test() uses (trainer.py:270-284): no matrix, two counters {correct, total}
    _, both = parallel.all_reduce_confusion(torch.zeros(0, dtype=torch.int64),
                                            torch.tensor([7 * (rank + 1), 100 * (rank + 1)], dtype=torch.int64))
    ok = ok and both.tolist() == [21, 300]
    out[rank] = ok
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
