"""GPU (-m gpu): the REAL N-rank data-parallel path (one process per rank under torch.distributed.run) against the
chunked n-replica CPU oracle (reference semantics: nn.DataParallel, trainer.py:120-122), and the
drop-in `main.py` / `Trainer` (main.py:46-58, trainer.py:132-284) at world sizes 1 and 2.

On a box with >= 2 GPUs the ranks use NCCL, one GPU each.  On a single-GPU box the two ranks share cuda:0 and
all-reduce through gloo (`CLK_DIST_BACKEND=gloo`): the same processes, kernels, buckets and hooks — only the
transport differs."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(nproc, script_args, timeout=900):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    if nproc > 1 and torch.cuda.device_count() < nproc:
        # all ranks on cuda:0, gloo transport
        env["CLK_DIST_BACKEND"] = "gloo"
        env["CLK_FORCE_LOCAL_RANK"] = "0"
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f"rc={r.returncode}\n--- stdout\n{r.stdout[-3000:]}\n--- stderr\n{r.stderr[-3000:]}"
    return r


@pytest.mark.parametrize("world", [2])
def test_two_real_ranks_match_the_chunked_oracle(lib_built, tmp_path, world):
    out = tmp_path / "ddp.json"
    _torchrun(world, ["scripts/ddp_parity.py", "--out", str(out)])
    res = json.load(open(out))
    print(res)
    assert res["world"] == world and res["replicas_bit_identical"]
    assert abs(res["loss_mean_over_ranks"] - res["loss_oracle"]) <= 1e-3 * res["loss_oracle"]
    assert res["grad_rel_l2"] <= 5e-2 and res["grad_cosine"] >= 0.998
    assert res["adam_update_sign_agreement"] >= 0.97
    assert max(res["bn_running_stats_rel"].values()) <= 2e-2 and res["num_batches_tracked"] == 1


@pytest.mark.parametrize("world", [1, 2])
def test_main_py_synthetic_epoch_and_trainer_test(lib_built, tmp_path, world):
    """`python main.py --mode train --synthetic ...` (the reference CLI, main.py:63-107) for two epochs, with the
    every-10th-iteration statistics block, the per-epoch checkpoint, a resume, and Trainer.test() — at world 1
    and under torchrun with 2 ranks (where the round-1 code raised in test())."""
    args = ["scripts/run_trainer_check.py", "--model_save_path", str(tmp_path / "model"), "--sample_save_path",
            str(tmp_path / "sample"), "--out", str(tmp_path / "res.json")]
    if world == 1:
        r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    else:
        _torchrun(world, args)
    res = json.load(open(tmp_path / "res.json"))
    print(res)
    assert res["world"] == world and res["replicas_identical_after_training"]
    assert 0.0 <= res["test_acc"] <= 100.0 and res["test_acc_ragged"] == pytest.approx(res["test_acc_expected_ragged"], abs=1e-9)
    assert res["checkpoint_keys"] == ["epoch", "model_state", "optimizer_state", "scheduler_state"]
    assert res["resumed_epoch"] == 3  # the reference stores epoch + 1 AFTER incrementing it (trainer.py:74,258)
    assert res["resumed_adam_step"] == res["adam_step_at_save"] > 0
    assert res["loss_after_resume_finite"] and res["weights_restored"] and res["module_prefix_checkpoint_loaded"]
    assert res["scheduler_last_epoch"] == 2  # LambdaLR state came back with the checkpoint (trainer.py:99)
