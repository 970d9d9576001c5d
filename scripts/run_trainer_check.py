#!/usr/bin/env python
"""Drive the drop-in CLI / Trainer end to end on synthetic data (tests/test_gpu_ddp.py runs this at world 1 and,
under torch.distributed.run, at world 2): main.py's parser and loaders (main.py:17-58), `Trainer.train_val()` for
two epochs (statistics block, sample dump, per-epoch checkpoint: trainer.py:132-265), `Trainer.test()` with a
RAGGED last validation batch (trainer.py:270-284), and a resume through `--continue_train` from the checkpoint the
run wrote — also re-written with DataParallel's `module.` key prefix (trainer.py:117-122)."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_save_path", required=True)
    ap.add_argument("--sample_save_path", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    import main as cli
    from continual_learning_b200 import parallel
    from continual_learning_b200.trainer import Trainer

    argv = ["--mode", "train", "--synthetic", "--synthetic_size", "44", "--train_batch_size", "4", "--val_batch_size", "3",
            "--h_image_size", "64", "--w_image_size", "64", "--n_iters", "2", "--num_workers", "0",
            "--model_save_path", a.model_save_path, "--sample_save_path", a.sample_save_path]
    cfg = cli.build_parser().parse_args(argv)
    os.makedirs(cfg.model_save_path, exist_ok=True)
    train_loader, val_loader = cli.get_loader(cfg)
    tr = Trainer(train_data_loader=train_loader, val_data_loader=val_loader, config=cfg)
    tr.train_val()
    # replicas must hold bit-identical weights after training (all-reduced gradients, identical Adam)
    w_all = torch.cat([p_.detach().flatten() for p_ in tr.model.parameters()])
    digest = torch.stack([w_all.double().sum(), w_all.double().abs().sum(), (w_all.double() * w_all.double()).sum()])
    replicas_identical = True
    if tr.world > 1:
        gathered = [torch.zeros_like(digest) for _ in range(tr.world)]
        dist.all_gather(gathered, digest)
        replicas_identical = all(torch.equal(g_, gathered[0]) for g_ in gathered)
    acc = tr.test()                       # 44 images in batches of 3: the last batch holds 2 (ragged at world 2)
    # the same sweep with EVERY shard evaluated locally (same sub-batch shapes, hence the same kernel instantiations
    # and bit-identical predictions): integer counts, so the all-reduced result of the sharded run must be identical
    correct = torch.zeros(1, device=tr.device, dtype=torch.int64)
    total = 0
    with torch.no_grad():
        for images, labels in val_loader:
            for r in range(tr.world):
                xi = parallel.shard_batch(images, r, tr.world, ragged=True)
                yi = parallel.shard_batch(labels, r, tr.world, ragged=True)
                if xi.shape[0]:
                    tr.model.evaluate_batch(xi.to(tr.device), yi.to(tr.device), correct=correct)
            total += labels.nelement()
    expected = 100 * float(correct) / float(total)
    if tr.world > 1:
        dist.barrier()                    # rank 0 has written latest_net_UNET_VOC.pth
    path = os.path.join(cfg.model_save_path, "latest_net_UNET_VOC.pth")
    ck = torch.load(path, map_location="cpu")
    adam_step = float(ck["optimizer_state"]["state"][0]["step"])

    # ---- resume (trainer.py:96-102 via --continue_train), then one more iteration of the hot loop
    cfg2 = cli.build_parser().parse_args(argv + ["--continue_train"])
    tr2 = Trainer(train_data_loader=train_loader, val_data_loader=val_loader, config=cfg2)
    p0 = next(tr2.model.parameters())
    resumed_step = float(tr2.optim.state[p0]["step"])
    same_w = all(torch.equal(v.cpu(), ck["model_state"][k]) for k, v in tr2.model.state_dict().items())
    x, y = next(iter(train_loader))
    x, y = x.to(tr2.device), y.to(tr2.device)
    if tr2.world > 1:
        x, y = parallel.shard_batch(x, tr2.rank, tr2.world), parallel.shard_batch(y, tr2.rank, tr2.world)
    out = tr2.model(x)
    tr2.reset_grad()
    loss = tr2.c_loss(out, y)
    loss.backward()
    tr2.optim.step()
    loss_ok = math.isfinite(float(loss)) and float(tr2.optim.state[p0]["step"]) == adam_step + 1

    # ---- a checkpoint written by the reference under nn.DataParallel carries `module.`-prefixed keys
    if tr.rank == 0:
        ck["model_state"] = {"module." + k: v for k, v in ck["model_state"].items()}
        torch.save(ck, os.path.join(cfg.model_save_path, "dp_net_UNET_VOC.pth"))
    if tr.world > 1:
        dist.barrier()
    cfg3 = cli.build_parser().parse_args(argv + ["--continue_train", "--which_epoch", "dp"])
    tr3 = Trainer(train_data_loader=train_loader, val_data_loader=val_loader, config=cfg3)
    prefixed_ok = all(torch.equal(a_.cpu(), b_.cpu()) for a_, b_ in zip(tr3.model.state_dict().values(),
                                                                        tr2_initial(ck)))
    if tr.rank == 0:
        res = {"world": tr.world, "test_acc": acc, "test_acc_ragged": acc, "test_acc_expected_ragged": expected,
               "replicas_identical_after_training": replicas_identical,
               "checkpoint_keys": sorted(k for k in ck.keys()), "adam_step_at_save": adam_step,
               "resumed_epoch": tr2.start_epoch, "resumed_adam_step": resumed_step, "weights_restored": same_w,
               "loss_after_resume_finite": loss_ok, "module_prefix_checkpoint_loaded": prefixed_ok,
               "scheduler_last_epoch": tr2.scheduler.last_epoch,
               "samples_written": sorted(os.listdir(os.path.join(cfg.sample_save_path, "generated")))[:3]}
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
    if tr.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def tr2_initial(ck):
    return [v for v in ck["model_state"].values()]


if __name__ == "__main__":
    main()
