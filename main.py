#!/usr/bin/env python
"""main.py — the reference CLI (main.py:63-107: same flags, same defaults) driving the B200 backend.

    python main.py --mode train --train_batch_size 16 --h_image_size 256 --w_image_size 256 --synthetic
    python -m torch.distributed.run --nproc-per-node 8 main.py ... --train_batch_size 128   # batch sharded by rank

Additive flags: --synthetic/--synthetic_size (seeded VOC-shaped data, no dataset on disk), --device_preprocess
(VOC files decoded on the host, everything per-pixel on the GPU), the continual-learning
options (--old_model_path, --num_old_classes, --distill_T, --distill_lambda) and --seed.
"""
import argparse
import os

import torch
from torch.utils.data import DataLoader, Dataset


class SyntheticVOC(Dataset):
    """VOC-shaped samples: x fp32 [3,H,W] in [-1,1], y int64 [H,W] in [0,20] (datasets/voc.py:127-144 contract)."""

    def __init__(self, n, h, w):
        self.n, self.h, self.w = n, h, w

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        from continual_learning_b200.synthetic import structured_batch
        x, y = structured_batch(i, 1, self.h, self.w)
        return x[0], y[0]


def get_loader(config):
    """train / "val" loaders (main.py:17-43; like the reference the val loader iterates the training set)."""
    if config.synthetic:
        ds = SyntheticVOC(config.synthetic_size, config.h_image_size, config.w_image_size)
    elif config.device_preprocess:
        # same files, same PIL decode; Pad / CenterCrop / ToTensor / Normalize / to_mask run on the GPU per batch
        from continual_learning_b200 import voc
        ds = voc.VOCDecoded(config.path, "train")
        size = (config.h_image_size, config.w_image_size)
        train = voc.DeviceBatches(ds, config.train_batch_size, size, shuffle=True, drop_last=True,
                                  num_workers=config.num_workers)
        val = voc.DeviceBatches(ds, config.val_batch_size, size, shuffle=False, num_workers=config.num_workers)
        return train, val
    else:
        from torchvision import transforms
        from datasets.voc import VOC  # the reference's dataset module (not part of this package)
        tf = transforms.Compose([transforms.Pad(10), transforms.CenterCrop((config.h_image_size, config.w_image_size)),
                                 transforms.ToTensor(), transforms.Normalize(mean=(0.5,) * 3, std=(0.5,) * 3)])
        ds = VOC(root=config.path, image_size=(config.h_image_size, config.w_image_size), dataset_type="train", transform=tf)
    # seeded shuffle: under torchrun every rank must draw the SAME global batch before taking its shard
    train = DataLoader(ds, batch_size=config.train_batch_size, shuffle=True, drop_last=True,
                       num_workers=config.num_workers, pin_memory=True,
                       generator=torch.Generator().manual_seed(config.seed))
    val = DataLoader(ds, batch_size=config.val_batch_size, shuffle=False, num_workers=config.num_workers, pin_memory=True)
    return train, val


def main(config):
    from continual_learning_b200.trainer import Trainer
    os.makedirs(config.model_save_path, exist_ok=True)
    os.makedirs(config.sample_save_path, exist_ok=True)
    if config.mode == "train":
        train_loader, val_loader = get_loader(config)
        Trainer(train_data_loader=train_loader, val_data_loader=val_loader, config=config).train_val()


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--mode", type=str, default="train", choices=["train"])
    p.add_argument("--model", type=str, default="unet", choices=["unet", "fcn8", "pspnet_avg", "pspnet_max", "dfnet"])
    p.add_argument("--dataset", type=str, default="voc", choices=["voc"])
    p.add_argument("--n_iters", type=int, default=10000)
    p.add_argument("--train_batch_size", type=int, default=2)
    p.add_argument("--val_batch_size", type=int, default=2)
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--lr_exp", type=float, default=0.9)
    p.add_argument("--beta1", type=float, default=5e-1)
    p.add_argument("--beta2", type=float, default=0.99)
    p.add_argument("--h_image_size", type=int, default=512)
    p.add_argument("--w_image_size", type=int, default=256)
    p.add_argument("--model_save_path", type=str, default="./model")
    p.add_argument("--sample_save_path", type=str, default="./sample")
    p.add_argument("--path", type=str, default="./dataset")
    p.add_argument("--log_step", type=int, default=1)
    p.add_argument("--val_step", type=int, default=1000)
    p.add_argument("--model_save_step", type=int, default=10, help="Saving epoch")
    p.add_argument("--sample_save_step", type=int, default=10, help="Saving epoch")
    p.add_argument("--continue_train", action="store_true", help="continue training: load the latest model")
    p.add_argument("--which_epoch", type=str, default="latest")
    p.add_argument("--num_workers", type=int, default=4)
    # additive
    p.add_argument("--synthetic", action="store_true", help="seeded synthetic VOC-shaped data instead of --path")
    p.add_argument("--synthetic_size", type=int, default=64)
    p.add_argument("--device_preprocess", action="store_true",
                   help="decode VOC files with PIL like datasets/voc.py, do the per-pixel work (crop, normalise, "
                        "palette -> label) on the GPU")
    p.add_argument("--old_model_path", type=str, default=None, help="checkpoint of the frozen previous-task network")
    p.add_argument("--num_old_classes", type=int, default=16)
    p.add_argument("--distill_T", type=float, default=2.0)
    p.add_argument("--distill_lambda", type=float, default=1.0)
    p.add_argument("--seed", type=int, default=0)
    return p


if __name__ == "__main__":
    config = build_parser().parse_args()
    if int(os.environ.get("RANK", "0")) == 0:
        print(config)
    main(config)
