"""Golden vectors for the data contract around the hot path, produced by the UNMODIFIED
reference: `datasets.voc.to_mask`, `datasets.voc.to_rgb` and the torchvision pipeline of main.py:17-23 /
datasets/voc.py:135-138.  Run in the build container:  python tests/golden/make_golden_voc.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

from datasets.voc import palette, to_mask, to_rgb  # /root/reference/datasets/voc.py

H, W = 48, 40
SIZES = [(70, 90), (30, 100), (20, 14), (37, 53), (48, 40), (28, 20), (29, 21), (64, 19), (101, 77)]


def main():
    rng = np.random.Generator(np.random.PCG64(11))
    transform = transforms.Compose([transforms.Pad(10), transforms.CenterCrop((H, W)), transforms.ToTensor(),
                                    transforms.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])  # main.py:17-22
    rec = {"h": np.int64(H), "w": np.int64(W), "n": np.int64(len(SIZES))}
    for k, (hs, ws) in enumerate(SIZES):
        img = rng.integers(0, 256, size=(hs, ws, 3), dtype=np.uint8)
        cls = rng.integers(0, 22, size=(hs // 4 + 1, ws // 4 + 1))          # blocky class map incl. void (21)
        cls = np.kron(cls, np.ones((4, 4), dtype=np.int64))[:hs, :ws]
        mask = np.asarray(palette, dtype=np.uint8)[cls]
        x = transform(Image.fromarray(img))                                  # voc.py:132-133
        m = transforms.Pad(10)(Image.fromarray(mask))                        # voc.py:136
        m = transforms.CenterCrop((H, W))(m)                                 # voc.py:137
        y = to_mask(m)                                                       # voc.py:138
        rec[f"img{k}"], rec[f"mask{k}"] = img, mask
        rec[f"x{k}"], rec[f"y{k}"] = x.numpy(), y.numpy()
    labels = torch.from_numpy(rng.integers(0, 25, size=(3, 9, 7)))           # includes indices >= 22
    rec["rgb_in"] = labels.numpy()
    rec["rgb_out"] = to_rgb(labels).numpy()                                  # voc.py:74-89
    np.savez_compressed(os.path.join(HERE, "voc_contract.npz"), **rec)
    print("wrote voc_contract.npz")


if __name__ == "__main__":
    main()
