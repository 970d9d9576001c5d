"""Seeded synthetic VOC-shaped batches: inputs for `main.py --synthetic`, `bench.py` and the tests.

x: fp32 [B,3,H,W] in [-1,1] (main.py:21-22 normalisation), y: int64 [B,H,W] in [0,20] (voc.py:56-72).
Generated with numpy PCG64 so fixtures do not depend on the torch RNG implementation.
"""
import numpy as np
import torch


def uniform_batch(seed, b, h, w, num_classes=21):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(-1.0, 1.0, size=(b, 3, h, w)).astype(np.float32)
    y = rng.integers(0, num_classes, size=(b, h, w), dtype=np.int64)
    return torch.from_numpy(x), torch.from_numpy(y)


def structured_batch(seed, b, h, w, num_classes=21, cell=32, noise=0.3):
    """class map at (H/cell x W/cell) nearest-upsampled; image = per-class colour + noise, clamped to [-1,1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gh, gw = max(1, h // cell), max(1, w // cell)
    grid = rng.integers(0, num_classes, size=(b, gh, gw), dtype=np.int64)
    y = np.repeat(np.repeat(grid, h // gh, axis=1), w // gw, axis=2)
    colours = np.random.Generator(np.random.PCG64(12345)).uniform(-0.8, 0.8, size=(num_classes, 3)).astype(np.float32)
    x = colours[y].transpose(0, 3, 1, 2) + noise * rng.standard_normal((b, 3, h, w)).astype(np.float32)
    return torch.from_numpy(np.clip(x, -1.0, 1.0).astype(np.float32)), torch.from_numpy(y)
