"""Data contract around the hot path: datasets/voc.py `to_mask`, `to_rgb` and the
Pad / CenterCrop / ToTensor / Normalize pipeline of main.py:17-23.

CPU part: the numpy oracle against golden vectors produced by the unmodified reference functions
(tests/golden/make_golden_voc.py), and the host-side crop arithmetic of the package against the oracle's.
GPU part (-m gpu): the CUDA kernels through the C ABI, bit-exact against the oracle and the goldens.
"""
import os

import numpy as np
import pytest
import torch

from oracle import voc_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "voc_contract.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_reproduces_the_reference_pipeline_bit_exactly(gold):
    h, w = int(gold["h"]), int(gold["w"])
    for k in range(int(gold["n"])):
        x, y = voc_ref.prepare_sample(gold[f"img{k}"], gold[f"mask{k}"], h, w)
        assert x.dtype == np.float32 and np.array_equal(x, gold[f"x{k}"]), k
        assert y.dtype == np.int64 and np.array_equal(y, gold[f"y{k}"]), k
    out = voc_ref.to_rgb(gold["rgb_in"])
    assert out.dtype == np.float64 and np.array_equal(out, gold["rgb_out"])


def test_oracle_to_mask_rejects_unknown_colours_like_list_index():
    bad = np.zeros((2, 2, 3), dtype=np.uint8)
    bad[1, 1] = (1, 2, 3)
    with pytest.raises(ValueError, match="not in list"):
        voc_ref.to_mask(bad)


def test_host_crop_arithmetic_matches_the_oracle():
    from continual_learning_b200 import voc
    for hs in (1, 7, 20, 28, 29, 47, 48, 49, 100, 375):
        for ws in (1, 14, 21, 39, 40, 41, 77, 500):
            for h, w in ((48, 40), (256, 256), (17, 33)):
                assert voc.crop_origin(hs, ws, h, w) == voc_ref.crop_origin(hs, ws, h, w), (hs, ws, h, w)


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_prepare_batch_matches_reference_goldens_bit_exactly(gold, lib_built):
    from continual_learning_b200 import voc
    h, w = int(gold["h"]), int(gold["w"])
    n = int(gold["n"])
    x, y = voc.prepare_batch([gold[f"img{k}"] for k in range(n)], [gold[f"mask{k}"] for k in range(n)], h, w)
    assert x.dtype == torch.float32 and y.dtype == torch.int64
    for k in range(n):
        assert np.array_equal(x[k].cpu().numpy(), gold[f"x{k}"]), k   # fp32, bit-equal
        assert np.array_equal(y[k].cpu().numpy(), gold[f"y{k}"]), k


@pytest.mark.gpu
def test_prepare_batch_voc_sized_samples_against_the_oracle(lib_built):
    from continual_learning_b200 import voc
    rng = np.random.Generator(np.random.PCG64(5))
    pal = np.asarray(voc_ref.PALETTE, dtype=np.uint8)
    imgs, masks = [], []
    for hs, ws in ((375, 500), (500, 333), (256, 256), (236, 236), (120, 640), (281, 500)):
        imgs.append(rng.integers(0, 256, size=(hs, ws, 3), dtype=np.uint8))
        masks.append(pal[rng.integers(0, 22, size=(hs, ws))])
    x, y = voc.prepare_batch(imgs, masks, 256, 256)
    for k in range(len(imgs)):
        xr, yr = voc_ref.prepare_sample(imgs[k], masks[k], 256, 256)
        assert np.array_equal(x[k].cpu().numpy(), xr) and np.array_equal(y[k].cpu().numpy(), yr), k
    assert int(y.min()) == 0 and int(y.max()) == 20  # void (21) became background
    # images only (the test-time loader), and the stand-alone to_mask
    x2, y2 = voc.prepare_batch(imgs, None, 256, 256)
    assert y2 is None and torch.equal(x2, x)
    assert np.array_equal(voc.to_mask(masks[0]).cpu().numpy(), voc_ref.to_mask(masks[0]))


@pytest.mark.gpu
def test_unknown_mask_colour_raises_like_the_reference(lib_built):
    from continual_learning_b200 import voc
    img = np.zeros((30, 30, 3), dtype=np.uint8)
    mask = np.zeros((30, 30, 3), dtype=np.uint8)
    mask[12, 17] = (1, 2, 3)
    with pytest.raises(ValueError, match="not in list"):
        voc.prepare_batch([img], [mask], 32, 32)


@pytest.mark.gpu
def test_to_rgb_matches_reference_golden_and_oracle(gold, lib_built):
    from continual_learning_b200 import voc
    out = voc.to_rgb(torch.from_numpy(gold["rgb_in"]).cuda())
    assert out.dtype == torch.float64 and np.array_equal(out.cpu().numpy(), gold["rgb_out"])
    rng = np.random.Generator(np.random.PCG64(6))
    labels = rng.integers(-2, 30, size=(4, 64, 48))
    assert np.array_equal(voc.to_rgb(torch.from_numpy(labels).cuda()).cpu().numpy(), voc_ref.to_rgb(labels))
