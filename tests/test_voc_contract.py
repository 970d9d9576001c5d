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


def _fake_voc_tree(root, sizes, seed=17):
    """a miniature VOC2012 directory (datasets/voc.py:91-113 layout) with random JPEG images and palette PNG masks"""
    from PIL import Image
    rng = np.random.Generator(np.random.PCG64(seed))
    base = os.path.join(root, "VOC2012")
    for sub in ("JPEGImages", "SegmentationClass", os.path.join("ImageSets", "Segmentation")):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    pal = np.asarray(voc_ref.PALETTE, dtype=np.uint8)
    names = []
    for k, (hs, ws) in enumerate(sizes):
        name = f"2007_{k:06d}"
        names.append(name)
        Image.fromarray(rng.integers(0, 256, size=(hs, ws, 3), dtype=np.uint8)).save(
            os.path.join(base, "JPEGImages", name + ".jpg"), quality=90)
        cls = np.kron(rng.integers(0, 22, size=(hs // 8 + 1, ws // 8 + 1)), np.ones((8, 8), dtype=np.int64))[:hs, :ws]
        Image.fromarray(pal[cls]).save(os.path.join(base, "SegmentationClass", name + ".png"))
    with open(os.path.join(base, "ImageSets", "Segmentation", "train.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return names


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout is only in the build container")
def test_decoded_dataset_plus_oracle_equals_the_reference_dataset_class(tmp_path):
    """live pin of the whole `VOC.__getitem__` (file lists, PIL decode, transforms, to_mask) where the reference exists:
    reference dataset == VOCDecoded (same files, same decode) + oracle.prepare_sample, bit for bit"""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from torchvision import transforms
        from datasets.voc import VOC
    finally:
        sys.path.remove("/root/reference")
    from continual_learning_b200 import voc
    sizes = [(60, 80), (33, 47), (90, 50), (20, 30)]
    _fake_voc_tree(str(tmp_path), sizes)
    h, w = 48, 40
    tf = transforms.Compose([transforms.Pad(10), transforms.CenterCrop((h, w)), transforms.ToTensor(),
                             transforms.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])  # main.py:17-22
    ref_ds = VOC(root=str(tmp_path), image_size=(h, w), dataset_type="train", transform=tf)
    ds = voc.VOCDecoded(str(tmp_path), "train")
    assert len(ds) == len(ref_ds) == len(sizes)
    for i in range(len(ds)):
        xr, yr = ref_ds[i]
        img, mask = ds[i]
        x, y = voc_ref.prepare_sample(img.numpy(), mask.numpy(), h, w)
        assert np.array_equal(x, xr.numpy()) and np.array_equal(y, yr.numpy()), i


@pytest.mark.gpu
def test_device_batches_from_files_match_the_oracle(tmp_path, lib_built):
    from continual_learning_b200 import voc
    sizes = [(375, 500), (500, 375), (300, 300), (120, 200), (281, 500), (256, 256)]
    _fake_voc_tree(str(tmp_path), sizes)
    ds = voc.VOCDecoded(str(tmp_path), "val")
    loader = voc.DeviceBatches(ds, batch_size=4, image_size=(256, 256), shuffle=False, drop_last=False)
    assert len(loader) == 2 and len(loader.dataset) == 6
    k = 0
    for x, y in loader:
        assert x.is_cuda and x.dtype == torch.float32 and y.dtype == torch.int64
        for b in range(x.shape[0]):
            img, mask = ds[k]
            xr, yr = voc_ref.prepare_sample(img.numpy(), mask.numpy(), 256, 256)
            assert np.array_equal(x[b].cpu().numpy(), xr) and np.array_equal(y[b].cpu().numpy(), yr), k
            k += 1
    assert k == 6
