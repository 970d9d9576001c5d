"""Device-side data contract of the reference's VOC pipeline.

The reference builds every sample in Python: PIL `Pad(10)` / `CenterCrop` / `ToTensor` / `Normalize` for the image
(main.py:17-23) and a per-pixel Python loop `to_mask` for the label map (datasets/voc.py:56-72, ~65k `list.index`
calls per 256x256 sample); its sample dump maps labels back to colours with `to_rgb` (datasets/voc.py:74-89,
trainer.py:193-194).  Once the training step runs at thousands of images per second these loops are the bottleneck
by orders of magnitude, so the same functions are offered here on the device, same names, same results (the image
tensor bit-equal in fp32, the label map and colours exact):

    prepare_batch(images, masks, h, w)  ==  torch.stack of VOC.__getitem__ over decoded uint8 RGB arrays
    to_mask(mask_rgb)                   ==  datasets.voc.to_mask
    to_rgb(labels)                      ==  datasets.voc.to_rgb

Decoding JPEG/PNG files stays on the host (PIL), as in the reference.
"""
import torch

from . import _lib

PAD = 10  # transforms.Pad(10): main.py:18, datasets/voc.py:136


def crop_origin(hs, ws, h, w, pad=PAD):
    """Pad(pad) followed by CenterCrop((h, w)) as one window: output pixel (i, j) reads source pixel
    (i + top, j + left), zero outside the hs x ws source.  torchvision `functional.center_crop` arithmetic
    (zero padding of too-small images, Python's round-half-to-even for the offsets)."""
    ih, iw = hs + 2 * pad, ws + 2 * pad
    off_t = off_l = 0
    if w > iw or h > ih:
        off_l = (w - iw) // 2 if w > iw else 0
        off_t = (h - ih) // 2 if h > ih else 0
        ih += off_t + ((h - ih + 1) // 2 if h > ih else 0)
        iw += off_l + ((w - iw + 1) // 2 if w > iw else 0)
        if w == iw and h == ih:
            return -off_t - pad, -off_l - pad
    return int(round((ih - h) / 2.0)) - off_t - pad, int(round((iw - w) / 2.0)) - off_l - pad


def _as_device_u8(a, dev):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("expected a uint8 [H, W, 3] RGB array")
    return t.contiguous().to(dev, non_blocking=True)


def prepare_batch(images, masks, h, w, device=None, pad=PAD):
    """images / masks: sequences of decoded uint8 [Hs, Ws, 3] RGB arrays (numpy or torch, any sizes; `masks` may be
    None).  Returns (x fp32 [B, 3, h, w] in [-1, 1], y int64 [B, h, w]) on the device: the batch the reference's
    DataLoader would have collated from `VOC.__getitem__` (datasets/voc.py:127-140)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _lib.ensure_device(dev.index)
    b = len(images)
    if masks is not None and len(masks) != b:
        raise ValueError("images and masks differ in length")
    keep, rows = [], []
    for k in range(b):
        im = _as_device_u8(images[k], dev)
        mk = _as_device_u8(masks[k], dev) if masks is not None else None
        if mk is not None and mk.shape != im.shape:
            raise ValueError("image and mask differ in size")
        hs, ws = int(im.shape[0]), int(im.shape[1])
        top, left = crop_origin(hs, ws, h, w, pad)
        rows.append([im.data_ptr(), mk.data_ptr() if mk is not None else 0, hs, ws, top, left, 0, 0])
        keep += [im, mk]
    items = torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)
    x = torch.empty((b, 3, h, w), device=dev, dtype=torch.float32)
    y = torch.empty((b, h, w), device=dev, dtype=torch.int64) if masks is not None else None
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    _lib.call("clk_voc_prepare_batch", items, b, h, w, x, y, err)
    if masks is not None and int(err.item()):  # also keeps `keep` alive until the kernel has run
        raise ValueError("a mask colour is not in list")  # what palette.index raises in the reference
    del keep
    return x, y


def to_mask(mask_rgb, device=None):
    """datasets.voc.to_mask: uint8 [H, W, 3] palette image -> int64 [H, W] class indices (void -> 0)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _lib.ensure_device(dev.index)
    mk = _as_device_u8(mask_rgb, dev)
    hs, ws = int(mk.shape[0]), int(mk.shape[1])
    items = torch.tensor([[0, mk.data_ptr(), hs, ws, 0, 0, 0, 0]], dtype=torch.int64).to(dev)
    y = torch.empty((1, hs, ws), device=dev, dtype=torch.int64)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    _lib.call("clk_voc_prepare_batch", items, 1, hs, ws, None, y, err)
    if int(err.item()):
        raise ValueError("a mask colour is not in list")
    return y[0]


def to_rgb(xs):
    """datasets.voc.to_rgb: int64 [B, H, W] class indices (device) -> float64 [B, 3, H, W] palette colours."""
    if not xs.is_cuda or xs.dtype != torch.int64 or xs.dim() != 3:
        raise ValueError("to_rgb expects an int64 [B, H, W] CUDA tensor")
    _lib.ensure_device(xs.device.index)
    xs = xs.contiguous()
    b, h, w = xs.shape
    out = torch.empty((b, 3, h, w), device=xs.device, dtype=torch.float64)
    _lib.call("clk_labels_to_rgb", xs, b, h * w, out)
    return out


# ------------------------------------------------------------------------------------------------
# Dataset side: the reference's file handling (datasets/voc.py:91-148) with the per-pixel work moved to the device.
class VOCDecoded(torch.utils.data.Dataset):
    """`datasets.voc.VOC` up to and including the PIL decode: same directory layout and file lists
    (`make_path`, datasets/voc.py:91-113 — like the reference, 'val' also reads train.txt), same
    `Image.open(...).convert('RGB')` (datasets/voc.py:129-130); returns the two decoded uint8 [H, W, 3] arrays.
    Pad / CenterCrop / ToTensor / Normalize / to_mask happen per batch in `prepare_batch`."""

    def __init__(self, root, dataset_type="train"):
        import os
        assert dataset_type in ["train", "val"], "dataset_type should be in train/val"
        img_path = os.path.join(root, "VOC2012", "JPEGImages")
        mask_path = os.path.join(root, "VOC2012", "SegmentationClass")
        with open(os.path.join(root, "VOC2012", "ImageSets", "Segmentation", "train.txt")) as f:
            names = [line.strip("\n") for line in f.readlines()]
        self.items = [(os.path.join(img_path, n + ".jpg"), os.path.join(mask_path, n + ".png")) for n in names]
        self.dataset_type = dataset_type

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        import numpy as np
        from PIL import Image
        name = self.items[i]
        image = np.asarray(Image.open(name[0]).convert("RGB"))
        mask = np.asarray(Image.open(name[1]).convert("RGB"))
        return torch.from_numpy(np.ascontiguousarray(image)), torch.from_numpy(np.ascontiguousarray(mask))


class DeviceBatches:
    """DataLoader-like iterable over a dataset of decoded (image, mask) uint8 pairs that yields the batches the
    reference's DataLoader would have produced, built on the device by `prepare_batch`:
    (x fp32 [B, 3, h, w] in [-1, 1], y int64 [B, h, w]), both CUDA tensors."""

    def __init__(self, dataset, batch_size, image_size, shuffle=False, drop_last=False, num_workers=0, device=None):
        self.dataset = dataset
        self.h, self.w = int(image_size[0]), int(image_size[1])
        self.device = device
        self.loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last,
                                                  num_workers=num_workers, collate_fn=list, pin_memory=False)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            images = [b[0] for b in batch]
            masks = [b[1] for b in batch]
            yield prepare_batch(images, masks, self.h, self.w, device=self.device)
