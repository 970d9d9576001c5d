"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's data contract around the
hot path.

  * `crop_origin`, `prepare_sample`: `transforms.Pad(10)` + `transforms.CenterCrop((h, w))` + `ToTensor` +
    `Normalize(0.5, 0.5)` on the image (main.py:17-23) and Pad + CenterCrop + `to_mask` on the mask
    (datasets/voc.py:135-138); CenterCrop arithmetic restated from torchvision 0.26 `functional.center_crop`
    (the third-party code the reference calls; not vendored in the reference).
  * `to_mask`: datasets/voc.py:56-72 (palette RGB -> class index, void -> 0, unknown colour raises ValueError).
  * `to_rgb`:  datasets/voc.py:74-89 (class index -> palette RGB, float64, indices >= 22 pass through unchanged).
Pinned by tests/golden/voc_contract.npz (outputs of the unmodified reference functions, tests/golden/make_golden_voc.py).
"""
import numpy as np

# datasets/voc.py:33-54
PALETTE = [[0, 0, 0], [128, 0, 0], [0, 128, 0], [128, 128, 0], [0, 0, 128], [128, 0, 128], [0, 128, 128],
           [128, 128, 128], [64, 0, 0], [192, 0, 0], [64, 128, 0], [192, 128, 0], [64, 0, 128], [192, 0, 128],
           [64, 128, 128], [192, 128, 128], [0, 64, 0], [128, 64, 0], [0, 192, 0], [128, 192, 0], [0, 64, 128],
           [224, 224, 192]]
PAD = 10  # main.py:18, datasets/voc.py:136


def crop_origin(hs, ws, h, w, pad=PAD):
    """(top, left): output pixel (i, j) of the h x w crop reads SOURCE pixel (i + top, j + left); positions outside
    the hs x ws source are the zero padding."""
    ih, iw = hs + 2 * pad, ws + 2 * pad
    off_t = off_l = 0
    if w > iw or h > ih:  # center_crop pads a too-small image with zeros first
        off_l = (w - iw) // 2 if w > iw else 0
        off_t = (h - ih) // 2 if h > ih else 0
        pr = (w - iw + 1) // 2 if w > iw else 0
        pb = (h - ih + 1) // 2 if h > ih else 0
        ih, iw = ih + off_t + pb, iw + off_l + pr
        if w == iw and h == ih:
            return -off_t - pad, -off_l - pad
    top = int(round((ih - h) / 2.0))   # Python round: half to even, as torchvision
    left = int(round((iw - w) / 2.0))
    return top - off_t - pad, left - off_l - pad


def _crop(src, h, w):
    hs, ws = src.shape[:2]
    top, left = crop_origin(hs, ws, h, w)
    out = np.zeros((h, w, 3), dtype=np.uint8)
    i0, i1 = max(0, -top), min(h, hs - top)
    j0, j1 = max(0, -left), min(w, ws - left)
    if i1 > i0 and j1 > j0:
        out[i0:i1, j0:j1] = src[i0 + top:i1 + top, j0 + left:j1 + left]
    return out


def to_mask(rgb):
    """uint8 [H, W, 3] -> int64 [H, W]"""
    flat = rgb.reshape(-1, 3).astype(np.int64)
    key = (flat[:, 0] << 16) | (flat[:, 1] << 8) | flat[:, 2]
    out = np.full(key.shape, -1, dtype=np.int64)
    for idx, (r, g, b) in enumerate(PALETTE):
        hit = (key == ((r << 16) | (g << 8) | b)) & (out < 0)
        out[hit] = 0 if idx == 21 else idx   # void -> background (voc.py:67-68)
    if (out < 0).any():
        bad = flat[np.argmax(out < 0)]
        raise ValueError(f"{list(bad)} is not in list")  # palette.index raises in the reference
    return out.reshape(rgb.shape[:2])


def prepare_sample(img_u8, mask_u8, h, w):
    """one `VOC.__getitem__` (datasets/voc.py:127-140) after decoding: uint8 HWC RGB image + mask ->
    (x float32 [3, h, w] in [-1, 1], y int64 [h, w])"""
    im = _crop(img_u8, h, w).astype(np.float32)
    x = (im / np.float32(255.0) - np.float32(0.5)) / np.float32(0.5)   # ToTensor, Normalize (main.py:20-21)
    return np.ascontiguousarray(x.transpose(2, 0, 1)), to_mask(_crop(mask_u8, h, w))


def to_rgb(xs):
    """int64 [B, H, W] -> float64 [B, 3, H, W]"""
    xs = np.asarray(xs)
    out = np.repeat(xs[..., None], 3, axis=-1).astype(np.float64)
    for j in range(22):
        out[xs == j] = PALETTE[j]
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))
