"""CPU restatement of the image preprocessing in front of the hot path.  TEST INFRASTRUCTURE ONLY.

Reference call sites: `groma/eval/run_groma.py:77-79`, `groma/eval/run_ddetr.py:43-45`, `groma/data/datasets/groma.py:94-96`:
    image = Image.open(f).convert('RGB').resize((448, 448))            # Pillow, default filter = BICUBIC, uint8
    image = vis_processor.preprocess(image, return_tensors='pt')['pixel_values']
with `vis_processor` = the DINOv2 `BitImageProcessor` run with do_resize=False, do_center_crop=False
(`run_ddetr.py:39-40`, `model_vqa.py:71`), i.e. rescale 1/255 + ImageNet mean/std, CHW float32.

The arithmetic lives in two third-party dependencies that are not vendored in /root/reference:
  * Pillow (`pyproject.toml` pulls it through torchvision/transformers; `src/libImaging/Resample.c`):
    two-pass separable resampling on uint8, horizontal pass first, coefficients computed in double, normalised, converted
    to 22-bit fixed point (PRECISION_BITS = 32-8-2), accumulators start at 1<<21, result clip8(acc >> 22).  Restated below
    from the published algorithm and PINNED bit-exactly against the Pillow installed in this image (tests/test_preprocess_cpu.py).
  * transformers==4.32.0 `image_transforms.rescale/normalize`: float32(uint8 * (1/255) in double), then
    (x - float32(mean)) / float32(std) in float32.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box and the bicubic filter (support 2).
    Returns (ksize, bounds int32 [out,2] = (first input index, tap count), kk int32 [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)       # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow `Image.resize((out_w, out_h))` (BICUBIC) on an HxWxC uint8 array; bit-exact."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape
    if (H, W) == (out_h, out_w):
        return img.copy()
    _, bh, kh = resample_coeffs(W, out_w)
    _, bv, kv = resample_coeffs(H, out_h)
    src = img
    if W != out_w:      # horizontal pass (only when needed, as ImagingResampleInner does)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        s64 = src.astype(np.int64)
        for xx in range(out_w):
            x0, n = bh[xx]
            acc = (s64[:, x0:x0 + n, :] * kh[xx, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp
    if H != out_h:      # vertical pass
        out = np.empty((out_h, out_w, C), dtype=np.uint8)
        s64 = src.astype(np.int64)
        for yy in range(out_h):
            y0, n = bv[yy]
            acc = (s64[y0:y0 + n, :, :] * kv[yy, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        src = out
    return src


def normalize_lut() -> np.ndarray:
    """float32 [3, 256]: transformers 4.32 rescale(1/255) then normalize(mean, std) of every possible byte."""
    v = np.arange(256, dtype=np.uint8)
    x = (v * (1 / 255)).astype(np.float32)                       # uint8 * python float -> float64, then astype(float32)
    mean = np.array(IMAGENET_MEAN, dtype=np.float32)
    std = np.array(IMAGENET_STD, dtype=np.float32)
    return ((x[None, :] - mean[:, None]) / std[:, None]).astype(np.float32)


def preprocess_ref(img: np.ndarray, size: int = 448) -> np.ndarray:
    """uint8 HxWx3 RGB -> float32 [3, size, size] pixel_values (resize + rescale + normalize)."""
    r = resize_bicubic_u8(img, size, size)
    lut = normalize_lut()
    return np.stack([lut[c][r[:, :, c]] for c in range(3)], 0)
