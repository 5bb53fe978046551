"""Device-side image preprocessing with the call shape of the reference's HF image processor.

The reference prepares every image on the CPU (`groma/eval/run_groma.py:77-79`, `groma/data/datasets/groma.py:94-96`):
    raw = Image.open(f).convert('RGB').resize((448, 448))                          # Pillow BICUBIC on uint8
    pixel_values = vis_processor.preprocess(raw, return_tensors='pt')['pixel_values']
`GromaImageProcessor.preprocess(images)` takes the un-resized RGB images (PIL / numpy HWC uint8 / torch uint8 HWC), ships
the raw bytes to the GPU and runs the resize (bit-identical to Pillow) + rescale + normalize there; the result is the
float32 [B, 3, 448, 448] cuda tensor `GromaModel.generate(images=...)` takes.  No CPU fallback.
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np
import torch

from . import ops as G

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _byte_table(mean: Sequence[float], std: Sequence[float], rescale_factor: float) -> np.ndarray:
    """float32 [3, 256]: what transformers 4.32 `rescale` then `normalize` return for each byte value
    (float32(uint8 * factor computed in double), then (x - float32(mean)) / float32(std) in float32)."""
    x = (np.arange(256, dtype=np.uint8) * float(rescale_factor)).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)[:, None]
    s = np.asarray(std, dtype=np.float32)[:, None]
    return ((x[None, :] - m) / s).astype(np.float32)


class GromaImageProcessor:
    """Mirror of the `BitImageProcessor` use in the reference scripts, with the PIL resize folded in."""

    model_input_names = ["pixel_values"]

    def __init__(self, size: int = 448, image_mean: Sequence[float] = IMAGENET_DEFAULT_MEAN,
                 image_std: Sequence[float] = IMAGENET_DEFAULT_STD, rescale_factor: float = 1 / 255, device: str = "cuda"):
        self.size = int(size)
        self.image_mean, self.image_std, self.rescale_factor = tuple(image_mean), tuple(image_std), float(rescale_factor)
        self.device = device
        self._lut = None

    def _table(self) -> torch.Tensor:
        if self._lut is None:
            self._lut = torch.from_numpy(_byte_table(self.image_mean, self.image_std, self.rescale_factor)).to(self.device)
        return self._lut

    @staticmethod
    def _to_hwc_u8(img) -> torch.Tensor:
        if isinstance(img, torch.Tensor):
            t = img
        else:
            if hasattr(img, "convert"):           # PIL.Image: same `.convert('RGB')` the reference applies
                img = np.array(img.convert("RGB"))            # writable copy
            t = torch.from_numpy(np.ascontiguousarray(img))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError(f"expected an RGB uint8 HWC image, got {tuple(t.shape)} {t.dtype}")
        return t.contiguous()

    def preprocess(self, images, return_tensors: str = "pt", **unused) -> dict:
        if return_tensors != "pt":
            raise ValueError("only return_tensors='pt' (cuda tensors) is supported")
        if not isinstance(images, (list, tuple)):
            images = [images]
        S = self.size
        out = torch.empty((len(images), 3, S, S), dtype=torch.float32, device=self.device)
        lut = self._table()
        for i, im in enumerate(images):
            t = self._to_hwc_u8(im)
            if not t.is_cuda:
                t = t.pin_memory().to(self.device, non_blocking=True)
            G.preprocess_image(t, lut, S, out_f32=out[i])
        return {"pixel_values": out}

    __call__ = preprocess
