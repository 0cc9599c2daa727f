"""GPU image path of PaliGemmaProcessor (processing_paligemma.py:13-73): bicubic resize + rescale + normalise + CHW in
two kernels per image (csrc/preprocess.cu), bit-exact with the reference's PIL / numpy path for 8-bit RGB (and L) images.

The host computes, once per (input size, output size), the per-output-pixel windows and 22-bit fixed-point weights
exactly as Pillow's Resample.c does (precompute_coeffs + normalize_coeffs_8bpc: Keys bicubic, a = -0.5, support
2 * max(scale, 1), weights normalised in double precision, rounded half away from zero), and the 256-entry value table
with the reference's own numpy expression.  There is no CPU fallback: without a GPU this raises.
"""
import math
from functools import lru_cache
from typing import List

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=64)
def _coefficients(in_size: int, out_size: int):
    """Device tensors (kk int32 [out, ksize], bounds int32 [out, 2]) + ksize for one axis; identity when sizes agree
    (Pillow skips the pass then, and a single weight of 1.0 reproduces the pixel exactly)."""
    if in_size == out_size:
        kk = np.full((out_size, 1), 1 << PRECISION_BITS, dtype=np.int32)
        bounds = np.stack([np.arange(out_size, dtype=np.int32), np.ones(out_size, dtype=np.int32)], 1)
        return torch.from_numpy(kk).cuda(), torch.from_numpy(np.ascontiguousarray(bounds)).cuda(), 1
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        x0 = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - x0
        w = [_bicubic((x + x0 - center + 0.5) * inv) for x in range(n)]
        total = 0.0
        for v in w:
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (x0, n)
    return torch.from_numpy(kk).cuda(), torch.from_numpy(bounds).cuda(), ksize


@lru_cache(maxsize=4)
def _value_table(scale_factor: float, mean: float, std: float):
    """uint8 -> float32 exactly as rescale() and normalise() of the reference compute it (processing_paligemma.py:21-33)."""
    u = np.arange(256, dtype=np.uint8)
    x = (u * scale_factor).astype(np.float32)
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    return torch.from_numpy(x.astype(np.float32)).cuda()


@torch.no_grad()
def process_images_gpu(images: List, image_size: int, scale_factor: float = 1 / 255.0, mean: float = 0.5, std: float = 0.5,
                       out: torch.Tensor = None) -> torch.Tensor:
    """images: PIL images (mode RGB or L) or uint8 HWC arrays / tensors -> float32 CUDA tensor [B, 3, S, S]."""
    _lib.require_device()
    L, st = _lib.lib(), _lib.stream()
    S = int(image_size)
    B = len(images)
    if out is None:
        out = torch.empty(B, 3, S, S, device="cuda", dtype=torch.float32)
    lut = _value_table(float(scale_factor), float(mean), float(std))
    for b, img in enumerate(images):
        if not isinstance(img, (np.ndarray, torch.Tensor)) and hasattr(img, "getbands"):  # PIL image
            if img.mode == "L":
                img = img.convert("RGB")  # channel replication commutes with the per-channel resample
            elif img.mode != "RGB":
                raise NotImplementedError(f"GPU image path handles 8-bit RGB / L images, got mode {img.mode!r}")
            img = np.asarray(img)
        t = torch.as_tensor(img)
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError("expected a uint8 [H, W, 3] image")
        src = t.contiguous().cuda(non_blocking=True)
        H, W = int(src.shape[0]), int(src.shape[1])
        kx, bx, ksx = _coefficients(W, S)
        ky, by, ksy = _coefficients(H, S)
        if W != S:
            tmp = torch.empty(H, S, 3, device="cuda", dtype=torch.uint8)
            _lib.check(L.pg_resample_h_u8(src.data_ptr(), tmp.data_ptr(), H, W, S, kx.data_ptr(), bx.data_ptr(), ksx, st), "pg_resample_h_u8")
        else:
            tmp = src
        _lib.check(L.pg_resample_v_u8_norm(tmp.data_ptr(), out[b].data_ptr(), H, S, S, ky.data_ptr(), by.data_ptr(), ksy,
                                           lut.data_ptr(), st), "pg_resample_v_u8_norm")
    return out
