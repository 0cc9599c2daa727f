"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the reference's image path, processing_paligemma.py:13-73:

    image.resize((S, S), resample=Image.Resampling.BICUBIC)   (:17-19)   -> np.array(image.convert("RGB"))   (:52)
    -> * 1/255 as float32  (:21-23, 56-59)  -> (x - 0.5) / 0.5  (:25-33, 64-68)  -> HWC -> CHW  (:71)

The resize itself lives in a third-party dependency that is not part of /root/reference: Pillow (no version is pinned by
the reference; this image ships Pillow 12.2.0).  Its published algorithm for 8-bit images (src/libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) is restated here in numpy:
per output pixel a window of `support = 2 * max(scale, 1)` input pixels either side of the centre, weights from the
Keys bicubic kernel (a = -0.5) normalised to sum 1, converted to 22-bit fixed point, accumulated in int32 starting from
1 << 21 and clipped to uint8; horizontal pass first, the vertical pass works on the uint8 result of the horizontal one.
Pinned against Pillow itself by tests/test_preprocess_cpu.py (bit-exact on every case).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int):
    """-> (kk int32 [out, ksize], bounds int32 [out, 2] = (first input index, count), ksize)"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


def resize_bicubic_u8(img: np.ndarray, size: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [size, size, C], bit-exact with PIL.Image.resize(..., BICUBIC) for 8-bit modes."""
    H, W, C = img.shape
    out = img
    if W != size:
        kk, b, _ = coefficients(W, size)
        tmp = np.zeros((H, size, C), dtype=np.uint8)
        for xx in range(size):
            x0, n = b[xx]
            acc = (1 << (PRECISION_BITS - 1)) + (out[:, x0:x0 + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
            tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        out = tmp
    if H != size:
        kk, b, _ = coefficients(H, size)
        res = np.zeros((size, out.shape[1], C), dtype=np.uint8)
        for yy in range(size):
            y0, n = b[yy]
            acc = (1 << (PRECISION_BITS - 1)) + (out[y0:y0 + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
            res[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        out = res
    return out


def process_image(img_rgb_u8: np.ndarray, size: int) -> np.ndarray:
    """The reference's process_images for one RGB uint8 image -> float32 [3, size, size]."""
    arr = resize_bicubic_u8(img_rgb_u8, size)
    x = (arr * (1 / 255.0)).astype(np.float32)                       # rescale (:21-23)
    x = (x - np.array([0.5, 0.5, 0.5], dtype=np.float32)) / np.array([0.5, 0.5, 0.5], dtype=np.float32)  # normalise (:25-33)
    return x.transpose(2, 0, 1)
