"""Writes tests/golden/preprocess_reference.npz: outputs of the UNMODIFIED reference process_images
(/root/reference/processing_paligemma.py:36-73, i.e. Pillow's bicubic resize + numpy rescale / normalise / transpose)
on small seeded synthetic images.  Run in the build container (needs /root/reference and Pillow)."""
import os
import sys

import numpy as np

CASES = [(37, 53, 32), (64, 48, 32), (20, 24, 32), (32, 40, 32), (90, 32, 32)]  # (H, W, output size)


def make_image(H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = np.stack([(xx * 255 // max(W - 1, 1)), (yy * 255 // max(H - 1, 1)), ((xx + yy) * 255 // max(H + W - 2, 1))], -1)
    noise = rng.integers(-40, 41, (H, W, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def main():
    from PIL import Image
    sys.path.insert(0, "/root/reference")
    import processing_paligemma as ref  # the unmodified reference (build container only)
    out = {}
    for i, (H, W, S) in enumerate(CASES):
        img = make_image(H, W, 100 + i)
        res = ref.process_images([Image.fromarray(img)], image_size=S, scale_factor=1 / 255.0, resampling=Image.Resampling.BICUBIC)
        out[f"case{i}"] = np.asarray(res[0], dtype=np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess_reference.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
