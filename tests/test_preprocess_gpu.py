"""GPU image path (csrc/preprocess.cu through the C ABI) against the oracle, the reference golden vectors and Pillow:
bit-exact (integer resample, table-mapped float values)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import pil_bicubic_oracle as O  # noqa: E402
from make_preprocess_golden import CASES, make_image  # noqa: E402


def test_gpu_matches_reference_golden():
    from paligemma_multimodal_system_b200.image_preprocess import process_images_gpu
    gold = np.load(os.path.join(ROOT, "tests", "golden", "preprocess_reference.npz"))
    for i, (H, W, S) in enumerate(CASES):
        got = process_images_gpu([make_image(H, W, 100 + i)], S).cpu().numpy()[0]
        assert np.array_equal(got, gold[f"case{i}"]), f"case {i}"


@pytest.mark.parametrize("H,W,S", [(480, 640, 224), (333, 1000, 224), (150, 200, 224), (224, 300, 224), (500, 224, 224),
                                   (224, 224, 224), (900, 1200, 448), (1, 5, 32), (2000, 3000, 896)])
def test_gpu_matches_oracle_and_pillow(H, W, S):
    from paligemma_multimodal_system_b200.image_preprocess import process_images_gpu
    img = np.random.default_rng(H * 7 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    out = torch.full((1, 3, S, S), float("nan"), device="cuda")
    got = process_images_gpu([img], S, out=out).cpu().numpy()[0]
    assert np.array_equal(got, O.process_image(img, S))
    Image = pytest.importorskip("PIL.Image")
    ref = np.array(Image.fromarray(img).resize((S, S), resample=Image.Resampling.BICUBIC))
    x = ((ref * (1 / 255.0)).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(got, x.transpose(2, 0, 1))


def test_gpu_batch_of_mixed_sizes_and_grey_images():
    from paligemma_multimodal_system_b200.image_preprocess import process_images_gpu
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    imgs = [Image.fromarray(rng.integers(0, 256, (120, 80, 3), dtype=np.uint8)),
            Image.fromarray(rng.integers(0, 256, (64, 200), dtype=np.uint8), mode="L"),
            rng.integers(0, 256, (300, 300, 3), dtype=np.uint8)]
    got = process_images_gpu(imgs, 224).cpu().numpy()
    for b, im in enumerate(imgs):
        pil = Image.fromarray(im) if isinstance(im, np.ndarray) else im
        ref = np.array(pil.resize((224, 224), resample=Image.Resampling.BICUBIC).convert("RGB"))  # the reference's order
        x = ((ref * (1 / 255.0)).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
        assert np.array_equal(got[b], x.transpose(2, 0, 1)), f"image {b}"
    with pytest.raises(NotImplementedError):
        process_images_gpu([Image.fromarray(rng.integers(0, 256, (8, 8, 4), dtype=np.uint8), mode="RGBA")], 32)
