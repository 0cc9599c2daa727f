"""Oracle of the image path (oracle/pil_bicubic_oracle.py) pinned against Pillow itself and against vectors produced by the
unmodified reference process_images (tests/golden/preprocess_reference.npz)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import pil_bicubic_oracle as O  # noqa: E402
from make_preprocess_golden import CASES, make_image  # noqa: E402


def test_oracle_matches_reference_golden():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "preprocess_reference.npz"))
    for i, (H, W, S) in enumerate(CASES):
        got = O.process_image(make_image(H, W, 100 + i), S)
        assert got.dtype == np.float32 and got.shape == (3, S, S)
        assert np.array_equal(got, gold[f"case{i}"]), f"case {i}"


@pytest.mark.parametrize("H,W,S", [(480, 640, 224), (333, 1000, 224), (150, 200, 224), (224, 300, 224), (500, 224, 224), (97, 61, 64)])
def test_oracle_resize_bit_exact_with_pillow(H, W, S):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(H * 7 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((S, S), resample=Image.Resampling.BICUBIC))
    assert np.array_equal(O.resize_bicubic_u8(img, S), ref)
