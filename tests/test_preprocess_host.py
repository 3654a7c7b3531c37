"""CPU tests for SURVEY §8 row f2: the preprocessing oracle (oracle/preprocess_oracle.py) is pinned
  * bit-exactly to Pillow's own bicubic resize (the third-party code CLIPImageProcessor 4.31 calls) on many sizes, and
  * to outputs of the reference's `process_images` (mm_utils.py:28-40) committed in tests/golden/preprocess.npz
    (written by tools/make_golden.py, pad and non-pad modes)."""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

PIL = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("h,w,oh,ow", [(50, 70, 33, 41), (100, 60, 336, 201), (37, 37, 80, 80), (500, 333, 336, 224),
                                        (64, 48, 64, 30), (30, 90, 45, 90), (640, 480, 448, 336), (336, 336, 336, 336),
                                        (1000, 700, 480, 336), (5, 9, 336, 604)])
def test_resize_is_bit_exact_with_pillow(h, w, oh, ow):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(PIL.fromarray(img).resize((ow, oh), PIL.BICUBIC))
    assert np.array_equal(po.resize_bicubic_u8(img, ow, oh), ref)


def test_process_images_matches_reference_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess.npz"))
    n = 0
    for key in z.files:
        if not key.startswith("out_"):
            continue
        _, size, k, mode = key.split("_")
        img = z[f"img_{size}_{k}"]
        got = po.clip_preprocess(img, int(size), pad=(mode == "pad"))
        assert got.dtype == np.float32 and np.array_equal(got, z[key]), key  # float32, bit for bit
        n += 1
    assert n == 8


def test_geometry_helpers():
    assert po.resize_target(480, 640, 336) == (336, 448)
    assert po.resize_target(640, 480, 336) == (448, 336)
    assert po.resize_target(300, 700, 336) == (336, 784)
    sq = po.expand2square(np.zeros((2, 6, 3), np.uint8), (122, 116, 104))
    assert sq.shape == (6, 6, 3) and (sq[0] == (122, 116, 104)).all() and (sq[2:4] == 0).all() and (sq[5] == (122, 116, 104)).all()
    lut = po.normalize_lut()
    assert lut.shape == (3, 256) and abs(lut[0, 0] + 0.48145466 / 0.26862954) < 1e-6
