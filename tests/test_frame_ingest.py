"""Frame-ingest oracle (oracle/frame_ingest.py) against golden vectors produced by the REAL
reference function (utils/dataset.py:47-69, generator: tests/golden/make_ingest_golden.py) and, where cv2 is importable, live."""
import hashlib
import os

import numpy as np
import pytest

from oracle import frame_ingest as fi

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "frame_ingest_v1.npz"))
N_CROPS = len(G["ref_noipp_sha256"])


def _u8(x):
    return np.rint((x * 0.5 + 0.5) * 255.0).astype(np.uint8)


def test_ingest_matches_reference_golden_bit_exact():
    for i in range(N_CROPS):
        x = fi.ingest_frame(G[f"crop{i}"])
        assert x.shape == (3, 224, 224) and x.dtype == np.float32
        assert -1.0 <= x.min() and x.max() <= 1.0
        digest = hashlib.sha256(np.ascontiguousarray(_u8(x)).tobytes()).hexdigest()
        assert digest == str(G["ref_noipp_sha256"][i]), f"crop {i} ({G[f'crop{i}'].shape}) differs from the reference"
        # the float arithmetic of ToTensor + Normalize, bit for bit, at sampled pixels
        assert np.array_equal(x[:, ::97, ::89], G["ref_f32_samples"][i])
    assert np.array_equal(_u8(fi.ingest_frame(G["crop0"])), G["ref_noipp_u8_crop0"])


def test_channel_order_is_kept_and_resize_rule_uses_height():
    crop = np.zeros((112, 112, 3), dtype=np.uint8)
    crop[..., 0] = 255                                     # cv2.imread's channel 0 is BLUE; it stays channel 0
    x = fi.ingest_frame(crop)
    assert np.all(x[0] == 1.0) and np.all(x[1] == -1.0) and np.all(x[2] == -1.0)
    with pytest.raises(ValueError):
        fi.ingest_frame(np.zeros((224, 200, 3), dtype=np.uint8))   # height 224 -> no resize -> the reference fails too


def test_ipp_wheels_differ_only_on_non_integer_cubic():
    """what the generator measured between an IPP-enabled cv2 and OpenCV's own path, per crop"""
    frac, mx = G["ipp_mismatch_fraction"], G["ipp_max_abs_diff"]
    assert mx.max() <= 1
    assert frac[3] > 0 and all(frac[i] == 0 for i in range(N_CROPS) if i != 3)


def test_live_against_cv2_random_sizes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    had_ipp = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for h in (30, 57, 111, 112, 223):
            im = rng.integers(0, 256, (h, h, 3), dtype=np.uint8)
            ref = cv2.resize(im, dsize=(224, 224), interpolation=cv2.INTER_CUBIC)
            got = fi.resize_cubic_u8(im)
            assert (got != ref).mean() < 5e-5 and np.abs(got.astype(int) - ref.astype(int)).max() <= 1
        for h in (225, 300, 448, 500, 672):
            im = rng.integers(0, 256, (h, h, 3), dtype=np.uint8)
            assert np.array_equal(fi.resize_area_u8(im), cv2.resize(im, dsize=(224, 224), interpolation=cv2.INTER_AREA))
        im = rng.integers(0, 256, (112, 112, 3), dtype=np.uint8)   # the 2x case is exact on every build
        cv2.ipp.setUseIPP(True)
        assert np.array_equal(fi.resize_cubic_u8(im), cv2.resize(im, dsize=(224, 224), interpolation=cv2.INTER_CUBIC))
    finally:
        cv2.ipp.setUseIPP(had_ipp)
