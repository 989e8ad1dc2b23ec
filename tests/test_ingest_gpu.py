"""SURVEY 8(f) row 1 on the device: uint8 crops -> cv2-exact resize -> ToTensor -> Normalize (utils/dataset.py:47-69), fused
with PatchEmbed's unfold. Byte/integer work + a fixed float formula: BIT-EXACT against oracle/frame_ingest.py (which is
pinned against the real reference function, tests/test_frame_ingest.py) and against the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _u8(x):
    return np.rint((x * 0.5 + 0.5) * 255.0).astype(np.uint8)


def test_ingest_matches_reference_golden_bit_exact():
    from facialmmt_b200.models import frame_ingest
    from oracle import frame_ingest as fi
    G = np.load(os.path.join(HERE, "golden", "frame_ingest_v1.npz"))
    n = len(G["ref_noipp_sha256"])
    for i in range(n):
        crop = G[f"crop{i}"]
        got = frame_ingest(torch.from_numpy(crop)[None]).cpu().numpy()[0]
        digest = hashlib.sha256(np.ascontiguousarray(_u8(got)).tobytes()).hexdigest()
        assert digest == str(G["ref_noipp_sha256"][i]), f"crop {i} {crop.shape}: device ingest differs from the reference"
        assert np.array_equal(got[:, ::97, ::89], G["ref_f32_samples"][i])
        assert np.array_equal(got, fi.ingest_frame(crop)), f"crop {i} {crop.shape}"


@pytest.mark.parametrize("h,w", [(112, 112), (30, 41), (57, 57), (111, 97), (223, 223), (224, 224), (225, 225), (300, 260),
                                 (448, 448), (448, 672), (500, 333), (672, 672), (1000, 1000)])
def test_ingest_random_sizes_bit_exact_vs_oracle(h, w):
    """cubic (h < 224, any width), copy (224), area integer ratios and general ratios; a batch of 3 crops per size."""
    from facialmmt_b200.models import frame_ingest
    from oracle import frame_ingest as fi
    rng = np.random.default_rng(h * 1000 + w)
    crops = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    crops[1, :, :, :] = np.where(rng.random((h, w, 1)) < 0.5, 0, 255).astype(np.uint8)      # saturating edges (overshoot clamps)
    got = frame_ingest(torch.from_numpy(crops)).cpu().numpy()
    ref = fi.ingest_frames(crops)
    assert np.array_equal(got, ref), f"{h}x{w}: {np.abs(got - ref).max()} max diff, {(got != ref).mean():.2e} of the pixels"


def test_ingest_rejects_what_the_reference_rejects(lib):
    from facialmmt_b200 import _lib
    from facialmmt_b200.models import frame_ingest
    with pytest.raises(_lib.FmmtError):
        frame_ingest(torch.zeros(1, 224, 200, 3, dtype=torch.uint8))    # height 224 -> no resize -> X[i,:] = x fails


def test_swin_forward_from_uint8_equals_forward_from_ingested_tensor():
    """The fused path (uint8 crops -> resize/normalise -> unfold inside the Swin forward) is the same function as
    reference-ingest followed by the fp32 forward: identical bits out."""
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.models import SwinForAffwildClassification
    from oracle import frame_ingest as fi
    cfg = FmmtConfig()
    sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    m = SwinForAffwildClassification(cfg, swin_chunk=3, swin_chunk_late=4)
    m.load_state_dict(sd)
    rng = np.random.default_rng(5)
    crops = rng.integers(0, 256, (7, 112, 112, 3), dtype=np.uint8)
    g = -torch.empty(7, 7).exponential_(generator=torch.Generator().manual_seed(2)).log()
    a = m.forward_full(torch.from_numpy(crops), g, want_feat=True)
    b = m.forward_full(torch.from_numpy(fi.ingest_frames(crops)), g, want_feat=True)
    m.check()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
