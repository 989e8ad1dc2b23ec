"""CPU: the oracle (oracle/facialmmt_oracle.py) against the golden vectors produced by the real reference
(tests/golden/make_golden.py). Tolerance 2e-4 max-abs on O(1) logits: both sides are fp32 CPU, only op order differs."""
import os

import numpy as np
import pytest
import torch

from facialmmt_b200 import synthetic as syn
from facialmmt_b200.config import FmmtConfig, TextConfig
from oracle import facialmmt_oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
TOL = 2e-4


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def _close(a, name, tol=TOL):
    b = torch.from_numpy(G[name])
    err = (a - b).abs().max().item()
    assert err < tol, f"{name}: max-abs {err}"


def test_swin_logits_and_features():
    cfg = FmmtConfig()
    sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    assert abs(_checksum(sd) - float(G["swin.weights_checksum"])) < 1e-3 * float(G["swin.weights_checksum"])
    frames = syn.synthetic_faces(4, 11)
    assert abs(float(frames.double().abs().sum()) - float(G["swin.input_checksum"])) < 1.0
    _close(orc.swin_features(sd, frames), "swin.feat512")
    _close(orc.swin_cls_logits(sd, frames), "swin.logits")


def test_multimodal_roberta_large_24_layers():
    cfg = FmmtConfig(text=TextConfig.roberta_large(24))
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    assert abs(_checksum(sd) - float(G["mm_rob24.weights_checksum"])) < 1e-6 * float(G["mm_rob24.weights_checksum"])
    b = syn.synthetic_batch(cfg, U=2, L=128, seed=21, n_frames=[160, 47], with_faces=False)
    probs = torch.from_numpy(G["mm_rob24.probs"])
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    t = orc.linear(orc.text_encoder(sd, b["text_ids"], b["text_mask"], "roberta"), sd["text_linear.weight"],
                   sd["text_linear.bias"])
    _close(t[:, ::16, ::64], "mm_rob24.text768_sample", 5e-4)
    out = orc.multimodal_forward(sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519,
                                 nm, b["idx_in_dia"], kind="roberta")
    _close(out, "mm_rob24.logits")


def test_multimodal_bert_2_layers_ragged():
    cfg = FmmtConfig(text=TextConfig.bert_large(2))
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    b = syn.synthetic_batch(cfg, U=3, L=64, seed=22, n_frames=[5, 160, 33], with_faces=False)
    probs = torch.from_numpy(G["mm_bert2.probs"])
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    out = orc.multimodal_forward(sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519,
                                 nm, b["idx_in_dia"], kind="bert")
    _close(out, "mm_bert2.logits")


def test_unimodal():
    cfg = FmmtConfig()
    sd = syn.unimodal_stress_state_dict(cfg.fusion, 1111)
    b = syn.synthetic_batch(cfg, U=3, L=16, seed=23, n_frames=[160, 9, 77], with_faces=False)
    _close(orc.unimodal_forward(sd, b["vision"], b["vision_mask"]), "uni.logits")


@pytest.mark.parametrize("name,seed", [("glue_mixed", 5), ("glue_none", 6)])
def test_filter_pack_matches_literal_reference_loop(name, seed):
    """Index work: bit-exact against the reference's own train.py loop at its batch size 1."""
    cfg = FmmtConfig()
    b = syn.synthetic_batch(cfg, U=1, L=16, seed=30 + seed, n_frames=[37], with_faces=False)
    probs = torch.from_numpy(G[f"{name}.probs"])
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    assert torch.equal(v519, torch.from_numpy(G[f"{name}.vision519"]))
    assert torch.equal(nm, torch.from_numpy(G[f"{name}.mask"]))
    if name == "glue_none":
        assert nm.sum() == 37 and torch.equal(nm, b["vision_mask"])
    else:
        assert 0 < nm.sum() < 37


def test_end_to_end_batch1():
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    b = syn.synthetic_batch(cfg, U=1, L=128, seed=41, n_frames=[12], with_faces=True)
    out = orc.evaluate_batch(swin_sd, sd, b, kind="roberta")
    _close(out, "e2e.logits")


def test_span_extract_edge_cases():
    H = 4
    t = torch.arange(2 * 20 * H, dtype=torch.float32).view(2, 20, H)
    sep = torch.zeros(2, 20)
    sep[0, [3, 9, 10]] = 1      # adjacent separators -> empty span for p=2 (roberta gap 2 gives n<0 -> 0)
    sep[1, [5]] = 1
    out, m = orc.span_extract(t, sep, torch.tensor([2, 3]), "roberta", max_len=4)
    assert m.sum() == 0 and out.abs().sum() == 0          # n<0 clamp; p >= #seps
    out, m = orc.span_extract(t, sep, torch.tensor([1, 0]), "bert", max_len=4)
    assert torch.equal(out[0, :4], t[0, 4:8]) and m[0].sum() == 4     # clamp to max_len
    assert torch.equal(out[1, :4], t[1, 1:5]) and m[1].sum() == 4
