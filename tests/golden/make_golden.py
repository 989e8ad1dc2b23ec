"""Generate tests/golden/golden_v1.npz by running the REAL reference (imported from /root/reference, CPU, fp32).

Run in the build container only:  python tests/golden/make_golden.py
Inputs and weights are regenerated from seeds by facialmmt_b200.synthetic (same torch build here and on the GPU
box), so only the reference OUTPUTS are stored (a few hundred KB), plus checksums of the regenerated weights/inputs
so that RNG drift is detected instead of silently mis-comparing.

Cases
  swin      SwinForAffwildClassification(frames, is_trg_task=False) -> logits (src/models.py:26-37) and the 512-d
            backbone features; 4 frames.
  mm_rob24  MultiModalTransformerForClassification.forward, RoBERTa-large architecture (24 layers), U=2, L=128.
  mm_bert2  same with the BERT-large architecture truncated to 2 layers, U=3, L=64, ragged frame counts.
  uni       meld_utt_transformer.forward, U=3.
  glue_*    the literal eval loop of train.py (multimodal_evaluate, lines located at run time and exec'd with stub
            models -- nothing is copied into this repository) at the reference's batch size 1: the packed
            (1,160,519) vision input and mask handed to the multimodal model, for a mixed case and for the
            "no frame passes" fallback.
  e2e       the same literal loop with the real Swin-cls and multimodal (2-layer text) models -> logits, U=1.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from facialmmt_b200 import synthetic as syn  # noqa: E402
from facialmmt_b200.config import FmmtConfig, TextConfig  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


def checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


literal_eval_loop = rh.literal_eval_loop


class _SwinStub(torch.nn.Module):
    def __init__(self, probs):
        super().__init__()
        self.probs = probs

    def forward(self, x, is_trg_task=None):
        return self.probs


class _CaptureMM(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.captured = None

    def forward(self, *a):
        self.captured = [t.clone() for t in a]
        return torch.zeros(a[5].shape[0], 7)


def main():
    rh.install_shims()
    torch.manual_seed(0)
    g = {}
    # ------------------------------------------------------------------ swin
    cfg = FmmtConfig()
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    swin = rh.build_swin_cls()
    swin.load_state_dict(swin_sd)
    frames = syn.synthetic_faces(4, 11)
    with torch.no_grad():
        g["swin.logits"] = swin(frames, is_trg_task=False).numpy()
        g["swin.feat512"] = swin.swin(frames).numpy()
    g["swin.weights_checksum"] = np.float64(checksum(swin_sd))
    g["swin.input_checksum"] = np.float64(frames.double().abs().sum())

    # ------------------------------------------------------------------ multimodal, RoBERTa-large 24 layers
    cfg24 = FmmtConfig(text=TextConfig.roberta_large(24))
    sd = syn.multimodal_stress_state_dict(cfg24, 1111)
    mm = rh.build_multimodal(rh.default_args("roberta-large"))
    mm.load_state_dict(sd, strict=False)
    b = syn.synthetic_batch(cfg24, U=2, L=128, seed=21, n_frames=[160, 47], with_faces=False)
    probs = torch.softmax(2.0 * torch.randn(207, 7, generator=torch.Generator().manual_seed(3)), -1)
    g["mm_rob24.probs"] = probs.numpy()
    from oracle import facialmmt_oracle as orc
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    with torch.no_grad():
        g["mm_rob24.logits"] = mm(b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519, nm,
                                  b["idx_in_dia"]).numpy()
        t = mm.text_linear(mm.roberta(b["text_ids"], b["text_mask"])[0])
        g["mm_rob24.text768_sample"] = t[:, ::16, ::64].numpy()
    g["mm_rob24.weights_checksum"] = np.float64(checksum(sd))
    del mm, sd

    # ------------------------------------------------------------------ multimodal, BERT-large truncated to 2 layers
    cfgb = FmmtConfig(text=TextConfig.bert_large(2))
    sdb = syn.multimodal_stress_state_dict(cfgb, 1111)
    mmb = rh.build_multimodal(rh.default_args("bert-large"), text_layers=2)
    mmb.load_state_dict(sdb, strict=False)
    bb = syn.synthetic_batch(cfgb, U=3, L=64, seed=22, n_frames=[5, 160, 33], with_faces=False)
    probs_b = torch.softmax(2.0 * torch.randn(198, 7, generator=torch.Generator().manual_seed(4)), -1)
    g["mm_bert2.probs"] = probs_b.numpy()
    v519b, nmb = orc.filter_pack(bb["vision"], bb["vision_mask"], bb["num_imgs"], probs_b, 0.2)
    with torch.no_grad():
        g["mm_bert2.logits"] = mmb(bb["text_ids"], bb["text_mask"], bb["sep_mask"], bb["audio"], bb["audio_mask"],
                                   v519b, nmb, bb["idx_in_dia"]).numpy()
    g["mm_bert2.weights_checksum"] = np.float64(checksum(sdb))

    # ------------------------------------------------------------------ unimodal V
    usd = syn.unimodal_stress_state_dict(cfg.fusion, 1111)
    uni = rh.build_unimodal()
    uni.load_state_dict(usd)
    ub = syn.synthetic_batch(cfg, U=3, L=16, seed=23, n_frames=[160, 9, 77], with_faces=False)
    with torch.no_grad():
        g["uni.logits"] = uni(ub["vision"], ub["vision_mask"]).numpy()

    # ------------------------------------------------------------------ literal eval glue (train.py), batch size 1
    make_loop = literal_eval_loop()
    args = argparse.Namespace(trg_batch_size=1, FacialEmoImpor_threshold=0.2, num_labels=7, trg_n_test=1, trg_n_valid=1)
    crit = lambda logits, y: torch.zeros(())  # noqa: E731
    for name, seed, sharp in (("glue_mixed", 5, 2.0), ("glue_none", 6, 0.0)):
        gb = syn.synthetic_batch(cfg, U=1, L=16, seed=30 + seed, n_frames=[37], with_faces=False)
        p = torch.softmax(sharp * torch.randn(37, 7, generator=torch.Generator().manual_seed(seed)), -1)
        faces = torch.zeros(1, 160, 3, 2, 2)   # only sliced, never read, by the stub
        batch = (gb["text_ids"], gb["text_mask"], gb["sep_mask"], gb["audio"], gb["audio_mask"], gb["vision"],
                 gb["vision_mask"], torch.zeros(1, dtype=torch.long), faces, gb["num_imgs"], gb["idx_in_dia"])
        cap = _CaptureMM()
        with torch.no_grad():
            make_loop(args, [batch])(_SwinStub(p), cap, crit, test=True)
        g[f"{name}.probs"] = p.numpy()
        g[f"{name}.vision519"] = cap.captured[5].numpy()
        g[f"{name}.mask"] = cap.captured[6].numpy()

    # ------------------------------------------------------------------ end to end, literal loop, real models, U=1
    cfg2 = FmmtConfig(text=TextConfig.roberta_large(2))
    sd2 = syn.multimodal_stress_state_dict(cfg2, 1111)
    mm2 = rh.build_multimodal(rh.default_args("roberta-large"), text_layers=2)
    mm2.load_state_dict(sd2, strict=False)
    eb = syn.synthetic_batch(cfg2, U=1, L=128, seed=41, n_frames=[12], with_faces=True)

    class _SwinWithNoise(torch.nn.Module):
        def forward(self, x, is_trg_task=None):  # explicit-noise form of F.gumbel_softmax (soft)
            return torch.softmax((swin(x, is_trg_task=False) + eb["gumbel"]) / 1.0, -1)

    batch = (eb["text_ids"], eb["text_mask"], eb["sep_mask"], eb["audio"], eb["audio_mask"], eb["vision"],
             eb["vision_mask"], torch.zeros(1, dtype=torch.long), eb["faces"], eb["num_imgs"], eb["idx_in_dia"])
    with torch.no_grad():
        _, results, _ = make_loop(args, [batch])(_SwinWithNoise(), mm2, crit, test=True)
        # cross-check the explicit-noise form against F.gumbel_softmax itself under a fixed RNG state
        torch.manual_seed(99)
        a = torch.nn.functional.gumbel_softmax(torch.zeros(5, 7), 1.0)
        torch.manual_seed(99)
        gn = -torch.empty(5, 7).exponential_().log()
        assert torch.allclose(a, torch.softmax(gn, -1), atol=1e-6)
    g["e2e.logits"] = results.numpy()

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
