"""fp32-grade mode (fmmt_config.precision = FMMT_PRECISION_FP32, `precision="fp32"` / `main.py --precision fp32`): the
north-star fp32 bar -- logits within 1e-3 ABSOLUTE of the fp32 reference arithmetic (oracle), argmax-exact. Every Linear
runs on the same tcgen05 kernels with split-bf16 x3 operands (16 mantissa bits per operand, fp32 accumulate); attention
cores, LayerNorm, GELU, softmax, residual streams are fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: "within 1e-3 fp32 ... on logits"


def test_swin_fp32_mode_stagewise_and_logits():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.models import SwinForAffwildClassification
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig()
    sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    m = SwinForAffwildClassification(cfg, swin_chunk=2, swin_chunk_late=3, precision="fp32")
    m.load_state_dict(sd)
    F = 5
    frames = syn.synthetic_faces(F, 77)
    col = {}
    ref_logits = orc.swin_cls_logits(sd, frames, collect=col)
    ref_feat = orc.swin_features(sd, frames)
    names = ["patch_embed"] + [f"layer{li}.block{bi}" for li, d in enumerate(cfg.swin.depths) for bi in range(d)]
    caps = {n: m.capture("swin." + n, col[n].numel()) for n in names}
    g = -torch.empty(F, 7).exponential_(generator=torch.Generator().manual_seed(1)).log()
    logits, probs, imp, feat = m.forward_full(frames.cuda(), g, want_feat=True)
    m.check()
    m.clear_captures()
    rep = [(n, (caps[n].cpu() - col[n].reshape(-1)).abs().max().item() / col[n].abs().max().item()) for n in names]
    print("\nfp32 mode, stage-wise max-abs error / max-abs value:", ", ".join(f"{n}={r:.1e}" for n, r in rep))
    for n, r in rep:
        assert r < 2e-4, (n, r)
    ferr = (feat.cpu() - ref_feat).abs().max().item()
    lerr = (logits.cpu() - ref_logits).abs().max().item()
    print(f"fp32 mode: feat512 err {ferr:.2e} (scale {ref_feat.abs().max():.2f}), aux logits err {lerr:.2e}")
    assert ferr < TOL and lerr < TOL
    assert torch.equal(logits.cpu().argmax(-1), ref_logits.argmax(-1))
    ref_probs = orc.gumbel_softmax_probs(ref_logits, g, 1.0)
    assert (probs.cpu() - ref_probs).abs().max().item() < TOL


@pytest.mark.parametrize("kind,layers,U,L", [("roberta", 2, 3, 128), ("bert", 2, 2, 64), ("roberta", 24, 2, 128),
                                             ("bert", 24, 2, 128)])
def test_multimodal_fp32_mode(kind, layers, U, L):
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.models import MultiModalTransformerForClassification
    from oracle import facialmmt_oracle as orc
    tc = TextConfig.roberta_large(layers) if kind == "roberta" else TextConfig.bert_large(layers)
    cfg = FmmtConfig(text=tc)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    m = MultiModalTransformerForClassification(cfg, precision="fp32")
    m.load_state_dict(sd)
    nf = [160, 47, 5][:U]
    b = syn.synthetic_batch(cfg, U=U, L=L, seed=21, n_frames=nf, with_faces=False)
    probs = torch.softmax(2.0 * torch.randn(sum(nf), 7, generator=torch.Generator().manual_seed(3)), -1)
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    col = {}
    ref = orc.multimodal_forward(sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519, nm,
                                 b["idx_in_dia"], kind=kind, collect=col)
    caps = {n: m.capture("mm." + n, col[n].numel()) for n in ("text", "audio", "vision", "ta", "fused")}
    got = m(b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519, nm, b["idx_in_dia"])
    m.check()
    m.clear_captures()
    rep = [(n, (t.cpu() - col[n].reshape(-1)).abs().max().item() / col[n].abs().max().item()) for n, t in caps.items()]
    err = (got.cpu() - ref).abs().max().item()
    print(f"\nfp32 mode {kind}-{layers}L U={U} L={L}: " + ", ".join(f"{n}={r:.1e}" for n, r in rep) +
          f"; logits err {err:.2e} (scale {ref.abs().max():.2f})")
    for n, r in rep:
        assert r < 5e-4, (n, r)
    assert err < TOL, err
    assert torch.equal(got.cpu().argmax(-1), ref.argmax(-1))


def test_unimodal_and_end_to_end_fp32_mode():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.evaluate import evaluate_batch
    from facialmmt_b200.models import (MultiModalTransformerForClassification, SwinForAffwildClassification,
                                       meld_utt_transformer)
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    usd = syn.unimodal_stress_state_dict(cfg.fusion, 1111)
    um = meld_utt_transformer(cfg, precision="fp32")
    um.load_state_dict(usd)
    ub = syn.synthetic_batch(cfg, U=3, L=16, seed=23, n_frames=[160, 9, 77], with_faces=False)
    uref = orc.unimodal_forward(usd, ub["vision"], ub["vision_mask"])
    ugot = um(ub["vision"], ub["vision_mask"]).cpu()
    um.check()
    assert (ugot - uref).abs().max().item() < TOL
    # end to end: Swin -> filter -> fusion, injected Gumbel noise, U=2
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    swin = SwinForAffwildClassification(cfg, precision="fp32")
    swin.load_state_dict(swin_sd)
    mm = MultiModalTransformerForClassification(cfg, precision="fp32")
    mm.load_state_dict(sd)
    b = syn.synthetic_batch(cfg, U=2, L=128, seed=41, n_frames=[12, 7], with_faces=True)
    col = {}
    ref = orc.evaluate_batch(swin_sd, sd, b, kind="roberta", collect=col)
    margin = ((col["probs"] ** 2).sum(-1) - 0.2).abs().min().item()
    batch = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], b["vision"], b["vision_mask"],
             torch.zeros(2, dtype=torch.long), b["faces"], b["num_imgs"], b["idx_in_dia"])
    got, inter = evaluate_batch(swin, mm, batch, 0.2, gumbel=b["gumbel"].cuda(), return_intermediates=True)
    swin.check(); mm.check()
    perr = (inter["probs"].cpu() - col["probs"]).abs().max().item()
    print(f"\nfp32 mode e2e: probs err {perr:.2e}, threshold margin {margin:.2e}")
    assert perr < TOL
    if margin > 2 * perr:
        assert torch.equal(inter["new_mask"].cpu(), col["new_mask"])
        err = (got.cpu() - ref).abs().max().item()
        print(f"fp32 mode e2e logits err {err:.2e}")
        assert err < TOL
        assert torch.equal(got.cpu().argmax(-1), ref.argmax(-1))
