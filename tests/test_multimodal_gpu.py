"""Fusion path (text encoder -> span -> audio/vision encoders -> cross-modal -> pooling -> logits) and the eval glue
through the C ABI against the CPU oracle. bf16 operands / fp32 accumulate; north_star tolerance for this mode: 1e-2
ABSOLUTE on the logits, argmax-exact (the fp32-grade mode is tested at 1e-3 in tests/test_precise_gpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL_STAGE = 4e-2


def logit_tol(ref):
    return 1e-2          # absolute (north_star: "1e-2 bf16 on logits")


def _mm(kind, layers):
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.models import MultiModalTransformerForClassification
    tc = TextConfig.roberta_large(layers) if kind == "roberta" else TextConfig.bert_large(layers)
    cfg = FmmtConfig(text=tc)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    m = MultiModalTransformerForClassification(cfg)
    m.load_state_dict(sd)
    return cfg, sd, m


@pytest.mark.parametrize("kind,layers,U,L", [("roberta", 2, 3, 128), ("bert", 2, 2, 64), ("roberta", 24, 2, 128)])
def test_multimodal_stagewise_and_logits(kind, layers, U, L):
    from facialmmt_b200 import synthetic as syn
    from oracle import facialmmt_oracle as orc
    cfg, sd, m = _mm(kind, layers)
    nf = [160, 47, 5][:U]
    b = syn.synthetic_batch(cfg, U=U, L=L, seed=21, n_frames=nf, with_faces=False)
    probs = torch.softmax(2.0 * torch.randn(sum(nf), 7, generator=torch.Generator().manual_seed(3)), -1)
    v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], b["num_imgs"], probs, 0.2)
    col = {}
    ref = orc.multimodal_forward(sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519,
                                 nm, b["idx_in_dia"], kind=kind, collect=col)
    caps = {n: m.capture("mm." + n, col[n].numel()) for n in ("text", "audio", "vision", "ta", "fused")}
    got = m(b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519, nm, b["idx_in_dia"])
    torch.cuda.synchronize()
    m.clear_captures()
    rep = []
    for n, t in caps.items():
        r = col[n].reshape(-1)
        g = t.cpu()
        assert torch.isfinite(g).all(), n
        rep.append((n, (g - r).abs().max().item() / r.abs().max().item()))
    err = (got.cpu() - ref).abs().max().item()
    print(f"\n{kind}-{layers}L U={U} L={L}: " + ", ".join(f"{n}={r:.2e}" for n, r in rep) +
          f"; logits err {err:.3e} (scale {ref.abs().max():.2f})")
    for n, r in rep:
        assert r < REL_TOL_STAGE, (n, r)
    assert err < logit_tol(ref), err
    assert torch.equal(got.cpu().argmax(-1), ref.argmax(-1))


def test_unimodal():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.models import meld_utt_transformer
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig()
    sd = syn.unimodal_stress_state_dict(cfg.fusion, 1111)
    m = meld_utt_transformer(cfg)
    m.load_state_dict(sd)
    b = syn.synthetic_batch(cfg, U=3, L=16, seed=23, n_frames=[160, 9, 77], with_faces=False)
    ref = orc.unimodal_forward(sd, b["vision"], b["vision_mask"])
    got = m(b["vision"], b["vision_mask"]).cpu()
    err = (got - ref).abs().max().item()
    print(f"\nunimodal logits err {err:.3e} (scale {ref.abs().max():.2f})")
    assert err < logit_tol(ref)
    assert torch.equal(got.argmax(-1), ref.argmax(-1))


@pytest.mark.parametrize("sharp,per_utt", [(2.0, True), (2.0, False), (0.0, True), (0.0, False), (0.7, True)])
def test_filter_pack_bit_exact(sharp, per_utt):
    """Index/byte work: bit-exact against the oracle (identical fp32 probabilities in, no arithmetic on the payload)."""
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.models import filter_pack
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig()
    nf = [37, 160, 1, 5]
    b = syn.synthetic_batch(cfg, U=4, L=16, seed=9, n_frames=nf, with_faces=False)
    probs = torch.softmax(sharp * torch.randn(sum(nf), 7, generator=torch.Generator().manual_seed(5)), -1)
    if sharp == 0.7:
        probs[37:197] = 1.0 / 7          # utterance 1: nothing passes while others do
    imp = (probs * probs).sum(-1)
    assert ((imp - 0.2).abs() > 1e-5).all()
    if per_utt:
        vs, ms, off = [], [], 0
        for u in range(4):
            v, mk = orc.filter_pack(b["vision"][u:u + 1], b["vision_mask"][u:u + 1], nf[u:u + 1], probs[off:off + nf[u]], 0.2)
            vs.append(v); ms.append(mk); off += nf[u]
        rv, rm = torch.cat(vs), torch.cat(ms)
    else:
        rv, rm = orc.filter_pack(b["vision"], b["vision_mask"], nf, probs, 0.2)
    gv, gm = filter_pack(b["vision"], b["vision_mask"], nf, probs, 0.2, per_utterance=per_utt)
    assert torch.equal(gv.cpu(), rv)
    assert torch.equal(gm.cpu(), rm)


def test_end_to_end_eval_batch():
    """Swin -> filter -> fusion with injected Gumbel noise vs the oracle, U=2 (per-utterance semantics)."""
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.evaluate import evaluate_batch
    from facialmmt_b200.models import MultiModalTransformerForClassification, SwinForAffwildClassification
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    swin = SwinForAffwildClassification(cfg)
    swin.load_state_dict(swin_sd)
    mm = MultiModalTransformerForClassification(cfg)
    mm.load_state_dict(sd)
    b = syn.synthetic_batch(cfg, U=2, L=128, seed=41, n_frames=[12, 7], with_faces=True)
    # the 0.2 threshold is a discontinuity: draw the injected Gumbel noise so that no frame sits on it
    frames = torch.cat([b["faces"][0, :12], b["faces"][1, :7]])
    z = orc.swin_cls_logits(swin_sd, frames)
    for s in range(200):
        g = -torch.empty(19, 7).exponential_(generator=torch.Generator().manual_seed(1000 + s)).log()
        imp = (orc.gumbel_softmax_probs(z, g, 1.0) ** 2).sum(-1)
        if (imp - 0.2).abs().min() > 0.02 and 0 < int((imp > 0.2).sum()) < 19:
            b["gumbel"] = g
            break
    else:
        pytest.fail("no noise draw keeps all frames clear of the threshold")
    col = {}
    ref = orc.evaluate_batch(swin_sd, sd, b, kind="roberta", collect=col)
    imp = (col["probs"] ** 2).sum(-1)
    margin = (imp - 0.2).abs().min().item()
    batch = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], b["vision"], b["vision_mask"],
             torch.zeros(2, dtype=torch.long), b["faces"], b["num_imgs"], b["idx_in_dia"])
    got, inter = evaluate_batch(swin, mm, batch, 0.2, gumbel=b["gumbel"].cuda(), return_intermediates=True)
    perr = (inter["probs"].cpu() - col["probs"]).abs().max().item()
    print(f"\ne2e: probs err {perr:.2e}, min |imp-0.2| margin {margin:.2e}")
    assert perr < 1e-2
    assert margin > 2 * perr
    assert torch.equal(inter["new_mask"].cpu(), col["new_mask"])       # identical keep decisions
    err = (got.cpu() - ref).abs().max().item()
    print(f"e2e logits err {err:.3e} (scale {ref.abs().max():.2f})")
    assert err < logit_tol(ref)
    assert torch.equal(got.cpu().argmax(-1), ref.argmax(-1))
