"""identical dialogues of a batch are encoded once (fmmt_multimodal_forward_dedup). MELD feeds every
utterance with its whole dialogue (src/meld_bert_extraText.py:65-130), so an eval batch repeats (ids, mask) rows; the
de-duplicated forward must give the SAME bits as the plain one (rows are independent in eval)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["roberta", "bert"])
def test_dedup_is_result_identical(kind):
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.models import MultiModalTransformerForClassification
    tc = TextConfig.roberta_large(3) if kind == "roberta" else TextConfig.bert_large(3)
    cfg = FmmtConfig(text=tc)
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    m = MultiModalTransformerForClassification(cfg)
    m.load_state_dict(sd)
    U, L = 6, 128
    b = syn.synthetic_batch(cfg, U=U, L=L, seed=77, n_frames=[160, 3, 40, 99, 160, 7], with_faces=False)
    # dialogues: utterances 0,1,2 = dialogue A (targets 0,1,3); 3 = dialogue B; 4,5 = dialogue C (targets 2,0)
    for dst, src in ((1, 0), (2, 0), (5, 4)):
        for k in ("text_ids", "text_mask", "sep_mask"):
            b[k][dst] = b[k][src]
    b["idx_in_dia"] = torch.tensor([0, 1, 3, 2, 2, 0])
    v519 = torch.cat([b["vision"], torch.softmax(torch.randn(U, 160, 7, generator=torch.Generator().manual_seed(1)), -1)], -1)
    args = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519, b["vision_mask"], b["idx_in_dia"])
    m.dedup_dialogues = False
    plain = m(*args)
    m.dedup_dialogues = True
    n0 = m.flops(reset=True)
    dedup = m(*args)
    f_dedup = m.flops(reset=True)
    m.dedup_dialogues = False
    m(*args)
    f_plain = m.flops(reset=True)
    m.check()
    assert torch.equal(plain, dedup)
    assert f_dedup < f_plain                       # 3 text rows instead of 6
    print(f"\n{kind}: dedup forward {f_dedup / 1e9:.1f} GFLOP vs {f_plain / 1e9:.1f} GFLOP, logits bit-identical")
