"""Text front-end against golden vectors produced by the REAL reference
(`src/meld_bert_extraText.py`, generator: tests/golden/make_text_golden.py) and, where /root/reference exists, live."""
import json
import os
import random
import sys

import pytest
import torch

from facialmmt_b200 import text_frontend as tf

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "text_frontend_v1.json")))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_text_golden import FakeTokenizer  # noqa: E402  (the stand-in tokenizer the vectors were made with)


def test_truncate_matches_reference_golden():
    for c in GOLD["truncate_cases"]:
        assert tf.truncate_longest_first(c["tokens"], c["max_length"]) == c["out"]


@pytest.mark.parametrize("plm", ["roberta-large", "bert-large"])
def test_encode_dialogues_matches_reference_golden(plm):
    kind = tf.kind_of("/some/path/" + plm)
    feats = tf.encode_dialogues(GOLD["dialogues"], FakeTokenizer(), kind)
    assert len(feats) == len(GOLD[plm])
    for f, g in zip(feats, GOLD[plm]):
        n = g["n"]
        assert len(f.input_ids) == len(f.input_mask) == len(f.sep_mask) == g["padded_len"] == 512
        assert f.input_ids[:n] == g["input_ids"] and f.sep_mask[:n] == g["sep_mask"]
        assert f.input_mask == [1] * n + [0] * (512 - n)
        assert not any(f.input_ids[n:]) and not any(f.sep_mask[n:])


@pytest.mark.parametrize("kind,bos,sep", [("roberta", 0, 2), ("bert", 101, 102)])
def test_ids_path_equals_token_path_and_spans_match_oracle(kind, bos, sep):
    """encode_dialogue (ids in) == encode_dialogues (strings in); the spans implied by sep_mask are the rows the oracle's
    span extraction (src/models.py:112-150) returns."""
    from oracle.facialmmt_oracle import span_extract
    tok = FakeTokenizer()
    dias = GOLD["dialogues"][:6]
    for utts in dias:
        ids = [tok.convert_tokens_to_ids(tok.tokenize(u)) for u in utts]
        a = tf.encode_dialogue(ids, kind, bos, sep)
        b = tf.encode_dialogues([utts], tok, kind)[0]
        assert a == b
        spans = tf.utterance_spans(a.sep_mask, kind)
        assert len(spans) == len(utts)
        L = 512
        hidden = torch.arange(L, dtype=torch.float32).view(1, L, 1).repeat(1, 1, 4)     # row index as the feature
        for p, (start, n) in enumerate(spans):
            out, mask = span_extract(hidden, torch.tensor([a.sep_mask]), torch.tensor([p]), kind)
            assert int(mask.sum()) == n
            if n:
                assert out[0, :n, 0].tolist() == [float(start + t) for t in range(n)]
            # the rows are exactly the utterance's own tokens (never a separator)
            assert all(a.input_ids[start + t] not in (bos, sep) or a.input_ids[start + t] >= 1000 for t in range(n))


def test_truncation_budget_and_errors():
    long = [[7] * 600, [8] * 600, [9] * 3]
    toks, sm = tf.assemble_dialogue(long, "roberta", 0, 2)
    assert len(toks) == (512 - 68) + 2 + 2 * 2 and sum(sm) == 3          # <s> A </s> (</s> B </s>) x 2
    toks, sm = tf.assemble_dialogue(long, "bert", 101, 102)
    assert len(toks) == (512 - 34) + 1 + 3 and sum(sm) == 3              # [CLS] A [SEP] B [SEP] C [SEP]
    with pytest.raises(ValueError):
        tf.kind_of("gpt2")
    with pytest.raises(ValueError):
        tf.assemble_dialogue([[1]], "t5", 0, 2)
    with pytest.raises(ValueError):                                        # 200 one-token utterances: separators overflow 512
        tf.encode_dialogue([[5]] * 200, "roberta", 0, 2)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference checkout (build container only)")
def test_truncate_live_against_reference_random():
    sys.path.insert(0, "/root/reference/src")
    import importlib
    ref = importlib.import_module("meld_bert_extraText")
    rng = random.Random(0)
    for _ in range(200):
        toks = [[rng.randint(0, 99) for _ in range(rng.choice([0, 1, 2, 5, 5, 9, 30]))] for _ in range(rng.randint(1, 10))]
        mx = rng.choice([0, 3, 10, 25, 200])
        assert tf.truncate_longest_first(toks, mx) == ref._truncate_seq_pair([list(t) for t in toks], mx)


def test_data_batches_follow_the_reference_tuple_and_share_dialogue_rows(tmp_path):
    """facialmmt_b200/data.py: dialogue JSON -> one sample per utterance, dialogue-level ids repeated per utterance (what the
    model forward de-duplicates), idx_in_dia = position, reference tuple order (utils/dataset.py:291-292)."""
    import json

    import torch
    from facialmmt_b200 import data as fdata
    from facialmmt_b200 import text_frontend as tf
    from facialmmt_b200.config import FmmtConfig
    dialogues = [{"utterances": [[11, 12, 13], [21, 22], [31]]}, {"utterances": [[41, 42, 43, 44]]}]
    path = tmp_path / "d.json"
    path.write_text(json.dumps(dialogues))
    ids, mask, sep, idx = fdata.encode_all(fdata.load_dialogues(str(path)), "roberta")
    assert idx == [0, 1, 2, 0] and len(ids) == 4
    assert ids[0] == ids[1] == ids[2] and ids[0] != ids[3]
    assert ids[0][:11] == [0, 11, 12, 13, 2, 2, 21, 22, 2, 2, 31]                 # <s> A </s></s> B </s></s> C </s>
    assert tf.utterance_spans(sep[0], "roberta") == [(1, 3), (6, 2), (10, 1)]
    cfg = FmmtConfig()
    feats = fdata.load_features("", cfg, 4, seed=1)
    batches = list(fdata.iter_batches(ids, mask, sep, idx, feats, batch_size=3))
    assert [b[0].shape[0] for b in batches] == [3, 1]
    b0 = batches[0]
    assert b0[0].shape == (3, 512) and b0[8].dtype == torch.uint8 and b0[8].shape[1:] == (160, 112, 112, 3)
    assert b0[10].tolist() == [0, 1, 2] and len(b0[9]) == 3
