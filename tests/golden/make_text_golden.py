"""Generates tests/golden/text_frontend_v1.json by running the REAL reference text front-end
(/root/reference/src/meld_bert_extraText.py: MELD.preprocess_data and _truncate_seq_pair) on seeded synthetic dialogues with
a deterministic stand-in tokenizer (whitespace split, ids = stable hash). Only the OUTPUTS are committed.
Usage (build container only): python tests/golden/make_text_golden.py"""
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class FakeTokenizer:
    """tokenize = whitespace split (a word of the form a-b-c splits into 3 sub-tokens); ids: specials fixed, others hashed."""
    SPECIAL = {"<s>": 0, "</s>": 2, "[CLS]": 101, "[SEP]": 102}

    def tokenize(self, text):
        out = []
        for w in text.split():
            out.extend(w.split("-"))
        return out

    def convert_tokens_to_ids(self, tokens):
        return [self.SPECIAL[t] if t in self.SPECIAL else 1000 + (sum(ord(c) * (i + 7) for i, c in enumerate(t)) % 20000)
                for t in tokens]


def synthetic_dialogues(seed):
    rng = random.Random(seed)
    words = ["oh", "my", "god", "joey", "what-are-you", "doing", "no", "yes", "i", "really-really", "pivot", "fine", "we",
             "were", "on", "a", "break", "how-you-doin"]
    dias = []
    for d in range(7):
        n_utt = rng.choice([1, 2, 5, 9, 24, 33])
        long_one = d in (3, 5)
        utts = []
        for u in range(n_utt):
            n_w = rng.randint(1, 12) if not long_one else rng.randint(20, 60)
            utts.append(" ".join(rng.choice(words) for _ in range(n_w)))
        dias.append(utts)
    dias.append(["a " * 300, "b " * 300, "c " * 5])          # forces longest-first truncation with ties
    dias.append(["x y z"] * 34)                                # the 34-utterance budget
    return dias


def run_reference(dias, plm):
    import importlib
    import pandas as pd
    sys.path.insert(0, os.path.join(REF, "src"))
    mod = importlib.import_module("meld_bert_extraText")
    mod.RobertaTokenizer = type("T", (), {"from_pretrained": staticmethod(lambda p: FakeTokenizer())})
    mod.BertTokenizer = type("T", (), {"from_pretrained": staticmethod(lambda p: FakeTokenizer())})
    with tempfile.TemporaryDirectory() as td:
        rows, texts = [], {}
        for d, utts in enumerate(dias):
            for u, t in enumerate(utts):
                rows.append({"Dialogue_ID": d, "Utterance_ID": u})
                texts[f"dia{d}_utt{u}"] = {"txt": [t]}
        pd.DataFrame(rows).to_csv(os.path.join(td, "test_sent_emo.csv"), index=False)
        json.dump(texts, open(os.path.join(td, "test_text.json"), "w"))
        feats = mod.MELD(td, f"/x/{plm}", td, "test").preprocess_data()
    return [{"input_ids": f.input_ids, "input_mask": f.input_mask, "sep_mask": f.sep_mask} for f in feats], mod


def main():
    dias = synthetic_dialogues(1111)
    out = {"dialogues": dias}
    for plm in ("roberta-large", "bert-large"):
        feats, mod = run_reference(dias, plm)
        # keep only the unpadded prefix (+ the length) to stay small
        out[plm] = [{"n": sum(f["input_mask"]), "input_ids": f["input_ids"][:sum(f["input_mask"])],
                     "sep_mask": f["sep_mask"][:sum(f["input_mask"])], "padded_len": len(f["input_ids"])} for f in feats]
    rng = random.Random(5)
    cases = []
    for _ in range(40):
        toks = [[rng.randint(0, 9) for _ in range(rng.choice([0, 1, 3, 7, 7, 20, 50]))] for _ in range(rng.randint(1, 8))]
        mx = rng.choice([5, 17, 40, 1000])
        got = mod._truncate_seq_pair([list(t) for t in toks], mx)
        cases.append({"tokens": toks, "max_length": mx, "out": got})
    out["truncate_cases"] = cases
    p = os.path.join(ROOT, "tests", "golden", "text_frontend_v1.json")
    json.dump(out, open(p, "w"))
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
