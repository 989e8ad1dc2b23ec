"""Host-side batch assembly for `main.py --doEval` on real inputs: the step that the
reference's dataset code performs before `multimodal_evaluate` sees a batch (utils/dataset.py:254-292 + src/meld_bert_extraText.py).

Two inputs, both optional (synthetic stand-ins fill what is absent, so the path runs offline):
  * dialogues: a JSON file `[{"utterances": ["text" | [token ids], ...]}, ...]` -> one eval sample per utterance, text encoded
    at DIALOGUE level with `text_frontend.encode_dialogue(s)` (`<s> A </s></s> B </s>` / `[CLS] A [SEP] B [SEP]`, longest-first
    truncation, sep_mask, zero padding), `batchUtt_in_dia_idx` = the utterance's position in its dialogue;
  * features: a `torch.save`d dict with per-utterance tensors `audio (N,La,Da)`, `audio_mask (N,La)`, `vision (N,160,512)`,
    `vision_mask (N,160)`, `faces` ((N,160,3,224,224) fp32 or (N,160,h,w,3) uint8 decoded crops), `num_imgs (N)`, `labels (N)`
    in the order of the flattened utterances (what utils/dataset.py:291-292 yields, stacked).
Consecutive utterances of a batch share their dialogue rows, which the model forward de-duplicates (models.py).
"""
from __future__ import annotations

import json
from typing import Dict, Iterator, List, Optional, Sequence

import torch

from . import text_frontend as tf


def load_dialogues(path: str) -> List[List]:
    with open(path) as f:
        raw = json.load(f)
    return [d["utterances"] if isinstance(d, dict) else d for d in raw]


def encode_all(dialogues: Sequence[Sequence], kind: str, tokenizer=None, max_seq_length: int = tf.MAX_SEQ_LENGTH):
    """-> per-utterance lists (ids, mask, sep_mask, idx_in_dia), dialogue-level features repeated per utterance like the
    reference's MELD dataset does (one InputFeatures per dialogue, indexed per utterance: utils/dataset.py:254-292)."""
    bos_id, sep_id = (0, 2) if kind == "roberta" else (101, 102)     # <s>, </s> / [CLS], [SEP] of the HF vocabularies
    ids, mask, sep, idx = [], [], [], []
    for utts in dialogues:
        if len(utts) and isinstance(utts[0], str):
            if tokenizer is None:
                raise ValueError("string utterances need a tokenizer (--tokenizer_path); token-id lists do not")
            feat = tf.encode_dialogues([utts], tokenizer, kind, max_seq_length)[0]
        else:
            feat = tf.encode_dialogue(utts, kind, bos_id, sep_id, max_seq_length)
        for p in range(len(utts)):
            ids.append(feat.input_ids); mask.append(feat.input_mask); sep.append(feat.sep_mask); idx.append(p)
    return ids, mask, sep, idx


def iter_batches(ids, mask, sep, idx, feats: Dict[str, torch.Tensor], batch_size: int) -> Iterator[tuple]:
    """Batches in the reference's tuple layout (utils/dataset.py:291-292), sequential order (the test loader is
    SequentialSampler: main.py:128-133)."""
    n = len(ids)
    t = lambda x: torch.tensor(x, dtype=torch.long)   # noqa: E731
    for lo in range(0, n, batch_size):
        hi = min(n, lo + batch_size)
        yield (t(ids[lo:hi]), t(mask[lo:hi]), t(sep[lo:hi]), feats["audio"][lo:hi], feats["audio_mask"][lo:hi],
               feats["vision"][lo:hi], feats["vision_mask"][lo:hi], feats["labels"][lo:hi], feats["faces"][lo:hi],
               [int(v) for v in feats["num_imgs"][lo:hi]], t(idx[lo:hi]))


def synthetic_features(cfg, n: int, seed: int, faces_u8: bool = True) -> Dict[str, torch.Tensor]:
    """MELD-shaped stand-ins for the per-utterance audio / vision / face inputs (no data is available offline)."""
    from . import synthetic as syn
    b = syn.synthetic_batch(cfg, U=n, L=8, seed=seed, with_faces=False)
    g = torch.Generator().manual_seed(seed)
    if faces_u8:
        faces = torch.randint(0, 256, (n, cfg.fusion.vision_len, 112, 112, 3), dtype=torch.uint8, generator=g)
    else:
        faces = torch.rand(n, cfg.fusion.vision_len, 3, 224, 224, generator=g) * 2 - 1
    return dict(audio=b["audio"], audio_mask=b["audio_mask"], vision=b["vision"], vision_mask=b["vision_mask"],
                faces=faces, num_imgs=torch.tensor([int(v) for v in b["num_imgs"]]),
                labels=torch.randint(0, cfg.fusion.num_labels, (n,), generator=g))


def load_features(path: Optional[str], cfg, n: int, seed: int) -> Dict[str, torch.Tensor]:
    if not path:
        return synthetic_features(cfg, n, seed)
    d = torch.load(path, map_location="cpu", weights_only=True)
    need = ("audio", "audio_mask", "vision", "vision_mask", "faces", "num_imgs", "labels")
    missing = [k for k in need if k not in d]
    if missing:
        raise KeyError(f"{path}: missing {missing}")
    if any(d[k].shape[0] != n for k in need):
        raise ValueError(f"{path}: every tensor needs one row per utterance ({n})")
    return d
