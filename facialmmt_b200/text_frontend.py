"""Dialogue text front-end of the T+A+V path: the step that produces `batch_text_input_ids / _mask / _sep_mask` for
`MultiModalTransformerForClassification.forward`.

Restates `src/meld_bert_extraText.py:22-45` (`_truncate_seq_pair`: longest-first truncation, one token at a time, ties
broken by the FIRST longest utterance) and `:65-130` (`MELD.preprocess_data`: `<s> A </s></s> B </s> ...` for RoBERTa with
the 1 of `sep_mask` on the closing `</s>` of every utterance, `[CLS] A [SEP] B [SEP] ...` for BERT with the 1 on every
`[SEP]`, zero padding of ids / mask / sep_mask to 512). The reference pops one token per pass over all utterances
(O(total^2)); here a heap gives the same result in O(pops * log n). Host-side integer work: no GPU involved.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass
from typing import List, Sequence

MAX_SEQ_LENGTH = 512          # meld_bert_extraText.py:9
SPECIAL_BUDGET = {"roberta": 34 * 2, "bert": 34}   # :92-94 (room for the separators of up to 34 utterances)


def truncate_longest_first(utterances: Sequence[Sequence], max_length: int) -> List[list]:
    """`_truncate_seq_pair` (meld_bert_extraText.py:22-45): while the total length exceeds `max_length`, drop the last token
    of the longest utterance; among equally long utterances the first one loses the token (stable `sorted(..., reverse=True)`)."""
    out = [list(u) for u in utterances]
    total = sum(len(u) for u in out)
    if total <= max_length:
        return out
    heap = [(-len(u), i) for i, u in enumerate(out)]   # longest first, lowest index first among ties
    heapq.heapify(heap)
    while total > max_length:
        neg, i = heapq.heappop(heap)
        out[i].pop()
        total -= 1
        heapq.heappush(heap, (neg + 1, i))
    return out


@dataclass
class DialogueFeatures:
    """`InputFeatures` of meld_bert_extraText.py:47-53."""
    input_ids: List[int]
    input_mask: List[int]
    sep_mask: List[int]


def kind_of(pretrained_path_or_name: str) -> str:
    """The reference branches on the last path component (`:67,70`; `src/models.py:49`)."""
    name = pretrained_path_or_name.rstrip("/").split("/")[-1]
    if name == "roberta-large":
        return "roberta"
    if name == "bert-large":
        return "bert"
    raise ValueError(f"unsupported text model: {name} (the reference knows roberta-large and bert-large)")


def assemble_dialogue(utterance_tokens: Sequence[Sequence], kind: str, bos, sep, max_seq_length: int = MAX_SEQ_LENGTH):
    """Tokens (or ids) and sep_mask of one dialogue before padding (meld_bert_extraText.py:91-112). `bos` / `sep` are the
    special tokens in the caller's vocabulary: ("<s>", "</s>") or ("[CLS]", "[SEP]"), or their ids."""
    if kind not in SPECIAL_BUDGET:
        raise ValueError(f"kind must be 'roberta' or 'bert', got {kind!r}")
    utts = truncate_longest_first(utterance_tokens, max_seq_length - SPECIAL_BUDGET[kind])
    tokens: list = []
    sep_mask: List[int] = []
    for num, utt in enumerate(utts):
        if num == 0:
            tokens = [bos] + utt + [sep]
            sep_mask = [0] * (len(tokens) - 1) + [1]
        elif kind == "roberta":            # <s> A </s></s> B </s>
            tokens += [sep] + utt + [sep]
            sep_mask += [0] * (len(utt) + 1) + [1]
        else:                              # [CLS] A [SEP] B [SEP]
            tokens += utt + [sep]
            sep_mask += [0] * len(utt) + [1]
    return tokens, sep_mask


def encode_dialogue(utterance_ids: Sequence[Sequence[int]], kind: str, bos_id: int, sep_id: int,
                    max_seq_length: int = MAX_SEQ_LENGTH) -> DialogueFeatures:
    """ids / attention mask / sep mask of one dialogue, zero-padded to `max_seq_length` (meld_bert_extraText.py:114-130).
    Note the reference pads the ids with 0, which is `<s>` for RoBERTa; the attention mask hides those positions."""
    ids, sep_mask = assemble_dialogue(utterance_ids, kind, bos_id, sep_id, max_seq_length)
    if len(ids) > max_seq_length:
        raise ValueError(f"dialogue of {len(ids)} tokens after truncation exceeds {max_seq_length} "
                         f"(more than {34} utterances? the reference reserves room for 34)")
    pad = [0] * (max_seq_length - len(ids))
    return DialogueFeatures(input_ids=list(ids) + pad, input_mask=[1] * len(ids) + pad, sep_mask=sep_mask + pad)


def encode_dialogues(dialogues: Sequence[Sequence[str]], tokenizer, kind: str,
                     max_seq_length: int = MAX_SEQ_LENGTH) -> List[DialogueFeatures]:
    """`MELD.preprocess_data` for already-loaded dialogues (lists of utterance strings). `tokenizer` needs `tokenize` and
    `convert_tokens_to_ids` (HF Roberta/BertTokenizer have both)."""
    bos, sep = ("<s>", "</s>") if kind == "roberta" else ("[CLS]", "[SEP]")
    feats = []
    for utts in dialogues:
        toks = [tokenizer.tokenize(u) for u in utts]
        tokens, sep_mask = assemble_dialogue(toks, kind, bos, sep, max_seq_length)
        ids = tokenizer.convert_tokens_to_ids(tokens)
        pad = [0] * (max_seq_length - len(ids))
        feats.append(DialogueFeatures(input_ids=list(ids) + pad, input_mask=[1] * len(ids) + pad, sep_mask=sep_mask + pad))
    return feats


def utterance_spans(sep_mask: Sequence[int], kind: str, max_len: int = 38):
    """(start, length) of every utterance inside the dialogue as `src/models.py:112-150` slices them (the device kernel
    `span_extract_kernel` does the same from `sep_mask` and the utterance index): host-side cross-check of the two ends."""
    seps = [i for i, v in enumerate(sep_mask) if v == 1]
    gap = 2 if kind == "roberta" else 1
    spans = []
    for p, s in enumerate(seps):
        if p == 0:
            start, n = 1, s - 1
        else:
            start, n = seps[p - 1] + gap, s - seps[p - 1] - gap
        spans.append((start, max(0, min(n, max_len))))
    return spans
