"""Eval entry mirroring train.py `multimodal_evaluate` / `unimodal_evaluate` (train.py:154-243, 275-292) and
utils/eval_metrics.py `eval_meld`, with the Python per-frame loops replaced by device kernels.

A batch is the tuple the reference DataLoader yields (utils/dataset.py:291-292):
  (text_ids, text_mask, sep_mask, audio, audio_mask, vision, vision_mask, label_ids, faces, num_imgs, idx_in_dia)
with faces (U, Lv, 3, 224, 224) fp32 and num_imgs the valid-frame count per utterance.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from .models import filter_pack


def gather_valid_frames(faces: torch.Tensor, num_imgs: Sequence[int]) -> torch.Tensor:
    """train.py:169-179: concatenate the first num_imgs[u] frames of every utterance (no copy when all are full)."""
    U, Lv = faces.shape[0], faces.shape[1]
    n = [int(x) for x in num_imgs]
    if all(x == Lv for x in n):
        return faces.reshape(U * Lv, *faces.shape[2:])
    return torch.cat([faces[u, :n[u]] for u in range(U)], dim=0)


def literal_batch_repack(num_imgs: Sequence[int], kept: Sequence[int], Lv: int):
    """Index plan of the reference's OWN re-pack loops for a batch of U > 1 utterances, bug included (train.py:188-215;
    SURVEY F7: `margin += n - 1` instead of `n`). `kept` = ascending global indices of the frames with importance > threshold.
    Returns (count[u], prob_row[u][k], vision_row[u][k]): slot k of utterance u takes the probability row `prob_row` (global
    frame index) and the vision feature row `vision_row` of utterance u (as the literal code indexes it: a negative row wraps
    like Python indexing does). Raises IndexError exactly where the reference's tensor indexing would. Pure host logic."""
    n = [int(x) for x in num_imgs]
    temp = list(int(k) for k in kept)
    counts, margin = [], 0
    for u in range(len(n)):                       # train.py:192-201
        real = 0
        for g in temp:
            if g < n[u] + margin:
                if real >= Lv:
                    raise IndexError(f"index {real} is out of bounds for dimension 0 with size {Lv}")
                real += 1
            else:
                break
        margin += n[u] - 1
        temp = temp[real:]
        counts.append(real)
    prob_row, vision_row, jj, margin = [], [], 0, 0
    for u in range(len(n)):                       # train.py:203-213
        pr, vr = [], []
        for _ in range(counts[u]):
            g = int(kept[jj])
            local = g - margin
            if local >= Lv or local < -Lv:
                raise IndexError(f"index {local} is out of bounds for dimension 0 with size {Lv}")
            pr.append(g)
            vr.append(local % Lv)
            jj += 1
        margin += n[u] - 1
        prob_row.append(pr)
        vision_row.append(vr)
    return counts, prob_row, vision_row


def _filter_pack_bug_compat(vision, vision_mask, n, probs, threshold):
    """(U,Lv,D+labels), (U,Lv) exactly as the reference's literal batch code produces them for U > 1 (bug_compat)."""
    U, Lv, D = vision.shape
    labels = probs.shape[1]
    imp = torch.diagonal(torch.mm(probs, probs.t()))          # train.py:183-184 (the N x N product, literally)
    kept = torch.nonzero(imp.gt(threshold)).squeeze(1).tolist()           # host sync: this mode is for compatibility runs
    dev = probs.device
    v = vision.to(dev, torch.float32)
    emo = torch.zeros(U, Lv, labels, device=dev)
    if len(kept) > 0:
        counts, prob_row, vision_row = literal_batch_repack(n, kept, Lv)
        new_v = torch.zeros_like(v)
        new_m = torch.zeros(U, Lv, device=dev)
        for u in range(U):
            c = counts[u]
            if c:
                new_v[u, :c] = v[u, torch.tensor(vision_row[u], device=dev)]
                emo[u, :c] = probs[torch.tensor(prob_row[u], device=dev)]
                new_m[u, :c] = 1
        return torch.cat([new_v, emo], -1), new_m
    jj = 0                                                    # train.py:223-232 fallback: nothing passes in the whole batch
    m = vision_mask.to(dev, torch.float32)
    mh = vision_mask.cpu()
    for u in range(U):
        for j in range(Lv):
            if mh[u, j] == 1:
                emo[u, j] = probs[jj]
                jj += 1
            else:
                break
    return torch.cat([v, emo], -1), m


def evaluate_batch(swin_model, multimodal_model, batch, threshold: float = 0.2, gumbel: Optional[torch.Tensor] = None,
                   per_utterance: bool = True, return_intermediates: bool = False, bug_compat: bool = False):
    """One iteration of multimodal_evaluate (train.py:164-234) -> (U, labels) logits on the GPU. Nothing here
    synchronises the device: Swin -> filter/pack -> fusion are stream-ordered kernels (call `.check()` on the two
    modules, or use multimodal_evaluate, before trusting the logits).

    Batch semantics at U > 1: `per_utterance=True` (default) gives every utterance the result the reference computes for
    it at its default `trg_batch_size=1` (main.py:56) -- the only mode its published W-F1 was produced in. The reference's
    literal batch code differs from that for U > 1 in two ways: the "no frame passes" fallback is decided for the whole
    batch (train.py:187,223), and its re-pack loop has the `margin += n-1` off-by-one (train.py:200,213).
    `per_utterance=False` reproduces the first (whole-batch fallback decision) but NOT the off-by-one.
    `bug_compat=True` (main.py --bug_compat 1) reproduces the literal batch code INCLUDING the off-by-one, through a host-side
    index plan (`literal_batch_repack`; one device->host sync per batch): only for bit-compatibility runs against the
    reference at trg_batch_size > 1."""
    (ids, mask, sep, audio, audio_mask, vision, vision_mask, _labels, faces, num_imgs, idx) = batch
    n = [int(x) for x in (num_imgs.tolist() if torch.is_tensor(num_imgs) else num_imgs)]
    frames = gather_valid_frames(faces.to("cuda", non_blocking=True), n)
    F = frames.shape[0]
    if gumbel is None:      # the draw F.gumbel_softmax makes internally (src/models.py:31-32)
        gumbel = -torch.empty(F, swin_model.num_labels, device="cuda").exponential_().log()
    _, probs, _ = swin_model.forward_full(frames, gumbel)
    cache = swin_model._out_cache if getattr(swin_model, "_graph", False) else None     # graph mode: stable buffers
    if bug_compat and len(n) > 1:
        v519, new_mask = _filter_pack_bug_compat(vision, vision_mask, n, probs, threshold)
    else:
        v519, new_mask = filter_pack(vision, vision_mask, n, probs, threshold, per_utterance, cache=cache)
    logits = multimodal_model(ids, mask, sep, audio, audio_mask, v519, new_mask, idx)
    if return_intermediates:
        return logits, dict(probs=probs, vision519=v519, new_mask=new_mask)
    return logits


def multimodal_evaluate(swin_model, multimodal_model, loader: Iterable, criterion=None, threshold: float = 0.2,
                        per_utterance: bool = True, bug_compat: bool = False):
    """-> (avg_loss, results (n,labels), truths (n,)) like train.py:154-243."""
    results, truths, total_loss, count = [], [], 0.0, 0
    for batch in loader:
        logits = evaluate_batch(swin_model, multimodal_model, batch, threshold, per_utterance=per_utterance,
                                bug_compat=bug_compat)
        labels = batch[7]
        if criterion is not None:
            total_loss += float(criterion(logits, labels.to(logits.device))) * logits.shape[0]
        results.append(logits)
        truths.append(labels)
        count += logits.shape[0]
    for m in (swin_model, multimodal_model):       # surface a kernel pipeline-watchdog event before the logits are used
        if hasattr(m, "check"):
            m.check()
    return total_loss / max(count, 1), torch.cat(results), torch.cat(truths)


def unimodal_evaluate(unimodal_model, loader: Iterable, criterion=None):
    """train.py:275-292; batches are (modality_feature, utterance_mask, labels)."""
    results, truths, total_loss, count = [], [], 0.0, 0
    for feats, utt_mask, labels in loader:
        logits = unimodal_model(feats, utt_mask)
        if criterion is not None:
            total_loss += float(criterion(logits, labels.to(logits.device))) * logits.shape[0]
        results.append(logits)
        truths.append(labels)
        count += logits.shape[0]
    if hasattr(unimodal_model, "check"):
        unimodal_model.check()
    return total_loss / max(count, 1), torch.cat(results), torch.cat(truths)


def eval_meld(results: torch.Tensor, truths: torch.Tensor, test: bool = False) -> float:
    """utils/eval_metrics.py:16-28: row argmax -> sklearn weighted F1 (per-class F1 printed when test=True)."""
    from sklearn.metrics import f1_score
    pred = np.argmax(results.detach().float().cpu().numpy(), axis=1)
    true = truths.detach().cpu().numpy()
    f1 = f1_score(true, pred, average="weighted")
    if test:
        print("**TEST** | f1 on each class (Neutral, Surprise, Fear, Sadness, Joy, Disgust, Anger): \n",
              f1_score(true, pred, average=None))
    return float(f1)
