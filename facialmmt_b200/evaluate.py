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


def evaluate_batch(swin_model, multimodal_model, batch, threshold: float = 0.2, gumbel: Optional[torch.Tensor] = None,
                   per_utterance: bool = True, return_intermediates: bool = False):
    """One iteration of multimodal_evaluate (train.py:164-234) -> (U, labels) logits on the GPU. Nothing here
    synchronises the device: Swin -> filter/pack -> fusion are stream-ordered kernels (call `.check()` on the two
    modules, or use multimodal_evaluate, before trusting the logits).

    Batch semantics at U > 1: `per_utterance=True` (default) gives every utterance the result the reference computes for
    it at its default `trg_batch_size=1` (main.py:56) -- the only mode its published W-F1 was produced in. The reference's
    literal batch code differs from that for U > 1 in two ways: the "no frame passes" fallback is decided for the whole
    batch (train.py:187,223), and its re-pack loop has the `margin += n-1` off-by-one (train.py:200,213).
    `per_utterance=False` reproduces the first (whole-batch fallback decision) but NOT the off-by-one."""
    (ids, mask, sep, audio, audio_mask, vision, vision_mask, _labels, faces, num_imgs, idx) = batch
    n = [int(x) for x in (num_imgs.tolist() if torch.is_tensor(num_imgs) else num_imgs)]
    frames = gather_valid_frames(faces.to("cuda", non_blocking=True), n)
    F = frames.shape[0]
    if gumbel is None:      # the draw F.gumbel_softmax makes internally (src/models.py:31-32)
        gumbel = -torch.empty(F, swin_model.num_labels, device="cuda").exponential_().log()
    _, probs, _ = swin_model.forward_full(frames, gumbel)
    cache = swin_model._out_cache if getattr(swin_model, "_graph", False) else None     # graph mode: stable buffers
    v519, new_mask = filter_pack(vision, vision_mask, n, probs, threshold, per_utterance, cache=cache)
    logits = multimodal_model(ids, mask, sep, audio, audio_mask, v519, new_mask, idx)
    if return_intermediates:
        return logits, dict(probs=probs, vision519=v519, new_mask=new_mask)
    return logits


def multimodal_evaluate(swin_model, multimodal_model, loader: Iterable, criterion=None, threshold: float = 0.2,
                        per_utterance: bool = True):
    """-> (avg_loss, results (n,labels), truths (n,)) like train.py:154-243."""
    results, truths, total_loss, count = [], [], 0.0, 0
    for batch in loader:
        logits = evaluate_batch(swin_model, multimodal_model, batch, threshold, per_utterance=per_utterance)
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
