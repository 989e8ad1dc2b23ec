"""Checkpoint ingestion for `main.py --doEval`.

The reference evaluates from whole-module pickles (`train.py:428-432`: `torch.load(...)` of the objects saved during training,
i.e. PyTorch-Lightning `_LiteModule` -> `DataParallel` -> module), and builds its Swin-cls model from an Aff-Wild pre-trained
backbone whose keys carry a `backbone.` prefix (`train.py:316-331`). Both become plain state_dicts with the reference's own key
names, which is what `fmmt_load_weight` takes. Host-side only.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping

import torch

_WRAPPER_PREFIXES = ("_forward_module.", "_module.", "module.")


def strip_wrappers(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Drops the prefixes that Lightning's `_LiteModule` and `nn.DataParallel` put in front of every key (any nesting)."""
    out = {}
    for k, v in state_dict.items():
        changed = True
        while changed:
            changed = False
            for p in _WRAPPER_PREFIXES:
                if k.startswith(p):
                    k = k[len(p):]
                    changed = True
        out[k] = v
    return out


def to_state_dict(obj) -> Dict[str, torch.Tensor]:
    """A pickled module, a `{'state_dict': ...}` checkpoint or a plain state_dict -> plain state_dict, wrappers stripped."""
    if hasattr(obj, "state_dict") and callable(obj.state_dict):
        obj = obj.state_dict()
    if isinstance(obj, Mapping) and "state_dict" in obj and isinstance(obj["state_dict"], Mapping):
        obj = obj["state_dict"]
    if not isinstance(obj, Mapping):
        raise TypeError(f"cannot read a state_dict out of {type(obj).__name__}")
    return strip_wrappers(obj)


def load_state_dict_file(path: str) -> Dict[str, torch.Tensor]:
    """`torch.load` of what the reference saves (`train.py:428-432` loads the same files). Pickled reference modules need
    the reference's classes importable; a state_dict file needs nothing."""
    return to_state_dict(torch.load(path, map_location="cpu", weights_only=False))


def remap_pretrained_backbone(model_keys: Iterable[str], pretrained: Mapping[str, torch.Tensor],
                              literal: bool = False) -> Dict[str, torch.Tensor]:
    """Aff-Wild pre-trained Swin (`backbone.*` keys) -> SwinForAffwildClassification keys (`swin.*`, `linear.*`), as
    `train.py:316-331` does; `classifier.*` is never taken from the checkpoint. `literal=True` keeps the reference's test
    `if k in pretrained_dict` on the UNprefixed model key (so only checkpoints that also carry the model's own key names match);
    the default takes every key whose `backbone.`-prefixed name exists, which is what that loop is for."""
    new = {}
    for k in model_keys:
        if k in ("classifier.weight", "classifier.bias"):
            continue
        src = "backbone." + (k[5:] if k.startswith("swin.") else k)
        if literal and k not in pretrained:
            continue
        if src in pretrained:
            new[k] = pretrained[src]
    return new
