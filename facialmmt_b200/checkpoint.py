"""Checkpoint ingestion for `main.py --doEval`.

The reference evaluates from whole-module pickles (`train.py:428-432`: `torch.load(...)` of the objects saved during training,
i.e. PyTorch-Lightning `_LiteModule` -> `DataParallel` -> module), and builds its Swin-cls model from an Aff-Wild pre-trained
backbone whose keys carry a `backbone.` prefix (`train.py:316-331`). Both become plain state_dicts with the reference's own key
names, which is what `fmmt_load_weight` takes. Host-side only.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping

import torch

_WRAPPER_PREFIXES = ("_forward_module.", "_original_module.", "_module.", "module.")


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


def load_state_dict_file(path: str, trust: bool = False) -> Dict[str, torch.Tensor]:
    """What the reference saves (`train.py:428-432` loads the same files with a bare `torch.load`).

    A plain state_dict / `{'state_dict': ...}` file is read with `weights_only=True` (no code execution). The reference's own
    format -- the whole pickled `_LiteModule(DataParallel(module))` -- cannot be read that way; it is unpickled only when the
    caller opts in (`trust=True`, `main.py --trust_checkpoint`, or FMMT_TRUST_CHECKPOINT=1), and then WITHOUT needing
    pytorch_lightning or the reference's classes importable: `StubUnpickler` rebuilds every unknown class as an inert shell
    that only keeps its attributes, and the tensors are collected from the `_parameters` / `_buffers` / `_modules` tree."""
    import os
    import pickle
    try:
        return to_state_dict(torch.load(path, map_location="cpu", weights_only=True))
    except (pickle.UnpicklingError, RuntimeError, AttributeError, ModuleNotFoundError) as e:
        if not (trust or os.environ.get("FMMT_TRUST_CHECKPOINT") == "1"):
            raise RuntimeError(
                f"{path} is not a plain state_dict (weights_only load failed: {type(e).__name__}). It is probably the "
                "reference's whole-module pickle (train.py:428-432); unpickling executes code from the file, so it needs "
                "--trust_checkpoint (or FMMT_TRUST_CHECKPOINT=1).") from e
    return load_pickled_module(path)


class _Shell:
    """Stand-in for any class the unpickler cannot (or should not) import: keeps attributes, runs no code."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots)
            if isinstance(state[0], dict):
                self.__dict__.update(state[0])
            self.__dict__.update(state[1])


def _make_stub_pickle_module():
    """A `pickle_module` for torch.load whose Unpickler resolves torch / collections / builtins normally and turns every
    other global (pytorch_lightning.lite.wrappers._LiteModule, the reference's src.models classes, ...) into `_Shell`."""
    import pickle
    import types

    safe_roots = ("torch", "collections", "builtins", "__builtin__", "numpy", "_codecs", "copyreg", "copy_reg")

    class StubUnpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in safe_roots:
                try:
                    return super().find_class(module, name)
                except (AttributeError, ModuleNotFoundError):
                    pass
            return type(name, (_Shell,), {"__module__": module})

    mod = types.ModuleType("fmmt_stub_pickle")
    mod.Unpickler = StubUnpickler
    mod.load = lambda f, **kw: StubUnpickler(f, **kw).load()
    mod.__name__ = "pickle"
    return mod


def _collect_module_tensors(obj, prefix: str, out: Dict[str, torch.Tensor]):
    d = getattr(obj, "__dict__", {})
    for kind in ("_parameters", "_buffers"):
        for k, v in (d.get(kind) or {}).items():
            if v is not None and k not in (d.get("_non_persistent_buffers_set") or ()):
                out[prefix + k] = v.data if hasattr(v, "data") else v
    for k, m in (d.get("_modules") or {}).items():
        if m is not None:
            _collect_module_tensors(m, prefix + k + ".", out)


def load_pickled_module(path: str) -> Dict[str, torch.Tensor]:
    """The reference's doEval checkpoints (`train.py:428-432`; saved by `utils/util.py:121-132` as whole modules after
    `Lite.setup`) -> plain state_dict with the reference key names, without importing Lightning or the reference."""
    obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_make_stub_pickle_module())
    if isinstance(obj, Mapping):
        return to_state_dict(obj)
    if hasattr(obj, "state_dict") and callable(getattr(obj, "state_dict")) and not isinstance(obj, _Shell):
        return to_state_dict(obj)
    out: Dict[str, torch.Tensor] = {}
    _collect_module_tensors(obj, "", out)
    if not out:
        raise TypeError(f"no tensors found in the pickled {type(obj).__name__}")
    return strip_wrappers(out)


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
