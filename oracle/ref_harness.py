"""TEST / BASELINE INFRASTRUCTURE ONLY -- imports the *real* reference (NUSTM/FacialMMT) on CPU.

Used (a) in the build container to pin oracle/facialmmt_oracle.py against the reference's own modules and to generate
the golden vectors under tests/golden/ (tests/golden/make_golden.py), from /root/reference; (b) by bench.py's CPU legs
(`--impl reference`, `cpu_baseline`) on the GPU box, from the unmodified copy that oracle/make_ref.py places in the
git-ignored baseline/_ref/ (/root/reference does not exist there). Never imported by the product package.

Three shims make the reference importable on CPU with the installed library versions:
  1. `timm` is not installed: a stand-in module provides DropPath / to_2tuple / trunc_normal_
     (Swin_Transformer.py:6). transformers must be imported first (it probes timm.__spec__).
  2. hard-coded `.cuda()` calls (src/models.py:114-115, modules/Transformer.py:213) become no-ops.
  3. `RobertaModel/BertModel.from_pretrained(path)` (src/models.py:73,76) builds a randomly initialised model of
     the roberta-large / bert-large architecture (no checkpoints and no network here).
"""
from __future__ import annotations

import argparse
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    env = os.environ.get("FMMT_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "src", "models.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models.py"))


_installed = False


def install_shims(text_layers: int | None = None):
    """Idempotent. `text_layers` optionally shrinks the text encoder depth (for fast golden cases)."""
    global _installed
    import torch
    import transformers  # noqa: F401  (must precede the timm stand-in)
    from transformers import BertConfig, BertModel, RobertaConfig, RobertaModel

    if not _installed:
        import torch.nn as nn

        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):  # eval(): identity
                assert not self.training, "stand-in DropPath supports eval() only"
                return x

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        timm_layers.DropPath = DropPath
        timm_layers.to_2tuple = to_2tuple
        timm_layers.trunc_normal_ = nn.init.trunc_normal_
        timm.models = timm_models
        timm_models.layers = timm_layers
        timm.__spec__ = types.SimpleNamespace(name="timm", loader=None, origin="stand-in", submodule_search_locations=[])
        sys.modules.setdefault("timm", timm)
        sys.modules.setdefault("timm.models", timm_models)
        sys.modules.setdefault("timm.models.layers", timm_layers)

        torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _installed = True

    def _roberta(cls, path, *a, **k):
        cfg = RobertaConfig(vocab_size=50265, hidden_size=1024, num_hidden_layers=text_layers or 24,
                            num_attention_heads=16, intermediate_size=4096, max_position_embeddings=514,
                            type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5,
                            hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
        return cls(cfg)

    def _bert(cls, path, *a, **k):
        cfg = BertConfig(vocab_size=30522, hidden_size=1024, num_hidden_layers=text_layers or 24,
                         num_attention_heads=16, intermediate_size=4096, max_position_embeddings=512,
                         type_vocab_size=2, pad_token_id=0, layer_norm_eps=1e-12)
        return cls(cfg)

    RobertaModel.from_pretrained = classmethod(_roberta)  # shim 3
    BertModel.from_pretrained = classmethod(_bert)


def default_args(plm: str = "roberta-large", audio_len: int = 160, vision_len: int = 160, audio_dim: int = 768,
                 vision_dim: int = 512, text_len: int = 38) -> argparse.Namespace:
    """The argparse fields the three model constructors read (main.py:16-103, :134-145 dataset-derived)."""
    return argparse.Namespace(
        num_labels=7, backbone_type="SwinTransformer",
        backbone_conf_file=os.path.join(REFERENCE_ROOT, "modules/SwinTransformer/swin_conf.yaml"),
        tau=1.0, FacialEmoImpor_threshold=0.2, choice_modality="T+A+V",
        pretrainedtextmodel_path="/nonexistent/" + plm, hidden_size=768, num_attention_heads=12,
        intermediate_size=3072, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12,
        audio_featExtr_dim=audio_dim, vision_featExtr_dim=vision_dim, audio_utt_Transformernum=5,
        vision_utt_Transformernum=2, get_text_utt_max_lens=text_len, get_audio_utt_max_lens=audio_len,
        get_vision_utt_max_lens=vision_len, crossmodal_num_heads_TA=12, crossmodal_layers_TA=2,
        crossmodal_attn_dropout_TA=0.1, crossmodal_num_heads_TA_V=12, crossmodal_layers_TA_V=2,
        crossmodal_attn_dropout_TA_V=0.1)


def build_swin_cls(args=None):
    install_shims()
    from src.models import SwinForAffwildClassification
    return SwinForAffwildClassification(args or default_args()).eval()


def build_multimodal(args=None, text_layers: int | None = None):
    install_shims(text_layers)
    from src.models import MultiModalTransformerForClassification
    return MultiModalTransformerForClassification(args or default_args()).eval()


def build_unimodal(args=None):
    install_shims()
    from src.models import meld_utt_transformer
    return meld_utt_transformer(args or default_args()).eval()


def literal_eval_loop():
    """The reference's own `multimodal_evaluate` (train.py:154-243) as a callable factory, obtained by exec'ing its source
    text at run time (train.py itself cannot be imported: it needs pytorch_lightning). Nothing is copied into the repo.
    `make(args, loader)` -> multimodal_evaluate(shareSwin_model, multimodal_model, criterion, test)."""
    import textwrap

    import torch
    lines = open(os.path.join(REFERENCE_ROOT, "train.py")).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("def multimodal_evaluate("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].strip().startswith("def "))
    src = textwrap.dedent("\n".join(lines[start:end]))

    def make(args, loader):
        ns = {"torch": torch, "args": args, "trg_test_loader": loader, "trg_valid_loader": loader}
        exec(compile(src, "reference:train.py:multimodal_evaluate", "exec"), ns)
        return ns["multimodal_evaluate"]
    return make
