"""Seeded synthetic weights and MELD-shaped inputs (there are no checkpoints or datasets offline).

* `*_state_dict_spec` list the reference's state_dict keys and shapes for this path (checked
  against the instantiated reference in tests/test_oracle_vs_reference.py when /root/reference is present).
* `stress_state_dict` fills them with a seeded "stress" initialisation: plain trunc_normal(std=.02) makes every
  softmax near-uniform and would hide indexing bugs, so weights are scaled to keep activations and attention
  logits O(1), biases / LayerNorm affine terms / BatchNorm statistics are non-trivial.
* `synthetic_batch` builds the input tuple the reference's DataLoader yields (utils/dataset.py:291-292).

Each tensor is drawn from its own generator seeded by (seed, crc32(key)), so the result does not depend on key
order and is identical here and on the GPU box (same torch build).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import FmmtConfig, FusionConfig, SwinConfig, TextConfig

Spec = "OrderedDict[str, Tuple[int, ...]]"


# ----------------------------------------------------------------------------------------------- key listings
def swin_cls_state_dict_spec(cfg: SwinConfig) -> Spec:
    """Keys of SwinForAffwildClassification (src/models.py:14-24; Swin_Transformer.py:434-494)."""
    s: Spec = OrderedDict()
    C0, ws = cfg.embed_dim, cfg.window_size
    N = ws * ws
    s["swin.patch_embed.proj.weight"] = (C0, cfg.in_chans, cfg.patch_size, cfg.patch_size)
    s["swin.patch_embed.proj.bias"] = (C0,)
    s["swin.patch_embed.norm.weight"] = (C0,)
    s["swin.patch_embed.norm.bias"] = (C0,)
    for li, (depth, heads, C, R) in enumerate(zip(cfg.depths, cfg.num_heads, cfg.dims, cfg.resolutions)):
        hid = int(C * cfg.mlp_ratio)
        for bi in range(depth):
            p = f"swin.layers.{li}.blocks.{bi}."
            s[p + "norm1.weight"] = (C,)
            s[p + "norm1.bias"] = (C,)
            s[p + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
            s[p + "attn.relative_position_index"] = (N, N)          # int64 buffer
            s[p + "attn.qkv.weight"] = (3 * C, C)
            s[p + "attn.qkv.bias"] = (3 * C,)
            s[p + "attn.proj.weight"] = (C, C)
            s[p + "attn.proj.bias"] = (C,)
            s[p + "norm2.weight"] = (C,)
            s[p + "norm2.bias"] = (C,)
            s[p + "mlp.fc1.weight"] = (hid, C)
            s[p + "mlp.fc1.bias"] = (hid,)
            s[p + "mlp.fc2.weight"] = (C, hid)
            s[p + "mlp.fc2.bias"] = (C,)
            if bi % 2 == 1 and R > ws:
                s[p + "attn_mask"] = ((R // ws) ** 2, N, N)         # float buffer of {0,-100}
        if li < len(cfg.depths) - 1:
            p = f"swin.layers.{li}.downsample."
            s[p + "reduction.weight"] = (2 * C, 4 * C)
            s[p + "norm.weight"] = (4 * C,)
            s[p + "norm.bias"] = (4 * C,)
    Cl, Rl = cfg.dims[-1], cfg.resolutions[-1]
    s["swin.output_layer.0.weight"] = (Cl,)
    s["swin.output_layer.0.bias"] = (Cl,)
    s["swin.output_layer.2.weight"] = (cfg.feat_dim, Rl * Rl * Cl)
    s["swin.output_layer.2.bias"] = (cfg.feat_dim,)
    s["swin.output_layer.3.weight"] = (cfg.feat_dim,)
    s["swin.output_layer.3.bias"] = (cfg.feat_dim,)
    s["swin.output_layer.3.running_mean"] = (cfg.feat_dim,)
    s["swin.output_layer.3.running_var"] = (cfg.feat_dim,)
    s["swin.output_layer.3.num_batches_tracked"] = ()
    s["linear.weight"] = (cfg.head_hidden, cfg.feat_dim)
    s["linear.bias"] = (cfg.head_hidden,)
    s["classifier.weight"] = (cfg.num_labels, cfg.head_hidden)
    s["classifier.bias"] = (cfg.num_labels,)
    return s


def _meld_trans_spec(s: Spec, prefix: str, layers: int, max_len: int, H: int, ffn: int):
    """MELDTransEncoder (modules/Transformer.py:196-204)."""
    s[prefix + "position_embeddings.weight"] = (max_len, H)
    for i in range(layers):
        p = f"{prefix}layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"transformer_self_attention.selfatt.{n}.weight"] = (H, H)
            s[p + f"transformer_self_attention.selfatt.{n}.bias"] = (H,)
        s[p + "transformer_self_attention.dense_norm.dense.weight"] = (H, H)
        s[p + "transformer_self_attention.dense_norm.dense.bias"] = (H,)
        s[p + "transformer_self_attention.dense_norm.LayerNorm.weight"] = (H,)
        s[p + "transformer_self_attention.dense_norm.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (ffn, H)
        s[p + "intermediate.dense.bias"] = (ffn,)
        s[p + "output.dense.weight"] = (H, ffn)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)


def _additive_attention_spec(s: Spec, prefix: str, H: int):
    """AdditiveAttention (modules/Transformer.py:8-21)."""
    s[prefix + "query_vector"] = (H,)
    s[prefix + "value.weight"] = (1, H)
    s[prefix + "value.bias"] = (1,)
    s[prefix + "P.weight"] = (H, H)
    s[prefix + "P.bias"] = (H,)
    s[prefix + "Q.weight"] = (H, H)
    s[prefix + "Q.bias"] = (H,)


def _cmt_spec(s: Spec, prefix: str, layers: int, H: int):
    """CrossModalTransformerEncoder (modules/CrossmodalTransformer.py:23-46,112-130)."""
    s[prefix + "version"] = (1,)
    s[prefix + "embed_positions._float_tensor"] = (1,)
    for i in range(layers):
        p = f"{prefix}layers.{i}."
        s[p + "self_attn.in_proj_weight"] = (3 * H, H)
        s[p + "self_attn.in_proj_bias"] = (3 * H,)
        s[p + "self_attn.out_proj.weight"] = (H, H)
        s[p + "self_attn.out_proj.bias"] = (H,)
        s[p + "fc1.weight"] = (4 * H, H)
        s[p + "fc1.bias"] = (4 * H,)
        s[p + "fc2.weight"] = (H, 4 * H)
        s[p + "fc2.bias"] = (H,)
        for j in (0, 1):
            s[p + f"layer_norms.{j}.weight"] = (H,)
            s[p + f"layer_norms.{j}.bias"] = (H,)
    s[prefix + "layer_norm.weight"] = (H,)
    s[prefix + "layer_norm.bias"] = (H,)


def text_state_dict_spec(t: TextConfig) -> Spec:
    """HF RobertaModel / BertModel keys (transformers==4.24.0 per requirements.txt:4), prefix roberta./bert."""
    s: Spec = OrderedDict()
    p = t.kind + "."
    D = t.hidden
    s[p + "embeddings.word_embeddings.weight"] = (t.vocab_size, D)
    s[p + "embeddings.position_embeddings.weight"] = (t.max_pos, D)
    s[p + "embeddings.token_type_embeddings.weight"] = (t.type_vocab, D)
    s[p + "embeddings.LayerNorm.weight"] = (D,)
    s[p + "embeddings.LayerNorm.bias"] = (D,)
    for i in range(t.layers):
        q = f"{p}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[q + f"attention.self.{n}.weight"] = (D, D)
            s[q + f"attention.self.{n}.bias"] = (D,)
        s[q + "attention.output.dense.weight"] = (D, D)
        s[q + "attention.output.dense.bias"] = (D,)
        s[q + "attention.output.LayerNorm.weight"] = (D,)
        s[q + "attention.output.LayerNorm.bias"] = (D,)
        s[q + "intermediate.dense.weight"] = (t.ffn, D)
        s[q + "intermediate.dense.bias"] = (t.ffn,)
        s[q + "output.dense.weight"] = (D, t.ffn)
        s[q + "output.dense.bias"] = (D,)
        s[q + "output.LayerNorm.weight"] = (D,)
        s[q + "output.LayerNorm.bias"] = (D,)
    return s  # the pooler is never used on the path (src/models.py:106 takes outputs[0])


def multimodal_state_dict_spec(cfg: FmmtConfig) -> Spec:
    """Keys of MultiModalTransformerForClassification (src/models.py:71-93)."""
    f, t = cfg.fusion, cfg.text
    s = text_state_dict_spec(t)
    H = f.hidden
    s["text_linear.weight"] = (H, t.hidden)
    s["text_linear.bias"] = (H,)
    s["audio_linear.weight"] = (H, f.audio_dim)
    s["audio_linear.bias"] = (H,)
    _meld_trans_spec(s, "audio_utt_transformer.", f.audio_layers, f.audio_len, H, f.ffn)
    s["vision_linear.weight"] = (H, f.vision_dim + f.num_labels)
    s["vision_linear.bias"] = (H,)
    _meld_trans_spec(s, "vision_utt_transformer.", f.vision_layers, f.vision_len, H, f.ffn)
    _additive_attention_spec(s, "attention.", H)
    _cmt_spec(s, "CrossModalTrans_TA.", f.cmt_layers_ta, H)
    _cmt_spec(s, "CrossModalTrans_TA_V.", f.cmt_layers_tav, H)
    s["classifier.weight"] = (f.num_labels, H)
    s["classifier.bias"] = (f.num_labels,)
    return s


def unimodal_state_dict_spec(f: FusionConfig) -> Spec:
    """Keys of meld_utt_transformer (src/models.py:192-207)."""
    s: Spec = OrderedDict()
    H = f.hidden
    s["modality_linear.weight"] = (H, f.vision_dim)
    s["modality_linear.bias"] = (H,)
    _meld_trans_spec(s, "utt_transformer.", f.vision_layers, f.vision_len, H, f.ffn)
    _additive_attention_spec(s, "attention.", H)
    s["classifier.weight"] = (f.num_labels, H)
    s["classifier.bias"] = (f.num_labels,)
    return s


# ----------------------------------------------------------------------------------------------- stress init
def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))


def _relative_position_index(ws: int) -> torch.Tensor:
    # index[i][j] = (y_i - y_j + ws-1) * (2ws-1) + (x_i - x_j + ws-1)   (Swin_Transformer.py:92-103)
    c = torch.arange(ws)
    yy, xx = torch.meshgrid(c, c, indexing="ij")
    y, x = yy.reshape(-1), xx.reshape(-1)
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def shift_attn_mask(R: int, ws: int, shift: int) -> torch.Tensor:
    """(nW, N, N) tensor of {0, -100}: region ids in shifted coordinates (Swin_Transformer.py:208-229)."""
    def region(p):
        return 0 if p < R - ws else (1 if p < R - shift else 2)
    rid = torch.tensor([[3 * region(r) + region(c) for c in range(R)] for r in range(R)])
    nw = R // ws
    w = rid.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)
    diff = w[:, None, :] - w[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


def stress_state_dict(spec: Spec, seed: int, swin_cfg: SwinConfig | None = None) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in spec.items():
        g = _gen(seed, key)
        parts = key.split(".")
        leaf = parts[-1]
        # the module that owns this tensor decides (NOT a substring of the path: `dense_norm.dense` is a Linear)
        owner = parts[-2] if len(parts) >= 2 else ""
        if owner.isdigit() and len(parts) >= 3:
            owner = parts[-3] + "." + owner
        is_norm = owner.lower() in ("norm", "norm1", "norm2", "layernorm", "layer_norm", "output_layer.0",
                                    "output_layer.3", "layer_norms.0", "layer_norms.1")
        if key.endswith("relative_position_index"):
            t = _relative_position_index(swin_cfg.window_size if swin_cfg else 7)
        elif key.endswith("attn_mask"):
            li = int(key.split(".")[2])
            R = swin_cfg.resolutions[li]
            t = shift_attn_mask(R, swin_cfg.window_size, swin_cfg.window_size // 2)
        elif leaf == "num_batches_tracked":
            t = torch.tensor(0, dtype=torch.int64)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "version":
            t = torch.tensor([2.0])
        elif leaf == "_float_tensor":
            t = torch.zeros(1)
        elif leaf == "relative_position_bias_table":
            t = 0.5 * torch.randn(shape, generator=g)
        elif leaf == "query_vector":
            t = torch.randn(shape, generator=g)
        elif "embeddings" in key and not is_norm:          # word / position / token-type / learned positions
            t = 0.5 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("bias", "in_proj_bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:                              # Linear / Conv / in_proj_weight: fan-in scaling
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / fan_in ** 0.5
        else:
            raise KeyError(f"no init rule for {key} {shape}")
        assert tuple(t.shape) == tuple(shape), (key, t.shape, shape)
        sd[key] = t
    return sd


def swin_cls_stress_state_dict(cfg: SwinConfig, seed: int = 1111):
    return stress_state_dict(swin_cls_state_dict_spec(cfg), seed, cfg)


def multimodal_stress_state_dict(cfg: FmmtConfig, seed: int = 1111):
    return stress_state_dict(multimodal_state_dict_spec(cfg), seed + 1)


def unimodal_stress_state_dict(cfg: FusionConfig, seed: int = 1111):
    return stress_state_dict(unimodal_state_dict_spec(cfg), seed + 2)


# ----------------------------------------------------------------------------------------------- inputs
def synthetic_faces(n_frames: int, seed: int, img: int = 224) -> torch.Tensor:
    """(n,3,img,img) fp32 in [-1,1]: uint8 crops at half resolution -> 2x bicubic upsample -> ToTensor+Normalize(.5,.5)
    (BASELINE.json "160x3x112x112 face stack"; utils/dataset.py:41-57). The same tensor feeds both sides."""
    g = _gen(seed, f"faces{n_frames}")
    raw = torch.randint(0, 256, (n_frames, 3, img // 2, img // 2), generator=g, dtype=torch.uint8)
    x = torch.nn.functional.interpolate(raw.float(), size=(img, img), mode="bicubic", align_corners=False)
    x = x.clamp_(0, 255) / 255.0
    return (x - 0.5) / 0.5


def synthetic_batch(cfg: FmmtConfig, U: int, L: int = 128, seed: int = 1111, n_frames=None, audio_valid: int = 100,
                    with_faces: bool = True) -> Dict[str, torch.Tensor]:
    """The tuple `multimodal_evaluate` unpacks (train.py:166-167), MELD-shaped, plus explicit Gumbel noise (F6)."""
    f, t = cfg.fusion, cfg.text
    g = _gen(seed, f"batch{U}x{L}")
    if n_frames is None:
        n_frames = [f.vision_len] * U
    n_frames = [int(n) for n in n_frames]
    assert len(n_frames) == U and all(1 <= n <= f.vision_len for n in n_frames)
    lo, hi = (3, 50000) if t.kind == "roberta" else (1000, 30000)
    hi = min(hi, t.vocab_size - 1)
    ids = torch.randint(lo, hi, (U, L), generator=g, dtype=torch.int64)
    mask = torch.zeros(U, L, dtype=torch.int64)
    sep = torch.zeros(U, L, dtype=torch.int64)
    idx = torch.zeros(U, dtype=torch.int64)
    cls_id, sep_id = (0, 2) if t.kind == "roberta" else (101, 102)
    for u in range(U):
        # dialogue of 4 utterances at fixed offsets (src/meld_bert_extraText.py:97-112): roberta
        # <s> A </s></s> B </s></s> C </s></s> D </s>; bert [CLS] A [SEP] B [SEP] C [SEP] D [SEP]
        ends = [L // 5, 2 * L // 5 + u % 3, 3 * L // 5, 4 * L // 5 - u % 2]
        ids[u, 0] = cls_id
        for e in ends:
            sep[u, e] = 1
            ids[u, e] = sep_id
            if t.kind == "roberta" and e != ends[-1]:
                ids[u, e + 1] = sep_id
        mask[u, : ends[-1] + 1] = 1
        ids[u, ends[-1] + 1:] = 0          # zero padding (meld_bert_extraText.py:121-124)
        idx[u] = (u * 7 + seed) % 4
    audio = torch.randn(U, f.audio_len, f.audio_dim, generator=g)
    audio_mask = torch.zeros(U, f.audio_len)
    audio_mask[:, : min(audio_valid, f.audio_len)] = 1
    vision = torch.randn(U, f.vision_len, f.vision_dim, generator=g)
    vision_mask = torch.zeros(U, f.vision_len)
    for u, n in enumerate(n_frames):
        vision_mask[u, :n] = 1
    F = sum(n_frames)
    # g = -log(E), E ~ Exp(1): the noise F.gumbel_softmax draws internally (src/models.py:31-32)
    gumbel = -torch.empty(F, f.num_labels).exponential_(generator=g).log()
    out = dict(text_ids=ids, text_mask=mask, sep_mask=sep, audio=audio, audio_mask=audio_mask, vision=vision,
               vision_mask=vision_mask, num_imgs=torch.tensor(n_frames, dtype=torch.int64), idx_in_dia=idx,
               gumbel=gumbel)
    if with_faces:
        faces = torch.zeros(U, f.vision_len, 3, cfg.swin.img_size, cfg.swin.img_size)
        for u, n in enumerate(n_frames):
            faces[u, :n] = synthetic_faces(n, seed * 131 + u, cfg.swin.img_size)
        out["faces"] = faces
    return out
