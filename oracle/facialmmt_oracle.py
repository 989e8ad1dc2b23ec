"""TEST INFRASTRUCTURE -- CPU restatement (fp32, torch-CPU tensor ops only) of the FacialMMT inference forward path.

This is the ORACLE for the CUDA path. It is not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it. It never imports facialmmt_b200's CUDA bindings and never
imports /root/reference; every function takes a plain state_dict with the reference's key names and cites the
reference lines it restates.

Pinning: the reference has no tests or golden vectors of its own. The oracle is pinned
against the reference ITSELF, imported in the build container (oracle/ref_harness.py):
tests/golden/make_golden.py runs the reference modules on seeded inputs/weights and commits the outputs under
tests/golden/; tests/test_oracle_golden.py checks this file against those vectors (max-abs <= 2e-4 on O(1) outputs),
and tests/test_oracle_vs_reference.py re-runs the live comparison whenever /root/reference is present.

Third-party arithmetic not under /root/reference: HF transformers RobertaModel / BertModel (pinned
transformers==4.24.0, requirements.txt:4; call sites src/models.py:73,76,101,104). `text_encoder` restates the
published BERT/RoBERTa encoder (post-LN, erf-GELU, additive key mask, RoBERTa position ids = pad_id + cumsum of
non-pad) and is pinned against the installed transformers 5.5 model in the same golden script.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# =============================================================================================== primitives
def layer_norm(x, w, b, eps):
    """nn.LayerNorm == TF-style LN of modules/Transformer.py:57-61: biased variance, eps inside the sqrt."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def gelu_erf(x):
    """modules/Transformer.py:119-124; nn.GELU() (Swin_Transformer.py:14); F.gelu (CrossmodalTransformer.py:157)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


# =============================================================================================== Swin-tiny
def swin_window_index(R: int, ws: int, shift: int) -> torch.Tensor:
    """Token index (y*R+x, in the un-shifted image) feeding window-order row (w*ws*ws + i).

    shifted[h,w] = x[(h+shift) mod R, (w+shift) mod R] (torch.roll(-shift), Swin_Transformer.py:244); windows are
    (wy, wx) row-major, in-window tokens (ty, tx) row-major (window_partition :43-44)."""
    nw = R // ws
    wy, wx, ty, tx = torch.meshgrid(torch.arange(nw), torch.arange(nw), torch.arange(ws), torch.arange(ws),
                                    indexing="ij")
    h = (wy * ws + ty + shift) % R
    w = (wx * ws + tx + shift) % R
    return (h * R + w).reshape(-1)


def swin_shift_mask(R: int, ws: int, shift: int) -> torch.Tensor:
    """(nW, N, N) of {0,-100} (Swin_Transformer.py:208-229), from region ids in shifted coordinates."""
    def region(p):
        return 0 if p < R - ws else (1 if p < R - shift else 2)
    rid = torch.tensor([[3 * region(r) + region(c) for c in range(R)] for r in range(R)])
    nw = R // ws
    w = rid.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)
    return (w[:, None, :] != w[:, :, None]).float() * -100.0


def swin_rel_bias(table: torch.Tensor, ws: int) -> torch.Tensor:
    """(nH, N, N): bias[h,i,j] = table[(y_i-y_j+ws-1)*(2ws-1) + (x_i-x_j+ws-1), h] (Swin_Transformer.py:92-103,126-129)."""
    c = torch.arange(ws)
    yy, xx = torch.meshgrid(c, c, indexing="ij")
    y, x = yy.reshape(-1), xx.reshape(-1)
    idx = (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)
    return table[idx.reshape(-1)].reshape(ws * ws, ws * ws, -1).permute(2, 0, 1).contiguous()


def swin_block(sd: SD, p: str, x: torch.Tensor, R: int, heads: int, ws: int, shift: int) -> torch.Tensor:
    """SwinTransformerBlock.forward (Swin_Transformer.py:233-270) + WindowAttention.forward (:113-143)."""
    Fn, T, C = x.shape
    N = ws * ws
    hd = C // heads
    idx = swin_window_index(R, ws, shift)
    h = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    win = h[:, idx, :].reshape(Fn * (T // N), N, C)
    qkv = linear(win, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    att = q @ k.transpose(-2, -1) + swin_rel_bias(sd[p + "attn.relative_position_bias_table"], ws)[None]
    if shift > 0:
        nW = T // N
        m = swin_shift_mask(R, ws, shift)
        att = (att.view(Fn, nW, heads, N, N) + m[None, :, None]).view(-1, heads, N, N)
    att = torch.softmax(att, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(-1, N, C)
    o = linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]).reshape(Fn, T, C)
    back = torch.empty_like(o)
    back[:, idx, :] = o                      # window_reverse + roll(+shift)  (:258-264)
    x = x + back
    h = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    h = gelu_erf(linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def swin_patch_merge(sd: SD, p: str, x: torch.Tensor, R: int) -> torch.Tensor:
    """PatchMerging.forward (Swin_Transformer.py:307-328): (2y,2x),(2y+1,2x),(2y,2x+1),(2y+1,2x+1) -> LN(4C) -> Linear."""
    Fn, T, C = x.shape
    x = x.view(Fn, R, R, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(Fn, -1, 4 * C)
    x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return linear(x, sd[p + "reduction.weight"])


def swin_features(sd: SD, frames: torch.Tensor, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), ws: int = 7,
                  patch: int = 4, collect: Optional[dict] = None) -> torch.Tensor:
    """SwinTransformer.forward (Swin_Transformer.py:515-541): (F,3,224,224) -> (F,512)."""
    Fn, Cin, Hh, Ww = frames.shape
    w = sd["swin.patch_embed.proj.weight"]
    C = w.shape[0]
    R = Hh // patch
    # Conv2d(k=4,s=4) as a per-patch GEMM, K index = c*16 + dy*4 + dx  (PatchEmbed :419)
    pt = frames.reshape(Fn, Cin, R, patch, R, patch).permute(0, 2, 4, 1, 3, 5).reshape(Fn, R * R, Cin * patch * patch)
    x = linear(pt, w.reshape(C, -1), sd["swin.patch_embed.proj.bias"])
    x = layer_norm(x, sd["swin.patch_embed.norm.weight"], sd["swin.patch_embed.norm.bias"], 1e-5)
    if collect is not None:
        collect["patch_embed"] = x
    for li, depth in enumerate(depths):
        wsl, shiftable = (ws, True) if R > ws else (R, False)     # :192-195
        for bi in range(depth):
            shift = ws // 2 if (bi % 2 == 1 and shiftable) else 0
            x = swin_block(sd, f"swin.layers.{li}.blocks.{bi}.", x, R, heads[li], wsl, shift)
            if collect is not None:
                collect[f"layer{li}.block{bi}"] = x
        if li < len(depths) - 1:
            x = swin_patch_merge(sd, f"swin.layers.{li}.downsample.", x, R)
            R //= 2
    # output_layer: LN -> flatten token-major -> Linear -> BatchNorm1d(eval)   (:491-494)
    x = layer_norm(x, sd["swin.output_layer.0.weight"], sd["swin.output_layer.0.bias"], 1e-5).reshape(Fn, -1)
    x = linear(x, sd["swin.output_layer.2.weight"], sd["swin.output_layer.2.bias"])
    x = (x - sd["swin.output_layer.3.running_mean"]) / torch.sqrt(sd["swin.output_layer.3.running_var"] + 1e-5)
    return x * sd["swin.output_layer.3.weight"] + sd["swin.output_layer.3.bias"]


def swin_cls_logits(sd: SD, frames: torch.Tensor, **kw) -> torch.Tensor:
    """SwinForAffwildClassification.forward without the sampling step (src/models.py:27-30)."""
    x = swin_features(sd, frames, **kw)
    x = torch.relu(linear(x, sd["linear.weight"], sd["linear.bias"]))
    return linear(x, sd["classifier.weight"], sd["classifier.bias"])


def gumbel_softmax_probs(logits: torch.Tensor, gumbel: torch.Tensor, tau: float) -> torch.Tensor:
    """F.gumbel_softmax(logits, tau, hard=False) with the noise made explicit (src/models.py:31-32)."""
    return torch.softmax((logits + gumbel) / tau, dim=-1)


# =============================================================================================== eval glue
def filter_pack(vision: torch.Tensor, vision_mask: torch.Tensor, num_imgs: Sequence[int], probs: torch.Tensor,
                threshold: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """train.py:183-232 with per-utterance index arithmetic (== the literal code at the reference's batch size 1; the
    literal `margin += n-1` for U>1 is a bug). Keep frames with sum_c p_c^2 > threshold, pack them to the
    front, append their 7 probabilities; if NO frame of the batch passes, keep the original inputs/mask (:223-232).
    Returns (vision519 (U,Lv,D+7), new_mask (U,Lv))."""
    U, Lv, D = vision.shape
    nl = probs.shape[1]
    imp = (probs * probs).sum(-1)            # diag(P P^T)  (:183-184)
    keep = imp > threshold
    emo = torch.zeros(U, Lv, nl)
    if bool(keep.any()):
        new_v = torch.zeros_like(vision)
        new_m = torch.zeros_like(vision_mask)
        off = 0
        for u in range(U):
            n = int(num_imgs[u])
            loc = torch.nonzero(keep[off:off + n]).squeeze(1)
            k = loc.numel()
            new_v[u, :k] = vision[u, loc]
            emo[u, :k] = probs[off + loc]
            new_m[u, :k] = 1
            off += n
        return torch.cat([new_v, emo], -1), new_m
    off = 0
    for u in range(U):
        n = int(num_imgs[u])
        emo[u, :n] = probs[off:off + n]
        off += n
    return torch.cat([vision, emo], -1), vision_mask.clone()


# =============================================================================================== text encoder (HF)
def text_encoder(sd: SD, ids: torch.Tensor, mask: torch.Tensor, kind: str, heads: int = 16, eps: Optional[float] = None,
                 pad_id: Optional[int] = None) -> torch.Tensor:
    """HF RobertaModel / BertModel last hidden state (outputs[0], src/models.py:99-106), eval mode."""
    p = kind + "."
    eps = eps if eps is not None else (1e-5 if kind == "roberta" else 1e-12)
    U, L = ids.shape
    if kind == "roberta":
        pad = 1 if pad_id is None else pad_id
        ne = (ids != pad).long()
        pos = torch.cumsum(ne, dim=1) * ne + pad        # create_position_ids_from_input_ids
    else:
        pos = torch.arange(L)[None].expand(U, L)
    x = sd[p + "embeddings.word_embeddings.weight"][ids] + sd[p + "embeddings.position_embeddings.weight"][pos] \
        + sd[p + "embeddings.token_type_embeddings.weight"][0]
    x = layer_norm(x, sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"], eps)
    D = x.shape[-1]
    hd = D // heads
    add = (1.0 - mask.float())[:, None, None, :] * torch.finfo(torch.float32).min
    i = 0
    while f"{p}encoder.layer.{i}.attention.self.query.weight" in sd:
        q_ = f"{p}encoder.layer.{i}."
        q = linear(x, sd[q_ + "attention.self.query.weight"], sd[q_ + "attention.self.query.bias"])
        k = linear(x, sd[q_ + "attention.self.key.weight"], sd[q_ + "attention.self.key.bias"])
        v = linear(x, sd[q_ + "attention.self.value.weight"], sd[q_ + "attention.self.value.bias"])
        q, k, v = (t.view(U, L, heads, hd).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd) + add, dim=-1) @ v
        a = a.transpose(1, 2).reshape(U, L, D)
        x = layer_norm(linear(a, sd[q_ + "attention.output.dense.weight"], sd[q_ + "attention.output.dense.bias"]) + x,
                       sd[q_ + "attention.output.LayerNorm.weight"], sd[q_ + "attention.output.LayerNorm.bias"], eps)
        h = gelu_erf(linear(x, sd[q_ + "intermediate.dense.weight"], sd[q_ + "intermediate.dense.bias"]))
        x = layer_norm(linear(h, sd[q_ + "output.dense.weight"], sd[q_ + "output.dense.bias"]) + x,
                       sd[q_ + "output.LayerNorm.weight"], sd[q_ + "output.LayerNorm.bias"], eps)
        i += 1
    return x


def span_extract(text768: torch.Tensor, sep_mask: torch.Tensor, idx_in_dia: torch.Tensor, kind: str,
                 max_len: int = 38) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/models.py:112-150 in closed form: s_j = positions where sep_mask==1, p = idx_in_dia[u].
    p==0: rows [1, 1+min(s_0-1, max_len)); p>0: start = s_{p-1}+2 (roberta) / +1 (bert), n = s_p - s_{p-1} - 2 / - 1."""
    U, L, H = text768.shape
    out = torch.zeros(U, max_len, H)
    m = torch.zeros(U, max_len)
    gap = 2 if kind == "roberta" else 1
    for u in range(U):
        s = torch.nonzero(sep_mask[u] == 1).squeeze(1).tolist()
        p = int(idx_in_dia[u])
        if p >= len(s):
            continue
        if p == 0:
            start, n = 1, s[0] - 1
        else:
            start, n = s[p - 1] + gap, s[p] - s[p - 1] - gap
        n = max(0, min(n, max_len))
        out[u, :n] = text768[u, start:start + n]
        m[u, :n] = 1
    return out, m


# =============================================================================================== fusion stack
def meld_trans_encoder(sd: SD, p: str, x: torch.Tensor, mask: torch.Tensor, heads: int = 12, eps: float = 1e-12):
    """MELDTransEncoder.forward (modules/Transformer.py:206-226) with SelfAttention (:87-116), Residual_Norm (:144-148),
    TransformerIntermediate (:132-135), Output_Residual_Norm (:158-162). `mask` is the 0/1 (U,L) mask; the additive
    form (1-m)*-10000 is src/models.py:156-157."""
    U, L, H = x.shape
    hd = H // heads
    add = (1.0 - mask.float())[:, None, None, :] * -10000.0
    x = x + sd[p + "position_embeddings.weight"][:L][None]
    i = 0
    while f"{p}layer.{i}.intermediate.dense.weight" in sd:
        q_ = f"{p}layer.{i}."
        a_ = q_ + "transformer_self_attention."
        q = linear(x, sd[a_ + "selfatt.query.weight"], sd[a_ + "selfatt.query.bias"])
        k = linear(x, sd[a_ + "selfatt.key.weight"], sd[a_ + "selfatt.key.bias"])
        v = linear(x, sd[a_ + "selfatt.value.weight"], sd[a_ + "selfatt.value.bias"])
        q, k, v = (t.view(U, L, heads, hd).permute(0, 2, 1, 3) for t in (q, k, v))
        s = q @ k.transpose(-1, -2) / math.sqrt(hd) + add
        c = (torch.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).reshape(U, L, H)
        x = layer_norm(linear(c, sd[a_ + "dense_norm.dense.weight"], sd[a_ + "dense_norm.dense.bias"]) + x,
                       sd[a_ + "dense_norm.LayerNorm.weight"], sd[a_ + "dense_norm.LayerNorm.bias"], eps)
        h = gelu_erf(linear(x, sd[q_ + "intermediate.dense.weight"], sd[q_ + "intermediate.dense.bias"]))
        x = layer_norm(linear(h, sd[q_ + "output.dense.weight"], sd[q_ + "output.dense.bias"]) + x,
                       sd[q_ + "output.LayerNorm.weight"], sd[q_ + "output.LayerNorm.bias"], eps)
        i += 1
    return x


def sinusoidal_positions(first_channel: torch.Tensor, dim: int) -> torch.Tensor:
    """SinusoidalPositionalEmbedding.forward + make_positions (modules/position_embedding.py:8-27,44-76), padding_idx 0,
    left_pad 0: position = t+1 where x[...,0] != 0 else 0 (-> zero row); table [sin(p f_j) | cos(p f_j)],
    f_j = exp(-j ln(10000)/(dim/2-1))."""
    U, L = first_channel.shape
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    ang = torch.arange(L + 1, dtype=torch.float)[:, None] * f[None]
    table = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    table[0] = 0
    pos = torch.where(first_channel != 0, torch.arange(1, L + 1)[None].expand(U, L), torch.zeros(U, L, dtype=torch.long))
    return table[pos]


def cmt_encoder(sd: SD, p: str, x_q: torch.Tensor, x_kv: torch.Tensor, heads: int = 12) -> torch.Tensor:
    """CrossModalTransformerEncoder.forward(x_in, x_in_k, x_in_v) with x_in_k is x_in_v
    (modules/CrossmodalTransformer.py:49-90,132-164; multihead_attention.py:51-135). Batch-first (U,L,H) here; the
    reference uses (L,U,H) -- attention never mixes the batch axis, so the layouts are equivalent."""
    U, Lq, H = x_q.shape
    Lk = x_kv.shape[1]
    hd = H // heads
    scale = math.sqrt(H)
    x = scale * x_q + sinusoidal_positions(x_q[:, :, 0], H)
    e_k = scale * x_kv + sinusoidal_positions(x_kv[:, :, 0], H)
    i = 0
    while f"{p}layers.{i}.fc1.weight" in sd:
        l_ = f"{p}layers.{i}."
        W, B = sd[l_ + "self_attn.in_proj_weight"], sd[l_ + "self_attn.in_proj_bias"]
        qn = layer_norm(x, sd[l_ + "layer_norms.0.weight"], sd[l_ + "layer_norms.0.bias"], 1e-5)
        kn = layer_norm(e_k, sd[l_ + "layer_norms.0.weight"], sd[l_ + "layer_norms.0.bias"], 1e-5)
        q = linear(qn, W[:H], B[:H]) * (hd ** -0.5)
        k = linear(kn, W[H:2 * H], B[H:2 * H])
        v = linear(kn, W[2 * H:], B[2 * H:])
        q = q.view(U, Lq, heads, hd).transpose(1, 2)
        k = k.view(U, Lk, heads, hd).transpose(1, 2)
        v = v.view(U, Lk, heads, hd).transpose(1, 2)
        a = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(U, Lq, H)
        x = x + linear(a, sd[l_ + "self_attn.out_proj.weight"], sd[l_ + "self_attn.out_proj.bias"])
        h = layer_norm(x, sd[l_ + "layer_norms.1.weight"], sd[l_ + "layer_norms.1.bias"], 1e-5)
        h = gelu_erf(linear(h, sd[l_ + "fc1.weight"], sd[l_ + "fc1.bias"]))
        x = x + linear(h, sd[l_ + "fc2.weight"], sd[l_ + "fc2.bias"])
        i += 1
    return layer_norm(x, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)


def additive_attention(sd: SD, p: str, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """AdditiveAttention.forward (modules/Transformer.py:23-45): softmax_mask(value(tanh(P x + Q q))) weighted sum."""
    qv = linear(sd[p + "query_vector"], sd[p + "Q.weight"], sd[p + "Q.bias"])
    s = torch.tanh(linear(x, sd[p + "P.weight"], sd[p + "P.bias"]) + qv)
    s = linear(s, sd[p + "value.weight"], sd[p + "value.bias"]).squeeze(-1)
    s = s.masked_fill(mask == 0, float("-inf"))
    a = torch.softmax(s, dim=-1)
    return (a[:, None, :] @ x).squeeze(1)


def multimodal_forward(sd: SD, text_ids, text_mask, sep_mask, audio, audio_mask, vision519, vision_mask, idx_in_dia,
                       kind: str = "roberta", text_heads: int = 16, heads: int = 12, text_len: int = 38,
                       collect: Optional[dict] = None) -> torch.Tensor:
    """MultiModalTransformerForClassification.forward (src/models.py:95-188) -> (U,7) logits."""
    t = text_encoder(sd, text_ids, text_mask, kind, text_heads)
    t = linear(t, sd["text_linear.weight"], sd["text_linear.bias"])
    txt, txt_mask = span_extract(t, sep_mask, idx_in_dia, kind, text_len)
    a = linear(audio, sd["audio_linear.weight"], sd["audio_linear.bias"])
    a = meld_trans_encoder(sd, "audio_utt_transformer.", a, audio_mask, heads)
    v = linear(vision519, sd["vision_linear.weight"], sd["vision_linear.bias"])
    v = meld_trans_encoder(sd, "vision_utt_transformer.", v, vision_mask, heads)
    ta = torch.cat([cmt_encoder(sd, "CrossModalTrans_TA.", txt, a, heads),
                    cmt_encoder(sd, "CrossModalTrans_TA.", a, txt, heads)], dim=1)          # :171-173
    out = torch.cat([cmt_encoder(sd, "CrossModalTrans_TA_V.", ta, v, heads),
                     cmt_encoder(sd, "CrossModalTrans_TA_V.", v, ta, heads)], dim=1)        # :176-179
    m = torch.cat([txt_mask, audio_mask.float(), vision_mask.float()], dim=1)               # :180-181
    if collect is not None:
        collect.update(text=txt, text_mask=txt_mask, audio=a, vision=v, ta=ta, fused=out)
    pooled = additive_attention(sd, "attention.", out, m)
    return linear(pooled, sd["classifier.weight"], sd["classifier.bias"])


def unimodal_forward(sd: SD, inputs: torch.Tensor, utt_mask: torch.Tensor, heads: int = 12) -> torch.Tensor:
    """meld_utt_transformer.forward (src/models.py:209-223)."""
    x = linear(inputs, sd["modality_linear.weight"], sd["modality_linear.bias"])
    x = meld_trans_encoder(sd, "utt_transformer.", x, utt_mask, heads)
    pooled = additive_attention(sd, "attention.", x, utt_mask)
    return linear(pooled, sd["classifier.weight"], sd["classifier.bias"])


def evaluate_batch(swin_sd: SD, mm_sd: SD, batch: Dict[str, torch.Tensor], kind: str = "roberta", tau: float = 1.0,
                   threshold: float = 0.2, swin_kw: Optional[dict] = None, per_utterance: bool = True,
                   collect: Optional[dict] = None) -> torch.Tensor:
    """One iteration of multimodal_evaluate (train.py:164-239) -> (U,7) logits. `per_utterance=True` runs the filter
    one utterance at a time (the reference's trg_batch_size=1 semantics; the batch-level fallback of :223 is decided
    per utterance then)."""
    swin_kw = swin_kw or {}
    n = [int(v) for v in batch["num_imgs"]]
    frames = torch.cat([batch["faces"][u, :n[u]] for u in range(len(n))], 0)          # :169-179
    probs = gumbel_softmax_probs(swin_cls_logits(swin_sd, frames, **swin_kw), batch["gumbel"], tau)   # :181
    if per_utterance:
        vs, ms, off = [], [], 0
        for u in range(len(n)):
            v, m = filter_pack(batch["vision"][u:u + 1], batch["vision_mask"][u:u + 1], n[u:u + 1],
                               probs[off:off + n[u]], threshold)
            vs.append(v); ms.append(m); off += n[u]
        v519, new_mask = torch.cat(vs, 0), torch.cat(ms, 0)
    else:
        v519, new_mask = filter_pack(batch["vision"], batch["vision_mask"], n, probs, threshold)
    if collect is not None:
        collect.update(probs=probs, vision519=v519, new_mask=new_mask)
    return multimodal_forward(mm_sd, batch["text_ids"], batch["text_mask"], batch["sep_mask"], batch["audio"],
                              batch["audio_mask"], v519, new_mask, batch["idx_in_dia"], kind=kind, collect=collect)
