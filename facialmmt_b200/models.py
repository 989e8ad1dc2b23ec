"""Host-side mirror of the reference's model API (src/models.py) on top of the C ABI.

Same class names, constructor fields, forward signatures and state_dict key names as the reference, so that
`train.py`-style eval code can switch over unchanged:

    SwinForAffwildClassification(args)(images_feature, is_trg_task=True)          src/models.py:14-37
    MultiModalTransformerForClassification(config)(ids, mask, sep, audio, ...)    src/models.py:41-188
    meld_utt_transformer(args)(inputs, utt_mask)                                  src/models.py:192-223

All arithmetic happens in libfacialmmt_b200.so (hand-written sm_100a CUDA). torch is used only for device
memory, streams and the Gumbel noise draw. Inference (eval) only; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int64, c_void_p
from typing import Dict, Optional

import torch

from . import _lib
from .config import FmmtConfig, FusionConfig, SwinConfig, TextConfig
from . import synthetic as _syn


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.FmmtError("facialmmt_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")


def _precision_code(precision: str) -> int:
    if precision in ("bf16", None):
        return _lib.PRECISION_BF16
    if precision in ("fp32", "f32", "float32"):
        return _lib.PRECISION_FP32
    raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")


def _cfg_c(cfg: FmmtConfig, model: int, swin_chunk: int = 0, swin_chunk_late: int = 0,
           precision: str = "bf16") -> _lib.FmmtConfigC:
    c = _lib.FmmtConfigC()
    c.precision = _precision_code(precision)
    s, t, f = cfg.swin, cfg.text, cfg.fusion
    c.model = model
    c.img_size, c.patch_size, c.in_chans, c.embed_dim = s.img_size, s.patch_size, s.in_chans, s.embed_dim
    c.num_stages = len(s.depths)
    for i in range(4):
        c.depths[i] = s.depths[i] if i < len(s.depths) else 0
        c.num_heads[i] = s.num_heads[i] if i < len(s.num_heads) else 0
    c.window_size = s.window_size
    if float(s.mlp_ratio) != int(s.mlp_ratio):
        raise ValueError("mlp_ratio must be an integer")
    c.mlp_ratio = int(s.mlp_ratio)
    c.feat_dim, c.head_hidden, c.num_labels = s.feat_dim, s.head_hidden, s.num_labels
    c.swin_chunk, c.swin_chunk_late = swin_chunk, swin_chunk_late
    c.text_kind = _lib.TEXT_ROBERTA if t.kind == "roberta" else _lib.TEXT_BERT
    c.vocab_size, c.text_hidden, c.text_layers, c.text_heads = t.vocab_size, t.hidden, t.layers, t.heads
    c.text_ffn, c.max_pos, c.type_vocab, c.pad_id, c.text_eps = t.ffn, t.max_pos, t.type_vocab, t.pad_id, t.eps
    c.hidden, c.heads, c.ffn, c.audio_dim, c.vision_dim = f.hidden, f.heads, f.ffn, f.audio_dim, f.vision_dim
    c.audio_layers, c.vision_layers = f.audio_layers, f.vision_layers
    c.cmt_layers_ta, c.cmt_heads_ta = f.cmt_layers_ta, f.cmt_heads_ta
    c.cmt_layers_tav, c.cmt_heads_tav = f.cmt_layers_tav, f.cmt_heads_tav
    c.text_len, c.audio_len, c.vision_len, c.eps = f.text_len, f.audio_len, f.vision_len, f.eps
    if model != _lib.MODEL_SWIN_CLS:
        c.num_labels = f.num_labels
    return c


class _Module:
    """Minimal nn.Module-like shell around one fmmt_handle."""

    _model_kind = 0

    def __init__(self, cfg: FmmtConfig, swin_chunk: int = 0, swin_chunk_late: int = 0, precision: str = "bf16"):
        self.cfg = cfg
        self.precision = "fp32" if _precision_code(precision) == _lib.PRECISION_FP32 else "bf16"
        self._lib = _lib.load()
        self._cfg_c = _cfg_c(cfg, self._model_kind, swin_chunk, swin_chunk_late, precision)
        self._h = c_void_p()
        _lib.check(self._lib.fmmt_create(ctypes.byref(self._cfg_c), ctypes.byref(self._h)), "fmmt_create")
        self._finalized = False
        self.training = False
        self._captures: Dict[str, torch.Tensor] = {}

    # -- nn.Module look-alikes used by eval code
    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("facialmmt_b200 implements the inference forward only")
        return self

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.fmmt_destroy(h)
            except Exception:
                pass
            self._h = c_void_p()

    # -- weights
    def _spec(self):
        raise NotImplementedError

    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts the reference's state_dict (same key names). Integer buffers are recomputed; extra keys
        (e.g. roberta.pooler.*, position_ids) are ignored; missing keys raise when strict."""
        _require_cuda()
        spec = self._spec()
        needed = [k for k, shp in spec.items()
                  if not (k.endswith("relative_position_index") or k.endswith("num_batches_tracked")
                          or k.endswith("attn_mask") or k.endswith(".version") or k.endswith("_float_tensor"))]
        missing = [k for k in needed if k not in state_dict]
        if missing and strict:
            raise KeyError(f"missing keys in state_dict: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        for k in needed:
            if k not in state_dict:
                continue
            v = state_dict[k]
            if tuple(v.shape) != tuple(spec[k]):
                raise ValueError(f"shape mismatch for {k}: got {tuple(v.shape)}, expected {tuple(spec[k])}")
            t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
            shape = (c_int64 * max(1, t.dim()))(*t.shape)
            _lib.check(self._lib.fmmt_load_weight(self._h, k.encode(), c_void_p(t.data_ptr()), shape, t.dim()),
                       f"fmmt_load_weight({k})")
        _lib.check(self._lib.fmmt_finalize(self._h), "fmmt_finalize")
        self._finalized = True
        return self

    def check(self):
        """Synchronise this module's last forward and raise FmmtError if a kernel pipeline watchdog fired in any forward
        since the previous check (include/facialmmt_b200.h fmmt_check). Call before trusting logits."""
        _lib.check(self._lib.fmmt_check(self._h), "fmmt_check")

    # -- diagnostics
    def capture(self, name: str, numel: int) -> torch.Tensor:
        """Ask the next forward to copy a named fp32 intermediate into a new device tensor (parity tests)."""
        t = torch.full((numel,), float("nan"), device="cuda", dtype=torch.float32)
        self._captures[name] = t
        _lib.check(self._lib.fmmt_set_capture(self._h, name.encode(), _lib.ptr(t), numel), "fmmt_set_capture")
        return t

    def clear_captures(self):
        self._captures.clear()
        _lib.check(self._lib.fmmt_set_capture(self._h, None, None, 0), "fmmt_set_capture")

    def set_graph(self, on: bool = True):
        """Replay repeated identical forwards (same tensors / sizes / stream) as one CUDA graph launch (fmmt_set_graph).
        While on, the module hands out the SAME output tensors for equal shapes (a graph replays fixed pointers): the result
        of a forward is overwritten by the next forward of the same shape."""
        _lib.check(self._lib.fmmt_set_graph(self._h, int(on)), "fmmt_set_graph")
        self._graph = bool(on)
        self._out_cache = {}

    def _out(self, key, shape):
        """Output tensor: fresh, or (graph mode) the persistent one for this key/shape."""
        if not getattr(self, "_graph", False):
            return torch.empty(*shape, device="cuda", dtype=torch.float32)
        k = (key, tuple(shape))
        t = self._out_cache.get(k)
        if t is None:
            t = self._out_cache[k] = torch.empty(*shape, device="cuda", dtype=torch.float32)
        return t

    def set_profile(self, on: bool = True):
        _lib.check(self._lib.fmmt_set_profile(self._h, int(on)), "fmmt_set_profile")

    def read_profile(self) -> dict:
        """{kernel key: {ms, flops, bytes, launches}} measured with CUDA events since set_profile(True)."""
        import json
        n = int(self._lib.fmmt_profile_read(self._h, None, 0))
        buf = ctypes.create_string_buffer(n + 16)
        self._lib.fmmt_profile_read(self._h, buf, n + 16)
        return json.loads(buf.value.decode())

    def flops(self, reset: bool = False) -> float:
        return float(self._lib.fmmt_flops(self._h, int(reset)))

    def device_bytes(self) -> int:
        return int(self._lib.fmmt_device_bytes(self._h))


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.to(device="cuda", dtype=torch.float32, non_blocking=True).contiguous()


def _i64(t: torch.Tensor) -> torch.Tensor:
    return t.to(device="cuda", dtype=torch.int64, non_blocking=True).contiguous()


def _swin_config_from_args(args) -> SwinConfig:
    """args.backbone_conf_file is the reference's swin_conf.yaml (backbone_def.py:17-19); defaults are Swin-tiny."""
    if isinstance(args, SwinConfig):
        return args
    cfg = SwinConfig()
    path = getattr(args, "backbone_conf_file", None)
    if path:
        try:
            import yaml
            with open(path) as f:
                conf = yaml.safe_load(f)[getattr(args, "backbone_type", "SwinTransformer")]
            cfg = SwinConfig(img_size=conf["img_size"], patch_size=conf["patch_size"], in_chans=conf["in_chans"],
                             embed_dim=conf["embed_dim"], depths=tuple(conf["depths"]),
                             num_heads=tuple(conf["num_heads"]), window_size=conf["window_size"],
                             mlp_ratio=conf["mlp_ratio"])
        except FileNotFoundError:
            pass
    cfg.num_labels = getattr(args, "num_labels", cfg.num_labels)
    return cfg


class SwinForAffwildClassification(_Module):
    """src/models.py:14-37. `forward(images_feature, is_trg_task)` returns raw logits, or -- when is_trg_task is truthy --
    the soft Gumbel-softmax distribution F.gumbel_softmax(logits, tau) (sampled: pass `gumbel=` to inject the noise)."""

    _model_kind = _lib.MODEL_SWIN_CLS

    def __init__(self, args=None, swin_chunk: int = 0, swin_chunk_late: int = 0, precision: str = None):
        precision = precision or getattr(args, "precision", "bf16")
        if isinstance(args, FmmtConfig):
            cfg = args
            self.tau = cfg.tau
        else:
            cfg = FmmtConfig(swin=_swin_config_from_args(args) if args is not None else SwinConfig())
            self.tau = float(getattr(args, "tau", 1.0)) if args is not None else 1.0
        self.num_labels = cfg.swin.num_labels
        super().__init__(cfg, swin_chunk, swin_chunk_late, precision)

    def _spec(self):
        return _syn.swin_cls_state_dict_spec(self.cfg.swin)

    def forward_full(self, images_feature: torch.Tensor, gumbel: Optional[torch.Tensor] = None, want_feat: bool = False):
        """-> (logits, probs, importance[, feat]); probs = softmax((logits + gumbel)/tau), importance = sum_c p_c^2."""
        _require_cuda()
        if not self._finalized:
            raise _lib.FmmtError("load_state_dict() must be called before forward()")
        s = self.cfg.swin
        u8 = images_feature.dtype == torch.uint8
        if u8:
            # decoded crops (F, h, w, 3) uint8 as cv2.imread yields them: resized + normalised on the device
            # (utils/dataset.py:47-69), see fmmt_swin_forward_u8
            x = images_feature.to(device="cuda", non_blocking=True).contiguous()
            if x.dim() != 4 or x.shape[3] != 3:
                raise ValueError(f"uint8 crops must be (F, h, w, 3), got {tuple(x.shape)}")
        else:
            x = _f32(images_feature)
            if x.dim() != 4 or tuple(x.shape[1:]) != (s.in_chans, s.img_size, s.img_size):
                # PatchEmbed asserts the input size (Swin_Transformer.py:417-418)
                raise ValueError(f"Input image size {tuple(x.shape)} doesn't match model "
                                 f"(*,{s.in_chans},{s.img_size},{s.img_size}).")
        F = x.shape[0]
        g = _f32(gumbel) if gumbel is not None else None
        if g is not None and tuple(g.shape) != (F, s.num_labels):
            raise ValueError("gumbel noise must have shape (frames, num_labels)")
        logits = self._out("logits", (F, s.num_labels))
        probs = self._out("probs", (F, s.num_labels))
        imp = self._out("imp", (F,))
        feat = self._out("feat", (F, s.feat_dim)) if want_feat else None
        if u8:
            _lib.check(self._lib.fmmt_swin_forward_u8(self._h, _lib.ptr(x), F, int(x.shape[1]), int(x.shape[2]), _lib.ptr(g),
                                                      float(self.tau), _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(imp),
                                                      _lib.ptr(feat), _lib.cur_stream()), "fmmt_swin_forward_u8")
        else:
            _lib.check(self._lib.fmmt_swin_forward(self._h, _lib.ptr(x), F, _lib.ptr(g), float(self.tau), _lib.ptr(logits),
                                                   _lib.ptr(probs), _lib.ptr(imp), _lib.ptr(feat), _lib.cur_stream()),
                       "fmmt_swin_forward")
        return (logits, probs, imp, feat) if want_feat else (logits, probs, imp)

    def forward(self, images_feature=None, is_trg_task=None, labels=None, criterion=None, gumbel=None):
        if is_trg_task and gumbel is None:
            # the draw F.gumbel_softmax makes internally (torch/nn/functional.py): -log(Exp(1))
            gumbel = -torch.empty(images_feature.shape[0], self.num_labels, device="cuda").exponential_().log()
        logits, probs, _ = self.forward_full(images_feature, gumbel if is_trg_task else None)
        out = probs if is_trg_task else logits
        if labels is not None:
            return criterion(out, labels.to(out.device))
        return out


def _fmmt_config_from_namespace(config, text_layers: Optional[int] = None) -> FmmtConfig:
    """Fields read by the reference constructor (src/models.py:45-69,196-202)."""
    if isinstance(config, FmmtConfig):
        return config
    path = getattr(config, "pretrainedtextmodel_path", "roberta-large")
    kind = "roberta" if str(path).rstrip("/").split("/")[-1] == "roberta-large" else "bert"   # src/models.py:49-52
    t = TextConfig.roberta_large() if kind == "roberta" else TextConfig.bert_large()
    if text_layers is not None:
        t.layers = text_layers
    g = lambda n, d: getattr(config, n, d)  # noqa: E731
    f = FusionConfig(hidden=g("hidden_size", 768), heads=g("num_attention_heads", 12), ffn=g("intermediate_size", 3072),
                     eps=g("layer_norm_eps", 1e-12), audio_dim=g("audio_featExtr_dim", 768),
                     vision_dim=g("vision_featExtr_dim", 512), audio_layers=g("audio_utt_Transformernum", 5),
                     vision_layers=g("vision_utt_Transformernum", 2), cmt_layers_ta=g("crossmodal_layers_TA", 2),
                     cmt_heads_ta=g("crossmodal_num_heads_TA", 12), cmt_layers_tav=g("crossmodal_layers_TA_V", 2),
                     cmt_heads_tav=g("crossmodal_num_heads_TA_V", 12), text_len=g("get_text_utt_max_lens", 38),
                     audio_len=g("get_audio_utt_max_lens", 160), vision_len=g("get_vision_utt_max_lens", 160),
                     num_labels=g("num_labels", 7))
    return FmmtConfig(text=t, fusion=f)


class MultiModalTransformerForClassification(_Module):
    """src/models.py:41-188: same 8-argument forward, returns (U, num_labels) fp32 logits on the GPU."""

    _model_kind = _lib.MODEL_MULTIMODAL

    def __init__(self, config, text_layers: Optional[int] = None, precision: str = None):
        cfg = _fmmt_config_from_namespace(config, text_layers)
        self.choice_modality = getattr(config, "choice_modality", "T+A+V")
        self.num_labels = cfg.fusion.num_labels
        self.text_pretrained_model = cfg.text.kind
        super().__init__(cfg, precision=precision or getattr(config, "precision", "bf16"))

    def _spec(self):
        return _syn.multimodal_state_dict_spec(self.cfg)

    dedup_dialogues = True   # encode identical (ids, mask) rows of a batch once (host tensors only; result-identical)

    def forward(self, batch_text_input_ids=None, batch_text_input_mask=None, batch_text_sep_mask=None,
                audio_inputs=None, audio_mask=None, vision_inputs=None, new_vision_mask=None,
                batchUtt_in_dia_idx=None):
        _require_cuda()
        if not self._finalized:
            raise _lib.FmmtError("load_state_dict() must be called before forward()")
        f = self.cfg.fusion
        U = vision_inputs.shape[0]                                   # utt_batch_size (src/models.py:112)
        row_of_utt = None
        if (self.dedup_dialogues and U > 1 and not batch_text_input_ids.is_cuda and not batch_text_input_mask.is_cuda
                and tuple(batch_text_input_ids.shape) == tuple(batch_text_input_mask.shape)
                and batch_text_input_ids.shape[0] == U):
            # MELD encodes an utterance with its whole dialogue (src/meld_bert_extraText.py:65-130): consecutive utterances of
            # an eval batch share their ids/mask rows. Host tensors (what the DataLoader yields) are compared for free.
            both = torch.cat([batch_text_input_ids.to(torch.int64), batch_text_input_mask.to(torch.int64)], dim=1)
            uniq, inverse = torch.unique(both, dim=0, return_inverse=True)
            if uniq.shape[0] < U:
                Lr = batch_text_input_ids.shape[1]
                batch_text_input_ids, batch_text_input_mask = uniq[:, :Lr].contiguous(), uniq[:, Lr:].contiguous()
                row_of_utt = inverse.to(torch.int32).to("cuda", non_blocking=True)
        ids, msk, sep = _i64(batch_text_input_ids), _i64(batch_text_input_mask), _i64(batch_text_sep_mask)
        Ud = ids.shape[0]
        L = ids.shape[1]
        if isinstance(batchUtt_in_dia_idx, (list, tuple)):
            batchUtt_in_dia_idx = torch.tensor(list(batchUtt_in_dia_idx))
        idx = _i64(batchUtt_in_dia_idx)
        a, am, v, vm = _f32(audio_inputs), _f32(audio_mask), _f32(vision_inputs), _f32(new_vision_mask)
        if tuple(ids.shape) != (Ud, L) or tuple(msk.shape) != (Ud, L) or tuple(sep.shape) != (U, L) or \
                (row_of_utt is None and Ud != U):
            raise ValueError("text tensors must be (U, L)")
        if tuple(a.shape) != (U, f.audio_len, f.audio_dim) or tuple(am.shape) != (U, f.audio_len):
            raise ValueError(f"audio must be (U,{f.audio_len},{f.audio_dim}) with mask (U,{f.audio_len})")
        if tuple(v.shape) != (U, f.vision_len, f.vision_dim + f.num_labels) or tuple(vm.shape) != (U, f.vision_len):
            raise ValueError(f"vision must be (U,{f.vision_len},{f.vision_dim + f.num_labels}) with mask (U,{f.vision_len})")
        if idx.numel() != U:
            raise ValueError("batchUtt_in_dia_idx must have U entries")
        logits = self._out("logits", (U, f.num_labels))
        if row_of_utt is not None:
            _lib.check(self._lib.fmmt_multimodal_forward_dedup(self._h, _lib.ptr(ids), _lib.ptr(msk), Ud, _lib.ptr(row_of_utt),
                                                               _lib.ptr(sep), _lib.ptr(a), _lib.ptr(am), _lib.ptr(v),
                                                               _lib.ptr(vm), _lib.ptr(idx), U, L, _lib.ptr(logits),
                                                               _lib.cur_stream()), "fmmt_multimodal_forward_dedup")
        else:
            _lib.check(self._lib.fmmt_multimodal_forward(self._h, _lib.ptr(ids), _lib.ptr(msk), _lib.ptr(sep), _lib.ptr(a),
                                                         _lib.ptr(am), _lib.ptr(v), _lib.ptr(vm), _lib.ptr(idx), U, L,
                                                         _lib.ptr(logits), _lib.cur_stream()), "fmmt_multimodal_forward")
        return logits


class meld_utt_transformer(_Module):  # noqa: N801  (reference class name)
    """src/models.py:192-223 (the --choice_modality V model)."""

    _model_kind = _lib.MODEL_UNIMODAL

    def __init__(self, args, precision: str = None):
        cfg = _fmmt_config_from_namespace(args)
        super().__init__(cfg, precision=precision or getattr(args, "precision", "bf16"))

    def _spec(self):
        return _syn.unimodal_state_dict_spec(self.cfg.fusion)

    def forward(self, inputs=None, utt_mask=None):
        _require_cuda()
        if not self._finalized:
            raise _lib.FmmtError("load_state_dict() must be called before forward()")
        f = self.cfg.fusion
        x, m = _f32(inputs), _f32(utt_mask)
        U = x.shape[0]
        if tuple(x.shape) != (U, f.vision_len, f.vision_dim) or tuple(m.shape) != (U, f.vision_len):
            raise ValueError(f"inputs must be (U,{f.vision_len},{f.vision_dim}) with utt_mask (U,{f.vision_len})")
        logits = self._out("logits", (U, f.num_labels))
        _lib.check(self._lib.fmmt_unimodal_forward(self._h, _lib.ptr(x), _lib.ptr(m), U, _lib.ptr(logits),
                                                   _lib.cur_stream()), "fmmt_unimodal_forward")
        return logits


def frame_ingest(crops_u8: torch.Tensor) -> torch.Tensor:
    """utils/dataset.py:47-69 on the device: uint8 crops (F, h, w, 3) (cv2.imread bytes) -> fp32 (F, 3, 224, 224), bit-exact
    with the reference's cv2 (non-IPP) resize + ToTensor + Normalize. The model forwards fuse this step (pass the uint8
    crops straight to SwinForAffwildClassification); this entry exists for parity tests and for callers that want the tensor."""
    _require_cuda()
    lib = _lib.load()
    x = crops_u8.to(device="cuda").contiguous()
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
        raise ValueError("crops must be uint8 (F, h, w, 3)")
    out = torch.empty(x.shape[0], 3, 224, 224, device="cuda", dtype=torch.float32)
    _lib.check(lib.fmmt_op_frame_ingest(_lib.ptr(x), int(x.shape[0]), int(x.shape[1]), int(x.shape[2]), _lib.ptr(out),
                                        _lib.cur_stream()), "fmmt_op_frame_ingest")
    return out


def filter_pack(vision_inputs: torch.Tensor, vision_mask: torch.Tensor, num_imgs, probs: torch.Tensor,
                threshold: float = 0.2, per_utterance: bool = True, cache: Optional[dict] = None):
    """Device-side restatement of train.py:183-232 -> (vision519 (U,Lv,D+labels), new_mask (U,Lv)). `cache`: a dict in which
    the outputs / offsets are kept across calls of equal shape (graph mode: stable pointers for the fusion forward)."""
    _require_cuda()
    lib = _lib.load()
    v, m, p = _f32(vision_inputs), _f32(vision_mask), _f32(probs)
    U, Lv, D = v.shape
    labels = p.shape[1]
    n = [int(x) for x in (num_imgs.tolist() if torch.is_tensor(num_imgs) else num_imgs)]
    if len(n) != U or sum(n) != p.shape[0] or any(x < 0 or x > Lv for x in n):
        raise ValueError("num_imgs must have one entry per utterance, each <= Lv, summing to probs.shape[0]")
    ck = ("filter_pack", U, Lv, D, labels, tuple(n))
    if cache is not None and ck in cache:
        off, out_v, out_m, scratch = cache[ck]
    else:
        off = torch.tensor([0] + list(torch.tensor(n).cumsum(0).tolist()), dtype=torch.int32).to("cuda", non_blocking=True)
        out_v = torch.empty(U, Lv, D + labels, device="cuda", dtype=torch.float32)
        out_m = torch.empty(U, Lv, device="cuda", dtype=torch.float32)
        scratch = torch.zeros(1, device="cuda", dtype=torch.int32)
        if cache is not None:
            cache[ck] = (off, out_v, out_m, scratch)
    _lib.check(lib.fmmt_filter_pack(_lib.ptr(v), _lib.ptr(m), _lib.ptr(off), p.shape[0], _lib.ptr(p), float(threshold),
                                    int(per_utterance), _lib.ptr(out_v), _lib.ptr(out_m), _lib.ptr(scratch), U, Lv, D,
                                    labels, _lib.cur_stream()), "fmmt_filter_pack")
    return out_v, out_m
