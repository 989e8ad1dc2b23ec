"""ctypes binding of libfacialmmt_b200.so -- the only way the Python host side reaches the CUDA kernels.

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError (FmmtError) is raised.
The declarations below mirror include/facialmmt_b200.h one to one (tests/test_abi.py checks every symbol).
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "libfacialmmt_b200.so"

MODEL_SWIN_CLS, MODEL_MULTIMODAL, MODEL_UNIMODAL = 1, 2, 3
TEXT_ROBERTA, TEXT_BERT = 0, 1
PRECISION_BF16, PRECISION_FP32 = 0, 1


class FmmtError(RuntimeError):
    pass


class FmmtConfigC(Structure):
    _fields_ = [
        ("model", c_int32),
        ("img_size", c_int32), ("patch_size", c_int32), ("in_chans", c_int32), ("embed_dim", c_int32),
        ("num_stages", c_int32), ("depths", c_int32 * 4), ("num_heads", c_int32 * 4),
        ("window_size", c_int32), ("mlp_ratio", c_int32),
        ("feat_dim", c_int32), ("head_hidden", c_int32), ("num_labels", c_int32),
        ("swin_chunk", c_int32), ("swin_chunk_late", c_int32),
        ("text_kind", c_int32), ("vocab_size", c_int32), ("text_hidden", c_int32), ("text_layers", c_int32),
        ("text_heads", c_int32), ("text_ffn", c_int32), ("max_pos", c_int32), ("type_vocab", c_int32),
        ("pad_id", c_int32), ("text_eps", c_float),
        ("hidden", c_int32), ("heads", c_int32), ("ffn", c_int32), ("audio_dim", c_int32), ("vision_dim", c_int32),
        ("audio_layers", c_int32), ("vision_layers", c_int32),
        ("cmt_layers_ta", c_int32), ("cmt_heads_ta", c_int32), ("cmt_layers_tav", c_int32), ("cmt_heads_tav", c_int32),
        ("text_len", c_int32), ("audio_len", c_int32), ("vision_len", c_int32),
        ("eps", c_float),
        ("precision", c_int32),
    ]


# name -> (restype, argtypes); the single source the ABI test compares with the header
SIGNATURES = {
    "fmmt_last_error": (c_char_p, []),
    "fmmt_version": (c_char_p, []),
    "fmmt_launch_count": (c_int64, []),
    "fmmt_create": (c_int, [POINTER(FmmtConfigC), POINTER(c_void_p)]),
    "fmmt_destroy": (None, [c_void_p]),
    "fmmt_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "fmmt_finalize": (c_int, [c_void_p]),
    "fmmt_swin_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "fmmt_swin_forward_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "fmmt_op_frame_ingest": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fmmt_filter_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "fmmt_multimodal_forward": (c_int, [c_void_p] * 9 + [c_int, c_int, c_void_p, c_void_p]),
    "fmmt_multimodal_forward_dedup": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fmmt_unimodal_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "fmmt_check": (c_int, [c_void_p]),
    "fmmt_set_graph": (c_int, [c_void_p, c_int]),
    "fmmt_set_capture": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "fmmt_set_profile": (c_int, [c_void_p, c_int]),
    "fmmt_profile_read": (c_int64, [c_void_p, c_void_p, c_int64]),
    "fmmt_debug_timeout": (ctypes.c_uint32, [c_int]),
    "fmmt_debug_umma": (c_int, [c_void_p, c_int, c_void_p, c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                ctypes.c_uint32, ctypes.c_uint32, c_int, c_int, c_int, c_int, c_void_p]),
    "fmmt_debug_mma_cycles": (c_double, [c_int, c_int]),
    "fmmt_debug_feed": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fmmt_debug_feed2": (c_double, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "fmmt_flops": (c_double, [c_void_p, c_int]),
    "fmmt_device_bytes": (c_int64, [c_void_p]),
    "fmmt_op_gemm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                             c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "fmmt_op_gemm_ln": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p, c_int, c_void_p]),
    "fmmt_op_layernorm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_float, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "fmmt_op_window_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_float, c_void_p]),
    "fmmt_op_swin_mlp_pack": (c_int, [c_void_p, c_void_p, c_void_p]),
    "fmmt_op_swin_mlp": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fmmt_op_swin_attn_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fmmt_op_swin_attn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "fmmt_op_swin_mlp_stream": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                        c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "fmmt_op_swin_mlp_pair": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                      c_void_p, c_int, c_void_p, c_void_p]),
    "fmmt_op_ln_qkv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
                               c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "fmmt_op_span_extract": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "fmmt_op_mha": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_float,
                            c_int, c_int, c_int, c_int, c_float, c_void_p]),
}


def _declare(lib):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args


def load(build_if_missing: bool = True):
    """Load the shared library (building it first if absent and nvcc is available). Raises if it cannot."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise FmmtError(f"{LIB_PATH} is missing; run `python -m facialmmt_b200.build`")
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(str(LIB_PATH))
    _declare(lib)
    _LIB = lib
    return lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().fmmt_last_error().decode(errors="replace")
        raise FmmtError(f"{what} failed with code {code}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor, or None."""
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
