"""ctypes binding of libfacialmmt_b200.so -- the only way the Python host side reaches the CUDA kernels.

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "libfacialmmt_b200.so"


class FmmtError(RuntimeError):
    pass


def _declare(lib):
    lib.fmmt_last_error.restype = c_char_p
    lib.fmmt_last_error.argtypes = []
    lib.fmmt_version.restype = c_char_p
    lib.fmmt_version.argtypes = []
    lib.fmmt_launch_count.restype = c_int64
    lib.fmmt_launch_count.argtypes = []
    lib.fmmt_op_gemm.restype = c_int
    lib.fmmt_op_gemm.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                 c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc is available). Raises if it cannot."""
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
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
