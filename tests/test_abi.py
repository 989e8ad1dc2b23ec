"""CPU: the C-ABI library builds/loads here (nvcc cross-compiles without a GPU) and exports exactly the symbols that
include/facialmmt_b200.h declares; the ctypes declarations cover every one of them. No compute call is made."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "facialmmt_b200.h")).read()
    return sorted(set(re.findall(r"FMMT_API\s+[\w\s\*]+?\b(fmmt_\w+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for must in ("fmmt_create", "fmmt_load_weight", "fmmt_finalize", "fmmt_swin_forward", "fmmt_filter_pack",
                 "fmmt_multimodal_forward", "fmmt_unimodal_forward", "fmmt_op_gemm", "fmmt_op_layernorm",
                 "fmmt_op_window_attention", "fmmt_op_mha", "fmmt_last_error", "fmmt_destroy"):
        assert must in syms


def test_library_exports_every_header_symbol_and_ctypes_covers_them():
    from facialmmt_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _header_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported by the .so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes declaration in facialmmt_b200/_lib.py"
    for name in _lib.SIGNATURES:
        assert name in _header_symbols(), f"{name} bound in _lib.py but missing from the header"
    assert lib.fmmt_version().startswith(b"facialmmt_b200")
    assert lib.fmmt_launch_count() == 0 or lib.fmmt_launch_count() > 0


def test_config_struct_matches_header_field_order():
    from facialmmt_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "facialmmt_b200.h")).read()
    body = txt[txt.index("typedef struct fmmt_config {"): txt.index("} fmmt_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body):
        for f in decl.split(","):
            fields.append(re.sub(r"\[\d+\]", "", f).strip())
    assert fields == [n for n, _ in _lib.FmmtConfigC._fields_]
    # 4-byte fields only, arrays of 4: size must be 4 * (scalars + 2*4 - 2)
    assert ctypes.sizeof(_lib.FmmtConfigC) == 4 * (len(fields) + 6)


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when there is no CUDA device (never route through the oracle)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200._lib import FmmtError
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.models import meld_utt_transformer
    cfg = FmmtConfig()
    m = meld_utt_transformer(cfg)
    with pytest.raises(FmmtError):
        m.load_state_dict(syn.unimodal_stress_state_dict(cfg.fusion, 1))
    with pytest.raises(FmmtError):
        m(torch.zeros(1, 160, 512), torch.ones(1, 160))
