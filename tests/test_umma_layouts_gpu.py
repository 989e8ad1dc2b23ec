"""Pins the tcgen05 shared-memory operand layouts the fused attention kernel depends on (csrc/attn_fused.cu) against a
plain matmul, through the layout probe fmmt_debug_umma: K-major SWIZZLE_128B (the GEMM's layout), K-major SWIZZLE_64B
(32-channel k-blocks, 64-byte rows: Q / K / LN tile / weights), MN-major SWIZZLE_64B (V as [key][head_dim])."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def idesc(m, n, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def desc_tpl(layout, sbo, lbo=16):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def img_kmajor(x: torch.Tensor, row_bytes: int) -> np.ndarray:
    """[R, K] (K * 2 == row_bytes) -> swizzled image: 16-byte chunk c of row r lives at chunk c ^ f(r)."""
    R, K = x.shape
    assert K * 2 == row_bytes
    bits = bf16_bits(x)
    out = np.zeros(R * row_bytes // 2, dtype=np.uint16)
    nchunk = row_bytes // 16
    for r in range(R):
        sw = (r & 7) if row_bytes == 128 else ((r >> 1) & 3)
        for c in range(nchunk):
            dst = r * row_bytes // 2 + ((c ^ sw) * 8)
            out[dst:dst + 8] = bits[r, c * 8:c * 8 + 8]
    return out


def run(lib, a_img, b_img, adesc, bdesc, idsc, ksteps, a_step, b_step, ncols, a_off=0, b_off=0):
    from facialmmt_b200 import _lib
    a = torch.from_numpy(a_img.view(np.int16)).cuda()
    b = torch.from_numpy(b_img.view(np.int16)).cuda()
    out = torch.full((128, ncols), float("nan"), device="cuda")
    _lib.check(lib.fmmt_debug_umma(_lib.ptr(a), a.numel() * 2, _lib.ptr(b), b.numel() * 2, adesc, bdesc, a_off, b_off, idsc,
                                   ksteps, a_step, b_step, ncols, _lib.ptr(out)), "fmmt_debug_umma")
    return out.cpu()


def _rand(r, c, seed):
    return torch.randn(r, c, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16).float()


def test_kmajor_sw128_reference_layout(lib):
    A, B = _rand(128, 64, 1), _rand(64, 64, 2)
    got = run(lib, img_kmajor(A, 128), img_kmajor(B, 128), desc_tpl(2, 1024), desc_tpl(2, 1024), idesc(128, 64), 4, 2, 2, 64)
    assert (got - A @ B.t()).abs().max() < 1e-3


def test_kmajor_sw64(lib):
    """Q_h K_h^T shape: M = 128, N = 128, K = 32 (two k-steps inside a 64-byte row)."""
    A, B = _rand(128, 32, 3), _rand(128, 32, 4)
    got = run(lib, img_kmajor(A, 64), img_kmajor(B, 64), desc_tpl(4, 512), desc_tpl(4, 512), idesc(128, 128), 2, 2, 2, 128)
    err = (got - A @ B.t()).abs().max().item()
    print("K-major SW64 err", err)
    assert err < 1e-3


def test_kmajor_sw64_row_offset(lib):
    """B operand that starts 32 rows into a taller SW64 tile (per-head weight rows inside a resident weight image)."""
    A, Bfull = _rand(128, 32, 5), _rand(96, 32, 6)
    got = run(lib, img_kmajor(A, 64), img_kmajor(Bfull, 64), desc_tpl(4, 512), desc_tpl(4, 512), idesc(128, 32), 2, 2, 2, 32,
              b_off=32 * 64)
    assert (got - A @ Bfull[32:64].t()).abs().max() < 1e-3


def test_mnmajor_sw64_b_operand(lib):
    """P V shape: A = P [128 x 64 keys] K-major SW128; B = V [64 keys x 32 dims] stored key-major (MN-major operand),
    64-byte rows with the 64-byte swizzle; one k-step = 16 keys = 1024 bytes."""
    P, V = _rand(128, 64, 7), _rand(64, 32, 8)
    ref = P @ V
    ok = []
    for lbo, sbo in ((16, 512), (512, 512), (512, 16), (1024, 512), (512, 1024)):
        got = run(lib, img_kmajor(P, 128), img_kmajor(V, 64), desc_tpl(2, 1024), desc_tpl(4, sbo, lbo), idesc(128, 32, 0, 1),
                  4, 2, 64, 32)
        err = (got - ref).abs().max().item()
        print(f"MN-major SW64 B operand, LBO={lbo} SBO={sbo}: err {err:.3e}")
        ok.append(err < 1e-3)
    assert ok[0] and ok[1], "attn_fused.cu builds its V descriptor with SBO = 512"
