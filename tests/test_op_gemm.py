"""Parity of the tcgen05 GEMM (fmmt_op_gemm, C ABI) against a plain fp32 torch matmul of the same bf16 operands.

Tolerance: operands are identical bf16 values on both sides, accumulation is fp32 on both -> only summation order
differs: |err| <= 2e-3 * sqrt(K)/16 on O(1) data; bf16 outputs add one rounding (rel 2^-8).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(lib, M, N, K, bias=False, act=0, residual=False, out16=False, row_map=False, block_n=0, lda_pad=0,
         out32=True):
    from facialmmt_b200._lib import check, cur_stream, ptr
    g = torch.Generator(device="cpu").manual_seed(M * 7919 + N * 31 + K)
    lda = K + lda_pad
    lda = (lda + 7) // 8 * 8
    A = torch.zeros(M, lda, dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = torch.zeros(N, lda, dtype=torch.bfloat16)
    W[:, :K] = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g) if bias else None
    R = torch.randn(M, N, generator=g) if residual else None
    rm = None
    period = 0
    if row_map:
        period = 49 if M % 49 == 0 else M
        rm = torch.randperm(period, generator=g).to(torch.int32)
    dev = "cuda"
    Ad, Wd = A.to(dev), W.to(dev)
    bd = b.to(dev) if bias else None
    Rd = R.to(dev) if residual else None
    rmd = rm.to(dev) if row_map else None
    o32 = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32) if out32 else None
    o16 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16) if out16 else None
    check(lib.fmmt_op_gemm(ptr(Ad), lda, ptr(Wd), lda, M, N, K, ptr(bd), act, ptr(Rd), N, ptr(o32), N,
                           ptr(o16), N, ptr(rmd), period, block_n, cur_stream()), "fmmt_op_gemm")
    torch.cuda.synchronize()
    ref = A[:, :K].float() @ W[:, :K].float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    elif act == 3:
        ref = torch.tanh(ref)
    if row_map:
        idx = (torch.arange(M) // period) * period + rm.long()[torch.arange(M) % period]
        full = torch.empty_like(ref)
        full[idx] = ref
        ref = full
    if residual:
        ref = ref + R
    tol = 2e-3 * max(1.0, (K ** 0.5) / 16)
    err = 0.0
    if out32:
        got = o32.cpu()
        assert torch.isfinite(got).all(), "unwritten outputs"
        err = (got - ref).abs().max().item()
        assert err < tol, f"fp32 out max err {err} (tol {tol}) M={M} N={N} K={K}"
    if out16:
        got16 = o16.float().cpu()
        assert torch.isfinite(got16).all(), "unwritten bf16 outputs"
        err16 = (got16 - ref).abs().max().item()
        assert err16 < tol + 0.02 * ref.abs().max().item(), f"bf16 out max err {err16} M={M} N={N} K={K}"
    return err


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("M,N,K", [
    (128, 96, 96), (256, 288, 96), (3136, 96, 384), (784 * 2, 576, 192), (196 * 3, 1152, 384),
    (49 * 5, 2304, 768), (49 * 5, 768, 3072), (100, 512, 37632), (8 * 128, 3072, 1024), (1000, 1024, 4096),
    (77, 768, 519), (130, 96, 48), (1, 768, 768), (3136 * 8, 384, 96),
])
def test_gemm_shapes(lib, M, N, K, generic):
    """generic=False: TMA-epilogue fast path (fp32 out); generic=True: register-path epilogue (block_n = -1)."""
    _run(lib, M, N, K, block_n=-1 if generic else 0)
    _run(lib, M, N, K, block_n=-1 if generic else 0, bias=True, residual=True)
    if N % 8 == 0:
        _run(lib, M, N, K, block_n=-1 if generic else 0, bias=True, act=1, out16=True, out32=False)


@pytest.mark.parametrize("bn", [32, 64, 96, 128, 192, 256])
def test_gemm_block_n(lib, bn):
    _run(lib, 500, 768, 320, block_n=bn, bias=True)
    _run(lib, 500, 768, 320, block_n=-bn, bias=True)
    if bn <= 192:
        _run(lib, 500, 768, 320, block_n=bn, bias=True, residual=True)
    if bn % 64 == 0:
        _run(lib, 500, 768, 320, block_n=bn, bias=True, act=1, out16=True, out32=False)


def test_gemm_many_tiles_per_cta(lib):
    """Persistent loop with > 2 tiles per CTA: exercises ring wrap-around of every barrier (fast + generic)."""
    for bn in (0, -1):
        _run(lib, 128 * 148 * 3 + 77, 96, 96, block_n=bn, bias=True, residual=True)
        _run(lib, 128 * 148 * 2 + 5, 384, 96, block_n=bn, bias=True, act=1, out16=True, out32=False)


def test_gemm_epilogues(lib):
    _run(lib, 49 * 16, 384, 384, bias=True, act=1, out16=True)
    _run(lib, 49 * 16, 384, 384, bias=True, residual=True, row_map=True)
    _run(lib, 49 * 16, 96, 96, bias=True, residual=True, row_map=True, out16=True)
    _run(lib, 300, 768, 768, bias=True, act=3)
    _run(lib, 300, 64, 512, bias=True, act=2)
    _run(lib, 300, 40, 128, bias=True)  # ragged N (scalar tail)


def test_gemm_rejects_bad_alignment(lib):
    from facialmmt_b200._lib import cur_stream, ptr
    A = torch.zeros(8, 12, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(8, 12, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(8, 8, device="cuda")
    rc = lib.fmmt_op_gemm(ptr(A), 12, ptr(W), 12, 8, 8, 12, None, 0, None, 0, ptr(o), 8, None, 0, None, 0, 0,
                          cur_stream())
    assert rc != 0
    assert b"invalid" in lib.fmmt_last_error()


@pytest.mark.parametrize("N,K,res,out16", [(96, 96, True, False), (96, 48, False, False), (128, 96, True, False),
                                           (288, 96, False, True), (384, 96, False, True), (96, 384, True, False)])
def test_gemm_streaming_many_tiles(lib, N, K, res, out16):
    """Stage-1 Swin shapes at chunk scale (784 m-tiles, > 5 per CTA), repeated: epilogue-bound streaming where the
    residual/out rings and the slab->group assignment cycle many times (regression for a ring-parity aliasing bug)."""
    for _ in range(4):
        _run(lib, 100352, N, K, bias=True, residual=res, out16=out16, out32=not out16, act=1 if out16 else 0)


@pytest.mark.parametrize("M,N,K,bn", [
    (256, 256, 64, 256), (256, 128, 256, 128), (300, 192, 256, 192), (128, 256, 384, 256), (1000, 384, 384, 0),
    (2048, 1536, 384, 0), (196 * 16, 1152, 384, 0), (196 * 16, 384, 1536, 0), (49 * 40, 3072, 768, 0),
    (1024, 4096, 1024, 0), (1030, 96, 512, 0), (31360, 1536, 384, 0), (5000, 200, 320, 0),
])
def test_gemm_cta_pair(lib, M, N, K, bn):
    """CTA-pair kernel (tcgen05 cta_group::2, 256-row tiles, block_n code 1000 + bn): same contract as the single-CTA
    kernels, incl. ragged M (second CTA of the last pair partly or wholly out of range), ragged N and K % 64 != 0."""
    code = 1000 + bn
    _run(lib, M, N, K, block_n=code)
    _run(lib, M, N, K, block_n=code, bias=True, residual=True)
    if N % 8 == 0:
        _run(lib, M, N, K, block_n=code, bias=True, act=1, out16=True, out32=False)
    assert lib.fmmt_debug_timeout(1) == 0, "pipeline wait timed out inside the CTA-pair GEMM"


def test_gemm_cta_pair_matches_single(lib):
    """same operands through both kernels: fp32 accumulation order over K is identical (k-blocks in order), so the
    results must agree bit for bit"""
    from facialmmt_b200._lib import check, cur_stream, ptr
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 3000, 768, 1536
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    outs = []
    for code in (999, 1000):
        o = torch.empty(M, N, device="cuda")
        check(lib.fmmt_op_gemm(ptr(A), K, ptr(W), K, M, N, K, None, 0, None, 0, ptr(o), N, None, 0, None, 0, code,
                               cur_stream()))
        outs.append(o)
    torch.cuda.synchronize()
    assert lib.fmmt_debug_timeout(1) == 0
    assert torch.equal(outs[0], outs[1])


def test_gemm_inplace_residual_stress(lib):
    """out_f32 == residual (how the engine applies every shortcut): the epilogue reads the TMA-loaded residual slab through
    the generic proxy and then lets the async proxy refill it; without a proxy fence before the release a late read saw the
    next slab's bytes in ~1 of 3 launches at these shapes (rows of one 16-byte chunk off by a whole residual)."""
    from facialmmt_b200._lib import check, cur_stream, ptr
    g = torch.Generator(device="cpu").manual_seed(3)
    N, K = 384, 1536
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    for M in (148 * 128, 31360):
        for rep in range(12):
            A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
            x = torch.randn(M, N, generator=g).cuda()
            y = x.clone()
            check(lib.fmmt_op_gemm(ptr(A), K, ptr(W), K, M, N, K, ptr(b), 0, ptr(y), N, ptr(y), N, None, 0, None, 0, 0,
                                   cur_stream()))
            torch.cuda.synchronize()
            ref = x + A.float() @ W.float().t() + b
            err = (y - ref).abs().max().item()
            assert err < 2e-2, f"M={M} rep={rep}: max err {err}"


@pytest.mark.parametrize("M,N,K", [(3136, 96, 48), (128 * 148 * 5 + 77, 96, 48), (1000, 256, 192), (260, 32, 64)])
def test_gemm_layernorm_epilogue(lib, M, N, K):
    """Linear + LayerNorm in one kernel (fmmt_op_gemm_ln; PatchEmbed = Conv2d(3,96,k4,s4) as a K = 48 GEMM + norm_layer(96),
    Swin_Transformer.py:402-412): same bf16 operands, fp32 accumulate, fp32 two-pass LayerNorm on both sides."""
    from facialmmt_b200._lib import check, cur_stream, ptr
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    gam = 1.0 + 0.2 * torch.randn(N, generator=g)
    bet = 0.1 * torch.randn(N, generator=g)
    ref = torch.nn.functional.layer_norm(A.float() @ W.float().t() + b, (N,), gam, bet, 1e-5)
    Ad, Wd, bd, gd, btd = (t.cuda() for t in (A, W, b, gam, bet))
    out = torch.full((M, N), float("nan"), device="cuda")
    first = None
    for rep in range(2):
        out.fill_(float("nan"))
        check(lib.fmmt_op_gemm_ln(ptr(Ad), K, ptr(Wd), K, M, N, K, ptr(bd), ptr(gd), ptr(btd), 1e-5, ptr(out), N, cur_stream()),
              "fmmt_op_gemm_ln")
        torch.cuda.synchronize()
        if rep == 0:
            first = out.clone()
    assert lib.fmmt_debug_timeout(1) == 0
    got = out.cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    print(f"\ngemm+LN {M}x{N}x{K}: err {err:.3e}")
    assert err < 2e-3                      # unit-variance rows: the accumulation-order noise of the GEMM, rescaled by 1/std
    assert torch.equal(first, out)
