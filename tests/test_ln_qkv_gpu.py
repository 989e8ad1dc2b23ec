"""LayerNorm-1 + roll/window gather fused into the qkv Linear (csrc/ln_qkv.cu, C = 192 / 384) through the C ABI against an fp32
torch restatement of Swin_Transformer.py:238-247 (norm1, roll, window_partition) + :119 (qkv Linear). The weights are the
same bf16 values on both sides; the normalised rows are bf16 operands of an fp32-accumulating MMA, so the bar is relative to
the output range. The gathered fp32 rows (the block's residual stream in window order) must be bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_case(lib, C, frames, R, shift, ragged=0, seed=0):
    from facialmmt_b200 import _lib
    from test_attn_fused_gpu import window_maps
    T = R * R
    M = frames * T - ragged          # ragged > 0 only without a gather (M % T == 0 is required with one)
    N = 3 * C
    g = torch.Generator().manual_seed(200 + seed)
    x = torch.randn(M, C, generator=g) * 1.5 + 0.3
    ln_g = 1.0 + 0.2 * torch.randn(C, generator=g)
    ln_b = 0.1 * torch.randn(C, generator=g)
    w = (torch.randn(N, C, generator=g) * 0.1).to(torch.bfloat16)
    b = torch.randn(N, generator=g) * 0.2
    gather = window_maps(R, shift)[0] if shift is not None else None
    xg = x if gather is None else x.view(frames, T, C)[:, gather.long()].reshape(-1, C)
    ref = torch.nn.functional.layer_norm(xg, (C,), ln_g, ln_b, 1e-5) @ w.float().t() + b

    xd, wd = x.cuda(), w.cuda()
    gd = gather.cuda() if gather is not None else None
    raw = torch.full((M, C), float("nan"), device="cuda") if gather is not None else None
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    vec = [t.cuda() for t in (ln_g, ln_b, b)]
    first = None
    for rep in range(2):             # second launch: bit-repeatable
        out.fill_(float("nan"))
        _lib.check(lib.fmmt_op_ln_qkv(_lib.ptr(xd), _lib.ptr(raw), M, C, T, _lib.ptr(gd), _lib.ptr(vec[0]), _lib.ptr(vec[1]), 1e-5,
                                      _lib.ptr(wd), C, _lib.ptr(vec[2]), N, _lib.ptr(out), N, 0, _lib.cur_stream()), "fmmt_op_ln_qkv")
        torch.cuda.synchronize()
        if rep == 0:
            first = out.clone()
    assert lib.fmmt_debug_timeout(1) == 0
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"\nln_qkv C={C} M={M} shift={shift}: err {err:.3e} / range {scale:.2f} = {err / scale:.2e}")
    assert err / scale < 1e-2
    assert torch.equal(first, out)
    if raw is not None:
        assert torch.equal(raw.cpu(), xg)


@pytest.mark.parametrize("C,R", [(192, 28), (384, 14)])
@pytest.mark.parametrize("frames,shift", [(1, 0), (3, 3), (21, 3)])
def test_ln_qkv_window_gather(lib, C, R, frames, shift):
    run_case(lib, C, frames, R, shift, seed=C + frames + shift)


@pytest.mark.parametrize("C,R", [(192, 28), (384, 14)])
def test_ln_qkv_identity_ragged_multiwave(lib, C, R):
    # no gather: M not a multiple of 128 (the last tile is clipped by the tensor map) and more tiles than SMs
    frames = 148 * 128 * 2 // (R * R) + 3
    run_case(lib, C, frames, R, None, ragged=37, seed=C)
