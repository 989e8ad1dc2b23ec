"""Fused Swin attention half-block (csrc/attn_fused.cu, C = 96 / 3 heads / 7x7 windows) through the C ABI against an fp32
torch restatement of Swin_Transformer.py:238-264 + :113-143: x_out = x[g] + proj(W-MSA(LN(x[g]))). bf16 operands, fp32
accumulate / softmax / residual: errors are stated relative to the output range (same bar as the un-fused kernels)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

C, H, N, WS = 96, 3, 49, 7


def window_maps(R, shift):
    """gather [T] (window-order row -> natural token, after torch.roll(-shift) + window_partition), rid [nW, 49], wflag [nW]"""
    nw = R // WS
    g, rid = [], []
    region = lambda p: 0 if p < R - WS else (1 if p < R - shift else 2)   # noqa: E731  (Swin_Transformer.py:208-229)
    for wy in range(nw):
        for wx in range(nw):
            for ty in range(WS):
                for tx in range(WS):
                    hh, ww = (wy * WS + ty + shift) % R, (wx * WS + tx + shift) % R
                    g.append(hh * R + ww)
                    rid.append(3 * region(wy * WS + ty) + region(wx * WS + tx))
    rid = torch.tensor(rid, dtype=torch.int8).view(nw * nw, N)
    wflag = (rid != rid[:, :1]).any(1).to(torch.int8)
    return torch.tensor(g, dtype=torch.int32), rid, wflag


def reference(x, T, gather, ln_g, ln_b, wqkv, bqkv, wproj, bproj, table, rid):
    F = x.shape[0] // T
    xg = x.view(F, T, C)[:, gather.long()].reshape(-1, C)
    h = torch.nn.functional.layer_norm(xg, (C,), ln_g, ln_b, 1e-5)
    qkv = (h @ wqkv.t() + bqkv).view(-1, N, 3, H, 32).permute(2, 0, 3, 1, 4)      # (3, B_, heads, N, 32)
    q, k, v = qkv[0] * (32 ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    coords = torch.stack(torch.meshgrid(torch.arange(WS), torch.arange(WS), indexing="ij")).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    idx = (rel[0] + WS - 1) * (2 * WS - 1) + rel[1] + WS - 1
    attn = attn + table[idx.view(-1)].view(N, N, H).permute(2, 0, 1)[None]
    if rid is not None:
        nW = rid.shape[0]
        r = rid.long()
        mask = torch.where(r[:, :, None] != r[:, None, :], -100.0, 0.0)          # (nW, 49, 49)
        attn = attn.view(-1, nW, H, N, N) + mask[None, :, None]
        attn = attn.view(-1, H, N, N)
    attn = attn.softmax(-1)
    o = (attn @ v).transpose(1, 2).reshape(-1, C)
    return xg + o @ wproj.t() + bproj


@pytest.mark.parametrize("frames,shift", [(1, 0), (1, 3), (5, 3), (3, 0)])
def test_fused_attention_half_block(lib, frames, shift):
    from facialmmt_b200 import _lib
    R = 56
    T = R * R
    M = frames * T
    g = torch.Generator().manual_seed(100 + frames + shift)
    x = torch.randn(M, C, generator=g) * 1.5 + 0.3
    ln_g = 1.0 + 0.2 * torch.randn(C, generator=g)
    ln_b = 0.1 * torch.randn(C, generator=g)
    wqkv = torch.randn(3 * C, C, generator=g) * 0.15
    bqkv = torch.randn(3 * C, generator=g) * 0.2
    wproj = torch.randn(C, C, generator=g) * 0.1
    bproj = torch.randn(C, generator=g) * 0.1
    table = torch.randn(169, H, generator=g)
    gather, rid, wflag = window_maps(R, shift)
    ref = reference(x, T, gather, ln_g, ln_b, wqkv, bqkv, wproj, bproj, table, rid if shift else None)

    img = torch.empty(73728, dtype=torch.uint8, device="cuda")
    tab = torch.empty(507, device="cuda")
    wq, wp, tb = wqkv.contiguous(), wproj.contiguous(), table.contiguous()
    _lib.check(lib.fmmt_op_swin_attn_pack(_lib.ptr(wq), _lib.ptr(wp), _lib.ptr(tb), _lib.ptr(img), _lib.ptr(tab)), "pack")
    xd = x.cuda()
    out = torch.full((M, C), float("nan"), device="cuda")
    gd, rd, wd = gather.cuda(), rid.cuda(), wflag.cuda()
    args = [t.cuda() for t in (ln_g, ln_b, bqkv, bproj)]
    for rep in range(2):      # second launch: bit-repeatable
        out.fill_(float("nan"))
        _lib.check(lib.fmmt_op_swin_attn(_lib.ptr(xd), _lib.ptr(out), M, T, _lib.ptr(gd), _lib.ptr(args[0]), _lib.ptr(args[1]),
                                         1e-5, _lib.ptr(img), _lib.ptr(tab), _lib.ptr(args[2]), _lib.ptr(args[3]),
                                         _lib.ptr(rd) if shift else None, _lib.ptr(wd) if shift else None, 64,
                                         _lib.cur_stream()), "fmmt_op_swin_attn")
        torch.cuda.synchronize()
        if rep == 0:
            first = out.clone()
    got = out.cpu()
    assert torch.isfinite(got).all()
    xg = x.view(frames, T, C)[:, gather.long()].reshape(-1, C)
    delta_ref, delta_got = ref - xg, got - xg                                     # the attention branch alone
    err = (delta_got - delta_ref).abs().max().item()
    scale = delta_ref.abs().max().item()
    print(f"\nfused attention frames={frames} shift={shift}: branch err {err:.3e} / range {scale:.2f} = {err / scale:.2e}")
    assert err / scale < 2e-2
    assert torch.equal(first, out)
