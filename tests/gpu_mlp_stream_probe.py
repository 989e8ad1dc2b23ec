"""Streamed fused Swin MLP (fmmt_op_swin_mlp_stream) vs the un-fused chain LN -> fc1+GELU -> fc2+residual: parity + timing.
Not a pytest."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr

lib = _lib.load()
g = torch.Generator().manual_seed(1)
MB = int(os.environ.get("MBIG", "0"))
for C, Mbig in ((192, MB or 50176), (384, MB or 31360)):
    H = 4 * C
    gam, bet = (1 + 0.2 * torch.randn(C, generator=g)).cuda(), (0.2 * torch.randn(C, generator=g)).cuda()
    w1 = (torch.randn(H, C, generator=g) / math.sqrt(C)).cuda().to(torch.bfloat16).contiguous()
    w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).cuda().to(torch.bfloat16).contiguous()
    b1, b2 = (0.3 * torch.randn(H, generator=g)).cuda(), (0.3 * torch.randn(C, generator=g)).cuda()
    COPIES = int(os.environ.get("COPIES", "1"))
    w1r, w2r = w1.repeat(COPIES, 1).contiguous(), w2.repeat(COPIES, 1).contiguous()

    def unfused(x, h16, hid16):
        M = x.shape[0]
        check(lib.fmmt_op_layernorm(ptr(x), C, M, 1, C, None, 0, 0, ptr(gam), ptr(bet), 1e-5, None, 0, ptr(h16), C, cur_stream()))
        check(lib.fmmt_op_gemm(ptr(h16), C, ptr(w1), C, M, H, C, ptr(b1), 1, None, 0, None, 0, ptr(hid16), H, None, 0, 0, cur_stream()))
        check(lib.fmmt_op_gemm(ptr(hid16), H, ptr(w2), H, M, C, H, ptr(b2), 0, ptr(x), C, ptr(x), C, None, 0, None, 0, 0, cur_stream()))

    def fused(x):
        check(lib.fmmt_op_swin_mlp_stream(ptr(x), x.shape[0], C, ptr(gam), ptr(bet), 1e-5, ptr(w1r), C, ptr(b1), ptr(w2r), H,
                                          ptr(b2), COPIES, cur_stream()))

    for M in (1000, Mbig):
        x0 = (torch.randn(M, C, generator=g) * 2 + 0.5).cuda()
        xa, xb = x0.clone(), x0.clone()
        h16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
        hid16 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
        unfused(xa, h16, hid16)
        fused(xb)
        torch.cuda.synchronize()
        to = lib.fmmt_debug_timeout(1)
        bf = lambda t: t.to(torch.bfloat16).float()
        hh = bf(torch.nn.functional.layer_norm(x0, (C,), gam, bet, 1e-5))
        ref = x0 + (bf(torch.nn.functional.gelu(hh @ w1.float().t() + b1)) @ w2.float().t() + b2)
        print(f"C={C} M={M}: fused vs un-fused {(xa - xb).abs().max().item():.3e}, fused vs torch {(xb - ref).abs().max().item():.3e}, "
              f"un-fused vs torch {(xa - ref).abs().max().item():.3e} timeout=0x{to:x} finite={bool(torch.isfinite(xb).all())}", flush=True)
        if to:
            break
    M = Mbig
    xs = [(torch.randn(M, C, generator=g)).cuda() for _ in range(4)]
    h16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    hid16 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    def fused_pair(x):
        check(lib.fmmt_op_swin_mlp_pair(ptr(x), x.shape[0], C, ptr(gam), ptr(bet), 1e-5, ptr(w1), C, ptr(b1), ptr(w2), H, ptr(b2),
                                        cur_stream()))

    for name, fn in (("un-fused", lambda x: unfused(x, h16, hid16)), ("fused", fused), ("fused CTA-pair", fused_pair)):
        for x in xs:
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            for x in xs:
                fn(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 4.0 * M * C * H
        print(f"C={C} {name}: {ms * 1e3:.1f} us per {M}-row half-block, {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
    print("timeout", hex(lib.fmmt_debug_timeout(1)), flush=True)
