"""Fused Swin MLP (fmmt_op_swin_mlp) vs the un-fused chain LN -> fc1+GELU -> fc2+residual: parity + timing.
Not a pytest. Usage: python tests/gpu_mlp_probe.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr

lib = _lib.load()
C, H = 96, 384
g = torch.Generator().manual_seed(1)
gam, bet = (1 + 0.2 * torch.randn(C, generator=g)).cuda(), (0.2 * torch.randn(C, generator=g)).cuda()
w1 = torch.randn(H, C, generator=g) / math.sqrt(C)
w2 = torch.randn(C, H, generator=g) / math.sqrt(H)
b1, b2 = (0.3 * torch.randn(H, generator=g)).cuda(), (0.3 * torch.randn(C, generator=g)).cuda()
img = torch.empty(147456, dtype=torch.uint8, device="cuda")
check(lib.fmmt_op_swin_mlp_pack(w1.data_ptr(), w2.data_ptr(), ptr(img)))
w1b, w2b = w1.cuda().to(torch.bfloat16).contiguous(), w2.cuda().to(torch.bfloat16).contiguous()


def unfused(x, h16, hid16):
    M = x.shape[0]
    check(lib.fmmt_op_layernorm(ptr(x), C, M, 1, C, None, 0, 0, ptr(gam), ptr(bet), 1e-5, None, 0, ptr(h16), C, cur_stream()))
    check(lib.fmmt_op_gemm(ptr(h16), C, ptr(w1b), C, M, H, C, ptr(b1), 1, None, 0, None, 0, ptr(hid16), H, None, 0, 0, cur_stream()))
    check(lib.fmmt_op_gemm(ptr(hid16), H, ptr(w2b), H, M, C, H, ptr(b2), 0, ptr(x), C, ptr(x), C, None, 0, None, 0, 0, cur_stream()))


def fused(x):
    check(lib.fmmt_op_swin_mlp(ptr(x), x.shape[0], ptr(gam), ptr(bet), 1e-5, ptr(img), ptr(b1), ptr(b2), cur_stream()))


for M in (128, 1000, 148 * 128, 200704):
    x0 = (torch.randn(M, C, generator=g) * 2 + 0.5).cuda()
    xa, xb = x0.clone(), x0.clone()
    h16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    hid16 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    unfused(xa, h16, hid16)
    fused(xb)
    torch.cuda.synchronize()
    to = lib.fmmt_debug_timeout(1)
    print(f"M={M}: fused vs un-fused max abs diff {(xa - xb).abs().max().item():.3e} (|x| max {xa.abs().max().item():.2f}) "
          f"timeout=0x{to:x} finite={bool(torch.isfinite(xb).all())}", flush=True)
    if to:
        sys.exit(1)

# timing on a working set larger than L2: 4 x 64-frame chunks, rotated
M = 200704
xs = [(torch.randn(M, C, generator=g)).cuda() for _ in range(4)]
h16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
hid16 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
for name, fn in (("un-fused", lambda x: unfused(x, h16, hid16)), ("fused", fused)):
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
    print(f"{name}: {ms * 1e3:.1f} us per 200704-row half-block, {fl / ms / 1e9:.0f} TFLOP/s, "
          f"algorithmic {8 * M * C / ms / 1e6:.0f} GB/s (x read + x write)", flush=True)
print("timeout", hex(lib.fmmt_debug_timeout(1)))
