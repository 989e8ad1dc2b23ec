"""Single-CTA GEMM timing on the shapes of the path with the epilogues the path uses. Not a pytest.
   FMMT_KBS=1 / 2 in the environment selects 2-D / forced 3-D (two k-blocks per TMA instruction) operand loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr

lib = _lib.load()
# (M, N, K, kind): kind 0 = bf16 out (qkv), 1 = bf16 out + GELU (fc1), 2 = fp32 out + residual (proj / fc2)
shapes = [(31360, 1152, 384, 0), (31360, 1536, 384, 1), (31360, 384, 1536, 2), (31360, 384, 384, 2),
          (7840, 2304, 768, 0), (7840, 3072, 768, 1), (7840, 768, 3072, 2), (7840, 768, 768, 2),
          (50176, 576, 192, 0), (50176, 768, 192, 1), (50176, 192, 768, 2), (50176, 192, 192, 2),
          (1024, 3072, 1024, 0), (1024, 4096, 1024, 1), (1024, 1024, 4096, 2), (1024, 1024, 1024, 2)]
tot = 0.0
for (M, N, K, kind) in shapes:
    nset = 3
    As = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    o16 = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nset)]
    o32 = [torch.randn(M, N, device="cuda") for _ in range(nset)] if kind == 2 else None
    def run(i):
        j = i % nset
        if kind == 2:
            check(lib.fmmt_op_gemm(ptr(As[j]), K, ptr(W), K, M, N, K, ptr(b), 0, ptr(o32[j]), N, ptr(o32[j]), N, None, 0,
                                   None, 0, 999, cur_stream()))
        else:
            check(lib.fmmt_op_gemm(ptr(As[j]), K, ptr(W), K, M, N, K, ptr(b), kind, None, 0, None, 0, ptr(o16[j]), N,
                                   None, 0, 999, cur_stream()))
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    tot += ms
    print(f"M={M} N={N} K={K} kind={kind}: {ms * 1e3:.1f}us {2.0 * M * N * K / ms / 1e9:.0f}TF", flush=True)
print(f"sum {tot * 1e3:.1f}us  timeout {hex(lib.fmmt_debug_timeout(1))}")
