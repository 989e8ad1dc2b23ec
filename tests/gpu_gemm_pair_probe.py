"""CTA-pair GEMM vs single-CTA GEMM vs cuBLAS: timing on the shapes of the path. Not a pytest."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr

lib = _lib.load()
shapes = [(31360, 1152, 384), (31360, 1536, 384), (31360, 384, 1536), (31360, 384, 384), (7840, 2304, 768),
          (7840, 3072, 768), (7840, 768, 3072), (50176, 768, 192), (50176, 192, 768), (1024, 4096, 1024),
          (1024, 1024, 4096), (1024, 3072, 1024), (8192, 8192, 8192)]
for (M, N, K) in shapes:
    # several operand sets, rotated, so that the working set exceeds L2 for the big shapes
    nset = 3
    As = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    os16 = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nset)]
    res = {}
    for name, code in (("single", 999), ("pair", 1000), ("pair256", 1256), ("pair192", 1192), ("pair128", 1128)):
        if code > 1000 and N < (code - 1000):
            continue
        def run(i):
            check(lib.fmmt_op_gemm(ptr(As[i % nset]), K, ptr(W), K, M, N, K, None, 0, None, 0, None, 0, ptr(os16[i % nset]), N,
                                   None, 0, code, cur_stream()))
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        if lib.fmmt_debug_timeout(1):
            print(f"M={M} N={N} K={K} {name}: TIMEOUT", flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(12):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        torch.matmul(As[i % nset], W.t(), out=os16[i % nset])
    e0.record()
    for i in range(12):
        torch.matmul(As[i % nset], W.t(), out=os16[i % nset])
    e1.record()
    torch.cuda.synchronize()
    res["cublas"] = e0.elapsed_time(e1) / 12
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: " + "  ".join(f"{k} {v * 1e3:.1f}us {fl / v / 1e9:.0f}TF" for k, v in res.items()), flush=True)
