"""Quick GEMM timing probe (not a pytest): python tests/gpu_gemm_perf.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr

lib = _lib.load()
shapes = [(3136 * 64, 288, 96), (3136 * 64, 384, 96), (3136 * 64, 96, 384), (784 * 64, 768, 192),
          (196 * 64, 1536, 384), (196 * 64, 384, 1536), (8192, 8192, 8192), (1024, 3072, 1024), (1024, 4096, 1024),
          (1024, 1024, 4096)]
for (M, N, K) in shapes:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    def run():
        check(lib.fmmt_op_gemm(ptr(A), K, ptr(W), K, M, N, K, None, 0, None, 0, None, 0, ptr(o16), N, None, 0, 0,
                               cur_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10):
        torch.matmul(A, W.t(), out=o16)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / 10
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    print(f"M={M} N={N} K={K}: ours {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s {by/ms/1e6:.0f} GB/s | cublas {ms_t:.3f} ms {fl/ms_t/1e9:.1f} TFLOP/s")
