"""Timing probe for the GEMM shapes of one bench step (not a pytest)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr
lib = _lib.load()
shapes = [(200704, 288, 96, 0, 1, 0), (200704, 384, 96, 0, 1, 1), (200704, 96, 96, 1, 0, 0), (200704, 96, 384, 1, 0, 0),
          (50176, 576, 192, 0, 1, 0), (50176, 768, 192, 0, 1, 1), (50176, 192, 768, 1, 0, 0),
          (31360, 1152, 384, 0, 1, 0), (31360, 1536, 384, 0, 1, 1), (31360, 384, 1536, 1, 0, 0), (31360, 384, 384, 1, 0, 0),
          (7840, 3072, 768, 0, 1, 1), (7840, 768, 3072, 1, 0, 0), (1024, 4096, 1024, 0, 1, 1), (1024, 1024, 4096, 1, 0, 0)]
tot = 0
sweep = os.environ.get("SWEEP", "0") == "1"
for (M, N, K, res, o16, act) in shapes:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    o32 = None if o16 else torch.empty(M, N, device="cuda")
    ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if o16 else None
    def timeit(bn):
        def run():
            check(lib.fmmt_op_gemm(ptr(A), K, ptr(W), K, M, N, K, ptr(b), act, ptr(R), N, ptr(o32), N, ptr(ob), N, None, 0, bn, cur_stream()))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20
    ms = timeit(0)
    if sweep:
        gran = 64 if o16 else 32
        res_s = []
        for bn in (32, 64, 96, 128, 160, 192, 224, 256):
            if bn % gran: continue
            try:
                res_s.append((bn, timeit(bn) * 1000))
            except Exception as e:
                res_s.append((bn, float("nan")))
        print("   sweep block_n:", ", ".join(f"{bn}:{t:.1f}" for bn, t in res_s))
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K) + (2 if o16 else 4) * M * N + (4 * M * N if res else 0)
    tot += ms
    print(f"M={M:6d} N={N:4d} K={K:4d} res={res} bf16out={o16} act={act}: {ms*1000:7.1f} us {fl/ms/1e9:7.1f} TF/s {by/ms/1e6:6.0f} GB/s")
print("sum", round(tot * 1000, 1), "us")
