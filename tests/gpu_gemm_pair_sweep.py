"""Pipeline-shape sweep of the CTA-pair GEMM (env overrides are read once per process -> one subprocess per config)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from facialmmt_b200 import _lib
    from facialmmt_b200._lib import check, cur_stream, ptr
    lib = _lib.load()
    out = []
    for (M, N, K) in [(31360, 1536, 384), (31360, 384, 1536), (7840, 3072, 768), (8192, 8192, 8192)]:
        As = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(3)]
        W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        os16 = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
        def run(i):
            check(lib.fmmt_op_gemm(ptr(As[i % 3]), K, ptr(W), K, M, N, K, None, 0, None, 0, None, 0, ptr(os16[i % 3]), N,
                                   None, 0, 1256, cur_stream()))
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
        ref = (As[2][:64].float() @ W.float().t())
        err = (os16[2][:64].float() - ref).abs().max().item() / ref.abs().max().item()
        out.append(f"{M}x{N}x{K} {ms * 1e3:.1f}us {2.0 * M * N * K / ms / 1e9:.0f}TF err{err:.1e}")
    print("  ".join(out), "timeout", hex(lib.fmmt_debug_timeout(1)), flush=True)
else:
    for kbs in (1, 2):
        for st in (2, 3, 4, 5, 6):
            for gr in (2,):
                env = dict(os.environ, FMMT_PAIR_KBS=str(kbs), FMMT_PAIR_STAGES=str(st), FMMT_PAIR_GROUPS=str(gr))
                r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=120)
                print(f"kbs={kbs} stages<={st} groups={gr}: {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
