"""Bench probe (not a test): timing + per-tile stamp trace of the LN + qkv kernel at the bench geometry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_attn_fused_gpu import window_maps

lib = _lib.load()
for C, R, frames in ((384, 14, 1280), (192, 28, 320)):
    T = R * R
    M, N = frames * T, 3 * C
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, C, generator=g).cuda()
    raw = torch.empty(M, C, device="cuda")
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    w = (torch.randn(N, C, generator=g) * 0.1).to(torch.bfloat16).cuda()
    vec = [torch.ones(C).cuda(), torch.zeros(C).cuda(), torch.zeros(N).cuda()]
    gd = window_maps(R, 3)[0].cuda()

    DBG = int(os.environ.get('DBG', '0'))

    def run(trace=None):
        _lib.check(lib.fmmt_op_ln_qkv(_lib.ptr(x), _lib.ptr(raw), M, C, T, _lib.ptr(gd), _lib.ptr(vec[0]), _lib.ptr(vec[1]), 1e-5,
                                      _lib.ptr(w), C, _lib.ptr(vec[2]), N, _lib.ptr(out), N, (1 if trace is not None else 0) | (DBG << 1),
                                      _lib.ptr(trace) if trace is not None else _lib.cur_stream()), "ln_qkv")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = (M + 127) // 128
    print(f"C={C} M={M}: {ms:.3f} ms/launch, {ms * 1e3 / (tiles / 148):.2f} us per tile per SM, "
          f"{(8 * M * C + 2 * M * N) / ms / 1e6:.0f} GB/s algorithmic, {2 * M * N * C / ms / 1e9:.0f} TFLOP/s")
    tr = torch.zeros(8 * 32, dtype=torch.int64, device="cuda")
    run(tr); torch.cuda.synchronize()
    t = tr.cpu().view(8, 32)
    base = t[0, 0].item()
    names = {0: "ln:start", 1: "ln:pass0", 2: "ln:pass1", 3: "ln:pass2", 4: "ln:pass3", 5: "ln:a_empty", 6: "ln:a_full",
             7: "mma:start", 8: "mma:a_full", 27: "ln:p1 loads issued", 28: "ln:p1 stats", 29: "ln:p1 written"}
    for j in range(5):
        names[9 + j] = f"mma:chunk{j} issued"; names[16 + j] = f"dr:d_full{j}"; names[22 + j] = f"dr:drained{j}"
    for i in (2, 3):
        ev = sorted((t[i, k].item() - base, names[k]) for k in names if t[i, k].item() != 0)
        print(f"  tile {i}: " + "  ".join(f"{n}@{c}" for c, n in ev))
