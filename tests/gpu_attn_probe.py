"""Bench probe (not a test): timing + per-tile stamp trace of the fused attention half-block at the bench geometry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_attn_fused_gpu import window_maps, C, H

lib = _lib.load()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 320
R, T = 56, 3136
M = frames * T
g = torch.Generator().manual_seed(1)
x = torch.randn(M, C, generator=g).cuda()
out = torch.empty(M, C, device="cuda")
wq = (torch.randn(288, C, generator=g) * 0.1).contiguous(); wp = (torch.randn(C, C, generator=g) * 0.1).contiguous()
tb = torch.randn(169, H, generator=g).contiguous()
img = torch.empty(73728, dtype=torch.uint8, device="cuda"); tab = torch.empty(507, device="cuda")
_lib.check(lib.fmmt_op_swin_attn_pack(_lib.ptr(wq), _lib.ptr(wp), _lib.ptr(tb), _lib.ptr(img), _lib.ptr(tab)), "pack")
vec = [torch.ones(C).cuda(), torch.zeros(C).cuda(), torch.zeros(288).cuda(), torch.zeros(C).cuda()]
for shift in (0, 3):
    gather, rid, wflag = window_maps(R, shift)
    gd, rd, wd = gather.cuda(), rid.cuda(), wflag.cuda()
    def run(trace=None):
        _lib.check(lib.fmmt_op_swin_attn(_lib.ptr(x), _lib.ptr(out), M, T, _lib.ptr(gd), _lib.ptr(vec[0]), _lib.ptr(vec[1]), 1e-5,
                                         _lib.ptr(img), _lib.ptr(tab), _lib.ptr(vec[2]), _lib.ptr(vec[3]),
                                         _lib.ptr(rd) if shift else None, _lib.ptr(wd) if shift else None,
                                         -64 if trace is not None else 64,
                                         _lib.ptr(trace) if trace is not None else _lib.cur_stream()), "attn")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = M // 98
    print(f"shift={shift} frames={frames}: {ms:.3f} ms/launch, {ms * 1e3 / (tiles / 148):.2f} us per tile per SM, "
          f"{8 * M * C / ms / 1e6:.0f} GB/s algorithmic")
    tr = torch.zeros(8 * 32, dtype=torch.int64, device="cuda")
    run(tr); torch.cuda.synchronize()
    t = tr.cpu().view(8, 32)
    base = t[0, 0].item()
    names = {0: "mma:start", 1: "mma:a_full", 2: "mma:qkv issued", 3: "mma:qkv_ready", 4: "mma:p_full0", 5: "mma:p_full1",
             6: "mma:p_full2", 7: "mma:o_smem_full", 8: "mma:proj issued", 31: "g:proj_full"}
    for h in range(3):
        for k, n in enumerate(["tile start", "qkv_full", "qkv drained", "s_full", "P written", "o_full", "O drained"]):
            names[10 + 7 * h + k] = f"g{h}:{n}"
    for i in (2, 3):
        ev = sorted((t[i, k].item() - base, names[k]) for k in names if t[i, k].item() != 0)
        print(f"  tile {i}: " + "  ".join(f"{n}@{c}" for c, n in ev))
