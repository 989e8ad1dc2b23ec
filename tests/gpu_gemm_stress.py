"""Race hunt: repeat the un-fused MLP chain (LN -> fc1+GELU -> fc2+residual) and check each GEMM against torch."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, cur_stream, ptr
lib = _lib.load()
g = torch.Generator().manual_seed(1)
C, H = 384, 1536
w1 = (torch.randn(H, C, generator=g) / math.sqrt(C)).cuda().to(torch.bfloat16).contiguous()
w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).cuda().to(torch.bfloat16).contiguous()
b1, b2 = (0.3 * torch.randn(H, generator=g)).cuda(), (0.3 * torch.randn(C, generator=g)).cuda()
bad = {"fc1": 0, "fc2": 0}
for M in (148 * 128, 31360, 148 * 128 + 64):
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
        h16 = torch.randn(M, C, generator=g).cuda().to(torch.bfloat16)
        x = torch.randn(M, C, generator=g).cuda()
        hid16 = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
        check(lib.fmmt_op_gemm(ptr(h16), C, ptr(w1), C, M, H, C, ptr(b1), 1, None, 0, None, 0, ptr(hid16), H, None, 0, 0, cur_stream()))
        y = x.clone()
        check(lib.fmmt_op_gemm(ptr(hid16), H, ptr(w2), H, M, C, H, ptr(b2), 0, ptr(y), C, ptr(y), C, None, 0, None, 0, 0, cur_stream()))
        torch.cuda.synchronize()
        ref1 = torch.nn.functional.gelu(h16.float() @ w1.float().t() + b1)
        e1 = (hid16.float() - ref1).abs().max().item()
        ref2 = x + hid16.float() @ w2.float().t() + b2
        e2 = (y - ref2).abs().max().item()
        if e1 > 0.1 or e2 > 0.1:
            bad["fc1"] += e1 > 0.1
            bad["fc2"] += e2 > 0.1
            rows1 = ((hid16.float() - ref1).abs().max(dim=1).values > 0.1).nonzero().flatten()
            rows2 = ((y - ref2).abs().max(dim=1).values > 0.1).nonzero().flatten()
            cols2 = ((y - ref2).abs().max(dim=0).values > 0.1).nonzero().flatten()
            print(f"M={M} rep={rep}: fc1 err {e1:.3e} (bad rows {rows1[:4].tolist()}..{len(rows1)}), fc2 err {e2:.3e} "
                  f"(bad rows {rows2[:4].tolist()}..{len(rows2)}, cols {cols2[:4].tolist()}..{len(cols2)})", flush=True)
print("KBS", os.environ.get("FMMT_KBS"), "bad:", bad, "timeout", hex(lib.fmmt_debug_timeout(1)), flush=True)
