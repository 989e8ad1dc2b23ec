python - <<'P'
import math, os, sys
sys.path.insert(0, '.')
import torch
from facialmmt_b200 import _lib
from facialmmt_b200._lib import check, ptr
lib = _lib.load()
g = torch.Generator().manual_seed(1)
for C, M in ((384, 31360), (192, 50176)):
    H = 4 * C
    gam, bet = torch.ones(C).cuda(), torch.zeros(C).cuda()
    w1 = (torch.randn(H, C, generator=g) / math.sqrt(C)).cuda().to(torch.bfloat16).contiguous()
    w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).cuda().to(torch.bfloat16).contiguous()
    b1, b2 = torch.zeros(H).cuda(), torch.zeros(C).cuda()
    x = torch.randn(M, C, generator=g).cuda()
    tr = torch.zeros(H // 64, 8, dtype=torch.int64, device='cuda')
    for _ in range(3):
        torch.cuda.synchronize()
        check(lib.fmmt_op_swin_mlp_stream(ptr(x), M, C, ptr(gam), ptr(bet), 1e-5, ptr(w1), C, ptr(b1), ptr(w2), H, ptr(b2), -1, ptr(tr)))
        torch.cuda.synchronize()
    t = tr.cpu()
    t0 = int(t[t > 0].min())
    print('C', C, 'columns: d1free w1p0 w1p1 w2land hidwait w1req w2req hidwritten (cycles from first stamp)')
    for j in range(H // 64):
        print(j, [int(v) - t0 if v > 0 else -1 for v in t[j].tolist()])
P
