python - <<'P'
import sys; sys.path.insert(0,'.')
from facialmmt_b200 import _lib
lib=_lib.load()
import torch; torch.zeros(1,device='cuda')
for mode in range(8):
    for n in (256, 128):
        print('mode', mode, 'N', n, 'cycles/MMA', round(lib.fmmt_debug_mma_cycles(n | (mode << 16), 2000), 2), flush=True)
P
