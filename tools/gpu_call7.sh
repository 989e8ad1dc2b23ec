python - <<'P'
import sys, ctypes; sys.path.insert(0,'.')
from facialmmt_b200 import _lib
lib=_lib.load()
import torch; torch.zeros(1,device='cuda')
out=(ctypes.c_double*2)()
for grid in (148, 74, 16):
  for mode in (0,1):
    for box in (128, 256):
      for ns in (1,2,3,4,6):
        if ns*box*128 > 150000: continue
        rc=lib.fmmt_debug_feed(4000, ns, box, mode, grid, ctypes.addressof(out))
        print(f'grid {grid} mode {mode} box_rows {box} stages {ns}: rc {rc}  {out[0]:.1f} B/clk/SM  ({out[0]*1.965*grid/1000:.2f} TB/s @1.965GHz)  cyc/MMA {out[1]:.1f}', flush=True)
print('timeout', hex(lib.fmmt_debug_timeout(1)))
P
