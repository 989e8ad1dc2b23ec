"""ncu driver (not a pytest): two Swin-cls passes over 16 frames (pass 1 = warm-up). Usage under gpurun:
   ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 53 -c 8 -o gpurun_out/x python tests/gpu_profile_swin.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facialmmt_b200 import synthetic as syn
from facialmmt_b200.config import FmmtConfig
from facialmmt_b200.models import SwinForAffwildClassification

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = FmmtConfig()
m = SwinForAffwildClassification(cfg, swin_chunk=F, swin_chunk_late=F)
m.load_state_dict(syn.swin_cls_stress_state_dict(cfg.swin, 1111))
x = torch.rand(F, 3, 224, 224, device="cuda") * 2 - 1
m(x, is_trg_task=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()   # ncu --profile-from-start off captures only this pass
m(x, is_trg_task=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
