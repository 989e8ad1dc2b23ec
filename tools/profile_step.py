#!/usr/bin/env python
"""One warm-up step + one step of the bench workload (BASELINE configs[1]: U=8, RoBERTa-large, 160 frames) for ncu.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
Prints the library's launch count of the second step, so that the launch list can be split into warm-up / measured."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from facialmmt_b200 import _lib, synthetic as syn  # noqa: E402
from facialmmt_b200.evaluate import evaluate_batch  # noqa: E402
from facialmmt_b200.models import MultiModalTransformerForClassification, SwinForAffwildClassification  # noqa: E402

U = int(os.environ.get("PROFILE_U", "8"))
steps = int(os.environ.get("PROFILE_STEPS", "2"))
cfg = bench.build_cfg()
swin = SwinForAffwildClassification(cfg)
swin.load_state_dict(syn.swin_cls_stress_state_dict(cfg.swin, 1111))
mm = MultiModalTransformerForClassification(cfg)
mm.load_state_dict(syn.multimodal_stress_state_dict(cfg, 1111))
b = bench.make_inputs(cfg, U, 128, 1111)
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
n = [int(x) for x in b["num_imgs"]]
lib = _lib.load()
for s in range(steps):
    c0 = lib.fmmt_launch_count()
    batch = (dev["text_ids"], dev["text_mask"], dev["sep_mask"], dev["audio"], dev["audio_mask"], dev["vision"],
             dev["vision_mask"], torch.zeros(U, dtype=torch.long), dev["faces"], n, dev["idx_in_dia"])
    out = evaluate_batch(swin, mm, batch, cfg.threshold, gumbel=dev["gumbel"])
    torch.cuda.synchronize()
    print(f"step {s}: {lib.fmmt_launch_count() - c0} library launches", flush=True)
swin.check(); mm.check()
