#!/usr/bin/env python
"""Upper bound of overlapping the fusion forward with the Swin forward of the same step (timing probe, results are NOT the
bench's: the fusion forward here consumes the PREVIOUS step's filtered vision features so that it can start beside Swin)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from facialmmt_b200 import synthetic as syn
from facialmmt_b200.evaluate import filter_pack, gather_valid_frames
from facialmmt_b200.models import MultiModalTransformerForClassification, SwinForAffwildClassification

U = 8
cfg = bench.build_cfg()
swin = SwinForAffwildClassification(cfg); swin.load_state_dict(syn.swin_cls_stress_state_dict(cfg.swin, 1111))
mm = MultiModalTransformerForClassification(cfg); mm.load_state_dict(syn.multimodal_stress_state_dict(cfg, 1111))
swin.set_graph(True); mm.set_graph(True)
b = bench.make_inputs(cfg, U, 128, 1111)
d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
n = [int(x) for x in b["num_imgs"]]
frames = gather_valid_frames(d["faces"], n)
s2 = torch.cuda.Stream()
state = {}

def serial():
    _, probs, _ = swin.forward_full(frames, d["gumbel"])
    v519, nm = filter_pack(d["vision"], d["vision_mask"], n, probs, cfg.threshold, True, cache=swin._out_cache)
    state["v"], state["m"] = v519, nm
    return mm(d["text_ids"], d["text_mask"], d["sep_mask"], d["audio"], d["audio_mask"], v519, nm, d["idx_in_dia"])

def overlapped():
    cur = torch.cuda.current_stream()
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        out = mm(d["text_ids"], d["text_mask"], d["sep_mask"], d["audio"], d["audio_mask"], state["v"], state["m"], d["idx_in_dia"])
    _, probs, _ = swin.forward_full(frames, d["gumbel"])
    filter_pack(d["vision"], d["vision_mask"], n, probs, cfg.threshold, True, cache=swin._out_cache)
    cur.wait_stream(s2)
    return out

for name, fn in (("serial", serial), ("fusion forward beside Swin", overlapped), ("serial", serial)):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per step")
swin.check(); mm.check()
