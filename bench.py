#!/usr/bin/env python
"""Headline benchmark: utterances/s of the FacialMMT T+A+V eval forward (Swin-tiny over 160-frame face stacks ->
frame filter -> RoBERTa-large text encoder -> audio/vision encoders -> CrossmodalTransformer fusion -> 7-way logits).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...           # the reference algorithm on the host cores (oracle port)

One "step" = one eval batch of U=8 utterances per GPU (BASELINE.json configs[1]); weak scaling across GPUs
(utterances are independent; the only exchange is an NCCL all-gather of the (U,7) logits). Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME = 9.018e9            # Swin-cls, 2*MAC
FLOP_TEXT_L128 = 79.1e9
FLOP_FUSION = 33.35e9


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/): mean of
    dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of gemm_bf16_tcgen05_tma_kernel."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_full_swin160_top_kernels.json")
    try:
        rows = [r for r in json.load(open(p)) if r["kernel"].startswith("gemm_bf16_tcgen05_tma_kernel")]
        if not rows:
            return None, None
        b = sum((r["dram_rd_MB"] + r["dram_wr_MB"]) * 1e6 for r in rows) / len(rows)
        return b, (f"mean over {len(rows)} captured launches (Swin stages 1-2 of one 160-frame pass, cold L2: ncu flushes "
                   f"caches between kernels), profiles/{os.path.basename(p)}")
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/fmmt_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_cfg(text_layers: int = 24):
    from facialmmt_b200.config import FmmtConfig, TextConfig
    return FmmtConfig(text=TextConfig.roberta_large(text_layers))


def make_inputs(cfg, U, L, seed, device_faces=True):
    """MELD-shaped synthetic batch; the 8x160 face stack is drawn on the GPU (uniform [-1,1], the value range of
    ToTensor+Normalize(.5,.5)) because bicubic-upsampling 1280 crops on the host would dominate start-up."""
    import torch
    from facialmmt_b200 import synthetic as syn
    b = syn.synthetic_batch(cfg, U=U, L=L, seed=seed, with_faces=False)
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = cfg.swin
    b["faces"] = torch.rand(U, cfg.fusion.vision_len, 3, s.img_size, s.img_size, device="cuda", generator=g) * 2 - 1
    return b


def cpu_reference_throughput(cfg, L, frames_sample=16, threads=None, swin_sd=None, mm_sd=None, repeats=1):
    """utterances/s of the reference algorithm on the host cores (oracle port = torch-CPU fp32 restatement of the
    reference modules). Bounded sample: Swin over `frames_sample` frames (scaled to 160) + one full fusion forward."""
    import torch
    from facialmmt_b200 import synthetic as syn
    from oracle import facialmmt_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    swin_sd = swin_sd or syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    mm_sd = mm_sd or syn.multimodal_stress_state_dict(cfg, 1111)
    b = syn.synthetic_batch(cfg, U=1, L=L, seed=5, with_faces=False)
    frames = syn.synthetic_faces(frames_sample, 3)
    g = -torch.empty(frames_sample, 7).exponential_().log()
    best = None
    with torch.no_grad():
        for _ in range(repeats + 1):     # first pass = warm-up
            t0 = time.perf_counter()
            z = orc.swin_cls_logits(swin_sd, frames)
            probs = orc.gumbel_softmax_probs(z, g, 1.0)
            t1 = time.perf_counter()
            p160 = probs.repeat(160 // frames_sample + 1, 1)[:160]
            v519, nm = orc.filter_pack(b["vision"], b["vision_mask"], [160], p160, cfg.threshold)
            orc.multimodal_forward(mm_sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], v519,
                                   nm, b["idx_in_dia"], kind=cfg.text.kind)
            t2 = time.perf_counter()
            t_utt = (t1 - t0) * (160.0 / frames_sample) + (t2 - t1)
            best = t_utt if best is None else min(best, t_utt)
    return 1.0 / best, torch.get_num_threads(), (f"Swin-cls over {frames_sample} frames scaled x{160 // frames_sample} "
                                                  f"+ one full T+A+V fusion forward (U=1, L={L}), fp32, best of {repeats}")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from facialmmt_b200 import synthetic as syn
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = build_cfg()
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    mm_sd = syn.multimodal_stress_state_dict(cfg, 1111)
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_reference_throughput(cfg, args.text_len, frames_sample=8, swin_sd=swin_sd, mm_sd=mm_sd,
                                                    repeats=1)
        if i >= args.warmup:
            vals.append(v)
    value = len(vals) / sum(1.0 / v for v in vals)
    out = {
        "impl": "reference", "metric": "utterances/sec (160-frame T+A+V fusion fwd)", "value": value,
        "unit": "utterances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * args.batch / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "utterances/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(args, world):
    return {"workload": f"T+A+V RoBERTa-large --doEval forward, U={args.batch} utterances/GPU, 160x3x224x224 face stack "
                        f"per utterance (synthetic), {args.text_len}-token dialogue text, 160x768 audio, 160x512 vision",
            "global_batch": args.batch * world, "text_len": args.text_len, "frames_per_utterance": 160,
            "parallelism": f"utterance sharding x{world} (NCCL all-gather of logits)",
            "l2": "inputs (770 MB of faces per step) exceed the 126 MB L2; no explicit flush"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from facialmmt_b200 import _lib, synthetic as syn
    from facialmmt_b200.evaluate import evaluate_batch
    from facialmmt_b200.models import MultiModalTransformerForClassification, SwinForAffwildClassification

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    cfg = build_cfg()
    U, L = args.batch, args.text_len
    swin = SwinForAffwildClassification(cfg, swin_chunk=args.swin_chunk, swin_chunk_late=args.swin_chunk_late)
    swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    mm_sd = syn.multimodal_stress_state_dict(cfg, 1111)
    swin.load_state_dict(swin_sd)
    mm = MultiModalTransformerForClassification(cfg)
    mm.load_state_dict(mm_sd)
    if not (rank == 0 and world == 1 and not args.no_cpu_baseline):
        del swin_sd, mm_sd

    b = make_inputs(cfg, U, L, seed=1111 + 1000 * rank)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    n_imgs = [int(x) for x in b["num_imgs"]]
    labels = torch.zeros(U, dtype=torch.long)
    gathered = torch.empty(world * U, cfg.fusion.num_labels, device="cuda") if world > 1 else None

    def step_local():
        batch = (dev["text_ids"], dev["text_mask"], dev["sep_mask"], dev["audio"], dev["audio_mask"], dev["vision"],
                 dev["vision_mask"], labels, dev["faces"], n_imgs, dev["idx_in_dia"])
        return evaluate_batch(swin, mm, batch, cfg.threshold, gumbel=dev["gumbel"])

    def step_device():
        logits = step_local()
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)
            return gathered
        return logits

    # ---- host-resident copy of the inputs for the end-to-end leg (pinned)
    host = {k: v.cpu().pin_memory() for k, v in dev.items() if torch.is_tensor(v)}
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host.items() if k != "num_imgs")
    out_host = torch.empty(world * U, cfg.fusion.num_labels).pin_memory()
    # End-to-end leg: inputs start in pinned HOST memory every step. The H2D copy of step i+1 runs on a copy stream
    # while step i computes (two device-side input sets), and the logits of step i are read back to the host inside the
    # timed region; the caller holds every result on the host when the clock stops.
    copy_stream = torch.cuda.Stream()
    dev_sets = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    out_hosts = [torch.empty(world * U, cfg.fusion.num_labels).pin_memory() for _ in range(2)]
    state = {"i": 0}

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the compute that last used this set has finished
            if dev_sets[slot] is None:
                dev_sets[slot] = {k: torch.empty_like(v, device="cuda") for k, v in host.items() if k != "num_imgs"}
            for k, v in dev_sets[slot].items():
                v.copy_(host[k], non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        i = state["i"]
        slot = i & 1
        if i == 0:
            stage_inputs(0)
        stage_inputs(slot ^ 1)                              # prefetch the next step's inputs behind this step's compute
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[slot])
        d = dev_sets[slot]
        batch = (d["text_ids"], d["text_mask"], d["sep_mask"], d["audio"], d["audio_mask"], d["vision"],
                 d["vision_mask"], labels, d["faces"], n_imgs, d["idx_in_dia"])
        logits = evaluate_batch(swin, mm, batch, cfg.threshold, gumbel=d["gumbel"])
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)
            logits = gathered
        out_hosts[slot].copy_(logits, non_blocking=True)
        consumed[slot].record(cur)
        state["i"] = i + 1
        return out_hosts[slot]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = int(lib.fmmt_launch_count())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        n1 = int(lib.fmmt_launch_count())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), n1 - n0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, gpu_launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_steps = max(2, args.steps // 2)
    e2e_ms = float("nan")
    if not args.no_e2e:
        e2e_ms, _ = timed(step_e2e, e2e_steps, 1)

    # ---- per-kernel event profile of one extra step (same stream) for the roofline object
    roof = None
    if rank == 0 and not args.no_e2e:
        swin.set_profile(True)
        mm.set_profile(True)
        step_local()                      # rank-local: no collective here (the other ranks are not in this step)
        torch.cuda.synchronize()
        prof = {}
        for k, v in list(swin.read_profile().items()) + list(mm.read_profile().items()):
            a = prof.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for f in a:
                a[f] += v[f]
        swin.set_profile(False)
        mm.set_profile(False)
        pk = peaks()
        # every tcgen05 GEMM-class launch: the Linear-layer GEMMs and the fused MLP kernels (LN + fc1 + GELU + fc2 + residual)
        gem = [v for k, v in prof.items() if k.startswith("gemm ") or k.startswith("mlp_fused")]
        g_ms = sum(v["ms"] for v in gem); g_fl = sum(v["flops"] for v in gem); g_n = sum(v["launches"] for v in gem)
        all_ms = sum(v["ms"] for v in prof.values())
        achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_tma_kernel + fused MLP kernels (every Linear layer of the step)",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"], "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1],
                "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "launches_per_step": g_n, "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1),
                "share_of_step_kernel_time": g_ms / all_ms if all_ms > 0 else None}
        if args.profile_out:
            json.dump(prof, open(args.profile_out, "w"), indent=1, sort_keys=True)

    utt = world * U * args.steps
    value = utt / (total_ms * 1e-3)
    e2e_value = world * U * e2e_steps / (e2e_ms * 1e-3)
    if rank == 0:
        pk = peaks()
        flop_per_utt = 160 * FLOP_PER_FRAME + (FLOP_TEXT_L128 if L == 128 else FLOP_TEXT_L128 * L / 128) + FLOP_FUSION
        out = {
            "metric": "utterances/sec (160-frame T+A+V fusion fwd)", "value": value, "unit": "utterances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "utterances/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": e2e_steps},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": roof,
            "path_tensor_frac": value / world * flop_per_utt / (pk["tf_sustained"] * 1e12),
        }
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(max(1, os.cpu_count() or 1))
            v, cores, sample = cpu_reference_throughput(cfg, L, frames_sample=8, swin_sd=swin_sd, mm_sd=mm_sd)
            out["cpu_baseline"] = {"value": v, "unit": "utterances/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="utterances per GPU per step (BASELINE.json configs[1]: 8)")
    ap.add_argument("--text-len", type=int, default=128)
    ap.add_argument("--swin-chunk", type=int, default=0)
    ap.add_argument("--swin-chunk-late", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel event profile of one step as JSON")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
