"""BASELINE INFRASTRUCTURE ONLY (bench.py's `--impl reference` arm and `cpu_baseline` leg; never the product path).

Times the reference's OWN CPU implementation of the hot path on the host cores: the unmodified reference modules
(`src/models.py` SwinForAffwildClassification + MultiModalTransformerForClassification, imported through
oracle/ref_harness.py from /root/reference or its byte-identical copy baseline/_ref/) driven by the reference's OWN eval
loop (`train.py:154-243` multimodal_evaluate, exec'd from its source text), fp32, eval(), all host threads.
Falls back to the oracle port (oracle/facialmmt_oracle.py) only where neither copy of the reference exists; the JSON
line says which (`cpu_baseline.kind` = "reference" | "port").

One "step" of this arm = ONE full utterance batch of `U` utterances x 160 frames through Swin -> Sum p^2 filter -> fusion
(U=1 by default = the reference's default `trg_batch_size`, main.py:56; a bounded sample of the GPU arm's U=8 step).
Nothing is extrapolated: ms_per_step is the measured wall time of the steps that ran.
"""
from __future__ import annotations

import argparse
import os
import statistics
import time

import torch


class _Timed(torch.nn.Module):
    """Wraps a reference module to accumulate the wall time of its forward (Swin / fusion split of BASELINE.md section 4)."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.seconds = 0.0

    def forward(self, *a, **k):
        t0 = time.perf_counter()
        out = self.inner(*a, **k)
        self.seconds += time.perf_counter() - t0
        return out


class ReferenceRunner:
    """Builds the models once (seeded stress weights = the GPU arm's weights) and runs utterance batches."""

    def __init__(self, cfg, text_len: int, plm: str = "roberta-large", threads: int | None = None):
        from facialmmt_b200 import synthetic as syn
        from oracle import ref_harness as rh
        torch.set_num_threads(threads or max(1, os.cpu_count() or 1))
        self.cfg, self.L = cfg, text_len
        self.kind = "reference" if rh.available() else "port"
        self.cores = torch.get_num_threads()
        self.swin_sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
        self.mm_sd = syn.multimodal_stress_state_dict(cfg, 1111)
        if self.kind == "reference":
            f = cfg.fusion
            args = rh.default_args(plm, audio_len=f.audio_len, vision_len=f.vision_len, audio_dim=f.audio_dim,
                                   vision_dim=f.vision_dim, text_len=f.text_len)
            swin = rh.build_swin_cls(args)
            swin.load_state_dict(self.swin_sd)
            mm = rh.build_multimodal(args, text_layers=cfg.text.layers)
            mm.load_state_dict(self.mm_sd, strict=False)
            self.swin, self.mm = _Timed(swin), _Timed(mm)
            self.loop_args = argparse.Namespace(trg_batch_size=1, FacialEmoImpor_threshold=cfg.threshold, num_labels=7,
                                                trg_n_test=1, trg_n_valid=1)
            self.make_loop = rh.literal_eval_loop()
        self.source = rh.REFERENCE_ROOT if self.kind == "reference" else "oracle/facialmmt_oracle.py"

    def batch(self, U: int, seed: int, faces=None):
        from facialmmt_b200 import synthetic as syn
        b = syn.synthetic_batch(self.cfg, U=U, L=self.L, seed=seed, with_faces=False)
        if faces is None:
            g = torch.Generator().manual_seed(seed)
            faces = torch.rand(U, self.cfg.fusion.vision_len, 3, 224, 224, generator=g) * 2 - 1
        b["faces"] = faces
        return b

    def swin_logits(self, frames: torch.Tensor) -> torch.Tensor:
        """Raw Swin-cls logits of the reference (is_trg_task falsy, src/models.py:30-37) -- parity check of the GPU arm."""
        with torch.no_grad():
            if self.kind == "reference":
                return self.swin.inner(frames, is_trg_task=False)
            from oracle import facialmmt_oracle as orc
            return orc.swin_cls_logits(self.swin_sd, frames)

    def step(self, b) -> dict:
        """One eval batch; returns seconds: total / swin / fusion / glue."""
        U = b["faces"].shape[0]
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.swin.seconds = self.mm.seconds = 0.0
            self.loop_args.trg_batch_size = U
            self.loop_args.trg_n_test = U
            batch = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], b["vision"],
                     b["vision_mask"], torch.zeros(U, dtype=torch.long), b["faces"], b["num_imgs"], b["idx_in_dia"])
            crit = torch.nn.CrossEntropyLoss()
            with torch.no_grad():
                _, results, _ = self.make_loop(self.loop_args, [batch])(self.swin, self.mm, crit, test=True)
            total = time.perf_counter() - t0
            return dict(total=total, swin=self.swin.seconds, fusion=self.mm.seconds,
                        glue=total - self.swin.seconds - self.mm.seconds, logits=results)
        from oracle import facialmmt_oracle as orc
        with torch.no_grad():
            n = [int(v) for v in b["num_imgs"]]
            frames = torch.cat([b["faces"][u, :n[u]] for u in range(U)], 0)
            z = orc.swin_cls_logits(self.swin_sd, frames)
            t1 = time.perf_counter()
            g = -torch.empty_like(z).exponential_().log()
            b2 = dict(b)
            b2["gumbel"] = g
            probs = orc.gumbel_softmax_probs(z, g, 1.0)
            vs, ms, off = [], [], 0
            for u in range(U):
                v, m = orc.filter_pack(b["vision"][u:u + 1], b["vision_mask"][u:u + 1], n[u:u + 1], probs[off:off + n[u]],
                                       self.cfg.threshold)
                vs.append(v); ms.append(m); off += n[u]
            t2 = time.perf_counter()
            logits = orc.multimodal_forward(self.mm_sd, b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"],
                                            b["audio_mask"], torch.cat(vs), torch.cat(ms), b["idx_in_dia"],
                                            kind=self.cfg.text.kind)
            t3 = time.perf_counter()
        return dict(total=t3 - t0, swin=t1 - t0, fusion=t3 - t2, glue=t2 - t1, logits=logits)


def summarize(runner: ReferenceRunner, times: list, U: int) -> dict:
    tot = [t["total"] for t in times]
    return {
        "value": U * len(tot) / sum(tot), "unit": "utterances/s", "cores": runner.cores, "kind": runner.kind,
        "sample": (f"{len(tot)} full eval batch(es) of U={U} utterance(s) x 160 frames (Swin over all 160 frames + filter + "
                   f"T+A+V fusion, L={runner.L}), fp32, {'the unmodified reference modules + train.py eval loop from ' + runner.source if runner.kind == 'reference' else 'oracle port'}; nothing extrapolated"),
        "median_s_per_batch": statistics.median(tot),
        "split_s_median": {k: statistics.median([t[k] for t in times]) for k in ("swin", "glue", "fusion")},
    }
