#!/usr/bin/env python
"""`main.py --doEval` drop-in for the reference CLI (main.py:16-103, train.py:424-435), inference only.

Keeps the reference flag names/defaults that matter for evaluation (--choice_modality, --plm_name, --load_*_path,
--trg_batch_size, --tau, --FacialEmoImpor_threshold, --seed, fusion sizes) and adds:
  --synthetic N     evaluate N seeded MELD-shaped synthetic utterances with seeded weights (no data/checkpoints here)
  --per_utterance   decide the filter fallback per utterance (default; == reference at its batch size 1)
Checkpoints: `--load_swin_path/--load_multimodal_path/--load_unimodal_path` accept a torch state_dict file with the
reference key names (or a pickled module exposing .state_dict()); without them seeded synthetic weights are used.
Multi-GPU: launch with torchrun; the utterance batch is sharded and logits all-gathered (facialmmt_b200.distributed).
"""
from __future__ import annotations

import argparse
import os

import torch


def build_parser():
    p = argparse.ArgumentParser(description="FacialMMT eval on B200 (facialmmt_b200)")
    p.add_argument("--num_labels", type=int, default=7)
    p.add_argument("--plm_name", type=str, default="roberta-large", choices=["roberta-large", "bert-large"])
    p.add_argument("--choice_modality", type=str, default="T+A+V", choices=["T+A+V", "V"])
    p.add_argument("--backbone_type", type=str, default="SwinTransformer")
    p.add_argument("--backbone_conf_file", type=str, default="")
    p.add_argument("--tau", type=float, default=1)
    p.add_argument("--FacialEmoImpor_threshold", type=float, default=0.2)
    p.add_argument("--trg_batch_size", type=int, default=1)
    p.add_argument("--crossmodal_layers_TA", type=int, default=2)
    p.add_argument("--crossmodal_num_heads_TA", type=int, default=12)
    p.add_argument("--crossmodal_layers_TA_V", type=int, default=2)
    p.add_argument("--crossmodal_num_heads_TA_V", type=int, default=12)
    p.add_argument("--audio_utt_Transformernum", type=int, default=5)
    p.add_argument("--vision_utt_Transformernum", type=int, default=2)
    p.add_argument("--hidden_size", type=int, default=768)
    p.add_argument("--num_attention_heads", type=int, default=12)
    p.add_argument("--intermediate_size", type=int, default=3072)
    p.add_argument("--layer_norm_eps", type=float, default=1e-12)
    p.add_argument("--seed", type=int, default=1111)
    p.add_argument("--doEval", type=int, default=1)
    p.add_argument("--load_unimodal_path", type=str, default="")
    p.add_argument("--load_multimodal_path", type=str, default="")
    p.add_argument("--load_swin_path", type=str, default="")
    p.add_argument("--synthetic", type=int, default=8, help="number of synthetic utterances to evaluate")
    p.add_argument("--text_len", type=int, default=128)
    p.add_argument("--text_layers", type=int, default=24)
    p.add_argument("--per_utterance", type=int, default=1)
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"],
                   help="bf16: bf16 operands / fp32 accumulate (logits within 1e-2 of the fp32 reference); fp32: split-bf16 x3 "
                        "operands + fp32 attention (within 1e-3)")
    p.add_argument("--dialogues", type=str, default="",
                   help="JSON [{'utterances': [text | [token ids], ...]}, ...]: evaluate every utterance with its dialogue encoded "
                        "as src/meld_bert_extraText.py does (facialmmt_b200.text_frontend); identical dialogues of a batch are "
                        "encoded once")
    p.add_argument("--features", type=str, default="", help="torch file with the per-utterance audio/vision/face tensors "
                                                            "(facialmmt_b200/data.py); synthetic stand-ins when absent")
    p.add_argument("--tokenizer_path", type=str, default="", help="HF tokenizer directory for string utterances")
    p.add_argument("--bug_compat", type=int, default=0,
                   help="1: reproduce the reference's literal batch re-pack at trg_batch_size > 1, off-by-one included "
                        "(train.py:200,213); default 0 = per-utterance semantics (== the reference at trg_batch_size 1)")
    p.add_argument("--trust_checkpoint", type=int, default=0,
                   help="allow unpickling the reference's whole-module checkpoints (train.py:428-432); executes pickle code")
    return p


_TRUST = {"on": False}


def _load_sd(path):
    from facialmmt_b200.checkpoint import load_state_dict_file
    return load_state_dict_file(path, trust=_TRUST["on"])


def main(argv=None):
    args = build_parser().parse_args(argv)
    _TRUST["on"] = bool(args.trust_checkpoint)
    if not args.doEval:
        raise SystemExit("facialmmt_b200 implements --doEval (inference) only")
    if args.trg_batch_size > 1 and not args.bug_compat:
        print(f"[facialmmt_b200] trg_batch_size={args.trg_batch_size}: utterances are evaluated with the reference's batch-size-1 "
              "semantics (its literal U>1 re-pack drops/shifts frames: train.py:200,213). Pass --bug_compat 1 for the literal code.")
    import torch.distributed as dist
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, FusionConfig, TextConfig
    from facialmmt_b200.distributed import gather_logits, shard_range
    from facialmmt_b200.evaluate import eval_meld, evaluate_batch
    from facialmmt_b200.models import (MultiModalTransformerForClassification, SwinForAffwildClassification,
                                       meld_utt_transformer)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    torch.manual_seed(args.seed)
    t = TextConfig.roberta_large(args.text_layers) if args.plm_name == "roberta-large" else TextConfig.bert_large(args.text_layers)
    f = FusionConfig(hidden=args.hidden_size, heads=args.num_attention_heads, ffn=args.intermediate_size,
                     eps=args.layer_norm_eps, audio_layers=args.audio_utt_Transformernum,
                     vision_layers=args.vision_utt_Transformernum, cmt_layers_ta=args.crossmodal_layers_TA,
                     cmt_heads_ta=args.crossmodal_num_heads_TA, cmt_layers_tav=args.crossmodal_layers_TA_V,
                     cmt_heads_tav=args.crossmodal_num_heads_TA_V, num_labels=args.num_labels)
    cfg = FmmtConfig(text=t, fusion=f, tau=args.tau, threshold=args.FacialEmoImpor_threshold)
    n = args.synthetic
    lo, hi = shard_range(n, rank, world)
    labels = torch.randint(0, args.num_labels, (n,), generator=torch.Generator().manual_seed(args.seed))
    results = []
    if args.choice_modality == "V":
        model = meld_utt_transformer(cfg, precision=args.precision)
        model.load_state_dict(_load_sd(args.load_unimodal_path) if args.load_unimodal_path
                              else syn.unimodal_stress_state_dict(cfg.fusion, args.seed))
        b = syn.synthetic_batch(cfg, U=n, L=16, seed=args.seed, with_faces=False)
        for u0 in range(lo, hi, args.trg_batch_size):
            u1 = min(hi, u0 + args.trg_batch_size)
            results.append(model(b["vision"][u0:u1], b["vision_mask"][u0:u1]))
    else:
        swin = SwinForAffwildClassification(cfg, precision=args.precision)
        swin.load_state_dict(_load_sd(args.load_swin_path) if args.load_swin_path
                             else syn.swin_cls_stress_state_dict(cfg.swin, args.seed))
        mm = MultiModalTransformerForClassification(cfg, precision=args.precision)
        mm.load_state_dict(_load_sd(args.load_multimodal_path) if args.load_multimodal_path
                           else syn.multimodal_stress_state_dict(cfg, args.seed))
        if args.dialogues:
            # real text path: dialogue-level encoding + per-utterance features (or synthetic stand-ins)
            from facialmmt_b200 import data as fdata
            tok = None
            if args.tokenizer_path:
                from transformers import AutoTokenizer
                tok = AutoTokenizer.from_pretrained(args.tokenizer_path)
            ids, msk, sep, idx = fdata.encode_all(fdata.load_dialogues(args.dialogues), cfg.text.kind, tok)
            n = len(ids)
            lo, hi = shard_range(n, rank, world)
            feats = fdata.load_features(args.features, cfg, n, args.seed)
            labels = feats["labels"]
            sl = lambda x: x[lo:hi]   # noqa: E731
            shard = {k: sl(v) for k, v in feats.items()}
            for batch in fdata.iter_batches(ids[lo:hi], msk[lo:hi], sep[lo:hi], idx[lo:hi], shard, args.trg_batch_size):
                results.append(evaluate_batch(swin, mm, batch, args.FacialEmoImpor_threshold,
                                              per_utterance=bool(args.per_utterance), bug_compat=bool(args.bug_compat)))
            u_range = range(0)
        else:
            u_range = range(lo, hi, args.trg_batch_size)
        for u0 in u_range:
            u1 = min(hi, u0 + args.trg_batch_size)
            b = syn.synthetic_batch(cfg, U=u1 - u0, L=args.text_len, seed=args.seed + u0, with_faces=True)
            batch = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], b["vision"],
                     b["vision_mask"], labels[u0:u1], b["faces"], b["num_imgs"], b["idx_in_dia"])
            results.append(evaluate_batch(swin, mm, batch, args.FacialEmoImpor_threshold,
                                          per_utterance=bool(args.per_utterance), bug_compat=bool(args.bug_compat)))
    local = torch.cat(results) if results else torch.zeros(0, args.num_labels, device="cuda")
    logits = gather_logits(local, n)
    if rank == 0:
        print("&" * 50)
        print("**TEST** | wg_av_f1 {:5.4f} ".format(eval_meld(logits, labels, test=True)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
