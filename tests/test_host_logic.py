"""CPU: host-side logic of the drop-in layer (config mapping, state-dict key listing, sharding + logits gather under
gloo with world_size 2, W-F1 metric, CLI flags)."""
import argparse
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_namespace_to_config_follows_reference_fields():
    from facialmmt_b200.models import _fmmt_config_from_namespace
    ns = argparse.Namespace(pretrainedtextmodel_path="/x/y/bert-large", hidden_size=768, num_attention_heads=12,
                            intermediate_size=3072, layer_norm_eps=1e-12, audio_featExtr_dim=768,
                            vision_featExtr_dim=512, audio_utt_Transformernum=5, vision_utt_Transformernum=2,
                            crossmodal_layers_TA=2, crossmodal_num_heads_TA=12, crossmodal_layers_TA_V=2,
                            crossmodal_num_heads_TA_V=12, get_text_utt_max_lens=38, get_audio_utt_max_lens=122,
                            get_vision_utt_max_lens=160, num_labels=7)
    cfg = _fmmt_config_from_namespace(ns)
    assert cfg.text.kind == "bert" and cfg.text.eps == 1e-12 and cfg.text.max_pos == 512   # src/models.py:49-52
    assert cfg.fusion.audio_len == 122 and cfg.fusion.vision_dim == 512
    ns.pretrainedtextmodel_path = "/x/roberta-large"
    assert _fmmt_config_from_namespace(ns).text.kind == "roberta"


def test_state_dict_specs_are_consistent():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    spec = syn.multimodal_state_dict_spec(cfg)
    assert spec["vision_linear.weight"] == (768, 519)
    assert spec["CrossModalTrans_TA.layers.0.self_attn.in_proj_weight"] == (2304, 768)
    sw = syn.swin_cls_state_dict_spec(cfg.swin)
    assert sw["swin.output_layer.2.weight"] == (512, 37632)
    assert "swin.layers.3.blocks.1.attn_mask" not in sw          # stage 4 never shifts (Swin_Transformer.py:192-195)
    assert sw["swin.layers.0.blocks.1.attn_mask"] == (64, 49, 49)
    sd = syn.stress_state_dict(syn.unimodal_state_dict_spec(cfg.fusion), 3)
    sd2 = syn.stress_state_dict(syn.unimodal_state_dict_spec(cfg.fusion), 3)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)            # seeded, order independent


def test_shard_range_partitions_exactly():
    from facialmmt_b200.distributed import shard_range
    for n in (1, 7, 8, 256, 13):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gloo_worker(rank, world, port, n, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from facialmmt_b200.distributed import gather_logits, shard_batch, shard_range
    full = torch.arange(n * 7, dtype=torch.float32).view(n, 7)          # stands in for per-utterance logits
    batch = tuple([full] * 9 + [list(range(n))] + [full])
    mine = shard_batch(batch, rank, world)
    lo, hi = shard_range(n, rank, world)
    assert mine[0].shape[0] == hi - lo and mine[9] == list(range(lo, hi))
    got = gather_logits(mine[0] * 2.0, n)
    ok = torch.equal(got, full * 2.0)
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, ok))


@pytest.mark.parametrize("n", [8, 5])
def test_sharded_logits_gather_world2_gloo(n):
    """N>1 path on CPU: 2 processes, gloo, ragged and even shards; gathered logits == single-process order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_eval_meld_weighted_f1():
    from facialmmt_b200.evaluate import eval_meld
    logits = torch.eye(7)[[0, 1, 2, 2, 4]]
    assert eval_meld(logits, torch.tensor([0, 1, 2, 2, 4])) == 1.0
    assert 0.0 < eval_meld(logits, torch.tensor([0, 1, 2, 3, 4])) < 1.0


def test_cli_keeps_reference_flag_names():
    sys.path.insert(0, ROOT)
    import main as cli
    a = cli.build_parser().parse_args(["--choice_modality", "T+A+V", "--plm_name", "bert-large", "--doEval", "1",
                                       "--FacialEmoImpor_threshold", "0.25", "--tau", "2", "--trg_batch_size", "4"])
    assert a.plm_name == "bert-large" and a.FacialEmoImpor_threshold == 0.25 and a.tau == 2 and a.trg_batch_size == 4


def test_checkpoint_ingestion_wrappers_and_backbone_remap(tmp_path):
    """train.py:428-432 pickles (_LiteModule -> DataParallel -> module) and the backbone.* remap of train.py:316-331."""
    import torch
    from facialmmt_b200.checkpoint import load_state_dict_file, remap_pretrained_backbone, strip_wrappers, to_state_dict

    inner = torch.nn.Sequential(torch.nn.Linear(3, 2))
    wrapped = torch.nn.Sequential()
    wrapped.add_module("_module", torch.nn.Sequential())
    wrapped._module.add_module("module", inner)                      # keys: _module.module.0.weight
    assert set(to_state_dict(wrapped)) == {"0.weight", "0.bias"}
    assert set(strip_wrappers({"module._module.module.a.b": 1})) == {"a.b"}
    p = tmp_path / "sd.pt"
    torch.save({"state_dict": {"module.linear.weight": torch.ones(2, 2)}}, p)
    assert set(load_state_dict_file(str(p))) == {"linear.weight"}
    with pytest.raises(TypeError):
        to_state_dict(3)

    model_keys = ["swin.patch_embed.proj.weight", "linear.weight", "classifier.weight", "classifier.bias", "swin.missing"]
    pre = {"backbone.patch_embed.proj.weight": torch.zeros(1), "backbone.linear.weight": torch.ones(1),
           "backbone.classifier.weight": torch.ones(1), "linear.weight": torch.full((1,), 2.0)}
    got = remap_pretrained_backbone(model_keys, pre)
    assert set(got) == {"swin.patch_embed.proj.weight", "linear.weight"} and got["linear.weight"].item() == 1.0
    # the literal reference loop, restated verbatim, agrees with literal=True
    ref = {}
    for k in model_keys:
        if k in pre:
            if k == 'classifier.weight' or k == 'classifier.bias':
                continue
            k_val = k[5:] if k[:5] == 'swin.' else k
            ref[k] = pre['backbone.' + k_val]
    assert set(remap_pretrained_backbone(model_keys, pre, literal=True)) == set(ref) == {"linear.weight"}


def test_pickled_lite_module_checkpoint_loads_without_lightning(tmp_path):
    """train.py:428-432 loads whole pickled `_LiteModule(DataParallel(module))` objects (utils/util.py:121-132 saves them after
    Lite.setup). checkpoint.load_state_dict_file must (a) refuse to unpickle them unless trusted, (b) read them WITHOUT
    pytorch_lightning or the reference's classes importable, giving the reference's own key names."""
    import sys
    import types

    import torch.nn as nn
    from facialmmt_b200 import checkpoint as ck

    pl = types.ModuleType("pytorch_lightning"); lite = types.ModuleType("pytorch_lightning.lite")
    wr = types.ModuleType("pytorch_lightning.lite.wrappers"); refmod = types.ModuleType("src_models_fake")

    class _LiteModule(nn.Module):
        def __init__(self, m):
            super().__init__()
            self._forward_module = m
            self._original_module = m

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(4, 3)
            self.bn = nn.BatchNorm1d(3)
            self.register_buffer("pos", torch.arange(5.0))

    _LiteModule.__module__ = "pytorch_lightning.lite.wrappers"; _LiteModule.__qualname__ = "_LiteModule"
    Inner.__module__ = "src_models_fake"; Inner.__qualname__ = "Inner"
    wr._LiteModule = _LiteModule; refmod.Inner = Inner
    names = {"pytorch_lightning": pl, "pytorch_lightning.lite": lite, "pytorch_lightning.lite.wrappers": wr,
             "src_models_fake": refmod}
    sys.modules.update(names)
    try:
        inner = Inner()
        obj = _LiteModule(nn.DataParallel(inner))
        path = str(tmp_path / "multimodal_fake.pt")
        torch.save(obj, path)
        want = {k: v.clone() for k, v in inner.state_dict().items()}
    finally:
        for k in names:
            sys.modules.pop(k, None)
    with pytest.raises(RuntimeError, match="trust_checkpoint"):
        ck.load_state_dict_file(path)                       # default: no code execution
    got = ck.load_state_dict_file(path, trust=True)         # neither Lightning nor the model class is importable now
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # a plain state_dict file needs no trust
    p2 = str(tmp_path / "sd.pt")
    torch.save({"module." + k: v for k, v in want.items()}, p2)
    got2 = ck.load_state_dict_file(p2)
    assert set(got2) == set(want)


def test_literal_batch_repack_reproduces_the_reference_bug():
    """SURVEY F7 toy batch n = [4, 3, 5], every frame kept: utterance 1 receives 2 of its 3 frames with local indices
    shifted by +1, utterance 2 is shifted by +2, 10 of 12 frames are consumed (train.py:192-213)."""
    from facialmmt_b200.evaluate import literal_batch_repack
    counts, prob_row, vision_row = literal_batch_repack([4, 3, 5], list(range(12)), Lv=160)
    assert counts == [4, 2, 4]
    assert prob_row == [[0, 1, 2, 3], [4, 5], [6, 7, 8, 9]]
    assert vision_row == [[0, 1, 2, 3], [1, 2], [1, 2, 3, 4]]
    # batch size 1 is exact
    assert literal_batch_repack([5], [0, 2, 4], 160) == ([3], [[0, 2, 4]], [[0, 2, 4]])
