"""CUDA-graph replay of repeated identical forwards (fmmt_set_graph): the replayed step must produce the same bits as the
directly launched one, and must read the CURRENT contents of the argument buffers (same pointers, new data)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graph_replay_is_bit_identical_and_reads_live_buffers():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.evaluate import evaluate_batch
    from facialmmt_b200.models import MultiModalTransformerForClassification, SwinForAffwildClassification
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    swin = SwinForAffwildClassification(cfg)
    swin.load_state_dict(syn.swin_cls_stress_state_dict(cfg.swin, 1111))
    mm = MultiModalTransformerForClassification(cfg)
    mm.load_state_dict(syn.multimodal_stress_state_dict(cfg, 1111))
    U = 2
    b = syn.synthetic_batch(cfg, U=U, L=128, seed=5, n_frames=[9, 14], with_faces=True)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    n = [int(x) for x in b["num_imgs"]]

    def step():
        batch = (dev["text_ids"], dev["text_mask"], dev["sep_mask"], dev["audio"], dev["audio_mask"], dev["vision"],
                 dev["vision_mask"], torch.zeros(U, dtype=torch.long), dev["faces"], n, dev["idx_in_dia"])
        return evaluate_batch(swin, mm, batch, cfg.threshold, gumbel=dev["gumbel"]).clone()

    ref1 = step()
    launches_direct = None
    lib = swin._lib
    c0 = lib.fmmt_launch_count(); step(); launches_direct = lib.fmmt_launch_count() - c0
    swin.set_graph(True); mm.set_graph(True)
    outs = [step() for _ in range(4)]            # direct, captured, replayed, replayed
    swin.check(); mm.check()
    for o in outs:
        assert torch.equal(o, ref1)
    c0 = lib.fmmt_launch_count(); step(); launches_graph = lib.fmmt_launch_count() - c0
    assert launches_graph == launches_direct      # the launch counter keeps counting kernels, not graph launches
    # new data behind the same pointers
    dev["audio"].mul_(0.5)
    dev["faces"].add_(0.05)
    got = step()
    swin.check(); mm.check()
    swin.set_graph(False); mm.set_graph(False)
    ref2 = step()
    assert torch.equal(got, ref2)
    assert not torch.equal(ref2, ref1)


def test_parallel_branches_match_the_single_stream_forward(monkeypatch):
    """The fusion forward runs the audio / vision encoders beside the text encoder and the two directions of each
    CrossmodalTransformer pair on side streams (Engine::fork_to). Every kernel computes the same values whatever runs beside
    it, so the logits must equal the single-stream forward (FMMT_NO_BRANCHES=1) bit for bit - eagerly and under graph replay,
    repeatedly (a missing dependency or aliased scratch buffer shows up as a difference on some repetition)."""
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.models import MultiModalTransformerForClassification
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    sd = syn.multimodal_stress_state_dict(cfg, 1111)
    U = 8
    b = syn.synthetic_batch(cfg, U=U, L=128, seed=9, n_frames=[160] * U, with_faces=False)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    v519 = torch.cat([dev["vision"], torch.rand(U, dev["vision"].shape[1], cfg.fusion.num_labels, device="cuda")], dim=-1).contiguous()

    def run(model):
        return model(dev["text_ids"], dev["text_mask"], dev["sep_mask"], dev["audio"], dev["audio_mask"], v519,
                     dev["vision_mask"], dev["idx_in_dia"]).clone()

    monkeypatch.setenv("FMMT_NO_BRANCHES", "1")
    single = MultiModalTransformerForClassification(cfg)
    single.load_state_dict(sd)
    ref = run(single)
    single.check()
    monkeypatch.delenv("FMMT_NO_BRANCHES")
    par = MultiModalTransformerForClassification(cfg)
    par.load_state_dict(sd)
    for rep in range(20):
        assert torch.equal(run(par), ref), f"eager repetition {rep}"
    par.set_graph(True)
    for rep in range(20):
        assert torch.equal(run(par), ref), f"graph repetition {rep}"
    par.check()


def test_alternating_swin_passes_match_the_single_stream_forward(monkeypatch):
    """Consecutive Swin passes alternate between two side streams (Engine::swin_body) and re-use each other's arena regions
    two passes later; the result must equal the single-stream forward bit for bit, eagerly and under graph replay."""
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from facialmmt_b200.models import SwinForAffwildClassification
    cfg = FmmtConfig(text=TextConfig.roberta_large(2))
    sd = syn.swin_cls_stress_state_dict(cfg.swin, 1111)
    F = 23                                                        # passes of 5 frames: 5 passes, the last one ragged
    g = torch.Generator().manual_seed(3)
    frames = (torch.rand(F, 3, 224, 224, generator=g) * 2 - 1).cuda()
    gum = -torch.log(torch.empty(F, 7).exponential_(generator=g)).cuda()
    monkeypatch.setenv("FMMT_NO_BRANCHES", "1")
    single = SwinForAffwildClassification(cfg, swin_chunk=3, swin_chunk_late=5)
    single.load_state_dict(sd)
    ref = [t.clone() for t in single.forward_full(frames, gum)]
    single.check()
    monkeypatch.delenv("FMMT_NO_BRANCHES")
    par = SwinForAffwildClassification(cfg, swin_chunk=3, swin_chunk_late=5)
    par.load_state_dict(sd)
    for graph in (False, True):
        par.set_graph(graph)
        for rep in range(10):
            out = par.forward_full(frames, gum)
            for a, b in zip(out, ref):
                assert torch.equal(a, b), f"graph={graph} repetition {rep}"
    par.check()
