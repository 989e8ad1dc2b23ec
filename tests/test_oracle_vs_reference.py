"""CPU, build container only: live comparison of the oracle with the reference modules imported from
/root/reference (skipped where the reference is absent, e.g. on the GPU box). Also checks that the state-dict key
listing in facialmmt_b200/synthetic.py equals the instantiated reference's."""
import pytest
import torch

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not present")


def test_swin_block_level_and_keys():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig()
    m = rh.build_swin_cls()
    ref_sd = m.state_dict()
    spec = syn.swin_cls_state_dict_spec(cfg.swin)
    assert set(spec) == set(ref_sd)
    for k, shp in spec.items():
        assert tuple(ref_sd[k].shape) == tuple(shp), k
    sd = syn.swin_cls_stress_state_dict(cfg.swin, 7)
    for k in sd:   # registered buffers must be reproduced exactly
        if k.endswith("relative_position_index") or k.endswith("attn_mask"):
            assert torch.equal(sd[k].to(ref_sd[k].dtype), ref_sd[k]), k
    m.load_state_dict(sd)
    frames = syn.synthetic_faces(2, 3)
    col = {}
    with torch.no_grad():
        x = m.swin.patch_embed(frames)
        orc.swin_features(sd, frames, collect=col)
        assert (x - col["patch_embed"]).abs().max() < 1e-5
        for li, layer in enumerate(m.swin.layers):
            for bi, blk in enumerate(layer.blocks):
                x = blk(x)
                err = (x - col[f"layer{li}.block{bi}"]).abs().max().item()
                assert err < 2e-4 * max(1.0, x.abs().max().item()), (li, bi, err)
            if layer.downsample is not None:
                x = layer.downsample(x)


def test_multimodal_keys_and_cmt_module():
    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig, TextConfig
    from oracle import facialmmt_oracle as orc
    cfg = FmmtConfig(text=TextConfig.roberta_large(1))
    mm = rh.build_multimodal(rh.default_args("roberta-large"), text_layers=1)
    ref_sd = mm.state_dict()
    spec = syn.multimodal_state_dict_spec(cfg)
    assert set(spec) <= set(ref_sd)
    assert all("pooler" in k for k in set(ref_sd) - set(spec))
    sd = syn.multimodal_stress_state_dict(cfg, 5)
    mm.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(1)
    q = torch.randn(2, 38, 768, generator=g)
    q[1, 20:] = 0            # zero rows -> pad positions in the sinusoidal embedding
    kv = torch.randn(2, 160, 768, generator=g)
    with torch.no_grad():
        ref = mm.CrossModalTrans_TA(q.transpose(0, 1), kv.transpose(0, 1), kv.transpose(0, 1)).transpose(0, 1)
    got = orc.cmt_encoder(sd, "CrossModalTrans_TA.", q, kv)
    assert (ref - got).abs().max() < 2e-4


def test_bug_compat_repack_equals_the_literal_reference_loop_at_batch_3():
    """--bug_compat: facialmmt_b200.evaluate._filter_pack_bug_compat against the reference's OWN multimodal_evaluate loop
    (train.py:154-243, exec'd from source) at trg_batch_size = 3, where the literal code drops / shifts frames."""
    import argparse

    from facialmmt_b200 import synthetic as syn
    from facialmmt_b200.config import FmmtConfig
    from facialmmt_b200.evaluate import _filter_pack_bug_compat
    rh.install_shims()
    cfg = FmmtConfig()
    n = [37, 5, 160]
    b = syn.synthetic_batch(cfg, U=3, L=16, seed=77, n_frames=n, with_faces=False)
    for sharp in (2.0, 0.0):
        probs = torch.softmax(sharp * torch.randn(sum(n), 7, generator=torch.Generator().manual_seed(9)), -1)

        class Swin(torch.nn.Module):
            def forward(self, x, is_trg_task=None):
                return probs

        class Cap(torch.nn.Module):
            def forward(self, *a):
                self.got = [t.clone() for t in a]
                return torch.zeros(a[5].shape[0], 7)

        cap = Cap()
        args = argparse.Namespace(trg_batch_size=3, FacialEmoImpor_threshold=0.2, num_labels=7, trg_n_test=3, trg_n_valid=3)
        faces = torch.zeros(3, 160, 3, 2, 2)
        batch = (b["text_ids"], b["text_mask"], b["sep_mask"], b["audio"], b["audio_mask"], b["vision"], b["vision_mask"],
                 torch.zeros(3, dtype=torch.long), faces, b["num_imgs"], b["idx_in_dia"])
        with torch.no_grad():
            rh.literal_eval_loop()(args, [batch])(Swin(), cap, lambda lo, y: torch.zeros(()), test=True)
        v519, mask = _filter_pack_bug_compat(b["vision"], b["vision_mask"], n, probs, 0.2)
        assert torch.equal(cap.got[5], v519), sharp
        assert torch.equal(cap.got[6].float(), mask), sharp
