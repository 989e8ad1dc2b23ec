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
