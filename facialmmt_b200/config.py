"""Static description of the model family on the path (dimensions only; no arithmetic).

Values mirror the reference: modules/SwinTransformer/swin_conf.yaml:4-22, main.py:62-82 (fusion flags), and the
HF `roberta-large` / `bert-large-uncased` configs the reference loads at src/models.py:73,76.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Tuple


@dataclass
class SwinConfig:
    img_size: int = 224
    patch_size: int = 4
    in_chans: int = 3
    embed_dim: int = 96
    depths: Tuple[int, ...] = (2, 2, 6, 2)
    num_heads: Tuple[int, ...] = (3, 6, 12, 24)
    window_size: int = 7
    mlp_ratio: float = 4.0
    feat_dim: int = 512       # output_layer Linear(49*768, 512)  (Swin_Transformer.py:493)
    head_hidden: int = 64     # src/models.py:21
    num_labels: int = 7

    @property
    def resolutions(self):
        r = self.img_size // self.patch_size
        return tuple(r // (2 ** i) for i in range(len(self.depths)))

    @property
    def dims(self):
        return tuple(self.embed_dim * (2 ** i) for i in range(len(self.depths)))


@dataclass
class TextConfig:
    kind: str = "roberta"          # "roberta" | "bert"   (src/models.py:49-52)
    vocab_size: int = 50265
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    max_pos: int = 514
    type_vocab: int = 1
    pad_id: int = 1
    eps: float = 1e-5

    @staticmethod
    def roberta_large(layers: int = 24) -> "TextConfig":
        return TextConfig("roberta", 50265, 1024, layers, 16, 4096, 514, 1, 1, 1e-5)

    @staticmethod
    def bert_large(layers: int = 24) -> "TextConfig":
        return TextConfig("bert", 30522, 1024, layers, 16, 4096, 512, 2, 0, 1e-12)


@dataclass
class FusionConfig:
    hidden: int = 768
    heads: int = 12
    ffn: int = 3072
    eps: float = 1e-12             # main.py:83 layer_norm_eps (TF-style LN)
    audio_dim: int = 768
    vision_dim: int = 512          # + num_labels emotion columns -> vision_linear (768, 519)
    audio_layers: int = 5
    vision_layers: int = 2
    cmt_layers_ta: int = 2
    cmt_heads_ta: int = 12
    cmt_layers_tav: int = 2
    cmt_heads_tav: int = 12
    text_len: int = 38             # TEXT_MAX_UTT_LEN (utils/dataset.py:24)
    audio_len: int = 160           # dataset-derived (main.py:134-145); synthetic default
    vision_len: int = 160
    num_labels: int = 7


@dataclass
class FmmtConfig:
    swin: SwinConfig = field(default_factory=SwinConfig)
    text: TextConfig = field(default_factory=TextConfig)
    fusion: FusionConfig = field(default_factory=FusionConfig)
    tau: float = 1.0               # main.py:41
    threshold: float = 0.2         # main.py:42 FacialEmoImpor_threshold

    def to_dict(self):
        return asdict(self)
