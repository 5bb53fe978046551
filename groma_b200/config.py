"""Size / hyper-parameter bundle of the Groma forward path.  Defaults = Groma-7B:
DINOv2-L, 1-level Deformable-DETR proposer (scripts/det_pretrain.sh:12-19), 3-level region encoder, Vicuna-7B."""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass
class PathConfig:
    # DINOv2-L (facebook/dinov2-large)
    image_size: int = 448
    patch: int = 14
    vit_hidden: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    vit_pos_grid: int = 37          # trained at 518 px
    vit_ln_eps: float = 1e-6
    # Deformable-DETR proposer (scripts/det_pretrain.sh:12-19)
    d_model: int = 256
    enc_layers: int = 6
    dec_layers: int = 6
    ddetr_heads: int = 8
    n_points: int = 4
    ddetr_ffn: int = 1024
    num_queries: int = 300
    # region encoder (groma/model/roi_align.py)
    fuse_rounds: int = 5
    gn_groups: int = 64
    roi_out: int = 14
    roi_sampling: int = 2
    pos_hidden: int = 256
    region_mid: int = 1024
    # LLaMA / Vicuna-7B
    llm_hidden: int = 4096
    llm_layers: int = 32
    llm_heads: int = 32
    llm_inter: int = 11008
    vocab: int = 32000
    num_new_token: int = 114
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_pos: int = 4096
    # selection
    nms_thres: float = 0.6
    box_score_thres: float = 0.15
    max_region_num: int = 100

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def head_dim(self) -> int:
        return self.llm_hidden // self.llm_heads

    def to_dict(self):
        return asdict(self)


def tiny_config(**kw) -> PathConfig:
    """A shape-faithful miniature (same op sequence, every dimension shrunk) the CPU oracle runs in seconds.
    Constraints kept: grid 32x32 (roi_align.py:283-284 asserts it), head dims 64 (ViT) / 32 (DDETR) / 128 (LLaMA),
    GroupNorm channels-per-group multiple of 8."""
    base = dict(vit_hidden=128, vit_layers=4, vit_heads=2, vit_mlp=256, enc_layers=2, dec_layers=6, d_model=64,
                ddetr_heads=2, ddetr_ffn=128, num_queries=60, fuse_rounds=2, gn_groups=8, region_mid=128, pos_hidden=64,
                llm_hidden=256, llm_layers=2, llm_heads=2, llm_inter=512, vocab=1000, max_pos=2048, max_region_num=20)
    base.update(kw)
    return PathConfig(**base)


# token ids of the synthetic tokenizer: the 14 DEFAULT_TOKENS missing from the LLaMA vocab, then <r0>..<r99>
# (groma/constants.py:5-25, groma/train/train.py:90-91).
NEW_TOKENS = ["[PAD]", "<sep>", "<img>", "</img>", "<roi>", "</roi>", "<p>", "</p>", "<image>", "<region>",
              "<refer_box>", "<ground_box>", "<refer_feat>", "[grounding]"] + [f"<r{i}>" for i in range(100)]


class SyntheticTokenizer:
    """Just enough of a tokenizer for GromaModel.init_special_token_id (groma/model/groma.py:136-144)."""

    def __init__(self, vocab: int):
        self.map = {t: vocab + i for i, t in enumerate(NEW_TOKENS)}
        self.pad_token_id = self.map["[PAD]"]

    def convert_tokens_to_ids(self, toks):
        return [self.map[t] for t in toks]
