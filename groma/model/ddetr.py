"""`groma.model.ddetr` surface: CustomDDETRConfig / CustomDDETRModel (reference groma/model/ddetr.py:48-200), backed by
the B200 engine.  CustomDDETRModel.forward(images) runs DINOv2 + the Deformable-DETR proposer and returns the
detector outputs the reference's callers read (eval/run_ddetr.py:49-50): logits {'coco','sa1b'} [B,300,1] and
pred_boxes [B,300,4]."""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from transformers import AutoConfig, AutoModel, DeformableDetrConfig, Dinov2Config, PretrainedConfig

from groma_b200.config import PathConfig


class CustomDDETRConfig(PretrainedConfig):
    model_type = "ddetr"

    def __init__(self, vis_encoder_cfg=None, zs_weight_path=None, vis_output_layer=-1, ddetr_cfg=None, **kwargs):
        super().__init__(**kwargs)
        if vis_encoder_cfg is None:
            self.vis_encoder_cfg = Dinov2Config()
        elif isinstance(vis_encoder_cfg, dict):
            self.vis_encoder_cfg = Dinov2Config(**vis_encoder_cfg)
        elif isinstance(vis_encoder_cfg, Dinov2Config):
            self.vis_encoder_cfg = vis_encoder_cfg
        else:
            raise NotImplementedError("currently only supports Dinov2Model as vis_encoder.")
        if ddetr_cfg is None:
            self.ddetr_cfg = DeformableDetrConfig()
        elif isinstance(ddetr_cfg, dict):
            self.ddetr_cfg = DeformableDetrConfig(**ddetr_cfg)
        elif isinstance(ddetr_cfg, DeformableDetrConfig):
            self.ddetr_cfg = ddetr_cfg
        else:
            raise NotImplementedError("currently only supports DeformableDetrTransformer as detector head.")
        self.zs_weight_path = zs_weight_path
        self.vis_output_layer = vis_output_layer

    def to_json_string(self, use_diff: bool = True) -> str:
        d = copy.deepcopy(self)
        if use_diff:
            d.vis_encoder_cfg = d.vis_encoder_cfg.to_diff_dict()
            d.ddetr_cfg = d.ddetr_cfg.to_diff_dict()
            d = d.to_diff_dict()
        else:
            d.vis_encoder_cfg = d.vis_encoder_cfg.to_dict()
            d.ddetr_cfg = d.ddetr_cfg.to_dict()
            d = d.to_dict()
        return json.dumps(d, indent=2, sort_keys=True, default=str) + "\n"


def perceiver_fields(pcfg: CustomDDETRConfig) -> dict:
    """PathConfig fields owned by the perceiver sub-config."""
    v, d = pcfg.vis_encoder_cfg, pcfg.ddetr_cfg
    if d.num_feature_levels != 1:
        raise NotImplementedError("the B200 path implements the 1-level proposer Groma ships (scripts/det_pretrain.sh)")
    if not (d.two_stage and d.with_box_refine):
        raise NotImplementedError("two_stage + with_box_refine is the only proposer variant on the path")
    return dict(patch=v.patch_size, vit_hidden=v.hidden_size, vit_layers=v.num_hidden_layers, vit_heads=v.num_attention_heads,
                vit_mlp=int(v.hidden_size * v.mlp_ratio), vit_pos_grid=v.image_size // v.patch_size, vit_ln_eps=v.layer_norm_eps,
                d_model=d.d_model, enc_layers=d.encoder_layers, dec_layers=d.decoder_layers, ddetr_heads=d.encoder_attention_heads,
                n_points=d.encoder_n_points, ddetr_ffn=d.encoder_ffn_dim, num_queries=d.two_stage_num_proposals)


@dataclass
class DetectionOutput:
    logits: Dict[str, torch.Tensor]
    pred_boxes: torch.Tensor
    loss: Optional[torch.Tensor] = None


class CustomDDETRModel(torch.nn.Module):
    """Detector-only entry (ddetr.py:169-196).  Shares the engine of a GromaModel when built through one
    (`GromaModel.perceiver`); `from_pretrained` on a detector checkpoint builds a detector-only engine."""
    config_class = CustomDDETRConfig

    def __init__(self, config: CustomDDETRConfig, engine=None):
        super().__init__()
        self.config = config
        self.engine = engine

    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "CustomDDETRModel":
        """Detector-only checkpoint written by the reference's `train_det.py` (keys `vis_encoder.*`, `input_proj.*`,
        `ddetr_transformer.*`; reference ddetr.py:98-155), as `eval/run_ddetr.py:41` loads it."""
        from groma_b200.checkpoint import ShardedStateDict, load_config_dict
        from groma_b200.config import PathConfig
        from groma_b200.engine import GromaEngine
        cd = load_config_dict(path)
        cd.pop("model_type", None)
        config = CustomDDETRConfig(**cd)
        return cls(config, engine=GromaEngine(PathConfig(**perceiver_fields(config)), ShardedStateDict(path), detector_only=True))

    def cuda(self, device=None):
        return self

    def eval(self):
        return self

    @torch.no_grad()
    def forward(self, images=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        if labels is not None:
            raise NotImplementedError("training losses are out of scope of the B200 forward path")
        if self.engine is None:
            raise RuntimeError("CustomDDETRModel has no engine bound (construct it via GromaModel.perceiver)")
        hs = self.engine.vit(images)
        pc, _, _, logits = self.engine.proposer(hs)
        q = self.engine.cfg.num_queries
        # the engine's proposer returns buffers owned by its CUDA graph: hand out copies, the caller may keep them across calls
        return DetectionOutput(logits={k: v.clone() for k, v in logits.items()}, pred_boxes=pc[:, :q].clone())


try:
    AutoConfig.register("ddetr", CustomDDETRConfig)
    AutoModel.register(CustomDDETRConfig, CustomDDETRModel)
except ValueError:
    pass
