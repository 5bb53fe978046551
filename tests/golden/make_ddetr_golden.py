"""Run the REFERENCE's own `DeformableDetrTransformer` (`/root/reference/groma/model/ddetr_transformer.py`, unmodified source:
extract_feature -> two-stage proposals -> DeformableDetrDecoderX -> heads, lines 484-728 and 77-202) on CPU and record its
outputs as a golden fixture for the oracle's proposer.

The file imports transformers==4.32.0 internals that the installed 5.x no longer has, so before the import the module path
`transformers.models.deformable_detr.modeling_deformable_detr` is pointed at a thin adapter that re-exports the 5.x classes
under their 4.32 call signatures (argument renames only -- position_embeddings / spatial_shapes_list / tuple returns -- plus the
4.32 helpers `_get_clones`, `build_position_encoding`); `mmcv.ops.bbox` (training-time NMS helper) is an empty stand-in.  The
5.x layer classes themselves are pinned against the oracle separately (tests/test_oracle_hf_pins_cpu.py); what THIS fixture
pins is the reference's wiring: level embedding + flattening, two-stage top-k and its `bbox_embed[-1]` head, `pos_trans` query
embeddings with the learned target, reference boxes that never advance through the decoder (T4), the recorded
`new_reference_points` feeding the final box refinement, and the coco / sa1b heads of the last level.

    python tests/golden/make_ddetr_golden.py          # writes tests/golden/ddetr_transformer_ref.pt
"""
import copy
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
MODPATH = "transformers.models.deformable_detr.modeling_deformable_detr"


def compat_module():
    from transformers.models.deformable_detr import modeling_deformable_detr as M
    c = types.ModuleType(MODPATH)
    for n in ("DeformableDetrPreTrainedModel", "DeformableDetrConfig", "DeformableDetrMLPPredictionHead", "DeformableDetrModelOutput",
              "DeformableDetrObjectDetectionOutput", "DeformableDetrDecoderOutput", "inverse_sigmoid"):
        setattr(c, n, getattr(M, n))

    def shapes_list(spatial_shapes):
        return [tuple(int(v) for v in s) for s in spatial_shapes.tolist()]

    class DeformableDetrEncoder(M.DeformableDetrEncoder):
        def forward(self, inputs_embeds=None, attention_mask=None, position_embeddings=None, spatial_shapes=None, level_start_index=None,
                    valid_ratios=None, output_attentions=None, output_hidden_states=None, return_dict=None):
            return super().forward(inputs_embeds=inputs_embeds, attention_mask=attention_mask, spatial_position_embeddings=position_embeddings,
                                   spatial_shapes=spatial_shapes, spatial_shapes_list=shapes_list(spatial_shapes),
                                   level_start_index=level_start_index, valid_ratios=valid_ratios)

    class DeformableDetrDecoderLayer(M.DeformableDetrDecoderLayer):
        def forward(self, hidden_states, position_embeddings=None, reference_points=None, spatial_shapes=None, level_start_index=None,
                    encoder_hidden_states=None, encoder_attention_mask=None, output_attentions=False):
            out = super().forward(hidden_states, object_queries_position_embeddings=position_embeddings, reference_points=reference_points,
                                  spatial_shapes=spatial_shapes, spatial_shapes_list=shapes_list(spatial_shapes),
                                  level_start_index=level_start_index, encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask)
            return out if isinstance(out, tuple) else (out,)

    class SinePosition432(nn.Module):           # 4.32: position_encoding(pixel_values, pixel_mask) -> [B, D, H, W]
        def __init__(self, d_model):
            super().__init__()
            self.inner = M.DeformableDetrSinePositionEmbedding(d_model // 2, normalize=True)

        def forward(self, pixel_values, pixel_mask):
            B, D, H, W = pixel_values.shape
            pos = self.inner(pixel_values.shape, pixel_values.device, pixel_values.dtype, mask=pixel_mask)
            return pos.reshape(B, H, W, D).permute(0, 3, 1, 2) if pos.dim() == 3 else pos

    class _Unused(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    c.DeformableDetrEncoder, c.DeformableDetrDecoderLayer, c.DeformableDetrDecoder = DeformableDetrEncoder, DeformableDetrDecoderLayer, M.DeformableDetrDecoder
    c.DeformableDetrHungarianMatcher, c.DeformableDetrLoss = _Unused, _Unused               # training only
    # only used for `isinstance(m, ...)` -> m._reset_parameters() at construction (5.x dropped that initialiser); every
    # parameter is overwritten by load_state_dict afterwards, so no instance needs to match
    c.DeformableDetrMultiscaleDeformableAttention = _Unused
    c.generalized_box_iou = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("training only"))
    c._get_clones = lambda module, N: nn.ModuleList([copy.deepcopy(module) for _ in range(N)])   # 4.32 definition
    c.build_position_encoding = lambda config: SinePosition432(config.d_model)              # 4.32: sine, d_model // 2 feats, normalize
    return c


def load_reference_module():
    bbox = types.ModuleType("mmcv.ops.bbox")
    bbox.bbox_overlaps = None
    ops, mmcv = types.ModuleType("mmcv.ops"), types.ModuleType("mmcv")
    ops.bbox, mmcv.ops = bbox, ops
    inject = {MODPATH: compat_module(), "mmcv": mmcv, "mmcv.ops": ops, "mmcv.ops.bbox": bbox}
    saved = {k: sys.modules.get(k) for k in inject}
    sys.modules.update(inject)
    try:
        spec = importlib.util.spec_from_file_location("ref_ddetr_transformer", "/root/reference/groma/model/ddetr_transformer.py")
        ref = importlib.util.module_from_spec(spec)
        sys.modules["ref_ddetr_transformer"] = ref          # transformers looks the defining module of a model class up by name
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


def case():
    """d_model 256 (the reference hard-codes the 256-wide box MLP), 8 heads x 32, 4 points, 1 level of 32x32, 2 + 3 layers."""
    from groma_b200.config import tiny_config
    from groma_b200.synth import make_state_dict
    cfg = tiny_config(vit_hidden=64, vit_heads=1, vit_mlp=128, vit_layers=4, d_model=256, ddetr_heads=8, ddetr_ffn=96, enc_layers=2,
                      dec_layers=3, num_queries=40, llm_hidden=128, llm_heads=1, llm_layers=1, llm_inter=64, vocab=64, region_mid=64,
                      pos_hidden=32)
    sd = make_state_dict(cfg, seed=9, perturb_norms=True)
    g = torch.Generator().manual_seed(13)
    hs = [torch.randn(2, cfg.grid * cfg.grid + 1, cfg.vit_hidden, generator=g) for _ in range(5)]
    return cfg, sd, hs


RENAME = [(".fc1.", ".mlp.fc1."), (".fc2.", ".mlp.fc2."), ("self_attn.out_proj", "self_attn.o_proj")]   # 4.32 -> 5.x submodule names


def run_reference(src):
    """src: [B, S, d_model] tokens after input_proj (the oracle's `ddetr_src` stage), 32x32 grid, all pixels valid."""
    from groma.model.groma import _ddetr_cfg
    ref = load_reference_module()
    cfg, sd, _ = case()
    dc = _ddetr_cfg(cfg)
    dc.num_labels = 1                                     # scripts/det_pretrain.sh: --num_classes 1
    torch.manual_seed(0)
    m = ref.DeformableDetrTransformer(dc).eval()
    pfx = "perceiver.ddetr_transformer."
    own = {}
    for k, v in sd.items():
        if k.startswith(pfx):
            n = k[len(pfx):]
            for a, b in RENAME:
                n = n.replace(a, b)
            own[n] = v.float()
    res = m.load_state_dict(own, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("decoder.bbox_embed.", "decoder.class_embed_")) for k in res.missing_keys), res.missing_keys   # aliases of the heads
    B, S, D = src.shape
    g = cfg.grid
    sources = [src.transpose(1, 2).reshape(B, D, g, g)]
    masks = [torch.ones(B, g, g, dtype=torch.bool)]
    with torch.no_grad():
        out = m(sources, masks, return_dict=True)
    return dict(pred_boxes=out.pred_boxes, coco=out.logits["coco"], sa1b=out.logits["sa1b"], init_reference=out.init_reference_points,
                enc_class=out.enc_outputs_class, memory=out.encoder_last_hidden_state,
                intermediate_reference_points=out.intermediate_reference_points, last_hidden=out.last_hidden_state)


def oracle_src():
    from oracle.groma_oracle import Oracle
    cfg, sd, hs = case()
    o = Oracle(cfg, sd, "fp32")
    pred, scores, logits = o.proposer(hs)
    return o, pred, scores, logits


if __name__ == "__main__":
    o, _, _, _ = oracle_src()
    out = run_reference(o.stages["ddetr_src"])
    out["memory"] = out["memory"][:, ::32].clone()       # every 32nd token keeps the fixture small
    torch.save({"outputs": out, "note": "reference DeformableDetrTransformer outputs on the oracle's ddetr_src of case()"},
               os.path.join(HERE, "ddetr_transformer_ref.pt"))
    print("wrote ddetr_transformer_ref.pt", {k: tuple(v.shape) for k, v in out.items()})
