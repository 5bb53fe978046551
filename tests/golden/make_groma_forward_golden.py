"""Run the REFERENCE's own `GromaModel.forward` prefill branch (`/root/reference/groma/model/groma.py:202-427`, unmodified
source, together with its own `ddetr.py`, `ddetr_transformer.py`, `roi_align.py`, `constants.py` and the vendored
`mmcv/ops/nms.py` wrapper) on CPU, and record what it returns as the golden fixture for the oracle's `forward_prefill`.

Adapters (leaf dependencies that cannot be imported in this image; none of them replaces logic of the reference repo):
  * transformers 4.32 call signatures around the installed 5.x Deformable-DETR classes  (make_ddetr_golden.compat_module)
  * mmcv.cnn.ConvModule / Linear / normal_init, mmdet BaseRoIExtractor, RoIAlign layer     (make_region_encoder_golden)
  * mmcv._ext.nms -> the reference's own C++ CPU kernel compiled into oracle/_ref (oracle/build_ref.py); the Python wrapper
    that calls it (score filter, sort, max_num: mmcv/mmcv/ops/nms.py:14-33,119-178) is the vendored file itself
  * the patch rows of the DINOv2 position table are made constant, so that 5.x's and 4.32's different
    `interpolate_pos_encoding` both return that constant and cannot influence the comparison
What this pins: the whole orchestration -- hidden-state selection, 2x2 space-to-depth order + bridge, mean of the last four
states + input_proj + channel LayerNorm, score fusion, box selection (NMS wrapper semantics, randperm on the global RNG,
argmax fallback), refer/ground matching and in-place id edits, region encoder call (hidden_states[-3:], CLS dropped),
placeholder expansion, truncation at the first pad, re-padding, the three masked_scatter_ splices, lm_head || extra_lm_head.

    python tests/golden/make_groma_forward_golden.py          # writes tests/golden/groma_forward_ref.pt
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _helper(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _load_as(modname, path, package=None):
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    if package:
        m.__package__ = package
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


def load_reference_stack():
    """Import the reference's groma.{constants, model.ddetr_transformer, model.roi_align, model.ddetr, model.groma} and
    mmcv.ops.nms under their own names, with the adapters in place.  Returns (groma.py module, modules to restore)."""
    reg, det = _helper("make_region_encoder_golden"), _helper("make_ddetr_golden")
    from oracle.build_ref import load_ref
    ext = load_ref()
    if ext is None:
        raise RuntimeError("oracle/_ref is not built (python -m oracle.build_ref)")
    saved = dict(sys.modules)
    for k in [k for k in sys.modules if k == "groma" or k.startswith("groma.")]:
        del sys.modules[k]                       # the repo's own `groma` shim must not shadow the reference package
    pkg = lambda name: types.ModuleType(name)
    mmcv, cnn, ops, utils, bbox = pkg("mmcv"), pkg("mmcv.cnn"), pkg("mmcv.ops"), pkg("mmcv.utils"), pkg("mmcv.ops.bbox")
    ops.__path__, mmcv.__path__ = [], []
    cnn.ConvModule, cnn.Linear = reg.ConvModule, nn.Linear
    cnn.normal_init = lambda m, mean=0, std=1, bias=0: nn.init.normal_(m.weight, mean, std)
    utils.deprecated_api_warning = lambda *a, **k: (lambda f: f)
    ext_loader = pkg("mmcv.utils.ext_loader")
    ext_loader.load_ext = lambda name, funcs: types.SimpleNamespace(nms=lambda boxes, scores, iou_threshold, offset: ext.nms(boxes, scores, float(iou_threshold), int(offset)),
                                                                    **{f: None for f in funcs if f != "nms"})
    utils.ext_loader = ext_loader
    bbox.bbox_overlaps = None
    mmcv.cnn, mmcv.ops, mmcv.utils, ops.bbox = cnn, ops, utils, bbox
    mmdet, models = pkg("mmdet"), pkg("mmdet.models")
    models.BaseRoIExtractor = reg.BaseRoIExtractor
    mmdet.models = models
    sys.modules.update({det.MODPATH: det.compat_module(), "mmcv": mmcv, "mmcv.cnn": cnn, "mmcv.ops": ops, "mmcv.utils": utils,
                        "mmcv.utils.ext_loader": ext_loader, "mmcv.ops.bbox": bbox, "mmdet": mmdet, "mmdet.models": models})
    _load_as("mmcv.ops.nms", f"{REF}/mmcv/mmcv/ops/nms.py", package="mmcv.ops")
    ops.nms = sys.modules["mmcv.ops.nms"]
    g, gm = pkg("groma"), pkg("groma.model")
    g.__path__, gm.__path__ = [], []
    sys.modules.update({"groma": g, "groma.model": gm})
    _load_as("groma.constants", f"{REF}/groma/constants.py")
    _load_as("groma.model.ddetr_transformer", f"{REF}/groma/model/ddetr_transformer.py")
    _load_as("groma.model.roi_align", f"{REF}/groma/model/roi_align.py")
    _load_as("groma.model.ddetr", f"{REF}/groma/model/ddetr.py")
    ref = _load_as("groma.model.groma", f"{REF}/groma/model/groma.py")
    return ref, saved


def restore(saved):
    for k in list(sys.modules):
        if k not in saved:
            del sys.modules[k]
    sys.modules.update(saved)


def case(box_score_thres=0.05):
    """Widths the reference hard-codes (box MLP 256, region encoder 256/1024/4096, 64 GN groups) with everything else small."""
    from groma_b200.config import SyntheticTokenizer, tiny_config
    from groma_b200.synth import make_state_dict
    cfg = tiny_config(vit_hidden=64, vit_heads=1, vit_mlp=128, vit_layers=4, d_model=256, ddetr_heads=8, ddetr_ffn=96, enc_layers=2,
                      dec_layers=3, num_queries=40, gn_groups=64, fuse_rounds=5, pos_hidden=256, region_mid=1024, llm_hidden=4096,
                      llm_heads=32, llm_layers=1, llm_inter=64, vocab=200, max_region_num=12, box_score_thres=box_score_thres, nms_thres=0.6)
    sd = make_state_dict(cfg, seed=21, perturb_norms=True)
    pe = sd["perceiver.vis_encoder.embeddings.position_embeddings"]
    pe[:, 1:] = pe[:, 1:2]                            # constant over the patch grid (see the module docstring)
    tok = SyntheticTokenizer(cfg.vocab)
    g = torch.Generator().manual_seed(23)
    B, T = 2, 22
    images = torch.randn(B, 3, 448, 448, generator=g)
    ids = torch.randint(10, cfg.vocab, (B, T), generator=g)
    ids[:, 2] = tok.map["<image>"]
    ids[:, 6] = tok.map["<region>"]
    ids[0, 9] = tok.map["<refer_box>"]; ids[0, 10] = tok.map["<refer_feat>"]
    ids[1, 12] = tok.map["<ground_box>"]
    ids[1, 17:] = tok.pad_token_id                   # ragged: row 1 is right-padded
    refer = [torch.tensor([[0.30, 0.40, 0.20, 0.25]]), torch.zeros(0, 4)]
    ground = [torch.zeros(0, 4), torch.tensor([[0.60, 0.55, 0.30, 0.20]])]
    return cfg, sd, tok, images, ids, refer, ground


def plain_ids(ids, tok):
    ids = ids.clone()
    for name in ("<refer_box>", "<refer_feat>", "<ground_box>"):
        ids[ids == tok.map[name]] = 11
    return ids


def run_reference(seed=1234, box_score_thres=0.05, with_user_boxes=True):
    cfg, sd, tok, images, ids, refer, ground = case(box_score_thres)
    if not with_user_boxes:      # plain prompt: no <refer_box>/<refer_feat>/<ground_box>, no user boxes
        ids = plain_ids(ids, tok)
        refer = ground = None
    import transformers as tr
    from groma.model.groma import _ddetr_cfg as shim_ddetr_cfg     # the repo's helper: builds a DeformableDetrConfig offline
    dc = shim_ddetr_cfg(cfg)
    dc.num_labels = 1
    ref, saved = load_reference_stack()
    try:
        det_mod = sys.modules["groma.model.ddetr"]
        vis = tr.Dinov2Config(hidden_size=cfg.vit_hidden, num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                              mlp_ratio=cfg.vit_mlp // cfg.vit_hidden, image_size=cfg.vit_pos_grid * cfg.patch, patch_size=cfg.patch,
                              layer_norm_eps=cfg.vit_ln_eps, attn_implementation="eager")
        llm = tr.LlamaConfig(hidden_size=cfg.llm_hidden, num_hidden_layers=cfg.llm_layers, num_attention_heads=cfg.llm_heads,
                             num_key_value_heads=cfg.llm_heads, intermediate_size=cfg.llm_inter, vocab_size=cfg.vocab,
                             rms_norm_eps=cfg.rms_eps, max_position_embeddings=cfg.max_pos, attn_implementation="eager")
        gc = ref.GromaConfig(llm_cfg=llm, perceiver_cfg=det_mod.CustomDDETRConfig(vis_encoder_cfg=vis, ddetr_cfg=dc),
                             num_new_token=cfg.num_new_token, nms_thres=cfg.nms_thres, box_score_thres=cfg.box_score_thres,
                             max_region_num=cfg.max_region_num)
        torch.manual_seed(0)
        m = ref.GromaModel(gc).eval()
        det = _helper("make_ddetr_golden")
        own = {}
        for k, v in sd.items():
            n = k
            if k.startswith("perceiver.ddetr_transformer."):
                for a, b in det.RENAME:
                    n = n.replace(a, b)
            own[n] = v.float()
        res = m.load_state_dict(own, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        ok_missing = ("perceiver.ddetr_transformer.decoder.bbox_embed.", "perceiver.ddetr_transformer.decoder.class_embed_",
                      "perceiver.vis_encoder.embeddings.mask_token", "perceiver.vis_encoder.layernorm.")
        assert all(k.startswith(ok_missing) or "rotary_emb" in k for k in res.missing_keys), res.missing_keys
        m.init_special_token_id(tok)
        ids_in = ids.clone()
        torch.manual_seed(seed)
        with torch.no_grad():
            out = m(input_ids=ids_in, images=images, refer_boxes=[r.clone() for r in refer] if refer is not None else None,
                    ground_boxes=[g.clone() for g in ground] if ground is not None else None,
                    labels=ids.clone(), use_cache=True, return_dict=True)
            # decode branch (groma.py:376-379): the reference reads `past_key_values[0][0].shape`, i.e. 4.32's tuple cache; give
            # the 5.x cache object that one accessor, nothing else
            pkv = out.past_key_values
            pkv.__class__ = type("IndexableCache", (type(pkv),), {"__getitem__": lambda self, i: (self.layers[i].keys, self.layers[i].values)})
            dec_logits, tokens = [], []
            tok_ids = out.logits[:, -1].argmax(-1)                        # HF greedy: last (padded) position of every row
            for _ in range(2):
                tokens.append(tok_ids.clone())
                T_now = pkv[0][0].shape[-2]
                step = m(input_ids=tok_ids[:, None], past_key_values=pkv, attention_mask=torch.ones(ids.shape[0], T_now + 1),
                         use_cache=True, return_dict=True)
                dec_logits.append(step.logits[:, 0].clone())
                tok_ids = step.logits[:, 0].argmax(-1)
            det = m.perceiver(images, return_dict=True)                # detector-only entry, ddetr.py:169-196 (eval/run_ddetr.py:49-50)
        vis_out = out.hidden_states[1]
        return dict(det_pred_boxes=det.pred_boxes, det_coco=det.logits["coco"], det_sa1b=det.logits["sa1b"],
                    logits=out.logits, loss=out.loss, input_ids_after=ids_in, selected_boxes=[b.clone() for b in vis_out["pred_boxes"]],
                    image_features=vis_out["image_features"], region_features=vis_out["region_features"],
                    decode_tokens=torch.stack(tokens, 1), decode_logits=torch.stack(dec_logits, 1))
    finally:
        restore(saved)


if __name__ == "__main__":
    out = run_reference()
    keep = dict(out)
    keep["logits"] = out["logits"][:, :, ::7].clone()          # every 7th vocabulary column keeps the fixture small
    keep["image_features"] = out["image_features"][:, ::16, ::16].clone()
    keep["region_features"] = out["region_features"][:, ::16].clone()
    keep["decode_logits"] = out["decode_logits"][:, :, ::7].clone()
    # second scenario: score threshold above every proposal score -> NMS returns nothing -> argmax-box fallback (groma.py:276-278)
    fb = run_reference(box_score_thres=0.9999, with_user_boxes=False)
    keep["fallback"] = dict(selected_boxes=fb["selected_boxes"], logits=fb["logits"][:, :, ::7].clone(), loss=fb["loss"],
                            decode_tokens=fb["decode_tokens"])
    torch.save({"outputs": keep, "note": "reference GromaModel.forward (prefill) outputs on case(); see this script"},
               os.path.join(HERE, "groma_forward_ref.pt"))
    print("wrote groma_forward_ref.pt", {k: tuple(v.shape) for k, v in keep.items() if torch.is_tensor(v)}, "fallback T =", keep["fallback"]["logits"].shape[1])
