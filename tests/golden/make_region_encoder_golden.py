"""Run the REFERENCE's own region encoder (`/root/reference/groma/model/roi_align.py`: MLVLROIQueryModule = MLVLFuseModule +
MlvlRoIExtractor, unmodified source) on CPU and record its outputs as a golden fixture for the oracle / CUDA path.

The module's leaf dependencies are not importable in this image (mmcv needs `addict` and a compiled `_ext`, mmdet is absent),
so three small stand-ins are injected before the import -- nothing of the reference's own logic is replaced:
  * mmcv.cnn.ConvModule  -> conv(bias=False) -> GroupNorm -> ReLU with submodules `.conv` / `.gn`, i.e. what
                            `mmcv/mmcv/cnn/bricks/conv_module.py:70-85,196-206` builds for norm_cfg=dict(type='GN') (bias='auto');
    mmcv.cnn.Linear      -> nn.Linear (mmcv's wrapper only adds empty-tensor handling); normal_init -> nn.init.normal_
  * mmdet.models.BaseRoIExtractor -> builds `roi_layers` exactly as `mmdet/.../base_roi_extractor.py:54-60`
                            (spatial_scale = 1 / stride per level)
  * the RoIAlign layer   -> torchvision.ops.roi_align(aligned=True, sampling_ratio) -- the CPU twin of mmcv's CUDA kernel
                            (`roi_align_cuda_kernel.cuh:17-108`; mmcv's own CPU kernel refuses Groma's cxcywh-as-xyxy boxes, T2).
What this pins: parameter names/shapes (load_state_dict strict), level order and up-sampling sizes, the coordinate channels,
the channel shuffle (which quarter comes from which neighbour), shared ConvModule per round, the box -> RoI conversion
(box * 448 used as xyxy, T1), strides 14/8, 14/4, 14/2 (T3), pconv sum + ReLU, (c, h, w) flatten order, the box MLP, updims.

    python tests/golden/make_region_encoder_golden.py          # writes tests/golden/region_encoder_ref.pt
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None, **kw):
        super().__init__()
        assert conv_cfg is None and norm_cfg["type"] == "GN"
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False)
        self.gn = nn.GroupNorm(norm_cfg["num_groups"], out_channels)
        self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.activate(self.gn(self.conv(x)))


class RoIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale, self.sampling_ratio, self.aligned = float(spatial_scale), int(sampling_ratio), aligned
        assert pool_mode == "avg"

    def forward(self, x, rois):
        return torchvision.ops.roi_align(x, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)


class BaseRoIExtractor(nn.Module):
    def __init__(self, roi_layer, out_channels, featmap_strides, init_cfg=None):
        super().__init__()
        cfg = dict(roi_layer)
        assert cfg.pop("type") == "RoIAlign"
        self.roi_layers = nn.ModuleList([RoIAlign(spatial_scale=1 / s, **cfg) for s in featmap_strides])
        self.out_channels, self.featmap_strides = out_channels, featmap_strides


def load_reference_module():
    mmcv, cnn = types.ModuleType("mmcv"), types.ModuleType("mmcv.cnn")
    cnn.ConvModule, cnn.Linear = ConvModule, nn.Linear
    cnn.normal_init = lambda m, mean=0, std=1, bias=0: (nn.init.normal_(m.weight, mean, std), nn.init.constant_(m.bias, bias) if getattr(m, "bias", None) is not None else None)
    mmcv.cnn = cnn
    mmdet, models = types.ModuleType("mmdet"), types.ModuleType("mmdet.models")
    models.BaseRoIExtractor = BaseRoIExtractor
    mmdet.models = models
    saved = {k: sys.modules.get(k) for k in ("mmcv", "mmcv.cnn", "mmdet", "mmdet.models")}
    sys.modules.update({"mmcv": mmcv, "mmcv.cnn": cnn, "mmdet": mmdet, "mmdet.models": models})
    try:
        spec = importlib.util.spec_from_file_location("ref_roi_align", "/root/reference/groma/model/roi_align.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


def case():
    """Shapes: the reference's code constants (5 rounds, 64 GN groups, 14x14 RoIs, 256/1024/4096 widths) with embed_dims 64."""
    from groma_b200.config import tiny_config
    from groma_b200.synth import make_state_dict
    cfg = tiny_config(vit_hidden=64, vit_heads=1, vit_mlp=128, gn_groups=64, fuse_rounds=5, pos_hidden=256, region_mid=1024,
                      llm_hidden=4096, llm_heads=32, llm_layers=1, llm_inter=64, vocab=64)
    sd = make_state_dict(cfg, seed=7, perturb_norms=True)
    g = torch.Generator().manual_seed(11)
    hs = [torch.randn(2, cfg.grid * cfg.grid + 1, cfg.vit_hidden, generator=g) for _ in range(3)]   # with the CLS row
    boxes = [torch.tensor([[0.50, 0.50, 0.40, 0.30], [0.25, 0.30, 0.10, 0.20], [0.80, 0.75, 0.35, 0.45], [0.10, 0.90, 0.15, 0.12],
                           [0.55, 0.20, 0.90, 0.30]]),
             torch.tensor([[0.40, 0.60, 0.20, 0.20], [0.95, 0.05, 0.08, 0.08], [0.30, 0.30, 0.55, 0.60]])]   # cxcywh in (0, 1)
    return cfg, sd, hs, boxes


def run_reference():
    ref = load_reference_module()
    cfg, sd, hs, boxes = case()
    torch.manual_seed(0)
    m = ref.MLVLROIQueryModule(embed_dims=cfg.vit_hidden, out_dims=cfg.llm_hidden, num_levels=3).eval()
    m.mlvl_fuse.generate_coordinate.__func__.__defaults__ = ("cpu",)      # the reference defaults this helper's device to 'cuda'
    res = m.load_state_dict({k[len("region_encoder."):]: v.float() for k, v in sd.items() if k.startswith("region_encoder.")}, strict=True)
    with torch.no_grad():
        out = m([h[:, 1:] for h in hs], boxes)                           # groma.py:311-313
    return [o.clone() for o in out]


if __name__ == "__main__":
    out = run_reference()
    torch.save({"outputs": out, "note": "MLVLROIQueryModule outputs of the reference source on case() (see this script)"},
               os.path.join(HERE, "region_encoder_ref.pt"))
    print("wrote region_encoder_ref.pt", [tuple(o.shape) for o in out], float(out[0].abs().max()))
