"""CPU restatement of the native ops on Groma's forward path (TEST INFRASTRUCTURE -- never imported by the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Each function cites the reference code it restates (paths relative to /root/reference).  Parity status:
  * nms_ref, roi_align_ref: PINNED against the reference's own golden vectors (mmcv/tests/test_ops/test_nms.py:13-20,
    mmcv/mmcv/ops/nms.py:139-150, mmcv/tests/test_ops/test_roi_align.py:14-32) and against the reference's C++ CPU
    kernels compiled from /root/reference into oracle/_ref (see oracle/build_ref.py, tests/test_oracle_pinned.py).
  * msda_ref: restates mmcv/mmcv/ops/multi_scale_deform_attn.py:93-150 (same algorithm the reference's own test
    mmcv/tests/test_ops/test_ms_deformable_attn.py:73-135 uses as ground truth for its CUDA kernel).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------- NMS
def nms_ref(boxes, scores, iou_threshold: float, offset: int = 0, score_threshold: float = 0.0, max_num: int = -1,
            iou_form: str = "cuda") -> np.ndarray:
    """mmcv.ops.nms indices (mmcv/mmcv/ops/nms.py:14-33 NMSop.forward + the native greedy sweep).

    iou_form 'cuda': suppress iff interS > thr*(Sa+Sb-interS) (mmcv/ops/csrc/common/cuda/nms_cuda_kernel.cuh:18-29)
    iou_form 'cpu' : suppress iff interS/(Sa+Sb-interS) > thr  (mmcv/ops/csrc/pytorch/cpu/nms.cpp:40-50)
    Sort is descending by score, ties broken by lower index first (stable)."""
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float32).reshape(-1)
    valid_inds = np.arange(len(scores))
    if score_threshold > 0:
        m = scores > np.float32(score_threshold)
        boxes, scores, valid_inds = boxes[m], scores[m], valid_inds[m]
    n = len(scores)
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable")
    b = boxes[order]
    off = np.float32(offset)
    thr = np.float32(iou_threshold)
    areas = (b[:, 2] - b[:, 0] + off) * (b[:, 3] - b[:, 1] + off)
    select = np.ones(n, dtype=bool)
    for i in range(n):
        if not select[i]:
            continue
        j = np.arange(i + 1, n)
        xx1 = np.maximum(b[i, 0], b[j, 0]); yy1 = np.maximum(b[i, 1], b[j, 1])
        xx2 = np.minimum(b[i, 2], b[j, 2]); yy2 = np.minimum(b[i, 3], b[j, 3])
        w = np.maximum(np.float32(0), xx2 - xx1 + off); h = np.maximum(np.float32(0), yy2 - yy1 + off)
        inter = (w * h).astype(np.float32)
        union = (areas[i] + areas[j] - inter).astype(np.float32)
        if iou_form == "cuda":
            sup = inter > thr * union
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                sup = (inter / union) > thr
        select[j] &= ~sup
    inds = order[select]
    if max_num > 0:
        inds = inds[:max_num]
    return valid_inds[inds].astype(np.int64)


# ----------------------------------------------------------------------------------------------------------- RoIAlign
def roi_align_ref(inp: torch.Tensor, rois: torch.Tensor, out_size: int, spatial_scale: float, sampling_ratio: int,
                  aligned: bool = True) -> torch.Tensor:
    """RoIAlign avg-pool with the semantics of the reference's CUDA kernel
    (mmcv/ops/csrc/common/cuda/roi_align_cuda_kernel.cuh:17-108, bilinear of common_cuda_helper.hpp:28-70):
    negative RoI extents allowed, samples with y<-1|y>H|x<-1|x>W contribute 0.
    inp [N,C,H,W] fp32, rois [K,5] fp32 -> [K,C,out,out] fp32.  sampling_ratio must be > 0."""
    assert sampling_ratio > 0
    inp = inp.float()
    rois = rois.float()
    N, C, H, W = inp.shape
    K = rois.shape[0]
    PH = PW = out_size
    g = sampling_ratio
    off = 0.5 if aligned else 0.0
    scale = torch.tensor(spatial_scale, dtype=torch.float32)
    start_w = rois[:, 1] * scale - off
    start_h = rois[:, 2] * scale - off
    end_w = rois[:, 3] * scale - off
    end_h = rois[:, 4] * scale - off
    rw, rh = end_w - start_w, end_h - start_h
    if not aligned:
        rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
    bin_h, bin_w = rh / PH, rw / PW
    ph = torch.arange(PH, dtype=torch.float32)
    iy = torch.arange(g, dtype=torch.float32)
    # y[k, ph, iy] = start_h + ph*bin_h + (iy+.5)*bin_h/g
    y = start_h[:, None, None] + ph[None, :, None] * bin_h[:, None, None] + (iy[None, None, :] + 0.5) * bin_h[:, None, None] / g
    x = start_w[:, None, None] + ph[None, :, None] * bin_w[:, None, None] + (iy[None, None, :] + 0.5) * bin_w[:, None, None] / g
    y = y.reshape(K, PH * g)  # sample rows
    x = x.reshape(K, PW * g)
    Y = y[:, :, None].expand(K, PH * g, PW * g)
    X = x[:, None, :].expand(K, PH * g, PW * g)
    valid = ~((Y < -1.0) | (Y > H) | (X < -1.0) | (X > W))
    Yc, Xc = Y.clamp(min=0.0), X.clamp(min=0.0)
    y_low, x_low = Yc.floor().long(), Xc.floor().long()
    top = y_low >= H - 1
    y_low = torch.where(top, torch.full_like(y_low, H - 1), y_low)
    y_high = torch.where(top, y_low, y_low + 1)
    Yc = torch.where(top, y_low.float(), Yc)
    right = x_low >= W - 1
    x_low = torch.where(right, torch.full_like(x_low, W - 1), x_low)
    x_high = torch.where(right, x_low, x_low + 1)
    Xc = torch.where(right, x_low.float(), Xc)
    ly, lx = Yc - y_low.float(), Xc - x_low.float()
    hy, hx = 1.0 - ly, 1.0 - lx
    bi = rois[:, 0].long()
    feat = inp[bi]  # [K,C,H,W]
    flat = feat.reshape(K, C, H * W)

    def gat(yy, xx):
        idx = (yy * W + xx).reshape(K, 1, -1).expand(K, C, -1)
        return torch.gather(flat, 2, idx).reshape(K, C, PH * g, PW * g)

    val = (hy * hx)[:, None] * gat(y_low, x_low) + (hy * lx)[:, None] * gat(y_low, x_high) \
        + (ly * hx)[:, None] * gat(y_high, x_low) + (ly * lx)[:, None] * gat(y_high, x_high)
    val = val * valid[:, None].float()
    # sum over the g x g samples of every bin in (iy, ix) order, then divide by count
    val = val.reshape(K, C, PH, g, PW, g).permute(0, 1, 2, 4, 3, 5).reshape(K, C, PH, PW, g * g)
    out = torch.zeros(K, C, PH, PW)
    for s in range(g * g):
        out = out + val[..., s]
    return out / float(max(g * g, 1))


# ----------------------------------------------------------------------------------------------------------- MSDA
def msda_ref(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor, attention_weights: torch.Tensor) -> torch.Tensor:
    """mmcv/mmcv/ops/multi_scale_deform_attn.py:93-150 (== HF modeling_deformable_detr.py:171-222).
    value [B,S,M,D]; sampling_locations [B,Q,M,L,P,2] in [0,1]; attention_weights [B,Q,M,L,P] -> [B,Q,M*D]."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, _, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([int(h) * int(w) for h, w in spatial_shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for level, (h, w) in enumerate(spatial_shapes):
        v = value_list[level].flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, int(h), int(w))
        gl = grids[:, :, :, level].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, gl, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(bs * num_heads, 1, num_queries, num_levels * num_points)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(bs, num_heads * embed_dims, num_queries)
    return out.transpose(1, 2).contiguous()


def msda_module_core_ref(value, proj, ref, level_hw, n_heads: int, n_points: int) -> torch.Tensor:
    """Softmax + sampling-location arithmetic of DeformableDetrMultiscaleDeformableAttention.forward
    (HF modeling_deformable_detr.py:586-610) followed by msda_ref.  proj [B,Q, nH*L*P*2 + nH*L*P] fp32."""
    B, Q = proj.shape[0], proj.shape[1]
    L = len(level_hw)
    n_off = n_heads * L * n_points * 2
    off = proj[..., :n_off].reshape(B, Q, n_heads, L, n_points, 2)
    aw = F.softmax(proj[..., n_off:].reshape(B, Q, n_heads, L * n_points), -1).reshape(B, Q, n_heads, L, n_points)
    if ref.shape[-1] == 2:
        norm = torch.tensor([[w, h] for h, w in level_hw], dtype=torch.float32)
        loc = ref[:, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, None, None, :2] + off / n_points * ref[:, :, None, None, None, 2:] * 0.5
    return msda_ref(value.float(), level_hw, loc, aw)


# ----------------------------------------------------------------------------------------------------------- misc
def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)
