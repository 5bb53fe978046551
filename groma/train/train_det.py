"""`groma.train.train_det.post_process` -- the detector entry's post-processing (reference groma/train/train_det.py:97-131):
per image the `top_k` highest sigmoid(coco-logit) (query, class) pairs, their boxes as absolute xyxy in the target image
size, filtered by `threshold`.  Training itself (the rest of the reference file) is out of scope of the forward path.

The top-k runs in the library's device kernel (`groma_topk_desc`, descending, ties by lower flat index); everything else
is index plumbing on the same device."""
from __future__ import annotations

from typing import List, Sequence, Union

import torch

from groma_b200 import ops as G


def post_process(outputs, target_sizes: Union[torch.Tensor, Sequence, None], threshold: float = 0.0, top_k: int = 100) -> List[dict]:
    out_logits, out_bbox = outputs.logits["coco"], outputs.pred_boxes
    if target_sizes is not None and len(out_logits) != len(target_sizes):
        raise ValueError("Make sure that you pass in as many target sizes as the batch dimension of the logits")
    B, Q, C = out_logits.shape
    prob = out_logits.float().sigmoid().reshape(B, Q * C).contiguous()
    k = min(top_k, Q * C)
    idx = G.topk_desc(prob, k)                                   # [B, k] int64 flat (query, class) indices
    scores = torch.gather(prob, 1, idx)
    topk_boxes = torch.div(idx, C, rounding_mode="floor")
    labels = idx % C
    cx, cy, w, h = out_bbox.float().unbind(-1)                   # center_to_corners_format
    boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
    boxes = torch.gather(boxes, 1, topk_boxes.unsqueeze(-1).repeat(1, 1, 4))
    if isinstance(target_sizes, (list, tuple)):
        img_h = torch.tensor([float(s[0]) for s in target_sizes])
        img_w = torch.tensor([float(s[1]) for s in target_sizes])
    else:
        img_h, img_w = target_sizes.unbind(1)
    scale = torch.stack([img_w, img_h, img_w, img_h], dim=1).to(boxes.device, boxes.dtype)
    boxes = boxes * scale[:, None, :]
    results = []
    for s, l, b in zip(scores, labels, boxes):
        keep = s > threshold
        results.append({"scores": s[keep], "labels": l[keep], "boxes": b[keep]})
    return results
