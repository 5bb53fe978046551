"""Generate tests/golden/post_process_ref.pt by executing the REFERENCE's own `post_process` (groma/train/train_det.py:97-131).

The reference module cannot be imported (it pulls in mmcv / mmdet / deepspeed at import time), but the function is
self-contained: its source segment is read from the reference file where it lies and executed in a namespace holding the three
names it uses (`torch`, `List`, transformers' `center_to_corners_format`).  Nothing of the reference is copied into the repo;
/root/reference does not exist on the GPU box, hence the committed fixture."""
import ast
import os
from types import SimpleNamespace
from typing import List

import torch
from transformers.image_transforms import center_to_corners_format

REF = "/root/reference/groma/train/train_det.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_post_process():
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "post_process")
    ns = {"torch": torch, "List": List, "center_to_corners_format": center_to_corners_format}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["post_process"]


def cases():
    g = torch.Generator().manual_seed(20)
    out = []
    for (B, Q, C, thr, k, sizes) in [(2, 300, 1, 0.0, 100, [[480, 640], [333, 500]]), (3, 40, 1, 0.45, 100, [[448, 448]] * 3),
                                     (1, 25, 3, 0.3, 20, [[600, 400]])]:
        out.append(dict(coco=torch.randn(B, Q, C, generator=g) * 2 - 1, boxes=torch.rand(B, Q, 4, generator=g) * 0.5 + 0.2,
                        sizes=torch.tensor(sizes, dtype=torch.float32), threshold=thr, top_k=k))
    return out


def run_reference():
    pp = reference_post_process()
    res = []
    for c in cases():
        o = SimpleNamespace(logits={"coco": c["coco"]}, pred_boxes=c["boxes"])
        res.append(pp(o, c["sizes"], threshold=c["threshold"], top_k=c["top_k"]))
        # list-of-pairs form of target_sizes must give the same answer (train_det.py:117-119)
        alt = pp(o, [list(map(float, s)) for s in c["sizes"].tolist()], threshold=c["threshold"], top_k=c["top_k"])
        assert all(torch.equal(a[k], b[k]) for a, b in zip(res[-1], alt) for k in a)
    return res


if __name__ == "__main__":
    torch.save({"cases": cases(), "results": run_reference()}, os.path.join(HERE, "post_process_ref.pt"))
    print("wrote post_process_ref.pt")
