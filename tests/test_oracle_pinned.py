"""Pin the oracle (CPU only): against every golden vector the reference's own tests hold for this path, against the
reference's C++ CPU kernels compiled from /root/reference into oracle/_ref (oracle/build_ref.py), and against the
committed model-level fixtures under tests/golden/.  No GPU, no compute through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch
import torchvision

from oracle import ops as O
from oracle import build_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ref():
    m = build_ref.load_ref()
    if m is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return m


# ---------------------------------------------------------------- reference golden vectors
def test_nms_golden_vectors():
    # mmcv/tests/test_ops/test_nms.py:13-20
    b = np.array([[6.0, 3.0, 8.0, 7.0], [3.0, 6.0, 9.0, 11.0], [3.0, 7.0, 10.0, 12.0], [1.0, 4.0, 13.0, 7.0]], dtype=np.float32)
    s = np.array([0.6, 0.9, 0.7, 0.2], dtype=np.float32)
    for form in ("cuda", "cpu"):
        assert O.nms_ref(b, s, 0.3, iou_form=form).tolist() == [1, 0, 3]
    # mmcv/mmcv/ops/nms.py:139-150 (docstring example: 3 kept)
    boxes = np.array([[49.1, 32.4, 51.0, 35.9], [49.3, 32.9, 51.0, 35.3], [49.2, 31.8, 51.0, 35.4], [35.1, 11.5, 39.1, 15.7],
                      [35.6, 11.8, 39.3, 14.2], [35.3, 11.5, 39.9, 14.5], [35.2, 11.7, 39.7, 15.7]], dtype=np.float32)
    scores = np.array([0.9, 0.9, 0.5, 0.5, 0.5, 0.4, 0.3], dtype=np.float32)
    assert len(O.nms_ref(boxes, scores, 0.6)) == 3


ROI_INPUTS = [([[[[1., 2.], [3., 4.]]]], [[0., 0., 0., 1., 1.]]),
              ([[[[1., 2.], [3., 4.]], [[4., 3.], [2., 1.]]]], [[0., 0., 0., 1., 1.]]),
              ([[[[1., 2., 5., 6.], [3., 4., 7., 8.], [9., 10., 13., 14.], [11., 12., 15., 16.]]]], [[0., 0., 0., 3., 3.]])]
ROI_OUTPUTS = [[[[[1.0, 1.25], [1.5, 1.75]]]], [[[[1.0, 1.25], [1.5, 1.75]], [[4.0, 3.75], [3.5, 3.25]]]],
               [[[[1.9375, 4.75], [7.5625, 10.375]]]]]


def test_roi_align_golden_vectors():
    # mmcv/tests/test_ops/test_roi_align.py:14-32,67-104 (forward values; pool 2x2, scale 1, sampling 2, avg, aligned)
    for (inp, rois), want in zip(ROI_INPUTS, ROI_OUTPUTS):
        got = O.roi_align_ref(torch.tensor(inp), torch.tensor(rois), 2, 1.0, 2, True)
        assert np.allclose(got.numpy(), np.array(want), atol=1e-6)


def test_msda_equivalence_setup():
    # the configuration of mmcv/tests/test_ops/test_ms_deformable_attn.py:73-135 (seed 3, N=1,M=2,D=2,Lq=2,L=2,P=2):
    # our restatement must agree with a direct per-sample bilinear evaluation (the CUDA kernel's definition)
    torch.manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    aw = torch.rand(N, Lq, M, L, P) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    got = O.msda_ref(value, shapes, loc, aw)
    want = torch.zeros(N, Lq, M * D)
    starts = [0, shapes[0][0] * shapes[0][1]]
    for q in range(Lq):
        for m in range(M):
            for l, (H, W) in enumerate(shapes):
                for p in range(P):
                    x, y = loc[0, q, m, l, p, 0] * W - 0.5, loc[0, q, m, l, p, 1] * H - 0.5
                    x0, y0 = int(np.floor(x)), int(np.floor(y))
                    for dy in (0, 1):
                        for dx in (0, 1):
                            xx, yy = x0 + dx, y0 + dy
                            if 0 <= xx < W and 0 <= yy < H:
                                wgt = (1 - abs(x - xx)) * (1 - abs(y - yy))
                                want[0, q, m * D:(m + 1) * D] += aw[0, q, m, l, p] * wgt * value[0, starts[l] + yy * W + xx, m]
    assert (got - want).abs().max() < 1e-9 + 1e-6 * want.abs().max()


# ---------------------------------------------------------------- the reference's own compiled CPU kernels
def test_nms_matches_compiled_reference(ref):
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 64, 300):
        ctr = torch.rand(n, 2, generator=g); wh = torch.rand(n, 2, generator=g) * 0.3 + 0.02
        if n > 20:
            ctr[n // 2:] = ctr[:n - n // 2] + 0.01 * torch.randn(n - n // 2, 2, generator=g)
            wh[n // 2:] = wh[:n - n // 2]
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).contiguous()
        scores = torch.rand(n, generator=g)
        for thr in (0.3, 0.6):
            want = ref.nms(boxes, scores, thr, 0)
            got = O.nms_ref(boxes.numpy(), scores.numpy(), thr, 0, iou_form="cpu")
            assert got.tolist() == want.tolist()
            # the CUDA-form inequality (what the GPU path implements) selects the same boxes on these inputs
            assert O.nms_ref(boxes.numpy(), scores.numpy(), thr, 0, iou_form="cuda").tolist() == want.tolist()
    assert ref.nms(torch.zeros(0, 4), torch.zeros(0), 0.5, 0).numel() == 0


def test_roi_align_matches_compiled_reference(ref):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 32, 32, generator=g)
    rois = torch.rand(30, 5, generator=g) * 300
    rois[:, 0] = torch.randint(0, 2, (30,), generator=g).float()
    rois[:, 3:] = rois[:, 1:3] + torch.rand(30, 2, generator=g) * 120   # non-negative extents: the CPU kernel asserts on negative
    for scale in (8 / 14, 4 / 14, 2 / 14):
        out = torch.zeros(30, 5, 14, 14)
        ay, ax = torch.zeros(0), torch.zeros(0)
        ref.roi_align_forward(x, rois, out, ay, ax, 14, 14, scale, 2, 1, True)
        got = O.roi_align_ref(x, rois, 14, scale, 2, True)
        assert (got - out).abs().max() < 1e-5


def test_reference_cpu_roi_align_rejects_negative_extents(ref):
    # the literal reference CPU path raises on the cxcywh-as-xyxy boxes Groma feeds it; the oracle follows the
    # CUDA kernel instead (same as torchvision's CPU roi_align), which is finite there.
    x = torch.randn(1, 3, 32, 32)
    rois = torch.tensor([[0.0, 200.0, 200.0, 50.0, 50.0]])
    out = torch.zeros(1, 3, 14, 14)
    with pytest.raises(Exception):
        ref.roi_align_forward(x, rois, out, torch.zeros(0), torch.zeros(0), 14, 14, 8 / 14, 2, 1, True)
    got = O.roi_align_ref(x, rois, 14, 8 / 14, 2, True)
    tv = torchvision.ops.roi_align(x, rois, 14, 8 / 14, 2, True)
    assert torch.isfinite(got).all() and (got - tv).abs().max() < 1e-6


# ---------------------------------------------------------------- committed model-level fixtures
def test_oracle_reproduces_golden_fixture():
    from tests.golden import make_golden
    path = os.path.join(GOLD, "tiny_forward.pt")
    if not os.path.exists(path):
        pytest.skip("fixture missing")
    gold = torch.load(path)
    out = make_golden.run_oracle()
    assert torch.equal(out["input_ids"], gold["input_ids"])
    assert torch.equal(out["new_tokens"], gold["new_tokens"])
    assert [x.tolist() for x in out["nms_inds"]] == [x.tolist() for x in gold["nms_inds"]]
    assert (out["pred_boxes"] - gold["pred_boxes"]).abs().max() < 1e-5
    assert (out["last_logits"] - gold["last_logits"]).abs().max() < 1e-4 * gold["last_logits"].abs().max()
