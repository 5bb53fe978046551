"""Op-level parity of every CUDA kernel (through the C ABI) against the oracle / a torch fp32 restatement.
Tolerances: integer outputs bit-exact; bf16 outputs within one bf16 rounding of the fp32 reference of the SAME
bf16 inputs (rel 2^-7 of the row scale); fp32 outputs 1e-4 relative."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ops as O  # noqa: E402


@pytest.fixture(scope="module")
def G():
    from groma_b200 import ops
    return ops


def dev(t):
    return t.cuda()


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (128, 256, 128, 256), (256, 512, 256, 128), (300, 200, 192, 0), (1000, 1024, 1024, 0),
    (4096, 4096, 1024, 256), (77, 96, 256, 0), (130, 16, 256, 16), (513, 40, 64, 0), (64, 64, 64, 64),
    (2048, 1024, 592, 0), (256, 1234, 320, 0), (4096, 256, 4096, 0), (200, 32, 1024, 32),
])
def test_gemm_plain(G, M, N, K, bn):
    a = rnd(M, K, seed=1).bfloat16()
    w = rnd(N, K, seed=2).bfloat16()
    ref = a.float() @ w.float().t()
    out = G.gemm(dev(a), dev(w), block_n=bn)
    assert out.shape == (M, N)
    assert rel_err(out, ref) < 6e-3
    out32 = G.gemm(dev(a), dev(w), out_f32=True, block_n=bn)
    assert rel_err(out32, ref) < 2e-5


def test_gemm_epilogues(G):
    M, N, K = 384, 640, 256
    a = rnd(M, K, seed=3).bfloat16(); w = rnd(N, K, seed=4, scale=0.1).bfloat16()
    bias = rnd(N, seed=5); gamma = rnd(N, seed=6); res = rnd(M, N, seed=7).bfloat16()
    base = a.float() @ w.float().t() + bias
    for act, fn in [(G.ACT_NONE, lambda x: x), (G.ACT_GELU, F.gelu), (G.ACT_RELU, F.relu)]:
        ref = fn(base) * gamma + res.float()
        out = G.gemm(dev(a), dev(w), bias=dev(bias), act=act, gamma=dev(gamma), residual=dev(res))
        assert rel_err(out, ref) < 6e-3, act
    # in-place residual (out aliases residual)
    r = dev(res.clone())
    G.gemm(dev(a), dev(w), bias=dev(bias), gamma=dev(gamma), residual=r, out=r)
    assert rel_err(r, base * gamma + res.float()) < 6e-3
    # swiglu: interleaved (gate, up) columns
    ref = F.silu(base[:, 0::2]) * base[:, 1::2]
    out = G.gemm(dev(a), dev(w), bias=dev(bias), act=G.ACT_SWIGLU)
    assert out.shape == (M, N // 2)
    assert rel_err(out, ref) < 6e-3


def test_gemm_strided_views(G):
    # A and out as column slices of wider buffers (row stride != width)
    M, N, K = 256, 128, 128
    big = rnd(M, 3 * K, seed=8).bfloat16()
    w = rnd(N, K, seed=9).bfloat16()
    a_view = dev(big)[:, K:2 * K]
    outbuf = torch.zeros((M, 2 * N), dtype=torch.bfloat16, device="cuda")
    G.gemm(a_view, dev(w), out=outbuf[:, N:])
    ref = big[:, K:2 * K].float() @ w.float().t()
    assert rel_err(outbuf[:, N:], ref) < 6e-3
    assert outbuf[:, :N].abs().max().item() == 0


@pytest.mark.parametrize("M,N,K,split", [(256, 512, 2048, 2), (100, 1024, 4096, 4), (16, 256, 1024, 3), (300, 130, 640, 5)])
def test_gemm_splitk(G, M, N, K, split):
    a = rnd(M, K, seed=10).bfloat16(); w = rnd(N, K, seed=11).bfloat16(); bias = rnd(N, seed=12)
    ref = a.float() @ w.float().t() + bias
    out = G.gemm_splitk(dev(a), dev(w), split, bias=dev(bias), out_f32=True)
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("M,N,K,split", [(16, 4096, 4096, 1), (16, 1024, 4096, 4), (8, 384, 512, 2), (48, 256, 256, 1), (1, 256, 512, 1)])
def test_gemm_swap_ab(G, M, N, K, split):
    x = rnd(M, K, seed=13).bfloat16(); w = rnd(N, K, seed=14).bfloat16()
    ws = torch.empty((split, N, M), dtype=torch.float32, device="cuda")
    G.gemm_swap_ab(dev(x), dev(w), ws, split_k=split)
    ref = x.float() @ w.float().t()
    got = ws.sum(0).t()
    assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("n_img,H,W,C,Cout,L", [(2, 8, 8, 64, 64, 1), (1, 14, 14, 128, 256, 3), (3, 32, 32, 64, 128, 1), (2, 16, 20, 192, 96, 1)])
def test_conv3x3_flat(G, n_img, H, W, C, Cout, L):
    xs = [rnd(n_img, C, H, W, seed=20 + l).bfloat16() for l in range(L)]
    ws = [rnd(Cout, C, 3, 3, seed=30 + l, scale=0.05).bfloat16() for l in range(L)]
    bias = rnd(Cout, seed=40)
    ref = sum(F.conv2d(x.float(), w.float(), padding=1) for x, w in zip(xs, ws)) + bias[None, :, None, None]
    ref = F.relu(ref).permute(0, 2, 3, 1).reshape(-1, Cout)
    hp, wp = H + 2, W + 2
    xpad = torch.zeros((L, n_img, hp, wp, C), dtype=torch.bfloat16)
    for l in range(L):
        xpad[l, :, 1:-1, 1:-1] = xs[l].permute(0, 2, 3, 1)
    wt = torch.cat([w.permute(0, 2, 3, 1).reshape(Cout, 9 * C) for w in ws], dim=1).contiguous()
    out = G.conv3x3_flat(dev(xpad.reshape(-1, C)), dev(wt), n_img, hp, wp, bias=dev(bias), act=G.ACT_RELU)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 6e-3
    # non-compact output keeps borders untouched
    full = torch.full((n_img * hp * wp, Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    G.conv3x3_flat(dev(xpad.reshape(-1, C)), dev(wt), n_img, hp, wp, bias=dev(bias), act=G.ACT_RELU, out=full, compact=False)
    full = full.reshape(n_img, hp, wp, Cout)
    assert rel_err(full[:, 1:-1, 1:-1].reshape(-1, Cout), ref) < 6e-3
    assert (full[:, 0] == 7).all() and (full[:, :, 0] == 7).all() and (full[:, -1] == 7).all() and (full[:, :, -1] == 7).all()


# ------------------------------------------------------------------------------------------------- attention
def attn_ref(q, k, v, causal, scale, q_pos0=0, kv_len=None):
    # q [B,Sq,H,D], k/v [B,H,Sk,D]
    B, Sq, H, D = q.shape
    Sk = k.shape[2]
    s = torch.einsum("bqhd,bhkd->bhqk", q.float(), k.float()) * scale
    mask = torch.zeros(B, 1, Sq, Sk, dtype=torch.bool)
    if causal:
        qi = torch.arange(Sq)[:, None] + q_pos0
        mask |= (torch.arange(Sk)[None, :] > qi)[None, None]
    if kv_len is not None:
        mask |= (torch.arange(Sk)[None, :] >= kv_len[:, None])[:, None, None, :]
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bhkd->bqhd", p, v.float()).reshape(B, Sq, H * D)


@pytest.mark.parametrize("B,H,Sq,Sk,D,causal", [
    (2, 4, 128, 128, 128, True), (1, 2, 300, 300, 32, False), (2, 3, 1025, 1025, 64, False), (2, 2, 200, 200, 128, True),
    (3, 4, 1, 333, 128, True), (1, 8, 70, 70, 64, True),
])
def test_attention(G, B, H, Sq, Sk, D, causal):
    q = rnd(B, Sq, H, D, seed=50).bfloat16(); k = rnd(B, H, Sk, D, seed=51).bfloat16(); v = rnd(B, H, Sk, D, seed=52).bfloat16()
    scale = 1.0 / math.sqrt(D)
    q_pos0 = Sk - Sq if causal else 0
    ref = attn_ref(q, k, v, causal, scale, q_pos0)
    out = G.attention(dev(q), dev(k), dev(v), causal=causal, scale=scale, q_pos0=q_pos0)
    assert rel_err(out, ref) < 1e-2


def test_attention_kvlen_and_strides(G):
    B, H, S, D = 3, 4, 190, 128
    qkv = rnd(B, S, 3 * H * D, seed=53).bfloat16()
    cache_cap = 256
    kc = torch.zeros(B, H, cache_cap, D, dtype=torch.bfloat16); vc = torch.zeros_like(kc)
    k = qkv[:, :, H * D:2 * H * D].reshape(B, S, H, D).permute(0, 2, 1, 3); v = qkv[:, :, 2 * H * D:].reshape(B, S, H, D).permute(0, 2, 1, 3)
    kc[:, :, :S] = k; vc[:, :, :S] = v
    kv_len = torch.tensor([190, 100, 7], dtype=torch.int32)
    qd = dev(qkv)
    q_view = qd[:, :, :H * D].unflatten(2, (H, D))  # row stride 3*H*D
    out = G.attention(q_view, dev(kc), dev(vc), causal=True, scale=0.1, kv_len=dev(kv_len), sk=S)
    ref = attn_ref(qkv[:, :, :H * D].reshape(B, S, H, D), kc[:, :, :S], vc[:, :, :S], True, 0.1, 0, kv_len.long())
    assert rel_err(out, ref) < 1e-2


# ------------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("rows,dim", [(33, 4096), (7, 256), (100, 1024)])
def test_rmsnorm(G, rows, dim):
    x = rnd(rows, dim, seed=60).bfloat16(); r = rnd(rows, dim, seed=61).bfloat16(); w = rnd(dim, seed=62)
    h = (x.float() + r.float()).bfloat16()
    hf = h.float()
    ref = w * (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float()
    h_out = torch.empty_like(dev(x))
    y = G.rmsnorm(dev(x), dev(w), 1e-5, residual=dev(r), h_out=h_out)
    assert torch.equal(h_out.cpu(), h)
    assert rel_err(y, ref) < 5e-3
    y2 = G.rmsnorm(dev(x), dev(w), 1e-5)
    xf = x.float()
    assert rel_err(y2, w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float()) < 5e-3


@pytest.mark.parametrize("rows,dim,eps", [(50, 1024, 1e-6), (300, 256, 1e-5), (10, 512, 1e-5)])
def test_layernorm(G, rows, dim, eps):
    x = rnd(rows, dim, seed=63).bfloat16(); r = rnd(rows, dim, seed=64).bfloat16(); w = rnd(dim, seed=65); b = rnd(dim, seed=66)
    ref = F.layer_norm(x.float(), (dim,), w, b, eps)
    assert rel_err(G.layernorm(dev(x), dev(w), dev(b), eps), ref) < 5e-3
    h = (x.float() + r.float()).bfloat16().float()
    assert rel_err(G.layernorm(dev(x), dev(w), dev(b), eps, residual=dev(r)), F.layer_norm(h, (dim,), w, b, eps)) < 5e-3


def test_groupnorm_relu(G):
    n_img, H, W, C, Gn = 2, 12, 12, 1024, 64
    x = rnd(n_img, C, H, W, seed=67).bfloat16(); gamma = rnd(C, seed=68); beta = rnd(C, seed=69)
    ref = F.relu(F.group_norm(x.float(), Gn, gamma, beta, 1e-5)).permute(0, 2, 3, 1).reshape(-1, C)
    out = G.groupnorm_relu(dev(x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()), dev(gamma), dev(beta), Gn, 1e-5, n_img)
    assert rel_err(out, ref) < 5e-3


def test_fuse_shuffle_gn_equals_apply_then_shuffle(G):
    """groupnorm_stats + fuse_shuffle_gn over raw maps == groupnorm_relu on every map followed by fuse_shuffle, bit for bit
    (the fusion rounds of the region encoder run the former; mmcv ConvModule norm + act, roi_align.py:118-126,150-178)."""
    B, C, Gn = 2, 256, 16
    gamma, beta = dev(rnd(C, seed=171)), dev(rnd(C, seed=172))
    raw = [dev((rnd(B, s, s, C, seed=173 + i) * 3).bfloat16()) for i, s in enumerate((24, 12, 6))]
    stats = [G.groupnorm_stats(m.reshape(-1, C), Gn, 1e-5, B) for m in raw]
    act = [G.groupnorm_relu(m.reshape(-1, C).clone(), gamma, beta, Gn, 1e-5, B).reshape(m.shape) for m in raw]
    for m, st, a in zip(raw, stats, act):     # the two halves of groupnorm_relu are the whole
        assert torch.equal(G.groupnorm_apply_relu(m.reshape(-1, C), st, gamma, beta, Gn, B).reshape(m.shape), a)
    for lvl in range(3):
        top, dn = min(lvl + 1, 2), max(lvl - 1, 0)
        want = G.fuse_shuffle(act[lvl], act[top], act[dn])
        got = G.fuse_shuffle_gn(raw[lvl], raw[top], raw[dn], stats[lvl], stats[top], stats[dn], gamma, beta, Gn)
        assert torch.equal(got, want), f"level {lvl}: {(got.float() - want.float()).abs().max().item()}"


# ------------------------------------------------------------------------------------------------- detection ops
def test_msda_matches_oracle(G):
    torch.manual_seed(3)
    for (B, Q, hw, ref_dim) in [(2, 1024, [(32, 32)], 2), (2, 300, [(32, 32)], 4), (1, 50, [(6, 4), (3, 2)], 2)]:
        nH, P, L = 8, 4, len(hw)
        S = sum(h * w for h, w in hw)
        value = rnd(B, S, nH, 32, seed=70).bfloat16()
        proj = rnd(B, Q, nH * L * P * 3, seed=71, scale=2.0)
        ref = torch.rand(B, Q, ref_dim, generator=torch.Generator().manual_seed(72))
        if ref_dim == 4:
            ref[..., 2:] *= 0.5
        want = O.msda_module_core_ref(value.float(), proj, ref, hw, nH, P)
        got = G.msda(dev(value), dev(proj.reshape(B * Q, -1).contiguous()), dev(ref), hw, nH, P)
        assert rel_err(got, want) < 6e-3


ROI_INPUTS = [([[[[1., 2.], [3., 4.]]]], [[0., 0., 0., 1., 1.]]),
              ([[[[1., 2.], [3., 4.]], [[4., 3.], [2., 1.]]]], [[0., 0., 0., 1., 1.]]),
              ([[[[1., 2., 5., 6.], [3., 4., 7., 8.], [9., 10., 13., 14.], [11., 12., 15., 16.]]]], [[0., 0., 0., 3., 3.]])]
ROI_OUTPUTS = [[[[[1.0, 1.25], [1.5, 1.75]]]], [[[[1.0, 1.25], [1.5, 1.75]], [[4.0, 3.75], [3.5, 3.25]]]],
               [[[[1.9375, 4.75], [7.5625, 10.375]]]]]


def test_roi_align_mmcv_golden(G):
    # mmcv/tests/test_ops/test_roi_align.py:14-32 (pool 2x2, scale 1.0, sampling 2, aligned avg)
    for (inp, rois), want in zip(ROI_INPUTS, ROI_OUTPUTS):
        x = torch.tensor(inp)  # [1,C,H,W]
        C = x.shape[1]
        xp = torch.zeros(1, x.shape[2], x.shape[3], 8)
        xp[..., :C] = x.permute(0, 2, 3, 1)
        out = G.roi_align(dev(xp.bfloat16()), dev(torch.tensor(rois)), 2, 1.0, 2, True)
        got = out.float().cpu()[..., :C].permute(0, 3, 1, 2)
        assert np.allclose(got.numpy(), np.array(want), atol=1e-3)


def test_roi_align_oracle_edge_cases(G):
    # negative extents, far out-of-bounds RoIs, zero border
    N, C, H, W = 2, 64, 32, 32
    x = rnd(N, C, H, W, seed=80).bfloat16()
    g = torch.Generator().manual_seed(81)
    rois = torch.rand(40, 5, generator=g) * 448
    rois[:, 0] = torch.randint(0, N, (40,), generator=g).float()
    rois[5:15, 3:] = rois[5:15, 1:3] * 0.3  # x2<x1, y2<y1
    for scale in (8 / 14, 2 / 14):
        want = O.roi_align_ref(x.float(), rois, 14, scale, 2, True).permute(0, 2, 3, 1)
        got = G.roi_align(dev(x.permute(0, 2, 3, 1).contiguous()), dev(rois), 14, scale, 2, True, pad=True).float().cpu()
        assert got.shape == (40, 16, 16, C)
        assert (got[:, 0] == 0).all() and (got[:, -1] == 0).all() and (got[:, :, 0] == 0).all() and (got[:, :, -1] == 0).all()
        assert (got[:, 1:-1, 1:-1] - want).abs().max() <= 2e-2 * want.abs().max().clamp(min=1e-3)
        assert torch.isfinite(got).all()


def test_roi_align_bench_shape_and_channel_loop(G):
    """The shared-tap kernel (fixed sampling grid) at the channel width of the bench (1024: one 16-byte vector per thread) and at
    a width that makes the channel loop wrap (C/8 > 256 threads), against the oracle; RoIs include the reference's
    cxcywh-fed-as-xyxy quirk (negative extents) and boxes that leave the map."""
    for C, H, W, pool, sr in ((1024, 16, 16, 14, 2), (2304, 8, 8, 7, 2), (64, 32, 32, 14, 3)):
        x = rnd(2, C, H, W, seed=90 + pool).bfloat16()
        g = torch.Generator().manual_seed(91)
        rois = torch.rand(24, 5, generator=g) * 448
        rois[:, 0] = torch.randint(0, 2, (24,), generator=g).float()
        rois[3:9, 3:] = rois[3:9, 1:3] * 0.4
        rois[9:12, 1:] += 400.0
        scale = H / 448.0
        want = O.roi_align_ref(x.float(), rois, pool, scale, sr, True).permute(0, 2, 3, 1)
        got = G.roi_align(dev(x.permute(0, 2, 3, 1).contiguous()), dev(rois), pool, scale, sr, True, pad=True).float().cpu()
        assert got.shape == (24, pool + 2, pool + 2, C)
        assert (got[:, 0] == 0).all() and (got[:, :, -1] == 0).all()
        assert (got[:, 1:-1, 1:-1] - want).abs().max() <= 2e-2 * want.abs().max().clamp(min=1e-3)


def test_nms_golden_and_oracle(G):
    # mmcv/tests/test_ops/test_nms.py:13-20 -> [1, 0, 3]
    b = torch.tensor([[6.0, 3.0, 8.0, 7.0], [3.0, 6.0, 9.0, 11.0], [3.0, 7.0, 10.0, 12.0], [1.0, 4.0, 13.0, 7.0]])
    s = torch.tensor([0.6, 0.9, 0.7, 0.2])
    keep, num, amax = G.nms_batched(dev(b[None]), dev(s[None]), 0.3, 0.0, -1)
    assert num.item() == 3 and keep[0, :3].tolist() == [1, 0, 3] and amax.item() == 1
    # docstring example mmcv/ops/nms.py:139-150 -> 3 kept
    boxes = torch.tensor([[49.1, 32.4, 51.0, 35.9], [49.3, 32.9, 51.0, 35.3], [49.2, 31.8, 51.0, 35.4], [35.1, 11.5, 39.1, 15.7],
                          [35.6, 11.8, 39.3, 14.2], [35.3, 11.5, 39.9, 14.5], [35.2, 11.7, 39.7, 15.7]])
    scores = torch.tensor([0.9, 0.9, 0.5, 0.5, 0.5, 0.4, 0.3])
    keep, num, _ = G.nms_batched(dev(boxes[None]), dev(scores[None]), 0.6, 0.0, -1)
    assert num.item() == 3 and keep[0, :3].tolist() == O.nms_ref(boxes.numpy(), scores.numpy(), 0.6).tolist()
    # random batches: clustered boxes, ties, score threshold, max_num, ragged counts
    g = torch.Generator().manual_seed(90)
    B, N = 5, 317
    ctr = torch.rand(B, N, 2, generator=g); wh = torch.rand(B, N, 2, generator=g) * 0.3 + 0.02
    ctr[:, 100:200] = ctr[:, :100] + 0.01 * torch.randn(B, 100, 2, generator=g)
    wh[:, 100:200] = wh[:, :100]
    bx = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).contiguous()
    sc = torch.rand(B, N, generator=g)
    sc[:, 300:305] = 1.0
    sc[:, 305:310] = 0.2
    counts = torch.tensor([317, 300, 1, 0, 250], dtype=torch.int32)
    for thr, sthr, mx in [(0.6, 0.0, 100), (0.6, 0.15, 100), (0.3, 0.5, -1), (0.6, 0.999, 100)]:
        keep, num, amax = G.nms_batched(dev(bx), dev(sc), thr, sthr, mx, counts=dev(counts))
        for i in range(B):
            n = counts[i].item()
            want = O.nms_ref(bx[i, :n].numpy(), sc[i, :n].numpy(), thr, 0, sthr, mx)
            assert num[i].item() == len(want)
            assert keep[i, :len(want)].tolist() == want.tolist()
            if n > 0:
                assert amax[i].item() == int(torch.argmax(sc[i, :n]))


def test_topk(G):
    sc = rnd(6, 1024, seed=95)
    sc[0, 10] = sc[0, 500]  # a tie
    idx = G.topk_desc(dev(sc), 300).cpu()
    for i in range(6):
        order = np.argsort(-sc[i].numpy(), kind="stable")[:300]
        assert idx[i].tolist() == order.tolist()


def test_ddetr_select_finalize(G):
    B, S, k, npf = 2, 1024, 300, 128
    delta = rnd(B, S, 4, seed=96); prop = rnd(S, 4, seed=97)
    topk = torch.stack([torch.randperm(S, generator=torch.Generator().manual_seed(98 + i))[:k] for i in range(B)])
    ref, pos = G.ddetr_select(dev(delta), dev(prop), dev(topk), npf)
    logits = torch.gather(delta + prop[None], 1, topk[..., None].expand(-1, -1, 4))
    want_ref = logits.sigmoid()
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
    p = (want_ref * (2 * math.pi))[:, :, :, None] / dim_t
    want_pos = torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).flatten(2)
    assert (ref.cpu() - want_ref).abs().max() < 1e-6
    assert (pos.float().cpu() - want_pos).abs().max() < 1e-2
    # finalize
    Q = 300
    d4 = rnd(B, Q, 4, seed=99); d5 = rnd(B, Q, 4, seed=100); r0 = torch.rand(B, Q, 4); coco = rnd(B, Q, seed=101); sa = rnd(B, Q, seed=102)
    Nmax = 320
    pc = torch.zeros(B, Nmax, 4, device="cuda"); px = torch.zeros(B, Nmax, 4, device="cuda"); sco = torch.zeros(B, Nmax, device="cuda")
    G.ddetr_finalize(dev(d4), dev(d5), dev(r0), dev(coco), dev(sa), pc, px, sco)

    def inv_sig(x, eps=1e-5):
        x = x.clamp(0, 1)
        return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))
    r1 = (d4 + inv_sig(r0)).sigmoid()
    want = (d5 + inv_sig(r1)).sigmoid()
    assert (pc[:, :Q].cpu() - want).abs().max() < 1e-5
    wx = torch.cat([want[..., :2] - 0.5 * want[..., 2:], want[..., :2] + 0.5 * want[..., 2:]], -1)
    assert (px[:, :Q].cpu() - wx).abs().max() < 1e-5
    ws = coco.sigmoid() ** 0.4 * sa.sigmoid() ** 0.6
    assert (sco[:, :Q].cpu() - ws).abs().max() < 1e-5


# ------------------------------------------------------------------------------------------------- resampling / plumbing
def test_upsample_coords_and_shuffle(G):
    B, g, C = 2, 32, 64
    tok = rnd(B, 1 + g * g, C, seed=110).bfloat16()
    fmap = tok[:, 1:].reshape(B, g, g, C).permute(0, 3, 1, 2).float()
    for Ho in (128, 64, 32):
        xs = torch.linspace(-1, 1, Ho); ys = torch.linspace(-1, 1, Ho)
        out = G.upsample_coords(dev(tok), 1, g, Ho, Ho, C + 64, dev(xs), dev(ys)).float().cpu()
        want = F.interpolate(fmap, size=(Ho, Ho), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        assert (out[..., :C] - want).abs().max() < 2e-2
        assert (out[..., C] - xs.bfloat16().float()[None, None, :]).abs().max() == 0
        assert (out[..., C + 1] - ys.bfloat16().float()[None, :, None]).abs().max() == 0
        assert out[..., C + 2:].abs().max() == 0
    # fuse shuffle
    C = 128
    maps = [rnd(B, s, s, C, seed=111 + i).bfloat16() for i, s in enumerate((16, 8, 4))]
    for lvl in range(3):
        top, dn = min(lvl + 1, 2), max(lvl - 1, 0)
        out = G.fuse_shuffle(dev(maps[lvl]), dev(maps[top]), dev(maps[dn])).float().cpu()
        s = maps[lvl].shape[1]
        nchw = [m.permute(0, 3, 1, 2).float() for m in maps]
        ft = F.interpolate(nchw[top][:, C // 2:][:, C // 4:], size=(s, s), mode="bilinear", align_corners=True)
        fd = F.interpolate(nchw[dn][:, C // 2:][:, :C // 4], size=(s, s), mode="bilinear", align_corners=True)
        want = torch.cat([nchw[lvl][:, :C // 2], ft, fd], 1).permute(0, 2, 3, 1)
        assert (out[:, 1:-1, 1:-1] - want).abs().max() < 2e-2
        assert out[:, 0].abs().max() == 0 and out[:, :, -1].abs().max() == 0


def test_vit_plumbing(G):
    B, S = 2, 56
    img = rnd(B, 3, S, S, seed=120)
    pt = G.vit_patchify(dev(img), 592).float().cpu()
    want = F.unfold(img, kernel_size=14, stride=14).transpose(1, 2).reshape(B * 16, 588)
    assert (pt[:, :588] - want.bfloat16().float()).abs().max() == 0 and pt[:, 588:].abs().max() == 0
    NP, C = 16, 64
    patch = rnd(B * NP, C, seed=121).bfloat16(); cls = rnd(C, seed=122); pos = rnd(NP + 1, C, seed=123)
    emb = G.vit_embed(dev(patch), dev(cls), dev(pos), B, NP).float().cpu()
    want = torch.cat([cls[None, None].expand(B, 1, C), patch.float().reshape(B, NP, C)], 1) + pos[None]
    assert (emb - want.bfloat16().float()).abs().max() == 0
    ts = [rnd(B, NP + 1, C, seed=124 + i).bfloat16() for i in range(4)]
    m = G.mean_tokens([dev(t) for t in ts], 1).float().cpu()
    want = torch.stack([t.float() for t in ts]).mean(0)[:, 1:]
    assert (m - want).abs().max() < 2e-2
    g = 4
    s2d = G.space_to_depth(dev(ts[0]), g).cpu()
    f = ts[0][:, 1:].reshape(B, g, g, C)
    want = torch.cat([f[:, 0::2, 0::2], f[:, 1::2, 0::2], f[:, 0::2, 1::2], f[:, 1::2, 1::2]], -1).reshape(B, 4, 4 * C)
    assert torch.equal(s2d, want)


@pytest.mark.parametrize("B,T,H,K,bn,pos0", [(2, 300, 4, 256, 256, 0), (3, 171, 2, 512, 256, 7), (2, 520, 6, 320, 512, 0)])
def test_gemm_qkv_rope_equals_gemm_then_rope_kv(G, B, T, H, K, bn, pos0):
    """groma_gemm_qkv_rope (RoPE + KV append in the GEMM epilogue) == groma_gemm_bf16 followed by groma_rope_kv, bit for bit:
    q rows, the appended K/V rows, and nothing else in the cache touched (ragged last row tile, pos0 > 0, both tile shapes)."""
    D, cap = 128, T + pos0 + 5
    x = dev(rnd(B * T, K, seed=140).bfloat16())
    w = dev((rnd(3 * H * D, K, seed=141) / math.sqrt(K)).bfloat16())
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(cap).float(), inv)
    cos_t, sin_t = dev(fr.cos().contiguous()), dev(fr.sin().contiguous())
    q0 = torch.empty(B * T, H * D, dtype=torch.bfloat16, device="cuda")
    k0 = torch.full((B, H, cap, D), 3.0, dtype=torch.bfloat16, device="cuda"); v0 = k0.clone()
    G.rope_kv(G.gemm(x, w), q0, k0, v0, cos_t, sin_t, B, T, H, D, pos0)
    q1 = torch.empty_like(q0)
    k1 = torch.full_like(k0, 3.0); v1 = k1.clone()
    G.gemm_qkv_rope(x, w, q1, k1, v1, cos_t, sin_t, B, T, H, D, pos0, block_n=bn)
    torch.cuda.synchronize()
    for name, a, b in (("q", q0, q1), ("k", k0, k1), ("v", v0, v1)):
        bad = (a != b)
        assert not bad.any(), f"{name}: {int(bad.sum())} of {a.numel()} differ, max |d| {(a.float() - b.float()).abs().max().item():.3e}, first {bad.nonzero()[0].tolist()}"
    assert (k1[:, :, :pos0] == 3.0).all() and (k1[:, :, pos0 + T:] == 3.0).all()


def test_gather_scatter_rope_argmax(G):
    V0, V1, D = 100, 14, 64
    t0 = rnd(V0, D, seed=130).bfloat16(); t1 = rnd(V1, D, seed=131).bfloat16()
    ids = torch.randint(0, V0 + V1, (57,), generator=torch.Generator().manual_seed(132))
    out = G.gather_rows(dev(ids), dev(t0), dev(t1), V0).cpu()
    want = torch.where((ids < V0)[:, None], t0[ids.clamp(max=V0 - 1)], t1[(ids - V0).clamp(min=0)])
    assert torch.equal(out, want)
    dst = torch.zeros(80, D, dtype=torch.bfloat16, device="cuda")
    pos = torch.randperm(80, generator=torch.Generator().manual_seed(133))[:20]
    G.scatter_rows(dev(pos), dev(t0[:20].contiguous()), dst)
    assert torch.equal(dst.cpu()[pos], t0[:20])
    # rope
    B, T, H, Dh, pos0, cap = 2, 9, 4, 128, 5, 32
    qkv = rnd(B * T, 3 * H * Dh, seed=134).bfloat16()
    inv = 1.0 / (10000 ** (torch.arange(0, Dh, 2).float() / Dh))
    fr = torch.outer(torch.arange(64).float(), inv)
    cos_t, sin_t = fr.cos().contiguous(), fr.sin().contiguous()
    q_out = torch.empty(B * T, H * Dh, dtype=torch.bfloat16, device="cuda")
    kc = torch.zeros(B, H, cap, Dh, dtype=torch.bfloat16, device="cuda"); vc = torch.zeros_like(kc)
    G.rope_kv(dev(qkv), q_out, kc, vc, dev(cos_t), dev(sin_t), B, T, H, Dh, pos0)
    x = qkv.float().reshape(B, T, 3, H, Dh)
    cos = torch.cat([cos_t, cos_t], -1)[pos0:pos0 + T][None, :, None]; sin = torch.cat([sin_t, sin_t], -1)[pos0:pos0 + T][None, :, None]

    def rot(t):
        return torch.cat([-t[..., Dh // 2:], t[..., :Dh // 2]], -1)
    qr = x[:, :, 0] * cos + rot(x[:, :, 0]) * sin
    kr = x[:, :, 1] * cos + rot(x[:, :, 1]) * sin
    assert (q_out.float().cpu().reshape(B, T, H, Dh) - qr).abs().max() < 3e-2
    assert (kc.float().cpu()[:, :, pos0:pos0 + T].permute(0, 2, 1, 3) - kr).abs().max() < 3e-2
    assert torch.equal(vc.cpu()[:, :, pos0:pos0 + T].permute(0, 2, 1, 3), qkv.reshape(B, T, 3, H, Dh)[:, :, 2])
    assert kc[:, :, :pos0].abs().max().item() == 0
    lg = rnd(5, 32114, seed=135)
    lg[2, 100] = lg[2].max() + 1; lg[2, 20000] = lg[2, 100]
    am = G.argmax(dev(lg)).cpu()
    assert am.tolist() == lg.argmax(-1).tolist() and am[2].item() == 100
