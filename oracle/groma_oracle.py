"""CPU restatement of Groma's forward hot path (TEST INFRASTRUCTURE -- the product never imports this).

Pure PyTorch on CPU tensors; no mmcv / mmdet / HF-model imports.  Each function cites the reference code it follows
(paths relative to /root/reference; $HF = transformers 4.32.0 pinned by pyproject.toml:19, read through the installed
5.5 sources where the arithmetic is unchanged -- see SURVEY.md section 8c).

PARITY STATUS: the native ops used here (oracle/ops.py: nms, roi_align, msda) are pinned against the reference's golden
vectors and its own compiled C++ CPU kernels (tests/test_oracle_pinned.py).  The third-party sub-models restated here
(Dinov2 stack, LLaMA prefill + cached decode, Deformable-DETR sine embedding / encoder layers / decoder layers) are pinned
against the transformers modules installed in this image on identical weights (tests/test_oracle_hf_pins_cpu.py).
The reference's own glue (GromaModel.forward prefill + decode branch, DeformableDetrTransformer / DeformableDetrDecoderX, region
selection, MLVLROIQueryModule) is pinned against outputs of the reference's own source files, executed in the authoring
container under leaf adapters and committed as fixtures (tests/golden/make_*_golden.py, tests/test_*_ref_cpu.py); the reference
holds no test or recorded output of its own for this level.

Precision modes (`prec`):
  'fp32' : plain fp32 everywhere = the reference's fp32 inference arithmetic (eval_rec.py:69).
  'bf16' : matrices rounded to bf16 once, activations rounded to bf16 at exactly the points where the B200 pipeline
           stores them (DESIGN.md "rounding points"), fp32 accumulation in between.  This is the oracle the 1e-3
           logits tolerance of BASELINE.json is checked against; 'fp32' is reported alongside for information.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .config import PathConfig
from . import ops as O


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class Oracle:
    def __init__(self, cfg: PathConfig, sd: Dict[str, torch.Tensor], prec: str = "bf16", prerounded: bool = False):
        """prerounded: `sd` holds fp32 tensors whose matrices already carry bf16-representable values (e.g. the bf16 arena of
        the GPU model copied back and widened): used as they are, so a 'bf16' and an 'fp32' oracle of Groma-7B size can share
        one 30 GB dict."""
        assert prec in ("fp32", "bf16")
        self.cfg = cfg
        self.prec = prec
        self.r = _bf if prec == "bf16" else (lambda x: x)
        # matrices (dim >= 2) are stored in bf16 on the device; vectors (bias / norm / LayerScale) stay fp32
        # (v.float() is a no-op view for fp32 inputs, so the 7B-size fp32 oracle does not duplicate its 30 GB of weights)
        rm = (lambda x: x) if prerounded else self.r
        self.sd = {k: (rm(v.float()) if v.dim() >= 2 else v.float()) for k, v in sd.items()}
        self.tok = None
        self.stages: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ helpers
    def W(self, name):
        return self.sd[name]

    def lin(self, x, prefix, act=None, out_round=True):
        """y = act(x @ W^T + b); rounded to bf16 on store unless out_round=False (fp32 outputs of the GEMM)."""
        y = x @ self.W(prefix + ".weight").t()
        b = self.sd.get(prefix + ".bias")
        if b is not None:
            y = y + b
        if act == "gelu":
            y = F.gelu(y)
        elif act == "relu":
            y = F.relu(y)
        return self.r(y) if out_round else y

    def ln(self, x, prefix, eps):
        d = x.shape[-1]
        return self.r(F.layer_norm(x, (d,), self.W(prefix + ".weight"), self.W(prefix + ".bias"), eps))

    def attention(self, q, k, v, scale, causal=False, q_pos0=0, kv_len=None):
        """q [B,Sq,H,D], k/v [B,Sk,H,D] -> [B,Sq,H*D].  fp32 softmax of (QK^T)*scale; probabilities rounded to bf16
        before the PV product as in the fused kernel; normalisation by the fp32 row sum."""
        B, Sq, H, D = q.shape
        Sk = k.shape[1]
        s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
        mask = torch.zeros(B, 1, Sq, Sk, dtype=torch.bool)
        if causal:
            mask = mask | (torch.arange(Sk)[None, :] > (torch.arange(Sq)[:, None] + q_pos0))[None, None]
        if kv_len is not None:
            mask = mask | (torch.arange(Sk)[None, :] >= kv_len[:, None])[:, None, None, :]
        s = s.masked_fill(mask, float("-inf"))
        m = s.max(-1, keepdim=True).values
        p = torch.exp(s - m)
        o = torch.einsum("bhqk,bkhd->bqhd", self.r(p), v) / p.sum(-1).permute(0, 2, 1)[..., None]
        return self.r(o.reshape(B, Sq, H * D))

    # ------------------------------------------------------------------ a1/a2: DINOv2 ($HF/models/dinov2/modeling_dinov2.py)
    def vit_pos_embed(self):
        """interpolate_pos_encoding of transformers 4.32.0: bicubic, scale_factor=((g+0.1)/G, (g+0.1)/G)."""
        cfg = self.cfg
        pe = self.sd["perceiver.vis_encoder.embeddings.position_embeddings"].float()
        G, g, dim = cfg.vit_pos_grid, cfg.grid, cfg.vit_hidden
        if g == G:
            return pe[0]
        cls_pe, patch_pe = pe[:, :1], pe[:, 1:]
        patch_pe = patch_pe.reshape(1, G, G, dim).permute(0, 3, 1, 2)
        sf = (g + 0.1) / G
        patch_pe = F.interpolate(patch_pe, scale_factor=(sf, sf), mode="bicubic", align_corners=False)
        assert patch_pe.shape[-1] == g and patch_pe.shape[-2] == g
        patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat([cls_pe, patch_pe], 1)[0]

    def vit(self, images: torch.Tensor) -> List[torch.Tensor]:
        cfg, r = self.cfg, self.r
        ve = "perceiver.vis_encoder."
        B = images.shape[0]
        H, nh = cfg.vit_hidden, cfg.vit_heads
        hd = H // nh
        # patch embedding = conv14/s14 as a GEMM over unfolded patches (k = c*196 + ky*14 + kx)
        patches = r(F.unfold(images.float(), kernel_size=cfg.patch, stride=cfg.patch).transpose(1, 2))  # [B, g*g, 588]
        w = self.W(ve + "embeddings.patch_embeddings.projection.weight").reshape(H, -1)
        emb = r(patches @ w.t() + self.W(ve + "embeddings.patch_embeddings.projection.bias"))
        pos = self.vit_pos_embed()
        cls = self.W(ve + "embeddings.cls_token").reshape(1, 1, H).expand(B, 1, H)
        x = r(torch.cat([cls + pos[:1], emb + pos[1:]], 1))
        hs = [x]
        for i in range(cfg.vit_layers):
            p = f"{ve}encoder.layer.{i}."
            y = self.ln(x, p + "norm1", cfg.vit_ln_eps)
            q = self.lin(y, p + "attention.attention.query").reshape(B, -1, nh, hd)
            k = self.lin(y, p + "attention.attention.key").reshape(B, -1, nh, hd)
            v = self.lin(y, p + "attention.attention.value").reshape(B, -1, nh, hd)
            a = self.attention(q, k, v, 1.0 / math.sqrt(hd))
            o = a @ self.W(p + "attention.output.dense.weight").t() + self.W(p + "attention.output.dense.bias")
            x = r(o * self.W(p + "layer_scale1.lambda1") + x)
            y = self.ln(x, p + "norm2", cfg.vit_ln_eps)
            h = self.lin(y, p + "mlp.fc1", act="gelu")
            o = h @ self.W(p + "mlp.fc2.weight").t() + self.W(p + "mlp.fc2.bias")
            x = r(o * self.W(p + "layer_scale2.lambda1") + x)
            hs.append(x)
        return hs

    # ------------------------------------------------------------------ a3: groma.py:227-237,361
    def image_tokens(self, last: torch.Tensor) -> torch.Tensor:
        B, _, d = last.shape
        g = self.cfg.grid
        f = last[:, 1:].reshape(B, g, g, d)
        f = torch.cat([f[:, 0::2, 0::2], f[:, 1::2, 0::2], f[:, 0::2, 1::2], f[:, 1::2, 1::2]], -1).reshape(B, g * g // 4, 4 * d)
        h = self.lin(f, "img_txt_bridge.0", act="gelu")
        return self.lin(h, "img_txt_bridge.2")

    # ------------------------------------------------------------------ a4..a9: proposer
    def sine_pos(self) -> torch.Tensor:
        """DeformableDetrSinePositionEmbedding(normalize=True, 128 feats) on an all-valid mask
        ($HF/models/deformable_detr/modeling_deformable_detr.py:340-389) -> [S, d_model] in (y-major token order)."""
        g, D = self.cfg.grid, self.cfg.d_model
        npf = D // 2
        ones = torch.ones(1, g, g)
        y_embed = ones.cumsum(1, dtype=torch.float32)
        x_embed = ones.cumsum(2, dtype=torch.float32)
        eps, scale = 1e-6, 2 * math.pi
        y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps) * scale
        x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps) * scale
        dim_t = torch.arange(npf, dtype=torch.float32)
        dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).reshape(g * g, D)

    def msda_block(self, prefix, query, value_src, ref, hw):
        """DeformableDetrMultiscaleDeformableAttention.forward ($HF ...:556-623), single level."""
        cfg = self.cfg
        nH, P = cfg.ddetr_heads, cfg.n_points
        B, Q, D = query.shape
        w = torch.cat([self.W(prefix + ".sampling_offsets.weight"), self.W(prefix + ".attention_weights.weight")], 0)
        b = torch.cat([self.W(prefix + ".sampling_offsets.bias"), self.W(prefix + ".attention_weights.bias")], 0)
        proj = query @ w.t() + b  # fp32, not rounded
        value = self.lin(value_src, prefix + ".value_proj").reshape(B, -1, nH, D // nH)
        samp = self.r(O.msda_module_core_ref(value, proj, ref, [hw], nH, P))
        return samp

    def proposer(self, hs: List[torch.Tensor], topk_override: Optional[torch.Tensor] = None):
        """topk_override (tests only): continue with another run's two-stage query selection so that two precisions of the
        oracle can be compared query by query (the selection itself is tie-sensitive)."""
        cfg, r = self.cfg, self.r
        dt = "perceiver.ddetr_transformer."
        g, D = cfg.grid, cfg.d_model
        B = hs[0].shape[0]
        S = g * g
        # groma.py:240-244 + ddetr.py:25-45,147-151 (1x1 conv + channel LN == per-token Linear + LN)
        x = r(torch.stack(hs[-4:]).mean(0)[:, 1:])
        w = self.W("perceiver.input_proj.0.0.weight").reshape(D, -1)
        src = r(x @ w.t() + self.W("perceiver.input_proj.0.0.bias"))
        src = self.ln(src, "perceiver.input_proj.0.1", 1e-6)
        pos = r(self.sine_pos() + self.W(dt + "level_embed")[0])          # ddetr_transformer.py:496-516
        # encoder reference points ($HF ...:950-978), valid ratios = 1
        lin = torch.linspace(0.5, g - 0.5, g, dtype=torch.float32) / g
        ry, rx = torch.meshgrid(lin, lin, indexing="ij")
        enc_ref = torch.stack((rx.reshape(-1), ry.reshape(-1)), -1)[None].expand(B, S, 2).contiguous()
        x = src
        for i in range(cfg.enc_layers):
            p = f"{dt}encoder.layers.{i}."
            q = r(x + pos)
            samp = self.msda_block(p + "self_attn", q, x, enc_ref, (g, g))
            h = r(samp @ self.W(p + "self_attn.output_proj.weight").t() + self.W(p + "self_attn.output_proj.bias") + x)
            x = self.ln(h, p + "self_attn_layer_norm", 1e-5)
            t = self.lin(x, p + "fc1", act="relu")
            h = r(t @ self.W(p + "fc2.weight").t() + self.W(p + "fc2.bias") + x)
            x = self.ln(h, p + "final_layer_norm", 1e-5)
        memory = x
        # two-stage proposals (ddetr_transformer.py:383-430,546-568)
        ctr = (torch.arange(g, dtype=torch.float32) + 0.5) / g
        gy, gx = torch.meshgrid(ctr, ctr, indexing="ij")
        prop = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.full((S,), 0.05), torch.full((S,), 0.05)], -1)
        valid = ((prop > 0.01) & (prop < 0.99)).all(-1)
        prop_logit = torch.log(prop / (1 - prop)).masked_fill(~valid[:, None], float("inf"))
        oq = memory.masked_fill(~valid[None, :, None], 0.0)
        eo = self.ln(self.lin(oq, dt + "enc_output"), dt + "enc_output_norm", 1e-5)
        cls = self.lin(eo, dt + "class_embed_enc", out_round=False)[..., 0]          # fp32 [B,S]
        nb = cfg.dec_layers
        t = self.lin(eo, f"{dt}bbox_embed.{nb}.layers.0", act="relu")
        t = self.lin(t, f"{dt}bbox_embed.{nb}.layers.1", act="relu")
        delta = self.lin(t, f"{dt}bbox_embed.{nb}.layers.2", out_round=False)      # fp32 [B,S,4]
        coord_logits = delta + prop_logit[None]
        topk = torch.stack([torch.from_numpy(np.argsort(-cls[b].numpy(), kind="stable")[:cfg.num_queries].copy()) for b in range(B)])
        if topk_override is not None:
            topk = topk_override
        tk = torch.gather(coord_logits, 1, topk[..., None].expand(-1, -1, 4))
        ref = tk.sigmoid()
        npf = D // 2
        dim_t = torch.arange(npf, dtype=torch.float32)
        dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
        pp = (ref * (2 * math.pi))[:, :, :, None] / dim_t
        pos512 = r(torch.stack((pp[..., 0::2].sin(), pp[..., 1::2].cos()), dim=4).flatten(2))
        pt = self.ln(self.lin(pos512, dt + "pos_trans"), dt + "pos_trans_norm", 1e-5)
        query_pos = pt[..., :D]
        h = r(self.W(dt + "query_position_embeddings.weight"))[None].expand(B, -1, -1)
        nH = cfg.ddetr_heads
        hd = D // nH
        keep = {}
        for i in range(cfg.dec_layers):   # DeformableDetrDecoderX.forward (ddetr_transformer.py:77-202); refs never advance (T4)
            p = f"{dt}decoder.layers.{i}."
            qk = r(h + query_pos)
            q = self.lin(qk, p + "self_attn.q_proj").reshape(B, -1, nH, hd)
            k = self.lin(qk, p + "self_attn.k_proj").reshape(B, -1, nH, hd)
            v = self.lin(h, p + "self_attn.v_proj").reshape(B, -1, nH, hd)
            a = self.attention(q, k, v, hd ** -0.5)
            h = self.ln(r(a @ self.W(p + "self_attn.out_proj.weight").t() + self.W(p + "self_attn.out_proj.bias") + h),
                        p + "self_attn_layer_norm", 1e-5)
            qc = r(h + query_pos)
            samp = self.msda_block(p + "encoder_attn", qc, memory, ref, (g, g))
            h = self.ln(r(samp @ self.W(p + "encoder_attn.output_proj.weight").t() + self.W(p + "encoder_attn.output_proj.bias") + h),
                        p + "encoder_attn_layer_norm", 1e-5)
            t = self.lin(h, p + "fc1", act="relu")
            h = self.ln(r(t @ self.W(p + "fc2.weight").t() + self.W(p + "fc2.bias") + h), p + "final_layer_norm", 1e-5)
            keep[i] = h
        L = cfg.dec_layers

        def bbox(i, hh):
            t = self.lin(hh, f"{dt}bbox_embed.{i}.layers.0", act="relu")
            t = self.lin(t, f"{dt}bbox_embed.{i}.layers.1", act="relu")
            return self.lin(t, f"{dt}bbox_embed.{i}.layers.2", out_round=False)

        def inv_sig(x, eps=1e-5):
            x = x.clamp(0, 1)
            return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))

        # ddetr_transformer.py:150-166,696-712: box refinement chains through the *recorded* new_reference_points
        r1 = (bbox(L - 2, keep[L - 2]) + inv_sig(ref)).sigmoid() if L >= 2 else ref
        pred = (bbox(L - 1, keep[L - 1]) + inv_sig(r1)).sigmoid()
        coco = self.lin(keep[L - 1], f"{dt}class_embed_coco.{L - 1}", out_round=False)[..., 0]
        sa1b = self.lin(keep[L - 1], f"{dt}class_embed_sa1b.{L - 1}", out_round=False)[..., 0]
        scores = coco.sigmoid() ** 0.4 * sa1b.sigmoid() ** 0.6                    # groma.py:247-249
        self.stages.update(dict(ddetr_src=src, memory=memory, enc_cls=cls, topk=topk, ref_init=ref, dec_last=keep[L - 1],
                                enc_obj_query=eo, prop_logit=prop_logit, topk_coord_logits=tk, pos512=pos512,
                                query_pos=query_pos, tgt=r(self.W(dt + "query_position_embeddings.weight"))[None].expand(B, -1, -1)))
        return pred, scores, {"coco": coco, "sa1b": sa1b}

    # ------------------------------------------------------------------ N1: detector post-processing (train/train_det.py:97-131)
    @staticmethod
    def post_process(coco_logits, pred_boxes, target_sizes, threshold=0.0, top_k=100):
        """coco_logits [B,Q,C], pred_boxes [B,Q,4] cxcywh in (0,1), target_sizes [B,2] (h, w).  torch.topk's order among
        equal values is unspecified; the restatement takes them by lower flat index (stable argsort)."""
        B, Q, C = coco_logits.shape
        prob = coco_logits.float().sigmoid().reshape(B, -1)
        k = min(top_k, prob.shape[1])
        idx = torch.stack([torch.from_numpy(np.argsort(-prob[b].numpy(), kind="stable")[:k].copy()) for b in range(B)])
        scores = torch.gather(prob, 1, idx)
        qi, labels = torch.div(idx, C, rounding_mode="floor"), idx % C
        bx = pred_boxes.float()
        xyxy = torch.cat([bx[..., :2] - 0.5 * bx[..., 2:], bx[..., :2] + 0.5 * bx[..., 2:]], -1)
        xyxy = torch.gather(xyxy, 1, qi[..., None].repeat(1, 1, 4))
        ts = torch.as_tensor(target_sizes, dtype=torch.float32)
        img_h, img_w = ts.unbind(1)
        xyxy = xyxy * torch.stack([img_w, img_h, img_w, img_h], 1)[:, None, :]
        return [{"scores": s[s > threshold], "labels": l[s > threshold], "boxes": b[s > threshold]} for s, l, b in zip(scores, labels, xyxy)]

    # ------------------------------------------------------------------ a10: region selection (groma.py:251-280)
    def select_regions(self, pred_boxes, scores, refer_boxes=None, ground_boxes=None, rng_draw=True):
        cfg = self.cfg
        B = pred_boxes.shape[0]
        selected, nms_inds_all = [], []
        for i in range(B):
            rb = refer_boxes[i] if refer_boxes is not None else torch.empty((0, 4))
            gb = ground_boxes[i] if ground_boxes is not None else torch.empty((0, 4))
            sc = torch.cat((scores[i], torch.ones(rb.shape[0]), torch.ones(gb.shape[0]) * 0.2))
            bx = torch.cat((pred_boxes[i], rb.float(), gb.float()))
            xyxy = torch.cat([bx[:, :2] - 0.5 * bx[:, 2:], bx[:, :2] + 0.5 * bx[:, 2:]], -1)
            inds = O.nms_ref(xyxy.numpy(), sc.numpy(), cfg.nms_thres, 0, cfg.box_score_thres, cfg.max_region_num)
            nms_inds_all.append(inds)
            if len(inds) > 0:
                bx = bx[torch.from_numpy(inds)]
                if rng_draw:
                    bx = bx[torch.randperm(len(bx))]      # global CPU RNG, exactly as groma.py:275
            else:
                mi = int(torch.max(sc, dim=0).indices)
                bx = bx[mi:mi + 1]
            selected.append(bx)
        return selected, nms_inds_all

    # ------------------------------------------------------------------ a12/a13: region encoder (groma/model/roi_align.py)
    def region_encoder(self, hs: List[torch.Tensor], boxes: List[torch.Tensor]) -> List[torch.Tensor]:
        cfg, r = self.cfg, self.r
        g, C = cfg.grid, cfg.vit_hidden
        B = hs[0].shape[0]
        re_ = "region_encoder."
        feats = [h[:, 1:].reshape(B, g, g, C).permute(0, 3, 1, 2) for h in hs[-3:]]
        sizes = [g * 4, g * 2, g]
        xs = []
        for l in range(3):   # roi_align.py:215-228 upsample, :118-126,180-189 coord concat + 1x1 conv
            s = sizes[l]
            f = r(F.interpolate(feats[l], size=(s, s), mode="bilinear", align_corners=True))
            xr = torch.linspace(-1, 1, s)
            yy, xx = torch.meshgrid(xr, xr, indexing="ij")
            coord = r(torch.stack([xx, yy], 0))[None].expand(B, 2, s, s)
            f = torch.cat([f, coord], 1)
            w = self.W(f"{re_}mlvl_fuse.input_conv.{l}.weight").reshape(C, C + 2)
            y = torch.einsum("bchw,oc->bohw", f, w) + self.W(f"{re_}mlvl_fuse.input_conv.{l}.bias")[None, :, None, None]
            xs.append(r(y))
        q = C // 4
        for k in range(cfg.fuse_rounds):   # _single_shuffle (roi_align.py:150-178) + ConvModule conv->GN->ReLU
            w = self.W(f"{re_}mlvl_fuse.fuse_convs.{k}.conv.weight")
            gw, gb = self.W(f"{re_}mlvl_fuse.fuse_convs.{k}.gn.weight"), self.W(f"{re_}mlvl_fuse.fuse_convs.{k}.gn.bias")
            new = []
            for l in range(3):
                top, dn = min(l + 1, 2), max(l - 1, 0)
                s = sizes[l]
                ft = r(F.interpolate(xs[top][:, 2 * q:][:, q:], size=(s, s), mode="bilinear", align_corners=True))
                fd = r(F.interpolate(xs[dn][:, 2 * q:][:, :q], size=(s, s), mode="bilinear", align_corners=True))
                fin = torch.cat([xs[l][:, :2 * q], ft, fd], 1)
                y = r(F.conv2d(fin, w, padding=1))
                new.append(r(F.relu(F.group_norm(y, cfg.gn_groups, gw, gb, 1e-5))))
            xs = new
        # MlvlRoIExtractor.forward (roi_align.py:274-327): cxcywh*image_size fed as xyxy (T1), strides 14/8,14/4,14/2 (T3)
        rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b.float() * cfg.image_size], 1) for i, b in enumerate(boxes)])
        allb = torch.cat(boxes).float()
        if len(rois) == 0:
            return [torch.zeros(0, cfg.llm_hidden) for _ in boxes]
        acc = 0
        for l in range(3):
            scale = (8, 4, 2)[l] / 14.0
            rf = r(O.roi_align_ref(xs[l], rois, cfg.roi_out, scale, cfg.roi_sampling, True))
            acc = acc + F.conv2d(rf, self.W(f"{re_}roi_align.pconvs.{l}.weight"), padding=1)
        bias = sum(self.W(f"{re_}roi_align.pconvs.{l}.bias") for l in range(3))
        fused = r(F.relu(acc + bias[None, :, None, None]))
        flat = self.lin(fused.flatten(1), re_ + "roi_align.flatten_linear")
        # pos_embedd (roi_align.py:254-261): first Linear has K=4 -> computed in fp32 on fp32 boxes
        p = r(F.relu(allb @ self.W(re_ + "roi_align.pos_embedd.0.weight").t() + self.W(re_ + "roi_align.pos_embedd.0.bias")))
        p = self.ln(p, re_ + "roi_align.pos_embedd.2", 1e-5)
        p = self.lin(p, re_ + "roi_align.pos_embedd.3", act="relu")
        p = self.ln(p, re_ + "roi_align.pos_embedd.5", 1e-5)
        z = r(flat + p)
        out = self.lin(z, re_ + "roi_align.updims")
        self.stages.update(dict(fused_maps=[x.permute(0, 2, 3, 1) for x in xs], roi_fused=fused, region_flat=flat))
        res, o = [], 0
        for b in boxes:
            res.append(out[o:o + len(b)])
            o += len(b)
        return res

    # ------------------------------------------------------------------ a11/a14: integer bookkeeping (groma.py:283-357)
    def init_special_token_id(self, tokenizer):
        self.tok = dict(pad=tokenizer.pad_token_id, img=tokenizer.convert_tokens_to_ids(["<image>"])[0],
                        reg=tokenizer.convert_tokens_to_ids(["<region>"])[0],
                        rbox=tokenizer.convert_tokens_to_ids(["<refer_box>"])[0],
                        rfeat=tokenizer.convert_tokens_to_ids(["<refer_feat>"])[0],
                        gbox=tokenizer.convert_tokens_to_ids(["<ground_box>"])[0],
                        box_idx=tokenizer.convert_tokens_to_ids([f"<r{i}>" for i in range(100)]))

    @staticmethod
    def box_iou(a, b):
        area = lambda t: (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
        lt = torch.max(a[:, None, :2], b[None, :, :2]); rb = torch.min(a[:, None, 2:], b[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        return inter / (area(a)[:, None] + area(b)[None, :] - inter)

    def match_refer_ground(self, input_ids, selected, refer_boxes, ground_boxes, labels=None):
        """groma.py:283-309 -- in-place on input_ids (and labels), returns refer_box_inds."""
        t = self.tok
        c2c = lambda b: torch.cat([b[:, :2] - 0.5 * b[:, 2:], b[:, :2] + 0.5 * b[:, 2:]], -1)
        refer_inds = []
        for i in range(input_ids.shape[0]):
            if (input_ids[i] == t["rbox"]).any():
                m = torch.max(self.box_iou(c2c(refer_boxes[i].float()), c2c(selected[i])), dim=-1).indices
                refer_inds.append(m)
                ids = torch.tensor(t["box_idx"])[m]
                mask = input_ids[i] == t["rbox"]
                input_ids[i].masked_scatter_(mask, ids)
            else:
                refer_inds.append(torch.zeros(0, dtype=torch.long))
            if (input_ids[i] == t["gbox"]).any():
                m = torch.max(self.box_iou(c2c(ground_boxes[i].float()), c2c(selected[i])), dim=-1).indices
                ids = torch.tensor(t["box_idx"])[m]
                mask = input_ids[i] == t["gbox"]
                input_ids[i].masked_scatter_(mask, ids)
                if labels is not None:
                    labels[i].masked_scatter_(mask, ids)
        return refer_inds

    def assemble_labels(self, input_ids: torch.Tensor, labels: torch.Tensor, num_regions: List[int], n_img_tokens: int):
        """groma.py:338-353: labels follow the same expansion with IGNORE_INDEX (-100) under the image / region placeholders."""
        t = self.tok
        new = []
        for i in range(input_ids.shape[0]):
            ids, lab = input_ids[i], labels[i]
            ip = int((ids == t["img"]).nonzero()[0]); rp = int((ids == t["reg"]).nonzero()[0])
            pp = (ids == t["pad"]).nonzero()
            pe = int(pp[0]) if len(pp) > 0 else len(ids)
            new.append(torch.cat((lab[:ip], torch.full((n_img_tokens,), -100, dtype=torch.long), lab[ip + 1:rp],
                                  torch.full((2 * num_regions[i],), -100, dtype=torch.long), lab[rp + 1:pe])))
        return torch.nn.utils.rnn.pad_sequence(new, batch_first=True, padding_value=-100)

    def assemble(self, input_ids: torch.Tensor, num_regions: List[int], n_img_tokens: int):
        """groma.py:317-357: expand <image>/<region> placeholders, cut at first pad, right-pad."""
        t = self.tok
        new = []
        for i in range(input_ids.shape[0]):
            ids = input_ids[i]
            ip = int((ids == t["img"]).nonzero()[0]); rp = int((ids == t["reg"]).nonzero()[0])
            pp = (ids == t["pad"]).nonzero()
            pe = int(pp[0]) if len(pp) > 0 else len(ids)
            assert ip < rp
            regs = torch.tensor([v for j in range(num_regions[i]) for v in (t["box_idx"][j], t["reg"])], dtype=torch.long)
            new.append(torch.cat((ids[:ip], torch.full((n_img_tokens,), t["img"], dtype=torch.long), ids[ip + 1:rp], regs, ids[rp + 1:pe])))
        out = torch.nn.utils.rnn.pad_sequence(new, batch_first=True, padding_value=t["pad"])
        return out, out.ne(t["pad"])

    # ------------------------------------------------------------------ a15..a18: LLaMA ($HF/models/llama/modeling_llama.py)
    def embed(self, ids):
        V = self.cfg.vocab
        e = self.W("llm.model.embed_tokens.weight")[ids.clamp(max=V - 1)]
        n = self.W("new_input_embs.weight")[(ids - V).clamp(min=0)]
        return torch.where((ids >= V)[..., None], n, e)

    def rms(self, x, name):
        v = x.pow(2).mean(-1, keepdim=True)
        return self.r(self.W(name + ".weight") * self.r(x * torch.rsqrt(v + self.cfg.rms_eps)))

    def rope_tables(self, n):
        D = self.cfg.head_dim
        inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, D, 2).float() / D))
        fr = torch.outer(torch.arange(n).float(), inv)
        return fr.cos(), fr.sin()

    def rope(self, t, pos0):
        # t [B,T,H,D]; rotate-half (modeling_llama.py:138-168), positions = arange
        B, T, H, D = t.shape
        cos, sin = self.rope_tables(pos0 + T)
        cos = torch.cat([cos, cos], -1)[pos0:pos0 + T][None, :, None]
        sin = torch.cat([sin, sin], -1)[pos0:pos0 + T][None, :, None]
        rot = torch.cat([-t[..., D // 2:], t[..., :D // 2]], -1)
        return self.r(t * cos + rot * sin)

    def llm(self, x, kv=None, kv_len=None, pos0=0):
        """x [B,T,hidden] -> (final-normed hidden, kv list).  kv: list of (k,v) [B,ctx,H,D] from earlier steps."""
        cfg, r = self.cfg, self.r
        B, T, Hd = x.shape
        nh, hd = cfg.llm_heads, cfg.head_dim
        new_kv = []
        for i in range(cfg.llm_layers):
            p = f"llm.model.layers.{i}."
            y = self.rms(x, p + "input_layernorm")
            q = r(y @ self.W(p + "self_attn.q_proj.weight").t()).reshape(B, T, nh, hd)
            k = r(y @ self.W(p + "self_attn.k_proj.weight").t()).reshape(B, T, nh, hd)
            v = r(y @ self.W(p + "self_attn.v_proj.weight").t()).reshape(B, T, nh, hd)
            q, k = self.rope(q, pos0), self.rope(k, pos0)
            if kv is not None:
                k = torch.cat([kv[i][0], k], 1); v = torch.cat([kv[i][1], v], 1)
            new_kv.append((k, v))
            a = self.attention(q, k, v, 1.0 / math.sqrt(hd), causal=True, q_pos0=pos0, kv_len=kv_len)
            x = r(a @ self.W(p + "self_attn.o_proj.weight").t() + x)
            y = self.rms(x, p + "post_attention_layernorm")
            gte = y @ self.W(p + "mlp.gate_proj.weight").t()
            up = y @ self.W(p + "mlp.up_proj.weight").t()
            x = r(r(F.silu(gte) * up) @ self.W(p + "mlp.down_proj.weight").t() + x)
        return self.rms(x, "llm.model.norm"), new_kv

    def logits(self, h):
        w = torch.cat([self.W("llm.lm_head.weight"), self.W("extra_lm_head.weight")], 0)
        return h @ w.t()   # fp32

    # ------------------------------------------------------------------ GromaModel.forward, prefill branch (groma.py:217-402)
    def forward_prefill(self, input_ids, images, refer_boxes=None, ground_boxes=None, selected_override=None, labels=None, topk_override=None):
        assert self.tok is not None
        hs = self.vit(images)
        self.stages["vit_last"] = hs[-1]
        self.stages["vit_hs"] = hs[-4:]
        img_tok = self.image_tokens(hs[-1])
        pred, scores, logits = self.proposer(hs, topk_override)
        if selected_override is not None:
            selected, nms_inds = selected_override, None
        else:
            selected, nms_inds = self.select_regions(pred, scores, refer_boxes, ground_boxes)
        refer_inds = self.match_refer_ground(input_ids, selected, refer_boxes, ground_boxes, labels)
        region = self.region_encoder(hs, selected)
        refer_feats = [rf[ind] for rf, ind in zip(region, refer_inds)]
        ids, mask = self.assemble(input_ids, [len(x) for x in region], img_tok.shape[1])
        x = self.embed(ids).clone()
        t = self.tok
        x[ids == t["img"]] = img_tok.reshape(-1, x.shape[-1])
        x[ids == t["reg"]] = torch.cat(region)
        if (ids == t["rfeat"]).any():
            x[ids == t["rfeat"]] = torch.cat(refer_feats)
        kv_len = mask.sum(1)
        h, kv = self.llm(x, kv_len=kv_len)
        out = dict(logits=self.logits(h), kv=kv, input_ids=ids, attention_mask=mask, pred_boxes=pred, scores=scores,
                   det_logits=logits, selected_boxes=selected, nms_inds=nms_inds, image_features=img_tok,
                   region_features=torch.cat(region), inputs_embeds=x)
        if labels is not None:      # groma.py:404-415: shifted cross entropy, mean over the non-ignored targets
            lab = self.assemble_labels(input_ids, labels, [len(x) for x in region], img_tok.shape[1])
            lg = out["logits"]
            out["labels"] = lab
            out["loss"] = F.cross_entropy(lg[:, :-1].reshape(-1, lg.shape[-1]).float(), lab[:, 1:].reshape(-1), ignore_index=-100)
        return out

    def forward_decode(self, token_ids, kv):
        """groma.py:376-402 decode branch: all-ones mask over past+1 (T7), positions = past length."""
        past = kv[0][0].shape[1]
        x = self.embed(token_ids)
        h, kv = self.llm(x, kv=kv, pos0=past)
        return self.logits(h), kv

    def generate(self, input_ids, images, max_new_tokens, refer_boxes=None, ground_boxes=None, selected_override=None):
        """Greedy search as the eval scripts drive it (eval/run_groma.py:82-95); no EOS stop (fixed-length for parity)."""
        out = self.forward_prefill(input_ids.clone(), images, refer_boxes, ground_boxes, selected_override)
        # HF greedy: next token from the LAST position of the padded batch (right padding)
        nxt = out["logits"][:, -1].argmax(-1)
        seq, kv, all_logits = [nxt], out["kv"], [out["logits"][:, -1]]
        for _ in range(max_new_tokens - 1):
            lg, kv = self.forward_decode(nxt[:, None], kv)
            nxt = lg[:, -1].argmax(-1)
            seq.append(nxt)
            all_logits.append(lg[:, -1])
        out["new_tokens"] = torch.stack(seq, 1)
        out["step_logits"] = torch.stack(all_logits, 1)
        return out
