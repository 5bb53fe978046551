"""Deterministic synthetic weights keyed by the reference's state-dict names (there is no network for checkpoints:
parity tests and bench.py run on random-init weights of the real architecture).

Names follow GromaModel.state_dict() of the reference with transformers 4.32 module names:
perceiver.vis_encoder.*, perceiver.input_proj.0.{0,1}, perceiver.ddetr_transformer.*, llm.model.*, llm.lm_head,
img_txt_bridge.{0,2}, region_encoder.{mlvl_fuse,roi_align}.*, extra_lm_head, new_input_embs.
Every tensor is drawn from its own generator seeded by (seed, crc32(name)), so the value of one tensor never depends
on which others exist; both the oracle and the product load the SAME dict in parity tests."""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

from .config import PathConfig


_DEVICE = "cpu"


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device=_DEVICE)
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def _randn(*shape, generator):
    return torch.randn(*shape, generator=generator, device=generator.device)


def make_state_dict(cfg: PathConfig, seed: int = 0, perturb_norms: bool = True, std: float = 0.02,
                    dtype=torch.float32, device: str = "cpu") -> "OrderedDict[str, torch.Tensor]":
    global _DEVICE
    _DEVICE = device  # device='cuda' draws directly on the GPU (different values than 'cpu'; used by bench.py only)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def normal(name, *shape, s=std):
        sd[name] = (_randn(*shape, generator=_gen(seed, name)) * s).to(dtype)

    def linear(prefix, out_f, in_f, bias=True, s=std):
        normal(prefix + ".weight", out_f, in_f, s=s)
        if bias:
            normal(prefix + ".bias", out_f, s=s)

    def norm(prefix, dim, bias=True):
        w = torch.ones(dim, device=_DEVICE)
        b = torch.zeros(dim, device=_DEVICE)
        if perturb_norms:
            w = w + _randn(dim, generator=_gen(seed, prefix + ".weight")) * 0.05
            b = b + _randn(dim, generator=_gen(seed, prefix + ".bias")) * 0.05
        sd[prefix + ".weight"] = w.to(dtype)
        if bias:
            sd[prefix + ".bias"] = b.to(dtype)

    H = cfg.vit_hidden
    # ---- DINOv2
    ve = "perceiver.vis_encoder."
    normal(ve + "embeddings.cls_token", 1, 1, H)
    normal(ve + "embeddings.mask_token", 1, H)
    normal(ve + "embeddings.position_embeddings", 1, cfg.vit_pos_grid ** 2 + 1, H)
    normal(ve + "embeddings.patch_embeddings.projection.weight", H, 3, cfg.patch, cfg.patch)
    normal(ve + "embeddings.patch_embeddings.projection.bias", H)
    for i in range(cfg.vit_layers):
        p = f"{ve}encoder.layer.{i}."
        norm(p + "norm1", H)
        for n in ("query", "key", "value"):
            linear(p + f"attention.attention.{n}", H, H)
        linear(p + "attention.output.dense", H, H)
        sd[p + "layer_scale1.lambda1"] = (torch.ones(H, device=_DEVICE) + (_randn(H, generator=_gen(seed, p + "ls1")) * 0.05 if perturb_norms else 0)).to(dtype)
        norm(p + "norm2", H)
        linear(p + "mlp.fc1", cfg.vit_mlp, H)
        linear(p + "mlp.fc2", H, cfg.vit_mlp)
        sd[p + "layer_scale2.lambda1"] = (torch.ones(H, device=_DEVICE) + (_randn(H, generator=_gen(seed, p + "ls2")) * 0.05 if perturb_norms else 0)).to(dtype)
    norm(ve + "layernorm", H)  # present in the checkpoint, unused by Groma

    # ---- input projection (1 level): Conv2d 1x1 + channel LayerNorm (groma/model/ddetr.py:147-151)
    D = cfg.d_model
    normal("perceiver.input_proj.0.0.weight", D, H, 1, 1)
    normal("perceiver.input_proj.0.0.bias", D)
    norm("perceiver.input_proj.0.1", D)

    # ---- Deformable-DETR transformer
    dt = "perceiver.ddetr_transformer."
    nH, P, L = cfg.ddetr_heads, cfg.n_points, 1

    def msda(prefix):
        # HF _reset_parameters: grid-pattern bias on the offsets (modeling_deformable_detr.py:833-851) + small noise
        linear(prefix + ".sampling_offsets", nH * L * P * 2, D)
        thetas = torch.arange(nH, dtype=torch.float32, device=_DEVICE) * (2.0 * math.pi / nH)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(nH, 1, 1, 2).repeat(1, L, P, 1)
        for i in range(P):
            grid[:, :, i, :] *= i + 1
        sd[prefix + ".sampling_offsets.bias"] = (grid.reshape(-1) + sd[prefix + ".sampling_offsets.bias"].float()).to(dtype)
        linear(prefix + ".attention_weights", nH * L * P, D, s=0.2)
        linear(prefix + ".value_proj", D, D, s=0.05)
        linear(prefix + ".output_proj", D, D, s=0.05)

    for i in range(cfg.enc_layers):
        p = f"{dt}encoder.layers.{i}."
        msda(p + "self_attn")
        norm(p + "self_attn_layer_norm", D)
        linear(p + "fc1", cfg.ddetr_ffn, D, s=0.05)
        linear(p + "fc2", D, cfg.ddetr_ffn, s=0.05)
        norm(p + "final_layer_norm", D)
    for i in range(cfg.dec_layers):
        p = f"{dt}decoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            linear(p + "self_attn." + n, D, D, s=0.05)
        norm(p + "self_attn_layer_norm", D)
        msda(p + "encoder_attn")
        norm(p + "encoder_attn_layer_norm", D)
        linear(p + "fc1", cfg.ddetr_ffn, D, s=0.05)
        linear(p + "fc2", D, cfg.ddetr_ffn, s=0.05)
        norm(p + "final_layer_norm", D)
    normal(dt + "level_embed", 1, D, s=1.0)
    normal(dt + "query_position_embeddings.weight", cfg.num_queries, D, s=1.0)
    linear(dt + "enc_output", D, D, s=0.05)
    norm(dt + "enc_output_norm", D)
    linear(dt + "pos_trans", 2 * D, 2 * D, s=0.05)
    norm(dt + "pos_trans_norm", 2 * D)
    bias_value = -math.log((1 - 0.01) / 0.01)
    linear(dt + "class_embed_enc", 1, D, s=0.1)
    sd[dt + "class_embed_enc.bias"] = torch.full((1,), bias_value, device=_DEVICE).to(dtype)
    for i in range(cfg.dec_layers):
        for n in ("class_embed_coco", "class_embed_sa1b"):
            linear(f"{dt}{n}.{i}", 1, D, s=0.1)
            sd[f"{dt}{n}.{i}.bias"] = torch.full((1,), bias_value + 3.0, device=_DEVICE).to(dtype)  # +3: scores spread around 0.1
    for i in range(cfg.dec_layers + 1):
        for j, (o, k) in enumerate([(D, D), (D, D), (4, D)]):
            # reference zero-inits the last layer (ddetr_transformer.py:334-335); small noise keeps boxes non-degenerate
            linear(f"{dt}bbox_embed.{i}.layers.{j}", o, k, s=0.1 if j == 2 else 0.05)

    # ---- bridge, region encoder
    T = cfg.llm_hidden
    linear("img_txt_bridge.0", T, 4 * H)
    linear("img_txt_bridge.2", T, T)
    re_ = "region_encoder."
    for l in range(3):
        normal(f"{re_}mlvl_fuse.input_conv.{l}.weight", H, H + 2, 1, 1)
        normal(f"{re_}mlvl_fuse.input_conv.{l}.bias", H)
    for k in range(cfg.fuse_rounds):
        normal(f"{re_}mlvl_fuse.fuse_convs.{k}.conv.weight", H, H, 3, 3, s=0.01)
        norm(f"{re_}mlvl_fuse.fuse_convs.{k}.gn", H)
    for l in range(3):
        normal(f"{re_}roi_align.pconvs.{l}.weight", H, H, 3, 3, s=0.01)
        normal(f"{re_}roi_align.pconvs.{l}.bias", H, s=0.01)
    linear(re_ + "roi_align.pos_embedd.0", cfg.pos_hidden, 4, s=0.5)
    norm(re_ + "roi_align.pos_embedd.2", cfg.pos_hidden)
    linear(re_ + "roi_align.pos_embedd.3", cfg.region_mid, cfg.pos_hidden, s=0.05)
    norm(re_ + "roi_align.pos_embedd.5", cfg.region_mid)
    linear(re_ + "roi_align.updims", T, cfg.region_mid)
    linear(re_ + "roi_align.flatten_linear", cfg.region_mid, H * cfg.roi_out ** 2, s=0.005)

    # ---- LLaMA
    normal("llm.model.embed_tokens.weight", cfg.vocab, T)
    for i in range(cfg.llm_layers):
        p = f"llm.model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            linear(p + "self_attn." + n, T, T, bias=False)
        linear(p + "mlp.gate_proj", cfg.llm_inter, T, bias=False)
        linear(p + "mlp.up_proj", cfg.llm_inter, T, bias=False)
        linear(p + "mlp.down_proj", T, cfg.llm_inter, bias=False)
        norm(p + "input_layernorm", T, bias=False)
        norm(p + "post_attention_layernorm", T, bias=False)
    norm("llm.model.norm", T, bias=False)
    linear("llm.lm_head", cfg.vocab, T, bias=False)
    linear("extra_lm_head", cfg.num_new_token, T, bias=False)
    normal("new_input_embs.weight", cfg.num_new_token, T)
    return sd
