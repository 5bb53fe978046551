"""Pin the oracle's restatement of the THIRD-PARTY arithmetic on the path (transformers: Dinov2, LLaMA, Deformable-DETR layers) against
the transformers modules installed in this image, on the same weights (fp32, miniature shapes, CPU).

The reference pins transformers==4.32.0 (`pyproject.toml:19`); the image has 5.x, whose modules compute the same
functions for these blocks (the one known difference, Dinov2 `interpolate_pos_encoding`, is avoided by feeding the trained
grid size).  The reference's own glue (`GromaModel.forward`, `DeformableDetrDecoderX`, two-stage head, region encoder) has no
counterpart to compare with and stays pinned only by the committed golden tensors."""
import math

import pytest
import torch

from oracle.config import tiny_config
from oracle.groma_oracle import Oracle
from oracle.weights import make_state_dict

tr = pytest.importorskip("transformers")


def close(a, b, tol=2e-4):
    a, b = a.float(), b.float()
    err = ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()
    assert err < tol, err


@pytest.fixture(scope="module")
def setup():
    cfg = tiny_config(image_size=518)                 # 37x37 patches = the trained position grid: no pos-embed resize (T11)
    sd = make_state_dict(cfg, seed=0, perturb_norms=True)
    return cfg, sd, Oracle(cfg, sd, "fp32")


def test_dinov2_hidden_states_match_hf(setup):
    cfg, sd, o = setup
    hc = tr.Dinov2Config(hidden_size=cfg.vit_hidden, num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                         mlp_ratio=cfg.vit_mlp // cfg.vit_hidden, image_size=518, patch_size=cfg.patch, layer_norm_eps=cfg.vit_ln_eps,
                         attn_implementation="eager")
    m = tr.Dinov2Model(hc).eval()
    pfx = "perceiver.vis_encoder."
    res = m.load_state_dict({k[len(pfx):]: v.float() for k, v in sd.items() if k.startswith(pfx)}, strict=False)
    assert not res.unexpected_keys and all(k.startswith(("embeddings.mask_token", "layernorm.")) for k in res.missing_keys), res
    g = torch.Generator().manual_seed(0)
    images = torch.randn(2, 3, 518, 518, generator=g)
    with torch.no_grad():
        want = m(pixel_values=images, output_hidden_states=True).hidden_states
    got = o.vit(images)
    assert len(got) == len(want) == cfg.vit_layers + 1
    for a, b in zip(got, want):
        close(a, b)


def test_llama_prefill_and_cached_decode_match_hf(setup):
    cfg, sd, o = setup
    hc = tr.LlamaConfig(hidden_size=cfg.llm_hidden, intermediate_size=cfg.llm_inter, num_hidden_layers=cfg.llm_layers,
                        num_attention_heads=cfg.llm_heads, num_key_value_heads=cfg.llm_heads, vocab_size=cfg.vocab,
                        rms_norm_eps=cfg.rms_eps, max_position_embeddings=cfg.max_pos, attn_implementation="eager")
    m = tr.LlamaForCausalLM(hc).eval()
    res = m.load_state_dict({k[len("llm."):]: v.float() for k, v in sd.items() if k.startswith("llm.")}, strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res
    g = torch.Generator().manual_seed(1)
    B, T = 2, 11
    x = torch.randn(B, T, cfg.llm_hidden, generator=g) * 0.5
    x1 = torch.randn(B, 1, cfg.llm_hidden, generator=g) * 0.5
    x2 = torch.randn(B, 1, cfg.llm_hidden, generator=g) * 0.5
    with torch.no_grad():
        w0 = m(inputs_embeds=x, use_cache=True)
        w1 = m(inputs_embeds=x1, past_key_values=w0.past_key_values, use_cache=True)
        w2 = m(inputs_embeds=x2, past_key_values=w1.past_key_values, use_cache=True)
    h0, kv = o.llm(x)
    close(o.logits(h0)[..., :cfg.vocab], w0.logits)
    h1, kv = o.llm(x1, kv=kv, pos0=T)
    close(o.logits(h1)[..., :cfg.vocab], w1.logits)
    h2, kv = o.llm(x2, kv=kv, pos0=T + 1)
    close(o.logits(h2)[..., :cfg.vocab], w2.logits)


def _ddetr_modules(cfg):
    from transformers.models.deformable_detr import modeling_deformable_detr as M
    from groma.model.groma import _ddetr_cfg
    return M, _ddetr_cfg(cfg)


def test_sine_position_embedding_matches_hf(setup):
    cfg, sd, o = setup
    M, _ = _ddetr_modules(cfg)
    g = cfg.grid
    pe = M.DeformableDetrSinePositionEmbedding(cfg.d_model // 2, normalize=True)
    want = pe(torch.Size([1, cfg.d_model, g, g]), "cpu", torch.float32, mask=torch.ones(1, g, g, dtype=torch.bool))
    want = want[0] if want.shape[1] == g * g else want[0].flatten(1).t()     # 5.x returns [B, S, D]; 4.32 returned [B, D, H, W]
    close(o.sine_pos(), want, tol=1e-6)


def _load(layer, sd, prefix, rename):
    out = {}
    for k, v in sd.items():
        if k.startswith(prefix):
            n = k[len(prefix):]
            for a, b in rename:
                n = n.replace(a, b)
            out[n] = v.float()
    res = layer.load_state_dict(out, strict=True)
    return res


def test_ddetr_encoder_layers_match_hf(setup):
    """ddetr_src --(6 x DeformableDetrEncoderLayer)--> memory, with the oracle's own position embeddings / reference points."""
    cfg, sd, o = setup
    M, dc = _ddetr_modules(cfg)
    g0 = torch.Generator().manual_seed(2)
    hs = o.vit(torch.randn(2, 3, 518, 518, generator=g0))
    o.proposer(hs)
    src, memory = o.stages["ddetr_src"], o.stages["memory"]
    g, B, S = cfg.grid, src.shape[0], src.shape[1]
    dt = "perceiver.ddetr_transformer."
    pos = (o.sine_pos() + sd[dt + "level_embed"][0].float())[None].expand(B, -1, -1)
    lin = torch.linspace(0.5, g - 0.5, g, dtype=torch.float32) / g
    ry, rx = torch.meshgrid(lin, lin, indexing="ij")
    ref = torch.stack((rx.reshape(-1), ry.reshape(-1)), -1)[None, :, None].expand(B, S, 1, 2)
    x = src
    with torch.no_grad():
        for i in range(cfg.enc_layers):
            layer = M.DeformableDetrEncoderLayer(dc).eval()
            _load(layer, sd, f"{dt}encoder.layers.{i}.", [("fc1.", "mlp.fc1."), ("fc2.", "mlp.fc2.")])
            x = layer(x, attention_mask=None, spatial_position_embeddings=pos, reference_points=ref,
                      spatial_shapes=torch.tensor([[g, g]]), spatial_shapes_list=[(g, g)], level_start_index=torch.tensor([0]))
            x = x[0] if isinstance(x, tuple) else x
    close(memory, x)


def test_ddetr_decoder_layers_match_hf(setup):
    """tgt --(6 x DeformableDetrDecoderLayer, 4-d reference boxes that never advance: ddetr_transformer.py:77-202, T4)--> last hidden."""
    cfg, sd, o = setup
    M, dc = _ddetr_modules(cfg)
    g0 = torch.Generator().manual_seed(3)
    hs = o.vit(torch.randn(2, 3, 518, 518, generator=g0))
    o.proposer(hs)
    memory, ref, want = o.stages["memory"], o.stages["ref_init"], o.stages["dec_last"]
    query_pos, tgt = o.stages["query_pos"], o.stages["tgt"]
    g, B = cfg.grid, memory.shape[0]
    dt = "perceiver.ddetr_transformer."
    h = tgt
    with torch.no_grad():
        for i in range(cfg.dec_layers):
            layer = M.DeformableDetrDecoderLayer(dc).eval()
            _load(layer, sd, f"{dt}decoder.layers.{i}.", [("fc1.", "mlp.fc1."), ("fc2.", "mlp.fc2."), ("self_attn.out_proj", "self_attn.o_proj")])
            h = layer(h, object_queries_position_embeddings=query_pos, reference_points=ref[:, :, None, :],
                      spatial_shapes=torch.tensor([[g, g]]), spatial_shapes_list=[(g, g)], level_start_index=torch.tensor([0]),
                      encoder_hidden_states=memory, encoder_attention_mask=None)
            h = h[0] if isinstance(h, tuple) else h
    close(want, h)


def test_two_stage_proposal_helpers_match_hf(setup):
    """gen_encoder_output_proposals / get_proposal_pos_embed (reference copies: ddetr_transformer.py:383-446) against the
    transformers implementations, called on a stub that carries the same enc_output / enc_output_norm parameters."""
    from types import SimpleNamespace
    cfg, sd, o = setup
    M, dc = _ddetr_modules(cfg)
    g0 = torch.Generator().manual_seed(4)
    o.proposer(o.vit(torch.randn(2, 3, 518, 518, generator=g0)))
    st = o.stages
    memory = st["memory"]
    B, S, D = memory.shape
    g = cfg.grid
    dt = "perceiver.ddetr_transformer."
    lin = torch.nn.Linear(D, D)
    ln = torch.nn.LayerNorm(D, eps=1e-5)
    lin.load_state_dict({"weight": sd[dt + "enc_output.weight"].float(), "bias": sd[dt + "enc_output.bias"].float()})
    ln.load_state_dict({"weight": sd[dt + "enc_output_norm.weight"].float(), "bias": sd[dt + "enc_output_norm.bias"].float()})
    stub = SimpleNamespace(enc_output=lin, enc_output_norm=ln, config=dc)
    with torch.no_grad():
        oq, props = M.DeformableDetrModel.gen_encoder_output_proposals(stub, memory, torch.zeros(B, S, dtype=torch.bool), torch.tensor([[g, g]]))
        pos = M.DeformableDetrModel.get_proposal_pos_embed(stub, st["topk_coord_logits"])
    close(st["enc_obj_query"], oq)
    finite = torch.isfinite(props[0])
    assert torch.equal(finite, torch.isfinite(st["prop_logit"]))                  # same (0.01, 0.99) validity mask
    close(st["prop_logit"][finite], props[0][finite], tol=1e-6)
    close(st["pos512"], pos, tol=1e-5)
