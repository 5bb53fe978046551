"""Drop-in boundary on the GPU: the call sequences the reference's callers make, executed
against this package on a miniature checkpoint directory, and the results that have an oracle compared with it.

  * eval/run_groma.py:36-116  -- from_pretrained(dir).cuda(), init_special_token_id, generate(..., generation_config=
                                model.generation_config) under inference_mode + autocast; stops at the checkpoint's EOS and pads
  * serve/model_worker.py:288-304, serve/cli.py:31-44 -- step-wise decode through forward(past_key_values=...)
  * groma.py:404-415          -- forward(labels=...) loss
  * ddetr.py:169-196 + train/train_det.py:97-131 -- detector-only checkpoint, forward + post_process vs the oracle
  * $HF modeling_llama.py rotary cache extension -- sequences longer than max_position_embeddings
"""
import json
import os
import re

import pytest
import torch

pytestmark = pytest.mark.gpu

from groma_b200.config import SyntheticTokenizer, tiny_config  # noqa: E402
from groma_b200.synth import make_state_dict  # noqa: E402
from oracle.groma_oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def nrel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


class WordTokenizer(SyntheticTokenizer):
    """Synthetic stand-in for the LLaMA sentencepiece tokenizer with the call shape run_groma.py uses: special tokens map to
    their ids, every other word hashes into the base vocabulary."""

    def __call__(self, prompts):
        from types import SimpleNamespace
        rows = []
        for p in prompts:
            ids = [1]
            for piece in re.findall(r"<[a-z_]+\d*>|\[[A-Za-z]+\]|[A-Za-z0-9']+|[^\sA-Za-z0-9]", p):
                ids.append(self.map[piece] if piece in self.map else 10 + (sum(ord(c) * (i + 7) for i, c in enumerate(piece)) % 900))
            rows.append(ids)
        return SimpleNamespace(input_ids=rows)


def _write_checkpoint(tmp_path, cfg, sd, gen_cfg=None):
    from groma.model.groma import GromaConfig
    from groma_b200.checkpoint import save_sharded
    cd = json.loads(GromaConfig.from_path_config(cfg).to_json_string())
    cd["path_overrides"] = {"gn_groups": cfg.gn_groups}
    (tmp_path / "config.json").write_text(json.dumps(cd))
    if gen_cfg is not None:
        (tmp_path / "generation_config.json").write_text(json.dumps(gen_cfg))
    save_sharded(sd, str(tmp_path), max_shard_bytes=1 << 22, fmt="safetensors")


def _run_groma_sequence(model_dir, tok, image_u8, max_new_tokens, loader):
    """The body of eval/run_groma.py:eval_model between loading and decoding, statement for statement."""
    from groma.constants import DEFAULT_TOKENS
    from groma.data.conversation import conv_templates
    from groma.utils import disable_torch_init
    from groma_b200.preprocess import GromaImageProcessor
    disable_torch_init()
    vis_processor = GromaImageProcessor()
    model = loader(model_dir).cuda()
    model.init_special_token_id(tok)
    model.config.box_score_thres = 0.0           # random-init scores sit near sigmoid(-4.6): keep some regions
    conversations = []
    instruct = "Here is an image with region crops from it. "
    instruct += "Image: {}. ".format(DEFAULT_TOKENS['image'])
    instruct += "Regions: {}.".format(DEFAULT_TOKENS['region'])
    answer = 'Thank you for the image! How can I assist you with it?'
    conversations.append((conv_templates['llava'].roles[0], instruct))
    conversations.append((conv_templates['llava'].roles[1], answer))
    conversations.append((conv_templates['llava'].roles[0], "Describe the image in detail."))
    conversations.append((conv_templates['llava'].roles[1], ''))
    prompt = conv_templates['llava'].get_prompt(conversations)
    inputs = tok([prompt])
    input_ids = torch.as_tensor(inputs.input_ids).cuda()
    image = vis_processor.preprocess(image_u8, return_tensors='pt')['pixel_values'].to('cuda')
    torch.manual_seed(0)
    with torch.inference_mode():
        with torch.autocast(device_type="cuda"):
            outputs = model.generate(input_ids, images=image, use_cache=True, do_sample=False, max_new_tokens=max_new_tokens,
                                     return_dict_in_generate=True, output_hidden_states=True,
                                     generation_config=model.generation_config)
    output_ids = outputs.sequences
    input_token_len = input_ids.shape[1]
    pred_boxes = outputs.hidden_states[0][-1]['pred_boxes'][0].cpu()
    box_idx_token_ids = model.box_idx_token_ids
    selected_box_inds = [box_idx_token_ids.index(id) for id in output_ids[0] if id in box_idx_token_ids]
    selected_box_inds = [x for x in selected_box_inds if x < len(pred_boxes)]
    n_diff_input_output = (input_ids != output_ids[:, :input_token_len]).sum().item()
    return model, input_ids, output_ids, pred_boxes, selected_box_inds, n_diff_input_output


@pytest.mark.parametrize("loader_name", ["GromaModel", "AutoModel"])
def test_run_groma_call_sequence_stops_at_eos(tmp_path, loader_name):
    from transformers import AutoModel
    from groma.model.groma import GromaModel
    loader = GromaModel.from_pretrained if loader_name == "GromaModel" else AutoModel.from_pretrained
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    tok = WordTokenizer(cfg.vocab)
    g = torch.Generator().manual_seed(3)
    image_u8 = torch.randint(0, 256, (300, 400, 3), dtype=torch.uint8, generator=g)
    # pass 1: no EOS in the vocabulary that the model could emit (id outside both heads' range is impossible -> use none)
    _write_checkpoint(tmp_path, cfg, sd, gen_cfg={"eos_token_id": None, "pad_token_id": tok.pad_token_id})
    model, input_ids, out_free, boxes, _, ndiff = _run_groma_sequence(str(tmp_path), tok, image_u8, 24, loader)
    assert isinstance(model, GromaModel) and ndiff == 0
    n_in = input_ids.shape[1]
    free = out_free[0, n_in:].tolist()
    assert len(free) == 24 and len(boxes) >= 1
    # pass 2: the checkpoint's generation_config.json names the token produced at step 5 as EOS -> generation must end there
    first = {}
    for i, t in enumerate(free):
        first.setdefault(t, i)
    stop_tok = free[5]
    stop_at = first[stop_tok]
    (tmp_path / "generation_config.json").write_text(json.dumps({"eos_token_id": [stop_tok, 999999], "pad_token_id": tok.pad_token_id,
                                                                 "bos_token_id": 1, "do_sample": True}))
    model2, _, out_eos, boxes2, sel, ndiff2 = _run_groma_sequence(str(tmp_path), tok, image_u8, 1024, loader)
    assert model2.generation_config.eos_token_id == [stop_tok, 999999]
    got = out_eos[0, n_in:].tolist()
    assert got == free[:stop_at + 1], (got, free)            # identical prefix, ends WITH the EOS token (HF greedy_search)
    assert ndiff2 == 0 and torch.equal(boxes2, boxes)
    assert all(0 <= s < len(boxes2) for s in sel)


def test_generate_pads_finished_rows_like_hf_greedy_search():
    from groma.model.groma import GromaConfig, GromaModel
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    tok = SyntheticTokenizer(cfg.vocab)
    model = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg)
    model.init_special_token_id(tok)
    g = torch.Generator().manual_seed(4)
    images = torch.randn(2, 3, 448, 448, generator=g).cuda()
    ids = torch.randint(10, cfg.vocab, (2, 20), generator=g)
    ids[:, 2], ids[:, 7] = tok.map["<image>"], tok.map["<region>"]
    boxes = [torch.rand(4, 4, generator=g) * 0.5 + 0.2, torch.rand(6, 4, generator=g) * 0.5 + 0.2]
    free = model.generate(ids.clone().cuda(), images=images, max_new_tokens=40, _selected_override=boxes)[:, 20:].cpu()
    # an EOS id that row 0 emits early and row 1 later (or never): row 0 must be padded after it, length = last finisher
    r0, r1 = free[0].tolist(), free[1].tolist()
    eos = r0[3]
    e0 = r0.index(eos)
    e1 = r1.index(eos) if eos in r1 else None
    out = model.generate(ids.clone().cuda(), images=images, max_new_tokens=40, eos_token_id=eos, _selected_override=boxes)[:, 20:].cpu()
    want_len = 40 if e1 is None else max(e0, e1) + 1
    assert out.shape[1] == want_len
    assert out[0, :e0 + 1].tolist() == r0[:e0 + 1] and (out[0, e0 + 1:] == tok.pad_token_id).all()
    n1 = want_len if e1 is None else e1 + 1
    assert out[1, :n1].tolist() == r1[:n1] and (out[1, n1:] == tok.pad_token_id).all()


def test_stepwise_forward_loop_equals_generate_and_grows_the_cache():
    """serve/model_worker.py:288-304: prefill forward(use_cache=True), then forward(last token, past_key_values) per step."""
    from groma.model.groma import GromaConfig, GromaModel
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    tok = SyntheticTokenizer(cfg.vocab)
    model = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg)
    model.init_special_token_id(tok)
    model.kv_headroom = 3                          # force two cache re-homings within 12 steps
    g = torch.Generator().manual_seed(5)
    images = torch.randn(2, 3, 448, 448, generator=g).cuda()
    ids = torch.randint(10, cfg.vocab, (2, 18), generator=g)
    ids[:, 1], ids[:, 6] = tok.map["<image>"], tok.map["<region>"]
    boxes = [torch.rand(3, 4, generator=g) * 0.5 + 0.2, torch.rand(3, 4, generator=g) * 0.5 + 0.2]
    n_new = 12
    want = model.generate(ids.clone().cuda(), images=images, max_new_tokens=n_new, _selected_override=boxes)[:, 18:].cpu()
    out = model.forward(input_ids=ids.clone().cuda(), images=images, use_cache=True, return_dict=True, _selected_override=boxes)
    cap0 = model.engine.kv_cap
    toks = [out.logits[:, -1].argmax(-1)]
    pkv = out.past_key_values
    for _ in range(n_new - 1):
        out = model.forward(input_ids=toks[-1][:, None], past_key_values=pkv, use_cache=True, return_dict=True)
        pkv = out.past_key_values
        toks.append(out.logits[:, -1].argmax(-1))
    assert model.engine.kv_cap > cap0              # the cache was re-homed, contents kept
    assert torch.equal(torch.stack(toks, 1).cpu(), want)
    with pytest.raises(RuntimeError, match="past_key_values"):      # a cache that is not this model's current one
        stale = tuple((k[:, :, :-2], v[:, :, :-2]) for k, v in pkv)
        model.forward(input_ids=toks[-1][:, None], past_key_values=stale, use_cache=True, return_dict=True)


def test_forward_with_labels_matches_the_oracle_loss():
    """groma.py:338-353 (label expansion under the placeholders), :404-415 (shifted cross-entropy), :305-309 (ground-box ids are
    written into the labels too)."""
    from groma.model.groma import GromaConfig, GromaModel
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    tok = SyntheticTokenizer(cfg.vocab)
    model = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg)
    model.init_special_token_id(tok)
    o = Oracle(cfg, sd, "bf16")
    o.init_special_token_id(tok)
    g = torch.Generator().manual_seed(6)
    images = torch.randn(2, 3, 448, 448, generator=g)
    ids = torch.randint(10, cfg.vocab, (2, 22), generator=g)
    ids[:, 2], ids[:, 8] = tok.map["<image>"], tok.map["<region>"]
    ids[0, 12] = tok.map["<ground_box>"]
    ids[1, 17:] = tok.pad_token_id
    labels = ids.clone()
    labels[:, :10] = -100
    labels[1, 17:] = -100
    ground = [torch.tensor([[0.4, 0.5, 0.2, 0.3]]), torch.zeros(0, 4)]
    boxes = [torch.rand(5, 4, generator=g) * 0.5 + 0.2, torch.rand(2, 4, generator=g) * 0.5 + 0.2]
    ids_o, lab_o = ids.clone(), labels.clone()
    want = o.forward_prefill(ids_o, images, None, ground, selected_override=boxes, labels=lab_o)
    ids_g, lab_g = ids.clone().cuda(), labels.clone().cuda()
    got = model.forward(input_ids=ids_g, labels=lab_g, images=images.cuda(), ground_boxes=ground, return_dict=True, _selected_override=boxes)
    assert torch.equal(ids_g.cpu(), ids_o) and torch.equal(lab_g.cpu(), lab_o)      # both edited in place identically
    assert not torch.equal(lab_o, labels)
    loss_g, loss_o = float(got.loss), float(want["loss"])
    print(f"loss gpu {loss_g:.6f} oracle {loss_o:.6f}")
    assert abs(loss_g - loss_o) <= 2e-3 * abs(loss_o)
    tup = model.forward(input_ids=ids.clone().cuda(), labels=labels.clone().cuda(), images=images.cuda(), ground_boxes=ground,
                        _selected_override=boxes)
    assert len(tup) == 3 and abs(float(tup[0]) - loss_g) < 1e-6                       # return_dict=False: (loss, logits, pkv)


def test_detector_checkpoint_forward_and_post_process_vs_oracle(tmp_path):
    """eval/run_ddetr.py:41-50 on a train_det.py-layout checkpoint, then train_det.py:97-131 post_process."""
    from groma.model.ddetr import CustomDDETRModel
    from groma.model.groma import GromaConfig
    from groma.train.train_det import post_process
    from groma_b200.checkpoint import save_sharded
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    det_sd = {k[len("perceiver."):]: v for k, v in sd.items() if k.startswith("perceiver.")}
    (tmp_path / "config.json").write_text(GromaConfig.from_path_config(cfg).perceiver_cfg.to_json_string())
    save_sharded(det_sd, str(tmp_path), max_shard_bytes=1 << 22, fmt="safetensors")
    det = CustomDDETRModel.from_pretrained(str(tmp_path)).cuda()
    g = torch.Generator().manual_seed(8)
    images = torch.randn(2, 3, 448, 448, generator=g)
    o = Oracle(cfg, sd, "bf16")
    pred_o, score_o, logits_o = o.proposer(o.vit(images))
    Q = cfg.num_queries
    det.engine.keep_stages = True
    out_own = det(images.cuda())                                   # the detector's own two-stage selection
    own = det.engine.stages["topk_own"].cpu()
    cls_g = det.engine.stages["enc_cls"].cpu()
    for b in range(2):                                             # its top-k is exactly the stable descending order of ITS scores
        assert own[b].tolist() == torch.argsort(-cls_g[b], stable=True)[:Q].tolist()
    e_cls = nrel(cls_g, o.stages["enc_cls"])
    det.engine.topk_override = o.stages["topk"]                    # same queries as the oracle -> comparable slot by slot
    out = det(images.cuda())
    det.engine.topk_override = None
    assert out.pred_boxes.shape == (2, Q, 4) and out.logits["coco"].shape == (2, Q, 1) and out.logits["sa1b"].shape == (2, Q, 1)
    e_coco, e_sa1b = nrel(out.logits["coco"][..., 0], logits_o["coco"]), nrel(out.logits["sa1b"][..., 0], logits_o["sa1b"])
    d_box = (out.pred_boxes.cpu() - pred_o).abs().max().item()
    common = sum(len(set(own[b].tolist()) & set(o.stages["topk"][b].tolist())) for b in range(2)) / (2 * Q)
    print(f"detector vs oracle: objectness nrel {e_cls:.2e}, own top-k overlap {common:.3f}; teacher-forced queries: coco nrel {e_coco:.2e}, "
          f"sa1b nrel {e_sa1b:.2e}, boxes max abs {d_box:.2e}")
    assert e_cls < 1.5e-2 and common > 0.9
    assert e_coco < 1.5e-2 and e_sa1b < 1.5e-2 and d_box < 5e-3     # bf16-stored decoder states; boxes are cxcywh in (0, 1)
    out = out_own
    # post_process on the detector's own output vs the oracle restatement on the same tensors: exact indices, fp32 boxes
    sizes = torch.tensor([[480.0, 640.0], [333.0, 500.0]])
    got = post_process(out, sizes.cuda(), threshold=0.0, top_k=50)
    want = Oracle.post_process(out.logits["coco"].cpu(), out.pred_boxes.cpu(), sizes, 0.0, 50)
    for a, b in zip(got, want):
        assert torch.equal(a["labels"].cpu(), b["labels"])
        assert torch.allclose(a["scores"].cpu(), b["scores"], rtol=0, atol=1e-6)     # device vs host sigmoid: last ulp
        assert torch.allclose(a["boxes"].cpu(), b["boxes"], rtol=0, atol=1e-4)
    # and on the committed reference-run fixture
    from types import SimpleNamespace
    fx = torch.load(os.path.join(HERE, "golden", "post_process_ref.pt"))
    for c, ref in zip(fx["cases"], fx["results"]):
        res = post_process(SimpleNamespace(logits={"coco": c["coco"].cuda()}, pred_boxes=c["boxes"].cuda()), c["sizes"].cuda(),
                           threshold=c["threshold"], top_k=c["top_k"])
        for a, b in zip(res, ref):
            assert torch.equal(a["labels"].cpu(), b["labels"])
            assert torch.allclose(a["scores"].cpu(), b["scores"], rtol=0, atol=1e-6)
            assert torch.allclose(a["boxes"].cpu(), b["boxes"], rtol=0, atol=1e-4)


def test_sequences_longer_than_max_position_embeddings_extend_the_rope_tables():
    from groma.model.groma import GromaConfig, GromaModel
    cfg = tiny_config(box_score_thres=0.0, max_pos=280)            # prefill is 256 + 2R + text > 280
    sd = make_state_dict(cfg, seed=0)
    tok = SyntheticTokenizer(cfg.vocab)
    model = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg)
    model.init_special_token_id(tok)
    o = Oracle(cfg, sd, "bf16")
    o.init_special_token_id(tok)
    g = torch.Generator().manual_seed(9)
    images = torch.randn(1, 3, 448, 448, generator=g)
    ids = torch.randint(10, cfg.vocab, (1, 30), generator=g)
    ids[:, 2], ids[:, 8] = tok.map["<image>"], tok.map["<region>"]
    boxes = [torch.rand(4, 4, generator=g) * 0.5 + 0.2]
    assert model.engine.rope_len == 280
    gen = model.generate(ids.clone().cuda(), images=images.cuda(), max_new_tokens=4, _selected_override=boxes, _keep_logits=True)
    assert model.engine.rope_len >= 256 + 8 + 28 + 4
    want = o.generate(ids.clone(), images, 4, selected_override=boxes)
    e = nrel(model._step_logits[0], want["step_logits"][:, 0])
    print(f"long-position first-step logits nrel {e:.2e}; tokens {gen[0, 30:].tolist()} oracle {want['new_tokens'][0].tolist()}")
    assert e < 1e-2
