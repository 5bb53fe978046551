"""from_pretrained on HF-layout checkpoints: a sharded Groma checkpoint and a detector-only checkpoint must
produce exactly what the same weights passed in memory produce (bit-identical: same kernels, same bf16 arena)."""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu

from groma_b200.config import tiny_config, SyntheticTokenizer  # noqa: E402
from groma_b200.synth import make_state_dict  # noqa: E402


def _inputs(cfg, tok):
    g = torch.Generator().manual_seed(1)
    images = torch.randn(2, 3, 448, 448, generator=g)
    ids = torch.randint(10, cfg.vocab, (2, 20), generator=g)
    ids[:, 2] = tok.map["<image>"]
    ids[:, 7] = tok.map["<region>"]
    return images, ids


@pytest.mark.parametrize("fmt,max_bytes", [("safetensors", 1 << 20), ("bin", 1 << 20), ("safetensors", 1 << 40)])
def test_groma_from_pretrained_matches_in_memory(tmp_path, fmt, max_bytes):
    from groma.model.groma import GromaConfig, GromaModel
    from groma_b200.checkpoint import save_sharded
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    tok = SyntheticTokenizer(cfg.vocab)
    hf_cfg = GromaConfig.from_path_config(cfg)
    cd = json.loads(hf_cfg.to_json_string())
    cd["path_overrides"] = {"gn_groups": cfg.gn_groups}     # the miniature uses 8 GroupNorm groups (reference constant: 64)
    (tmp_path / "config.json").write_text(json.dumps(cd))
    names = save_sharded(sd, str(tmp_path), max_shard_bytes=max_bytes, fmt=fmt)
    assert (len(names) > 1) == (max_bytes == 1 << 20)
    ref = GromaModel(hf_cfg, state_dict=sd, path_config=cfg)
    got = GromaModel.from_pretrained(str(tmp_path)).cuda()
    assert got._path_cfg == cfg                      # config.json round trip reproduces every dimension of the path
    images, ids = _inputs(cfg, tok)
    outs = []
    for m in (ref, got):
        m.init_special_token_id(tok)
        m.config.box_score_thres = 0.0
        torch.manual_seed(0)
        o = m.generate(ids.clone().cuda(), images=images.cuda(), max_new_tokens=6, return_dict_in_generate=True, output_hidden_states=True)
        outs.append(o)
    assert torch.equal(outs[0].sequences, outs[1].sequences)
    pa, pb = outs[0].hidden_states[0][-1]["pred_boxes"], outs[1].hidden_states[0][-1]["pred_boxes"]   # per-image box lists
    assert len(pa) == len(pb) and all(torch.equal(x, y) for x, y in zip(pa, pb))


def test_detector_checkpoint_from_pretrained(tmp_path):
    from groma.model.ddetr import CustomDDETRModel
    from groma.model.groma import GromaConfig, GromaModel
    from groma_b200.checkpoint import save_sharded
    cfg = tiny_config(box_score_thres=0.0)
    sd = make_state_dict(cfg, seed=0)
    hf_cfg = GromaConfig.from_path_config(cfg)
    det_sd = {k[len("perceiver."):]: v for k, v in sd.items() if k.startswith("perceiver.")}   # train_det.py key layout
    (tmp_path / "config.json").write_text(hf_cfg.perceiver_cfg.to_json_string())
    save_sharded(det_sd, str(tmp_path), max_shard_bytes=1 << 20, fmt="safetensors")
    det = CustomDDETRModel.from_pretrained(str(tmp_path)).cuda()
    full = GromaModel(hf_cfg, state_dict=sd, path_config=cfg)
    images, _ = _inputs(cfg, SyntheticTokenizer(cfg.vocab))
    a, b = det(images.cuda()), full.perceiver(images.cuda())
    assert torch.equal(a.pred_boxes, b.pred_boxes)
    for k in ("coco", "sa1b"):
        assert torch.equal(a.logits[k], b.logits[k])
    with pytest.raises(KeyError):                     # a detector-only engine has no LLaMA weights to decode with
        det.engine.w["head.w"]
