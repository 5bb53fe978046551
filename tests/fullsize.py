"""Stage-by-stage comparison of the B200 path with the CPU oracle at Groma-7B dimensions (BASELINE.json configs[0]: one image,
32-token prompt; bench.py --check runs it on a 512-token prompt).  TEST INFRASTRUCTURE: imported by
tests/test_fullsize_gpu.py and by bench.py's --check leg, never by the product.

One synthetic state dict (bf16 values) is loaded into the GPU model, widened to fp32 on the host and shared by two oracles:
  'bf16' -- rounds activations where the GPU pipeline stores them (DESIGN.md "rounding points"): the parity target;
  'fp32' -- same weights, no activation rounding: the reference's fp32 arithmetic, used to measure how far bf16 storage
            alone moves every stage (the noise floor the 1e-3 logits tolerance of BASELINE.json has to be read against).
Where an integer decision sits between two stages (two-stage top-k, NMS keep set, greedy tokens) the downstream comparison is
teacher-forced with the oracle's decision so that every stage is asserted unconditionally; the decision itself is checked
bit-exactly on the GPU's OWN inputs (its scores -> its top-k, its boxes/scores -> its NMS keep list and shuffled boxes).
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch

from oracle import ops as O
from oracle.groma_oracle import Oracle


def nrel(a, b) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


def rmsrel(a, b) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-9)).item()


def widen_state_dict(sd) -> Dict[str, torch.Tensor]:
    """bf16 (device or host) state dict -> fp32 host tensors with the same (bf16-representable) values."""
    return {k: v.detach().to("cpu").float() for k, v in sd.items()}


def make_inputs(cfg, tok, n_text: int, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(1, 3, cfg.image_size, cfg.image_size, generator=g)
    ids = torch.randint(1000 if cfg.vocab > 2000 else 10, cfg.vocab, (1, n_text), generator=g)
    ids[:, 4] = tok.map["<image>"]
    ids[:, n_text // 2] = tok.map["<region>"]
    return images, ids


def run_fullsize_check(model, cfg, sd_f32: Dict[str, torch.Tensor], tok, n_text: int = 32, n_new: int = 8, seed: int = 7,
                       fp32_floor: bool = True, log=print) -> dict:
    """Returns a flat dict of distances / flags (see keys below); raises nothing on mismatch -- callers assert."""
    eng = model.engine
    res: dict = {"n_text": n_text, "n_new": n_new}
    images, ids = make_inputs(cfg, tok, n_text, seed)
    t0 = time.time()
    ob = Oracle(cfg, sd_f32, "bf16", prerounded=True)
    ob.init_special_token_id(tok)
    torch.manual_seed(1234)
    want = ob.generate(ids.clone(), images, n_new)
    res["oracle_bf16_s"] = time.time() - t0
    sel_o = want["selected_boxes"]
    R = len(sel_o[0])
    res["regions"] = R
    res["prefill_tokens"] = int(want["input_ids"].shape[1])
    log(f"[fullsize] bf16 oracle: {res['oracle_bf16_s']:.1f}s, R={R}, T={res['prefill_tokens']}")
    hs_o = ob.stages["vit_hs"]

    # ---------------- vision stages on the GPU's own data flow
    eng.keep_stages = True
    eng.topk_override = None
    hs_g = eng.vit(images.cuda())
    for k in range(1, 5):
        res[f"vit_hidden[-{k}]_nrel"] = nrel(hs_g[-k], hs_o[-k])
        res[f"vit_hidden[-{k}]_rms"] = rmsrel(hs_g[-k], hs_o[-k])
    res["image_tokens_nrel"] = nrel(eng.image_tokens(hs_g[-1]), want["image_features"])
    pc, px, sc, _ = eng.proposer(hs_g)
    st = eng.stages
    res["ddetr_src_nrel"] = nrel(st["ddetr_src"], ob.stages["ddetr_src"])
    res["memory_nrel"] = nrel(st["memory"], ob.stages["memory"])
    res["memory_rms"] = rmsrel(st["memory"], ob.stages["memory"])
    cls_g = st["enc_cls"].cpu()
    res["enc_cls_max_abs"] = (cls_g - ob.stages["enc_cls"]).abs().max().item()
    Q = cfg.num_queries
    own = st["topk_own"].cpu()
    res["topk_is_stable_argsort_of_own_scores"] = bool(own[0].tolist() == np.argsort(-cls_g[0].numpy(), kind="stable")[:Q].tolist())
    res["topk_overlap_with_oracle"] = len(set(own[0].tolist()) & set(ob.stages["topk"][0].tolist())) / Q
    # NMS + shuffle on the model's OWN proposals: bit-exact vs the oracle's NMS restatement fed the same fp32 boxes/scores
    torch.manual_seed(99)
    sel_g = eng.select_regions(pc.clone(), px.clone(), sc.clone(), None, None, cfg.nms_thres, cfg.box_score_thres, cfg.max_region_num)
    torch.manual_seed(99)
    sel_ref, inds_ref = ob.select_regions(pc.cpu()[:, :Q], sc.cpu()[:, :Q])
    keep, num = eng.stages["nms_keep"], eng.stages["nms_num"]
    res["nms_keep_exact_on_own_proposals"] = bool(int(num[0]) == len(inds_ref[0]) and keep[0, :len(inds_ref[0])].tolist() == inds_ref[0].tolist())
    res["selected_boxes_exact_on_own_proposals"] = bool(torch.equal(sel_g[0], sel_ref[0]))
    res["own_regions"] = int(num[0])
    # teacher-forced top-k: decoder / heads comparable query by query
    eng.topk_override = ob.stages["topk"]
    pc2, px2, sc2, lg2 = eng.proposer(hs_g)
    eng.topk_override = None
    res["ref_init_max_abs"] = (eng.stages["ref_init"].cpu() - ob.stages["ref_init"]).abs().max().item()
    res["dec_last_nrel"] = nrel(eng.stages["dec_last"], ob.stages["dec_last"])
    res["pred_boxes_max_abs"] = (pc2.cpu()[:, :Q] - want["pred_boxes"]).abs().max().item()
    res["scores_max_abs"] = (sc2.cpu()[:, :Q] - want["scores"]).abs().max().item()
    # NMS on the teacher-forced proposals vs the oracle's keep list on ITS proposals (equal unless a score gap / IoU sits inside bf16 noise)
    torch.manual_seed(1234)
    sel_tf = eng.select_regions(pc2.clone(), px2.clone(), sc2.clone(), None, None, cfg.nms_thres, cfg.box_score_thres, cfg.max_region_num)
    n_tf = int(eng.stages["nms_num"][0])
    res["nms_keep_equal_to_oracle_after_teacher_forced_topk"] = bool(n_tf == len(want["nms_inds"][0]) and
                                                                     eng.stages["nms_keep"][0, :n_tf].tolist() == want["nms_inds"][0].tolist())
    # ---------------- region encoder on the oracle's boxes
    reg_g = eng.region_encoder(hs_g, sel_o)
    for l in range(3):
        res[f"fused_map{l}_nrel"] = nrel(eng.stages["fused_maps"][l], ob.stages["fused_maps"][l])
    res["roi_fused_nrel"] = nrel(eng.stages["roi_fused"], ob.stages["roi_fused"].permute(0, 2, 3, 1))
    res["region_flat_nrel"] = nrel(eng.stages["region_flat"], ob.stages["region_flat"])
    res["region_features_nrel"] = nrel(reg_g, want["region_features"])
    res["region_features_rms"] = rmsrel(reg_g, want["region_features"])
    eng.keep_stages = False
    eng.stages = {}
    # ---------------- full forward with the oracle's regions: assembled ids, logits at every position, KV
    out = model.forward(input_ids=ids.clone().cuda(), images=images.cuda(), use_cache=True, return_dict=True, _selected_override=sel_o)
    res["assembled_ids_exact"] = bool(torch.equal(model._last["ids"], want["input_ids"]))
    lg = out.logits.float().cpu()
    res["logits_nrel_all_positions"] = nrel(lg, want["logits"])
    res["logits_nrel_last_position"] = nrel(lg[:, -1], want["logits"][:, -1])
    res["logits_rms_vs_bf16_oracle"] = rmsrel(lg, want["logits"])
    res["kv_last_layer_k_nrel"] = nrel(out.past_key_values[-1][0].permute(0, 2, 1, 3), want["kv"][-1][0])
    del out
    # ---------------- greedy tokens (eager decode is covered by the miniature tests; this is the CUDA-graph step at real shapes)
    gen = model.generate(ids.clone().cuda(), images=images.cuda(), max_new_tokens=n_new, _selected_override=sel_o, _keep_logits=True)
    new = gen[0, n_text:].cpu()
    wt, sl = want["new_tokens"][0], want["step_logits"][0]
    res["tokens_gpu"], res["tokens_oracle"] = new.tolist(), wt.tolist()
    first_diff, margin_rel = None, None
    for t in range(n_new):
        if int(new[t]) != int(wt[t]):
            top2 = sl[t].topk(2).values
            first_diff, margin_rel = t, ((top2[0] - top2[1]) / sl[t].abs().max()).item()
            break
    res["tokens_first_divergence"] = first_diff
    res["tokens_divergence_oracle_margin_rel"] = margin_rel
    res["tokens_equal"] = first_diff is None
    stepl = torch.stack([x.float().cpu()[0] for x in model._step_logits])
    n_cmp = n_new if first_diff is None else first_diff + 1
    res["decode_step_logits_nrel"] = nrel(stepl[:n_cmp], sl[:n_cmp])
    # ---------------- the bf16-storage noise floor at this size
    if fp32_floor:
        t0 = time.time()
        of = Oracle(cfg, sd_f32, "fp32", prerounded=True)
        of.init_special_token_id(tok)
        torch.manual_seed(1234)
        wf = of.forward_prefill(ids.clone(), images, selected_override=sel_o)
        res["oracle_fp32_s"] = time.time() - t0
        res["floor_logits_nrel_bf16_vs_fp32_oracle"] = nrel(want["logits"], wf["logits"])
        res["floor_logits_rms_bf16_vs_fp32_oracle"] = rmsrel(want["logits"], wf["logits"])
        res["logits_nrel_gpu_vs_fp32_oracle"] = nrel(lg, wf["logits"])
        res["logits_rms_gpu_vs_fp32_oracle"] = rmsrel(lg, wf["logits"])
        res["floor_vit_hidden[-1]_nrel"] = nrel(hs_o[-1], of.stages["vit_last"])
        res["floor_memory_nrel"] = nrel(ob.stages["memory"], of.stages["memory"])
        res["floor_region_features_nrel"] = nrel(want["region_features"], wf["region_features"])
        log(f"[fullsize] fp32 oracle: {res['oracle_fp32_s']:.1f}s")
    return res


# the bars tests/test_fullsize_gpu.py asserts and bench.py --check reports against
BARS = {
    "bf16_stage_nrel": 1.5e-2,        # bf16-stored stages: one bf16 ulp at the largest magnitude is 2^-8 .. 2^-7
    "boxes_max_abs": 5e-3,            # cxcywh in (0,1) computed in fp32 from bf16 decoder states
    "logits_vs_floor": 1.25,          # rms(gpu - bf16 oracle) <= 1.25 x rms(bf16 oracle - fp32 oracle)
    "token_margin_rel": 1e-2,
}


def verdict(res: dict) -> list:
    """List of violated bars (empty = green)."""
    bad = []
    b = BARS["bf16_stage_nrel"]
    for k, v in res.items():
        if k.endswith("_nrel") and not k.startswith("floor_") and not k.startswith("logits_") and not k.startswith("decode_step") and v >= b:
            bad.append(f"{k}={v:.3e} >= {b}")
    for k in ("topk_is_stable_argsort_of_own_scores", "nms_keep_exact_on_own_proposals", "selected_boxes_exact_on_own_proposals",
              "assembled_ids_exact"):
        if not res[k]:
            bad.append(f"{k} is False")
    for k in ("pred_boxes_max_abs", "scores_max_abs", "ref_init_max_abs"):
        if res[k] >= BARS["boxes_max_abs"]:
            bad.append(f"{k}={res[k]:.3e} >= {BARS['boxes_max_abs']}")
    if res["logits_nrel_all_positions"] >= 1e-2:
        bad.append(f"logits_nrel_all_positions={res['logits_nrel_all_positions']:.3e} >= 1e-2")
    if "floor_logits_rms_bf16_vs_fp32_oracle" in res and res["logits_rms_vs_bf16_oracle"] > BARS["logits_vs_floor"] * res["floor_logits_rms_bf16_vs_fp32_oracle"]:
        bad.append(f"logits rms vs bf16 oracle {res['logits_rms_vs_bf16_oracle']:.3e} > {BARS['logits_vs_floor']} x floor "
                   f"{res['floor_logits_rms_bf16_vs_fp32_oracle']:.3e}")
    if not res["tokens_equal"] and res["tokens_divergence_oracle_margin_rel"] >= BARS["token_margin_rel"]:
        bad.append(f"greedy tokens diverge at step {res['tokens_first_divergence']} with oracle margin {res['tokens_divergence_oracle_margin_rel']:.3e}")
    return bad
