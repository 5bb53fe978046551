"""GPU-vs-oracle parity at Groma-7B dimensions (BASELINE.json configs[0]: one 448x448 image + 32-token prompt, random-init
DINOv2-L + 6+6-layer proposer + 3-level region encoder + Vicuna-7B-shaped LLaMA).  This is the configuration the benchmark
executes: cta_group::2 GEMMs (C >= 512, llm_hidden >= 2048), the 27-tap RoI conv at C = 1024, flatten K = 200 704 split-K,
8-head MSDA, GroupNorm(64), the real decode split-K factors and the CUDA-graph decode step -- none of which the miniature
pipeline tests reach.  Stage by stage (tests/fullsize.py); every distance is printed."""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fullsize_stage_parity_vs_oracle():
    from groma.model.groma import GromaConfig, GromaModel
    from groma_b200.config import PathConfig, SyntheticTokenizer
    from groma_b200.synth import make_state_dict
    from tests.fullsize import run_fullsize_check, verdict, widen_state_dict
    free, total = torch.cuda.mem_get_info()
    if total < 60 << 30:
        pytest.skip("needs a >= 60 GB GPU for the Groma-7B-shaped weights")
    cfg = PathConfig(box_score_thres=0.0)
    tok = SyntheticTokenizer(cfg.vocab)
    sd = make_state_dict(cfg, seed=0, perturb_norms=True, dtype=torch.bfloat16, device="cuda")
    model = GromaModel(GromaConfig.from_path_config(cfg), state_dict=sd, path_config=cfg)
    model.init_special_token_id(tok)
    sd_f32 = widen_state_dict(sd)
    del sd
    torch.cuda.empty_cache()
    res = run_fullsize_check(model, cfg, sd_f32, tok, n_text=32, n_new=8)
    print("[fullsize] " + json.dumps(res))
    bad = verdict(res)
    assert not bad, "\n".join(bad)
