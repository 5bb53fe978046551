"""Host-side config plumbing (no GPU): GromaConfig <-> config.json <-> PathConfig, detector sub-config, decode split-K picks."""
import json

import pytest

from groma.model.ddetr import CustomDDETRConfig, perceiver_fields
from groma.model.groma import GromaConfig
from groma_b200.config import PathConfig, tiny_config


def test_groma_config_json_round_trip_reproduces_the_path_config():
    for cfg in (tiny_config(box_score_thres=0.0), PathConfig()):
        h = GromaConfig.from_path_config(cfg)
        cd = json.loads(h.to_json_string())
        assert cd["model_type"] == "groma"
        cd.pop("model_type")
        back = GromaConfig(**cd).to_path_config(fuse_rounds=cfg.fuse_rounds, gn_groups=cfg.gn_groups, pos_hidden=cfg.pos_hidden,
                                                region_mid=cfg.region_mid)
        assert back == cfg


def test_groma_7b_defaults_match_the_survey_dimensions():
    p = PathConfig()
    assert (p.vit_hidden, p.vit_layers, p.vit_heads, p.patch) == (1024, 24, 16, 14)
    assert (p.d_model, p.enc_layers, p.dec_layers, p.num_queries, p.n_points) == (256, 6, 6, 300, 4)
    assert (p.llm_hidden, p.llm_layers, p.llm_heads, p.llm_inter, p.vocab, p.num_new_token) == (4096, 32, 32, 11008, 32000, 114)
    assert (p.nms_thres, p.max_region_num) == (0.6, 100)


def test_detector_subconfig_fields_and_rejections():
    cfg = tiny_config()
    h = GromaConfig.from_path_config(cfg)
    pd = json.loads(h.perceiver_cfg.to_json_string())
    pd.pop("model_type", None)
    pc = CustomDDETRConfig(**pd)
    f = perceiver_fields(pc)
    assert all(getattr(cfg, k) == v for k, v in f.items())
    pc.ddetr_cfg.num_feature_levels = 4            # the multi-level proposer is not on the path: must be refused loudly
    with pytest.raises(NotImplementedError):
        perceiver_fields(pc)


def test_mutable_thresholds_flow_into_the_path_config():
    h = GromaConfig.from_path_config(tiny_config())
    h.box_score_thres, h.nms_thres, h.max_region_num = 0.3, 0.5, 7     # run_groma.py mutates these on model.config
    p = h.to_path_config()
    assert (p.box_score_thres, p.nms_thres, p.max_region_num) == (0.3, 0.5, 7)
