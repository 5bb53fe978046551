"""Checkpoint ingestion (groma_b200.checkpoint): HF directory layouts, laziness, error behaviour.  CPU only."""
import json
import os

import pytest
import torch

from groma_b200.checkpoint import BIN_INDEX, SAFE_INDEX, ShardedStateDict, save_sharded


def _sd():
    g = torch.Generator().manual_seed(0)
    return {f"llm.model.layers.{i}.w": torch.randn(64, 32, generator=g).to(torch.bfloat16 if i % 2 else torch.float32) for i in range(6)}


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
@pytest.mark.parametrize("max_bytes", [1 << 30, 9000])
def test_roundtrip_single_and_sharded(tmp_path, fmt, max_bytes):
    sd = _sd()
    names = save_sharded(sd, str(tmp_path), max_shard_bytes=max_bytes, fmt=fmt)
    sharded = max_bytes < (1 << 20)
    assert (len(names) > 1) == sharded
    assert os.path.exists(tmp_path / (SAFE_INDEX if fmt == "safetensors" else BIN_INDEX)) == sharded
    if sharded:
        idx = json.load(open(tmp_path / (SAFE_INDEX if fmt == "safetensors" else BIN_INDEX)))
        assert set(idx["weight_map"]) == set(sd) and idx["metadata"]["total_size"] == sum(t.numel() * t.element_size() for t in sd.values())
        assert all(n.startswith("model-0000" if fmt == "safetensors" else "pytorch_model-0000") for n in names)
    view = ShardedStateDict(str(tmp_path))
    assert set(view) == set(sd) and len(view) == len(sd) and sorted(view.shards()) == sorted(names)
    for k, t in sd.items():
        assert k in view
        got = view[k]
        assert got.dtype == t.dtype and torch.equal(got, t)
    assert "nope" not in view
    with pytest.raises(KeyError):
        view["nope"]


def test_shards_open_lazily(tmp_path):
    sd = _sd()
    save_sharded(sd, str(tmp_path), max_shard_bytes=9000, fmt="safetensors")
    view = ShardedStateDict(str(tmp_path))
    assert view._safe_handles == {}                      # index json only: no shard touched yet
    view["llm.model.layers.0.w"]
    assert len(view._safe_handles) == 1


def test_missing_shard_and_empty_dir_fail_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        ShardedStateDict(str(tmp_path))
    names = save_sharded(_sd(), str(tmp_path), max_shard_bytes=9000, fmt="safetensors")
    os.remove(tmp_path / names[-1])
    with pytest.raises(FileNotFoundError):
        ShardedStateDict(str(tmp_path))
