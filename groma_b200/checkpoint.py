"""Checkpoint ingestion: the on-disk format on the input side of the path.

The reference loads weights through HF `PreTrainedModel.from_pretrained` (`groma/eval/run_groma.py:43-61`,
`groma/model/groma.py:72-108`): a directory with `config.json` and either `model.safetensors` /
`model-0000x-of-0000y.safetensors` + `model.safetensors.index.json`, or `pytorch_model.bin` /
`pytorch_model-0000x-of-0000y.bin` + `pytorch_model.bin.index.json` (`weight_map`: parameter name -> shard file).

`ShardedStateDict` is a read-only mapping over such a directory that opens a shard only when one of its tensors is asked
for, so the engine packs parameter by parameter straight into its bf16 device arena and host memory never holds more than
one tensor (safetensors) or one shard (.bin) -- not the 30 GB fp32 state dict.  `save_sharded` writes the same layout (used
by the tests and by tools that export synthetic weights).
"""
from __future__ import annotations

import glob
import json
import os
from collections.abc import Mapping
from typing import Dict, Iterator, List, Optional

import torch

SAFE_INDEX, BIN_INDEX = "model.safetensors.index.json", "pytorch_model.bin.index.json"


class ShardedStateDict(Mapping):
    def __init__(self, path: str):
        self.path = path
        self.weight_map: Dict[str, str] = {}
        self._bin_cache: Optional[tuple] = None          # (file name, state dict) of the most recently used .bin shard
        self._safe_handles: Dict[str, object] = {}
        for index in (SAFE_INDEX, BIN_INDEX):
            ip = os.path.join(path, index)
            if os.path.exists(ip):
                with open(ip) as f:
                    self.weight_map = dict(json.load(f)["weight_map"])
                break
        else:
            files = sorted(glob.glob(os.path.join(path, "*.safetensors"))) or sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
            if not files:
                raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin (or their index json) under {path}")
            for fp in files:
                for k in self._keys_of(os.path.basename(fp)):
                    self.weight_map[k] = os.path.basename(fp)
        missing = sorted({f for f in self.weight_map.values() if not os.path.exists(os.path.join(path, f))})
        if missing:
            raise FileNotFoundError(f"index lists shard files that do not exist under {path}: {missing}")

    # ---- shard access
    def _safe(self, fname: str):
        h = self._safe_handles.get(fname)
        if h is None:
            from safetensors import safe_open
            h = self._safe_handles[fname] = safe_open(os.path.join(self.path, fname), framework="pt", device="cpu")
        return h

    def _bin(self, fname: str) -> Dict[str, torch.Tensor]:
        if self._bin_cache is None or self._bin_cache[0] != fname:
            self._bin_cache = (fname, torch.load(os.path.join(self.path, fname), map_location="cpu", weights_only=True, mmap=True))
        return self._bin_cache[1]

    def _keys_of(self, fname: str) -> List[str]:
        return list(self._safe(fname).keys()) if fname.endswith(".safetensors") else list(self._bin(fname).keys())

    # ---- Mapping
    def __getitem__(self, key: str) -> torch.Tensor:
        fname = self.weight_map[key]      # KeyError names the missing parameter
        return self._safe(fname).get_tensor(key) if fname.endswith(".safetensors") else self._bin(fname)[key]

    def shape(self, key: str) -> tuple:
        """Shape of a parameter without reading its data (safetensors header; .bin shards are mmap-loaded)."""
        fname = self.weight_map[key]
        if fname.endswith(".safetensors"):
            return tuple(self._safe(fname).get_slice(key).get_shape())
        return tuple(self._bin(fname)[key].shape)

    def __contains__(self, key) -> bool:
        return key in self.weight_map

    def __iter__(self) -> Iterator[str]:
        return iter(self.weight_map)

    def __len__(self) -> int:
        return len(self.weight_map)

    def shards(self) -> List[str]:
        return sorted(set(self.weight_map.values()))


def save_sharded(state_dict: Mapping, path: str, max_shard_bytes: int = 1 << 30, fmt: str = "safetensors") -> List[str]:
    """Write `state_dict` the way HF `save_pretrained` lays it out (shards + index json when more than one shard)."""
    if fmt not in ("safetensors", "bin"):
        raise ValueError(fmt)
    os.makedirs(path, exist_ok=True)
    shards: List[Dict[str, torch.Tensor]] = [{}]
    size = 0
    total = 0
    for k, t in state_dict.items():
        nb = t.numel() * t.element_size()
        if shards[-1] and size + nb > max_shard_bytes:
            shards.append({})
            size = 0
        shards[-1][k] = t.detach().cpu().contiguous()
        size += nb
        total += nb
    stem, ext = ("model", "safetensors") if fmt == "safetensors" else ("pytorch_model", "bin")
    names = [f"{stem}.{ext}"] if len(shards) == 1 else [f"{stem}-{i + 1:05d}-of-{len(shards):05d}.{ext}" for i in range(len(shards))]
    for name, sh in zip(names, shards):
        if fmt == "safetensors":
            from safetensors.torch import save_file
            save_file(sh, os.path.join(path, name), metadata={"format": "pt"})
        else:
            torch.save(sh, os.path.join(path, name))
    if len(shards) > 1:
        index = {"metadata": {"total_size": total}, "weight_map": {k: n for n, sh in zip(names, shards) for k in sh}}
        with open(os.path.join(path, SAFE_INDEX if fmt == "safetensors" else BIN_INDEX), "w") as f:
            json.dump(index, f, indent=2)
    return names


def load_config_dict(path: str) -> dict:
    with open(os.path.join(path, "config.json")) as f:
        return json.load(f)
