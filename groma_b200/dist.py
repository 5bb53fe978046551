"""Multi-GPU plumbing of the forward path: the image batch is split contiguously across ranks,
weights are replicated, and the only exchange is one fixed-shape all-gather of the generated ids and the selected boxes
(this mirrors the reference's own rank-sharded eval, groma/eval/eval_rec.py:80,122-124).  Backend-agnostic
(`nccl` on the GPUs, `gloo` in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, end) image range of `rank`; the first (global_batch % world) ranks get one extra image."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def replayed_randperms(local_counts: Sequence[int], group=None) -> List[torch.Tensor]:
    """torch.randperm draws for this rank's images that are bit-identical to a single-process run over the whole batch.

    The reference shuffles the kept boxes of every image with torch.randperm on the global CPU generator, in image order
    (groma/model/groma.py:275).  randperm(n) advances the generator by an n-dependent amount, so a rank can
    only reproduce its own draws by replaying everyone's: all-gather the per-image keep counts (a few bytes), then draw
    for every image of the global batch in order and keep the local slice.  Images with count 0 draw nothing, exactly as
    the reference's `if len(nms_inds) > 0` branch.  All ranks must hold the same generator state on entry."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [torch.randperm(n) if n > 0 else torch.zeros(0, dtype=torch.long) for n in local_counts]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda") if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(local_counts)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [s.cpu() for s in sizes]
    mx = int(max(int(s) for s in sizes))
    pad = torch.full((mx,), -1, dtype=torch.int64)
    pad[:len(local_counts)] = torch.tensor(list(local_counts), dtype=torch.int64)
    allc = [torch.zeros(mx, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, pad.to(dev), group=group)
    allc = [a.cpu() for a in allc]
    out: List[torch.Tensor] = []
    for r in range(world):
        for j in range(int(sizes[r])):
            n = int(allc[r][j])
            p = torch.randperm(n) if n > 0 else torch.zeros(0, dtype=torch.long)
            if r == rank:
                out.append(p)
    return out


def gather_outputs(sequences: torch.Tensor, boxes: Sequence[torch.Tensor], max_regions: int, group=None):
    """All-gather (sequences [B_loc, L] int64, per-image boxes [R_i, 4]) -> (sequences [B_glob, L], boxes [B_glob, max_regions, 4],
    counts [B_glob]).  Fixed shapes so it is a single collective per tensor; B_loc must be equal on all ranks."""
    B = sequences.shape[0]
    dev = sequences.device
    bx = torch.zeros((B, max_regions, 4), dtype=torch.float32, device=dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    for i, b in enumerate(boxes):
        n = min(len(b), max_regions)
        bx[i, :n] = b[:n].to(dev)
        cnt[i] = n
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sequences, bx, cnt
    world = dist.get_world_size(group)
    seq_all = torch.empty((world * B, sequences.shape[1]), dtype=sequences.dtype, device=dev)
    box_all = torch.empty((world * B, max_regions, 4), dtype=torch.float32, device=dev)
    cnt_all = torch.empty((world * B,), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(seq_all, sequences.contiguous(), group=group)
    dist.all_gather_into_tensor(box_all, bx, group=group)
    dist.all_gather_into_tensor(cnt_all, cnt, group=group)
    return seq_all, box_all, cnt_all
