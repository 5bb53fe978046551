"""world_size-2 gloo tests of the multi-GPU host logic (groma_b200/dist.py): batch sharding, RNG replay that keeps the
randperm draws bit-identical to a single-process run, and the final fixed-shape all-gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from groma_b200.dist import gather_outputs, replayed_randperms, shard_range


def test_shard_range_covers_batch():
    for gb in (1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_range(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


COUNTS = [5, 0, 17, 100, 3, 1, 0, 42]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(len(COUNTS), world, rank)
    torch.manual_seed(7)
    perms = replayed_randperms(COUNTS[s:e])
    seq = torch.arange((e - s) * 6, dtype=torch.int64).reshape(e - s, 6) + 1000 * rank
    boxes = [torch.full((min(n, 4), 4), float(rank * 10 + i)) for i, n in enumerate(COUNTS[s:e])]
    seq_all, box_all, cnt_all = gather_outputs(seq, boxes, max_regions=4)
    q.put((rank, [p.tolist() for p in perms], seq_all.tolist(), box_all[:, 0, 0].tolist(), cnt_all.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replay_and_gather():
    world, port = 2, 29500 + (os.getpid() % 1000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(7)
    single = [torch.randperm(n).tolist() if n > 0 else [] for n in COUNTS]
    got = res[0][1] + res[1][1]
    assert got == single                                   # bit-identical to the single-process draw order
    assert res[0][2] == res[1][2] and len(res[0][2]) == len(COUNTS)
    assert res[0][2][0][0] == 0 and res[0][2][4][0] == 1000    # rank order preserved
    assert res[0][4] == [min(n, 4) for n in COUNTS]


def test_single_process_fallbacks():
    torch.manual_seed(3)
    a = replayed_randperms([4, 0, 2])
    torch.manual_seed(3)
    assert a[0].tolist() == torch.randperm(4).tolist() and a[1].numel() == 0 and a[2].tolist() == torch.randperm(2).tolist()
    seq, bx, cnt = gather_outputs(torch.zeros(2, 3, dtype=torch.int64), [torch.ones(2, 4), torch.zeros(0, 4)], 5)
    assert seq.shape == (2, 3) and bx.shape == (2, 5, 4) and cnt.tolist() == [2, 0]
