"""CPU-only, world_size 2 over gloo: the N>1 plumbing of bench.py (contiguous sharding, max-over-ranks
timing, summed work).  The data path itself has no collective (reads partition by index)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_range(1001, world, rank)
    my_ms = 10.0 + 5.0 * rank                      # rank 1 is the slow one
    (mx, e2e), (cells,) = bench.reduce_over_ranks(dist, [my_ms, 2 * my_ms], [float(hi - lo)], "cpu")
    dist.barrier()
    q.put((rank, lo, hi, mx, e2e, cells))
    dist.destroy_process_group()


def test_sharding_and_reductions_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, mx0, e0, c0), (r1, lo1, hi1, mx1, e1, c1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)          # ceil(1001/2) = 501, contiguous, disjoint, complete
    assert mx0 == mx1 == 15.0 and e0 == e1 == 30.0              # MAX over ranks
    assert c0 == c1 == 1001.0                                   # work summed over ranks


def test_shard_range_edge_cases():
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.shard_range(5, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    assert bench.shard_range(0, 4, 2) == (0, 0)
    cover = [bench.shard_range(103, 4, r) for r in range(4)]
    assert cover[0][0] == 0 and cover[-1][1] == 103 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
