"""The N > 1 path on CPU: world_size-2 gloo. The range shards embarrassingly: ranks own disjoint,
job-aligned, contiguous sub-ranges whose union is the whole range, the only cross-rank traffic is the final
counter reduction. The per-rank compute here is the oracle standing in for the GPU (test infrastructure)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_shards_partition_the_range():
    import bench

    for world in (1, 2, 4, 8):
        shards = [bench.shard_of(r, world) for r in range(world)]
        assert shards[0][0] == bench.RANGE_S
        for (s0, n0), (s1, _) in zip(shards, shards[1:]):
            assert s0 + n0 == s1 and n0 % (1 << 21) == 0  # whole reference jobs
        assert sum(n for _, n in shards) == bench.RANGE_KEYS


def test_config4_shards_are_job_aligned_and_disjoint():
    """BASELINE configs[3]: the range of Makefile:58 cut into one shard per GPU (bench.py's add_endo_blf leg)"""
    import bench

    for world in (1, 2, 4, 8):
        shards = [bench.shard71_of(r, world) for r in range(world)]
        assert shards[0][0] == 0x400000000000000000
        for (s0, n0), (s1, _) in zip(shards, shards[1:]):
            assert s0 + n0 == s1 and n0 % (1 << 21) == 0
        last_s, last_n = shards[-1]
        assert last_s + last_n <= 0x7FFFFFFFFFFFFFFFFF and 0x7FFFFFFFFFFFFFFFFF - (last_s + last_n) < world << 21
        # every rank's timed prefix (3 steps of 2^32 keys) lies inside its own shard
        assert all(n > 3 << 32 for _, n in shards)


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import ecloop_b200.host as H

    # a small range split the same way bench.py splits the 2^40 one: contiguous, 2048-aligned
    range_s, total = 0x8000, 2048 * 24
    per = total // world
    start = range_s + rank * per
    flt = H.load_filter(ROOT / "tests" / "golden" / "btc-puzzles-hash")
    oflt = O.filter_from_text_file(ROOT / "tests" / "golden" / "btc-puzzles-hash")
    n, hits = O.add_span(start, 1, per, O.A33, oflt)
    found = [H.calc_priv(start, 1, k, e) for k, e, kd, h, _ in hits if flt.check_exact(tuple(int(h[i:i + 8], 16) for i in range(0, 40, 8)))]
    t = torch.tensor([per, len(found)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the only cross-rank step: counters
    mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
    gathered = [None] * world
    dist.all_gather_object(gathered, found)
    if rank == 0:
        out_q.put((t.tolist(), mx.item(), sorted(k for g in gathered for k in g)))
    dist.destroy_process_group()


def test_world2_gloo_union_equals_single_rank():
    import oracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (total_keys, n_found), mx, keys = res
    oflt = O.filter_from_text_file(ROOT / "tests" / "golden" / "btc-puzzles-hash")
    n, hits = O.add_span(0x8000, 1, 2048 * 24, O.A33, oflt)
    assert total_keys == 2048 * 24 and mx == 2.0
    assert keys == sorted(h[4] for h in hits) and n_found == len(hits) == 1  # puzzle 16 (c936) lies in this window
