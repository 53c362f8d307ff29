"""CPU, world_size 2 over gloo: the block partition bench.py uses for N > 1 (no data-path collective;
SURVEY.md 8e).  Every rank derives its own block range, ranges tile the stream, and the gathered
per-rank sizes reproduce the ordered concatenation offsets the host computes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from bench import shard_blocks, reduce_max_time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 13
    lo, hi = shard_blocks(total, rank, world)
    sizes = torch.zeros(total, dtype=torch.int64)
    sizes[lo:hi] = torch.arange(lo, hi) * 10 + 7          # pretend compressed sizes
    dist.all_reduce(sizes)                                # test-only gather; the data path has no collective
    t = reduce_max_time(1.0 + rank, "cpu")
    q.put((rank, lo, hi, sizes.tolist(), t))
    dist.destroy_process_group()


def test_block_partition_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    (r0, lo0, hi0, s0, t0), (r1, lo1, hi1, s1, t1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == 13 and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    assert s0 == s1 == [i * 10 + 7 for i in range(13)]
    assert t0 == t1 == 2.0                                  # max over ranks


def test_shard_blocks_tiles_any_world():
    sys.path.insert(0, ROOT)
    from bench import shard_blocks
    for total in (0, 1, 7, 8192):
        for world in (1, 2, 3, 8):
            cuts = [shard_blocks(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
