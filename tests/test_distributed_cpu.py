"""N>1 host logic on CPU: world_size-2 gloo processes shard a job with no data-path collective and
agree on the max-over-ranks time."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vspbfr_b200.sharding import max_over_ranks, micro_batches, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_micro_batches_cover_slice():
    assert list(micro_batches(3, 20, 8)) == [(3, 11), (11, 19), (19, 20)]
    assert list(micro_batches(5, 5, 8)) == []
    with pytest.raises(ValueError):
        list(micro_batches(0, 4, 0))


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # stand-in for the hot path: each rank "restores" its own slice independently
    data = torch.arange(total, dtype=torch.float32)
    mine = data[lo:hi] * 2
    t = max_over_ranks(0.5 + rank)           # slowest rank defines the step time
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine.tolist()))   # test-only check, not part of the data path
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put((t, gathered))


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, total = 2, 11
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == pytest.approx(1.5)
    flat = []
    for lo, hi, vals in gathered:
        assert len(vals) == hi - lo
        flat += vals
    assert flat == [2.0 * i for i in range(total)]
