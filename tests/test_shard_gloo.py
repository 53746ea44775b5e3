"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (contiguous batch shards, no data-path collective,
units summed and time taken as the max over ranks)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libxsmm_b200.shard import aggregate, shard_range, weak_batch  # noqa: E402


def test_shard_range_covers_without_overlap():
    for total in (0, 1, 7, 8192, 65536, 1000000, 999984):
        for world in (1, 2, 3, 4, 8):
            for granule in (1, 4, 16):
                pieces = [shard_range(total, world, r, granule) for r in range(world)]
                assert pieces[0][0] == 0 and pieces[-1][1] == total
                for (b0, e0), (b1, e1) in zip(pieces, pieces[1:]):
                    assert e0 == b1 and b0 <= e0
                sizes = [e - b for b, e in pieces]
                assert all(b % granule == 0 for b, _ in pieces)
                assert max(sizes[:-1] + [sizes[-1] - total % granule]) - min(sizes[:-1] + [sizes[-1] - total % granule]) <= granule
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    assert weak_batch(65536, 8) == (65536, 524288)
    # BASELINE configs[3]: 8192 m_blocks over 8 GPUs in groups of 4 m_blocks (one M=128 MMA operand) -> 1024 each
    assert [shard_range(8192, 8, r, 4) for r in range(8)] == [(r * 1024, (r + 1) * 1024) for r in range(8)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(65536 + 3, world, rank)
        # stand-in for the per-rank pass: checksum of the tile ids this rank owns, a fake per-rank time
        local = torch.arange(b, e, dtype=torch.float64).sum().item()
        units, ms = aggregate(dist, e - b, 1.0 + rank)
        t = torch.tensor([local], dtype=torch.float64)
        dist.all_reduce(t)
        # the strong-scaling cut of bench.py: BCSC m_blocks in groups of 4, fsspmdm columns in groups of 16; the optional gather of the
        # C ranges (here: the ids of the units) must reassemble the whole job in order
        for total, granule in ((8192, 4), (1000000, 16)):
            sb, se = shard_range(total, world, rank, granule)
            assert sb % granule == 0 and (se - sb) * world >= total - granule * world
            size = (total + world - 1) // world + granule
            mine = torch.full((size,), -1, dtype=torch.int64); mine[:se - sb] = torch.arange(sb, se)
            parts = [torch.empty(size, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(parts, mine)
            whole = torch.cat([p_[p_ >= 0] for p_ in parts])
            assert torch.equal(whole, torch.arange(total)), (total, granule)
        q.put((rank, units, ms, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_ranks_sum_units_and_take_max_time():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    total = 65536 + 3
    for rank, units, ms, chk in got:
        assert units == total and ms == 2.0                      # sum of units, max of the per-rank times
        assert chk == total * (total - 1) / 2                    # every tile owned exactly once
