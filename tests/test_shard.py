"""CPU: cost-balanced sharding and the result gather (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from optimalinterpolation_b200.shard import gather_results, imbalance, lpt_partition


def test_lpt_is_a_partition_and_balanced():
    rng = np.random.default_rng(0)
    counts = rng.integers(150, 1800, size=2389)
    for w in (1, 2, 4, 8):
        parts = lpt_partition(counts, w)
        allc = np.sort(np.concatenate(parts))
        assert np.array_equal(allc, np.arange(len(counts)))
        assert imbalance(counts, parts) < 0.01
    assert all(np.array_equal(a, b) for a, b in zip(lpt_partition(counts, 4), lpt_partition(counts, 4)))


def test_lpt_ragged_and_empty():
    parts = lpt_partition([5, 0, 0, 7, 1], 8)          # more ranks than cells
    assert sum(len(p) for p in parts) == 5 and sum(len(p) == 0 for p in parts) >= 3
    parts = lpt_partition([], 2)
    assert all(len(p) == 0 for p in parts)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = (np.arange(37) * 13 % 29 + 1) * 20
    parts = lpt_partition(counts, world)
    mine = parts[rank]
    local = np.stack([mine * 1.0 + 0.25 * k for k in range(8)], axis=1)
    if rank == 1 and len(local):
        local[0, 2] = np.nan                         # per-cell failures travel as data
    full = gather_results(local, mine, len(counts), parts)
    q.put((rank, full))
    dist.barrier(); dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert np.array_equal(got[0], got[1], equal_nan=True)
    full = got[0]
    assert np.isnan(full).sum() == 1
    ok = ~np.isnan(full[:, 2])
    assert np.array_equal(full[ok, 0], np.arange(37)[ok] * 1.0) and np.allclose(full[ok, 7], np.arange(37)[ok] + 1.75)
