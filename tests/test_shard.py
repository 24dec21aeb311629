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


def _queue_worker(rank, name, n_items, n_runs, q, barrier):
    import time
    import cg_driver
    L = cg_driver._lib()
    sq = L.sq_new()
    assert L.sq_attach(sq, name.encode()) == 0
    rng = np.random.default_rng(rank)
    taken = []
    for run in range(n_runs):
        barrier.wait()                                             # the collective that ends every step of a real job
        L.sq_begin(sq, n_items[run])
        mine = []
        while True:
            i = L.sq_take(sq, int(rng.integers(0, 2)))          # a random end each time, like the bulk / small-first groups
            if i < 0:
                break
            mine.append(int(i))
            if rng.uniform() < 0.05:
                time.sleep(0.0005)
        taken.append(mine)
        time.sleep(0.01 * rank)                                    # ranks leave a run at different times
    L.sq_free(sq)
    q.put((rank, taken))


def test_shared_work_list_every_cell_claimed_exactly_once():
    """csrc/oi_shared_queue.h (host build): four processes claim from both ends of one list in POSIX shared memory over
    three consecutive runs of different lengths (generations): every index is claimed exactly once per run, no rank
    ever sees an index of another run."""
    import cg_driver
    cg_driver._lib()                                               # build before the workers race to do it
    name = f"/oi_b200_test_{os.getpid()}"
    n_items = [5000, 1, 1237]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    barrier = ctx.Barrier(4)
    procs = [ctx.Process(target=_queue_worker, args=(r, name, n_items, len(n_items), q, barrier)) for r in range(4)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(4))
    [p.join(60) for p in procs]
    cg_driver._lib().sq_unlink(name.encode())
    for run, n in enumerate(n_items):
        allc = sorted(i for r in range(4) for i in got[r][run])
        assert allc == list(range(n)), (run, len(allc))
    assert sum(len(got[r][0]) > 0 for r in range(4)) >= 2          # the work really was shared


def _owned_worker(rank, world, port, q):
    import torch.distributed as dist
    from optimalinterpolation_b200.shard import gather_owned
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 23
    owner = (np.arange(n) * 7 % 3 == 0).astype(int)              # cells of rank 1 where True, rank 0 elsewhere
    owned = owner == rank
    owned[5] = True                                                 # a cell without observations: "owned" by both
    out = np.full((n, 8), np.nan)
    out[owned] = np.arange(n)[owned, None] + 0.125 * np.arange(8)[None, :] + 100 * rank
    out[5] = np.nan
    full, counts = gather_owned(out, owned)
    q.put((rank, full, counts))
    dist.barrier(); dist.destroy_process_group()


def test_gather_owned_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_owned_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(2))}
    [p.join(60) for p in procs]
    assert np.array_equal(got[0][0], got[1][0], equal_nan=True)
    full = got[0][0]
    owner = (np.arange(23) * 7 % 3 == 0).astype(int)
    for c in range(23):
        if c == 5:
            assert np.isnan(full[c]).all()
        else:
            assert np.allclose(full[c], c + 0.125 * np.arange(8) + 100 * owner[c])
    assert list(got[0][1]) == [int((owner == 0).sum()) + (1 if owner[5] != 0 else 0), int((owner == 1).sum()) + (1 if owner[5] != 1 else 0)]
