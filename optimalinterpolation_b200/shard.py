"""Cost-balanced assignment of grid cells to GPUs: a static LPT split (``lpt_partition`` + ``gather_results``) and the
merge for the dynamic shared work list (``gather_owned``; the list itself lives in csrc/oi_shared_queue.h).

Replaces the reference's round-robin ``split(container, count)`` (GPR_CS2S3.py:18-23, :250-256):
cells are independent, inputs are replicated on every rank (as in the reference, where every rank
loads all data, :201-246) and only cell indices are sharded.  Cost model: one NLML evaluation is
~n^3 flops, so cells are assigned by LPT (longest processing time first) greedy on n_c^3.
"""
from __future__ import annotations

import heapq

import numpy as np


def lpt_partition(counts, world_size: int):
    """Return a list of ``world_size`` int64 index arrays (each sorted ascending) whose sum of
    counts**3 is balanced.  Deterministic: ties broken by cell index, then by rank."""
    counts = np.asarray(counts, dtype=np.int64)
    cost = counts.astype(np.float64) ** 3
    order = np.lexsort((np.arange(len(counts)), -cost))      # descending cost, stable in index
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    parts = [[] for _ in range(world_size)]
    for c in order:
        load, r = heapq.heappop(heap)
        parts[r].append(int(c))
        heapq.heappush(heap, (load + cost[c], r))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def imbalance(counts, parts) -> float:
    """max rank cost / mean rank cost - 1."""
    cost = np.asarray(counts, dtype=np.float64) ** 3
    loads = np.array([cost[p].sum() for p in parts])
    return float(loads.max() / loads.mean() - 1.0) if loads.mean() > 0 else 0.0


def gather_results(local_out: np.ndarray, part: np.ndarray, n_cells: int, parts=None):
    """The only collective of the path: gather every rank's (n_local, 8) result rows into the
    full (n_cells, 8) field on all ranks (reference: COMM.gather, GPR_CS2S3.py:262).  Uses
    torch.distributed (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        full = np.full((n_cells, local_out.shape[1]), np.nan)
        full[part] = local_out
        return full
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sizes = [len(p) for p in parts] if parts is not None else None
    if sizes is None:
        t = torch.tensor([len(part)], device=dev, dtype=torch.int64)
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        sizes = [int(v.item()) for v in lst]
    width = local_out.shape[1] + 1
    mx = max(sizes)
    buf = torch.full((mx, width), float("nan"), dtype=torch.float64, device=dev)
    if len(part):
        buf[:len(part), 0] = torch.from_numpy(part.astype(np.float64)).to(dev)
        buf[:len(part), 1:] = torch.from_numpy(np.ascontiguousarray(local_out)).to(dev)
    allbuf = torch.empty((world, mx, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf.view(world * mx, width), buf)
    allbuf = allbuf.cpu().numpy()
    full = np.full((n_cells, local_out.shape[1]), np.nan)
    for r in range(world):
        k = sizes[r]
        idx = allbuf[r, :k, 0].astype(np.int64)
        full[idx] = allbuf[r, :k, 1:]
    return full


def gather_owned(out: np.ndarray, owned: np.ndarray):
    """Merge for the dynamic mode (``Handle.set_shared_queue``): every rank ran the SAME cells through one shared
    cost-sorted work list and holds the rows of the cells it computed (``owned``), NaN elsewhere.  One all-gather of the
    (n_cells, 1 + 8) arrays; each cell's row is taken from the rank that owns it (the lowest such rank: cells without
    observations are "owned" by every rank).  Returns the full field and the per-rank owned counts."""
    import torch
    import torch.distributed as dist
    out = np.ascontiguousarray(out, dtype=np.float64)
    owned = np.asarray(owned, dtype=bool)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if not owned.all():
            raise ValueError("gather_owned: a single process must own every cell")
        return out.copy(), np.array([int(owned.sum())])
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n, w = out.shape
    buf = torch.empty((n, w + 1), dtype=torch.float64, device=dev)
    buf[:, 0] = torch.from_numpy(owned.astype(np.float64)).to(dev)
    buf[:, 1:] = torch.from_numpy(out).to(dev)
    allbuf = torch.empty((world, n, w + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf.view(world * n, w + 1), buf)
    allbuf = allbuf.cpu().numpy()
    own = allbuf[:, :, 0] > 0.5                                  # (world, n)
    if not own.any(axis=0).all():
        raise RuntimeError(f"gather_owned: {int((~own.any(axis=0)).sum())} cells were computed by no rank")
    first = np.argmax(own, axis=0)                               # lowest owning rank per cell
    full = allbuf[first, np.arange(n), 1:]
    return full, own.sum(axis=1)
