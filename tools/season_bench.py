"""BASELINE.json configs[3]: a synthetic winter season on the GPUs of one box, sharded BY DAY (season.run_season_sharded).

    python tools/season_bench.py --days 16                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/season_bench.py --days 16 [--optimiser lbfgs] [--cell-stride 8]

Each day = day setup (window flattening, ice cells, prior mean) + pass 1 (fit + predict) + hyperparameter smoothing +
pass 2 (predict with the smoothed fields), GPR_CS2S3.py:201-336.  --cell-stride k keeps every k-th ice cell of each day
(same n histogram) to bound the run time of the CG parity mode.  Prints one JSON line on rank 0: wall time (max over
ranks, barrier on both sides), days, cells, sustained cells/s."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ap = argparse.ArgumentParser()
ap.add_argument("--days", type=int, default=16)
ap.add_argument("--cell-stride", type=int, default=1)
ap.add_argument("--optimiser", default="cg", choices=["cg", "lbfgs"])
ap.add_argument("--no-smooth-pass", action="store_true")
args = ap.parse_args()
import torch
import torch.distributed as dist
from optimalinterpolation_b200.season import make_synthetic_season, run_season_sharded
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
obs, sie, x, y = make_synthetic_season(args.days)
if args.cell_stride > 1:                       # thin the ice mask: every k-th ice cell (row-major), the same cells every day
    ids = np.where(~np.isnan(sie[:, :, 0]))
    keep = np.zeros(len(ids[0]), bool); keep[::args.cell_stride] = True
    sie[ids[0][~keep], ids[1][~keep], :] = np.nan
fast = args.optimiser == "lbfgs"
kw = dict(device=local, smooth_pass=not args.no_smooth_pass)
if fast:
    kw.update(optimiser=1, grad_convention=1)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = run_season_sharded(obs, sie, x, y, days=range(args.days), **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
if rank == 0:
    tm = out["_timing"]
    cells = sum(v["cells"] for v in tm.values())
    print(json.dumps({"workload": f"synthetic season, {args.days} days, every {args.cell_stride}-th ice cell, both passes" if not args.no_smooth_pass else "pass 1 only",
                      "optimiser": args.optimiser, "n_gpus": world, "days": len(tm), "cells": cells, "wall_s": wall, "cells_per_s": cells / wall,
                      "days_per_hour": len(tm) / wall * 3600, "seconds_per_day_mean": float(np.mean([v["seconds"] for v in tm.values()])),
                      "seconds_per_day_max": float(np.max([v["seconds"] for v in tm.values()])),
                      "nonfinite_cells": sum(v["nonfinite"] for v in tm.values()), "sharding": "by day, days[rank::world]"}))
if world > 1:
    dist.destroy_process_group()
