#!/usr/bin/env python
"""bench.py -- GP cells/sec (fit + predict) on the synthetic 25 km pan-Arctic day (BASELINE.json).

A step = one pass of the hot path (H2D of the day's observations and the step's cell coordinates -> neighbour gather
-> lockstep CG fit -> posterior -> D2H of the result rows -> gather of the field over the ranks) over one batch of
cells, through the ONE C-ABI call a user makes (oi_gpr_day, host buffers in / host buffers out).  The batch is
``2 x --gpus N`` stripes of the day (stripe s = every 16th ice cell starting at s, so each stripe has the day's
n-histogram; ~1195 cells per stripe, 2 stripes = 1/8 day per GPU and step, per-GPU work fixed => weak scaling; at N=8 one
step is the whole day sharded over the 8 GPUs = BASELINE.json configs[2]).  Cells of a step are sharded over ranks by
LPT on n^3 (optimalinterpolation_b200/shard.py); the only collective is the final gather of the result rows.

There is ONE timed loop of exactly --steps steps (after --warmup identical steps):
  e2e   : cells/s of that loop as the caller sees it (CUDA events on the launching stream and the wall clock,
          whichever is longer; max over ranks)
  value : the same steps with the host<->device copies taken out: per step, the device time the library measures with
          CUDA events on its launching stream from the first gather kernel to the last result (oi_stats.ms_gather +
          ms_total; max over ranks per step)
After the loop, while the wall-clock budget (--budget-s, default 575 s from process start) allows and only at N=1:
the CPU baseline sample, a single-stream pass that times each kernel family, and the WHOLE day in one call
(BASELINE.json configs[1]; ~8 steps' worth, reported under "full_day").

  --workload cfg5  : BASELINE.json configs[4] (12.5 km lattice, 500 km radius, n ~ 1200...5400), 64 cells per GPU and step
  --workload realmask : the same tracks over the real ice mask of a QuickLook day (17 697 cells, n 2...2326)
  --impl reference : the reference's CPU path (oracle port, see oracle/) on the host cores, whole fits, time-boxed
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

T_START = time.time()

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: one hardware queue per stream group

METRIC = "GP cells/sec (fit+predict), 25km Arctic day"
N_STRIPES = 16
CFG5_CELLS_PER_GPU = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="day25", choices=["day25", "cfg5", "realmask"])
    ap.add_argument("--optimiser", default="cg", choices=["cg", "lbfgs"],
                    help="cg = the reference's scipy-CG restatement (parity mode, the headline); lbfgs = exact-gradient L-BFGS fast mode")
    ap.add_argument("--stripes-per-gpu", type=int, default=2)
    ap.add_argument("--sharding", default="dynamic", choices=["dynamic", "lpt"],
                    help="N>1: dynamic = one cost-sorted work list shared by the ranks (POSIX shared memory), lpt = static LPT split on n^3")
    ap.add_argument("--max-active", type=int, default=0)
    ap.add_argument("--budget-s", type=float, default=float(os.environ.get("OI_BENCH_BUDGET_S", 575)),
                    help="wall-clock budget of the whole run; the optional passes after the timed loop only start while it allows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-day", action="store_true", help="skip the single whole-day pass (N=1 only)")
    ap.add_argument("--no-family-pass", action="store_true", help="skip the single-stream pass that times each kernel family")
    ap.add_argument("--groups", type=int, default=0, help="lockstep stream groups (0 = library default)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([v.strip() for v in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_workload(args, world):
    """The synthetic day and the cell indices of one step (deterministic; both arms use it)."""
    from optimalinterpolation_b200.synthetic import make_day, make_day_cfg5, make_day_real_mask
    if args.workload == "realmask":
        day = make_day_real_mask(os.path.join(ROOT, "tests", "golden", "quicklook_icemask.npz"))
        nc = len(day.X)
        cells = np.sort(np.concatenate([np.arange(s, nc, N_STRIPES) for s in range(min(world * args.stripes_per_gpu, N_STRIPES))]))
    elif args.workload == "cfg5":
        day = make_day_cfg5()
        cells = np.unique(np.linspace(0, len(day.X) - 1, CFG5_CELLS_PER_GPU * world).round().astype(np.int64))
    else:
        day = make_day()
        nc = len(day.X)
        cells = np.sort(np.concatenate([np.arange(s, nc, N_STRIPES) for s in range(min(world * args.stripes_per_gpu, N_STRIPES))]))
    return day, cells


def make_config(args, world, day, cells, counts_step):
    """The ``config`` object of the JSON line: identical for both arms (they are given the same workload)."""
    from optimalinterpolation_b200.shard import lpt_partition, imbalance
    parts = lpt_partition(counts_step, world)
    if args.workload == "cfg5":
        wl = (f"{len(cells)} cells ({CFG5_CELLS_PER_GPU} per GPU) of the synthetic 12.5 km / 500 km day (BASELINE.json configs[4]: 640x640 "
              f"lattice, {len(day.X)} ice cells, {day.z.size} obs, 9 days)")
    elif args.workload == "realmask":
        wl = (f"{world * args.stripes_per_gpu}/16 stripes of the synthetic 25 km day over the REAL ice mask of the reference's QuickLook product "
              f"of 2019-01-28 (SURVEY.md 8d option B: {len(day.X)} ice cells, {day.z.size} obs, r=300 km, 9 days)")
    else:
        wl = (f"{world * args.stripes_per_gpu}/16 stripes of the synthetic 25 km pan-Arctic day "
              f"(SURVEY.md 8d: 320x320 lattice, {len(day.X)} ice cells, {day.z.size} obs, r=300 km, 9 days)")
    opt = ("scipy-CG restatement (reference gradient convention), x0 as GPR_CS2S3.py:217" if args.optimiser == "cg" else
           "FAST MODE, not the parity mode: exact-gradient L-BFGS (m=8) on the device, x0 as GPR_CS2S3.py:217")
    return {"workload": wl, "cells_per_step": int(len(cells)), "n_obs": int(day.z.size),
            "n_min_median_max": [int(counts_step.min()), int(np.median(counts_step)), int(counts_step.max())],
            "optimiser": opt,
            "sharding": (f"dynamic: one cost-sorted work list shared by the {world} ranks (largest cells claimed from the front, smallest "
                         f"from the back; 64-bit cursor word in POSIX shared memory), every rank holds all {len(cells)} cells")
            if world > 1 and args.sharding == "dynamic" else
            f"LPT on n^3 over {world} ranks, imbalance {imbalance(counts_step, parts):.4f}",
            "cache": "per-iteration working set (sum of n_pad^2*8 B over active cells, GBs) is far larger than the 126 MB L2; no L2 flush needed"}, parts


def boxed_plan(total_steps, budget_s=200.0):
    """(cells per core, time box in s per cell) of one reference-arm step so that ``total_steps`` steps fit the budget."""
    per_step = min(36.0, budget_s / max(total_steps, 1))
    cpc = 2 if per_step >= 20.0 else 1
    return cpc, max(2.0, 0.8 * per_step / cpc)


def boxed_sample_text(s, cores, cpc, box, steps):
    return (f"whole GPR3D fits (scipy CG from x0 + prediction) of the reference path, one process per core with 1 BLAS thread, "
            f"{cores} cores: per step {cores * cpc} cells at evenly spaced quantiles of the day's n-distribution "
            f"(cells of tests/golden/day_fit_sample_1k.npz, a different offset every step), each fit time-boxed at {box:.1f} s; "
            f"{s['n_finished']} of {s['n_sample']} fits over {steps} step(s) finished inside the box and are measured whole, the others "
            f"cost (their own measured s/evaluation) x (their recorded evaluation count + 1/3), the reference optimiser being "
            f"deterministic; n {s['n_min']}..{s['n_max']}, {s['evals_timed']} evaluations timed, mean cost {s['mean_cost_s']:.1f} core-s per cell "
            f"(+-{100 * (s['sem_rel'] or 0):.0f} % s.e.m.), nfev mean {s['nfev_mean']:.0f}")


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path (oracle port) on the host cores: whole fits, time-boxed."""
    if rank != 0:
        return
    import warnings
    warnings.simplefilter("ignore")
    from oracle import cpu_baseline
    from scipy.spatial import cKDTree
    day, cells = make_workload(args, world)
    counts_step = np.asarray(cKDTree(np.c_[day.x_train, day.y_train]).query_ball_point(
        day.X[cells], r=day.radius_km * 1000.0, return_length=True))
    config, _ = make_config(args, world, day, cells, counts_step)
    if args.workload != "day25":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm's fit fixture covers the 25 km day only"}))
        return
    cores = os.cpu_count() or 1
    fx = cpu_baseline.load_fixture()
    cpc, box = boxed_plan(args.warmup + args.steps)
    rows, wall = [], []
    for s in range(args.warmup + args.steps):
        r = cpu_baseline.run_boxed_fits(day, step_index=s, cores=cores, cells_per_core=cpc, box_s=box, fixture=fx)
        if s >= args.warmup:
            rows += r["rows"]; wall.append(r["wall_s"])
    sm = cpu_baseline.summarise_boxed(rows, cores, fixture=fx)
    v = sm["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": len(cells) / v * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": boxed_sample_text(sm, cores, cpc, box, args.steps)},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_step_is": "the time the host cores need for one step at the measured rate (cells_per_step / value); the sampling itself "
                          f"took {float(np.mean(wall)) * 1e3:.0f} ms per step",
        "sample_wall_ms_per_step": float(np.mean(wall)) * 1e3,
        "value_all_fixture_cells": sm.get("value_all_fixture_cells"),
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import optimalinterpolation_b200 as oi
    from optimalinterpolation_b200.shard import gather_results, gather_owned

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # a non-default stream: the library launches on it and the events below are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def left():
        return args.budget_s - (time.time() - T_START)

    day, cells = make_workload(args, world)
    Xstep = day.X[cells]
    h = oi.Handle(local_rank)
    h.set_stream(stream.cuda_stream)

    # pinned host staging of the inputs (every step copies from these)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    px, py, pt, pz = map(pinned, (day.x_train, day.y_train, day.t_train, day.z))

    # planning (untimed, deterministic): counts of the step's cells -> LPT shard
    h.set_observations(px, py, pt, pz)
    h.set_cells(Xstep)
    counts_step = h.gather_neighbours(day.radius_km * 1000.0).copy()
    config, parts = make_config(args, world, day, cells, counts_step)
    dynamic = world > 1 and args.sharding == "dynamic"
    if dynamic:
        # every rank runs ALL cells of the step through one shared cost-sorted work list (csrc/oi_shared_queue.h): the same
        # fresh segment name on every rank, attached before the first run
        name = [f"/oi_b200_bench_{os.getpid()}_{int(time.time())}"]
        dist.broadcast_object_list(name, src=0)
        ok = torch.ones(1, device=dev)
        try:
            h.set_shared_queue(name[0])
        except Exception as e:                       # no POSIX shared memory in this container: every rank falls back together
            print(f"rank {rank}: shared work list unavailable ({e}); static LPT split instead", file=sys.stderr)
            ok[0] = 0
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok[0]) < 0.5:
            h.set_shared_queue(None)
            dynamic = False
            from optimalinterpolation_b200.shard import imbalance
            config["sharding"] = f"LPT on n^3 over {world} ranks, imbalance {imbalance(counts_step, parts):.4f} (shared memory unavailable: fell back from the dynamic list)"
    if dynamic:
        mine = np.arange(len(cells))
    else:
        mine = parts[rank]
    Xmine = pinned(Xstep[mine])
    fast = args.optimiser == "lbfgs"
    params = h.make_params(day.radius_km * 1000.0, day.T_mid, day.mean, day.x0, mode=0, max_active=args.max_active,
                           n_groups=args.groups, optimiser=1 if fast else 0, grad_convention=1 if fast else 0)

    def step():
        res = h.gpr_day(px, py, pt, pz, Xmine, params)
        st = h.stats()
        if dynamic:
            full, owned_counts = gather_owned(res["out"], h.get_owned())
            res["owned_counts"] = owned_counts
        else:
            full = gather_results(res["out"], mine, len(cells), parts)
        return res, st, full

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    agg, dev_ms = {}, []
    for _ in range(args.steps):
        res, st, full = step()
        dev_ms.append(st["ms_gather"] + st["ms_total"])
        for k, v in st.items():
            if not isinstance(v, list):
                agg[k] = agg.get(k, 0) + v
    e1.record(stream); barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    ev_ms = e0.elapsed_time(e1)
    tt = torch.tensor([ev_ms, wall * 1e3] + dev_ms, device=dev, dtype=torch.float64)
    launches = torch.tensor([agg["n_launches"]], device=dev, dtype=torch.float64)
    rank_ms = torch.zeros(world, device=dev, dtype=torch.float64); rank_ms[rank] = sum(dev_ms) / args.steps
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        dist.all_reduce(rank_ms, op=dist.ReduceOp.SUM)
    e2e_ms = max(float(tt[0]), float(tt[1]))
    ms = float(tt[2:].sum())                                  # device-resident time of the K steps, max over ranks per step
    value = len(cells) * args.steps / (ms * 1e-3)
    e2e_value = len(cells) * args.steps / (e2e_ms * 1e-3)
    step_s = e2e_ms * 1e-3 / args.steps
    h2d = 4 * day.z.size * 8 + Xmine.size * 8
    d2h = len(mine) * (64 + 12)
    status_hist = np.bincount(res["status"][h.get_owned()] if dynamic else res["status"], minlength=6).tolist()

    solo = world == 1
    # ---------------- CPU baseline: bounded sample of whole fits on the host cores (rank 0, N=1 only) ----------------
    cpu = None
    if solo and not args.no_cpu_baseline and args.workload == "day25" and left() > 45:
        import warnings
        warnings.simplefilter("ignore")
        from oracle import cpu_baseline
        cores = os.cpu_count() or 1
        cpc, box = 2, 9.0
        r = cpu_baseline.run_boxed_fits(day, step_index=0, cores=cores, cells_per_core=cpc, box_s=box)
        sm = cpu_baseline.summarise_boxed(r["rows"], cores)
        cpu = {"value": sm["value"], "unit": "cells/s", "cores": cores, "kind": "port",
               "sample": boxed_sample_text(sm, cores, cpc, box, 1) + f"; {r['wall_s']:.0f} s wall",
               "value_all_fixture_cells": sm.get("value_all_fixture_cells")}
    # ---------------- single-stream pass: device time of each kernel family (diagnostic, untimed) ----------------
    # With several stream groups the families of different groups overlap, so their per-stream times do not
    # add up; one extra step with n_groups=1 gives the per-family numbers that the ncu launch list is compared to.
    fam1 = None
    if solo and not args.no_family_pass and left() > 1.6 * step_s + 10:
        p1 = h.make_params(day.radius_km * 1000.0, day.T_mid, day.mean, day.x0, mode=0, max_active=args.max_active, n_groups=1,
                           optimiser=1 if fast else 0, grad_convention=1 if fast else 0)
        h.set_cells(Xmine); h.gather_neighbours(day.radius_km * 1000.0); h.run(p1)
        fam1 = h.stats()
    # ---------------- the whole day in one call on one GPU (BASELINE.json configs[1]) ----------------
    full_day = None
    day_est = step_s * len(day.X) / max(len(cells), 1) * 1.05
    if solo and not args.no_full_day and args.workload == "day25" and left() > day_est + 10:
        Xday = pinned(day.X)
        t0 = time.perf_counter(); e0.record(stream)
        res_d = h.gpr_day(px, py, pt, pz, Xday, params)
        e1.record(stream); torch.cuda.synchronize()
        sec = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
        st_d = h.stats()
        full_day = {"cells": int(len(day.X)), "seconds": sec, "cells_per_s": len(day.X) / sec,
                    "tflops": st_d["flops"] / sec * 1e-12, "evals": int(st_d["n_evals"]), "iterations": int(st_d["n_iterations"]),
                    "finite_frac": float(np.isfinite(res_d["out"][:, 0]).mean()),
                    "status_hist": np.bincount(res_d["status"], minlength=6).tolist(),
                    "how": "one oi_gpr_day call with host buffers (H2D + gather + fit + predict + D2H), single pass, not part of the timed steps"}
    elif solo and not args.no_full_day and args.workload == "day25":
        full_day = {"skipped": f"needs ~{day_est:.0f} s, {left():.0f} s of the {args.budget_s:.0f} s budget left"}
    if rank == 0:
        # measured FP64 peak (MEASURED_PEAKS.json holds no FP64 entry): cuBLAS DGEMM 8192^3
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(4):
            p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
            p0.record(); a @ b; p1.record(); torch.cuda.synchronize(); best = min(best, p0.elapsed_time(p1))
        peak_tf = 2 * 8192 ** 3 / best * 1e-9
        del a, b
        # dominant kernel family = the FP64 DMMA tile kernels (one gemm_nt_stream core: k_chol_update, k_chol_panel,
        # k_trtri, k_lauum_trace).  achieved = their algorithmic flops / the device time of the whole
        # lockstep region of the timed steps (a lower bound: build, substitutions, finalize and host gaps included).
        ach = agg["flops_factor"] / agg["ms_total"] * 1e-9
        dmma_launches = agg["launches_chol"] + agg["launches_trtri"] + agg["launches_lauum"]
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("dmma_tile_kernels")
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "FP64 DMMA tile kernels (k_chol_update+k_chol_panel+k_trtri+k_lauum_trace)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA issue-rate microbenchmark: 37.1 TFLOP/s, profiles/r01_fp64_peak_microbench.txt)",
                    "flops_per_launch": agg["flops_factor"] / max(dmma_launches, 1),
                    "ms_per_launch": agg["ms_total"] / max(dmma_launches, 1),
                    "how": f"{int(st['n_groups'])} stream groups overlap their kernels, so the time base is the device time of the "
                           "whole lockstep region of the timed steps (CUDA events on the launching stream), not a sum of launches",
                    "whole_step_tflops": agg["flops"] / (ms * 1e-3) * 1e-12 if world == 1 else None}
        if fam1 is not None:
            names = {"chol": "k_chol_update+k_chol_panel", "trtri": "k_trtri", "lauum": "k_lauum_trace"}
            roofline["families_single_stream"] = {
                names[k]: {"tflops": fam1["flops_" + k] / fam1["ms_" + k] * 1e-9, "ms": fam1["ms_" + k],
                           "share": fam1["ms_" + k] / fam1["ms_total"], "launches": int(fam1["launches_" + k])}
                for k in names}
            roofline["families_single_stream"]["other (k_build, k_fwd, k_alpha, k_finalize)"] = {
                "ms": fam1["ms_build"] + fam1["ms_fwd"] + fam1["ms_alpha"] + fam1["ms_finalize"],
                "share": (fam1["ms_build"] + fam1["ms_fwd"] + fam1["ms_alpha"] + fam1["ms_finalize"]) / fam1["ms_total"]}
            roofline["share_of_step"] = sum(fam1["ms_" + k] for k in names) / fam1["ms_total"]
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "how": "the timed loop itself: every step is one oi_gpr_day call with pinned host buffers + the NCCL gather of the rows"},
            "value_is": "the same timed steps without the host<->device copies: oi_stats.ms_gather + ms_total (CUDA events on the "
                        "library's launching stream, first gather kernel to last result), max over ranks per step",
            "gpu_launches": int(launches.item()),
            "nfev_mean": float(res["nfev"][h.get_owned()].mean() if dynamic else res["nfev"].mean()),
            "cells_per_rank_last_step": [int(v) for v in res["owned_counts"]] if dynamic else [len(p) for p in parts], "evals_per_step": agg["n_evals"] / args.steps,
            "iterations_per_step": agg["n_iterations"] / args.steps,
            "status_hist_rank0": status_hist, "finite_frac": float(np.isfinite(full[:, 0]).mean()),
            "per_rank_device_ms_per_step": [float(v) for v in rank_ms.tolist()],
            "limiter": ("FP64 DMMA issue rate; the ranks draw from one shared work list and end within one cell's run time of each other "
                        "(per_rank_device_ms_per_step)" if dynamic else
                        "the step ends with the slowest rank's optimiser tail: LPT balances n^3 but a cell costs n^3 x its evaluation "
                        "count (65...2200); see per_rank_device_ms_per_step") if world > 1 else
                       "FP64 DMMA issue rate in the bulk, dependent-launch latency of the last long optimiser runs in the tail",
            "roofline": roofline,
            "run_s": time.time() - T_START,
        }
        if full_day is not None:
            line["full_day"] = full_day
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dynamic:
        dist.barrier()
        h.set_shared_queue(None)
        if rank == 0:
            h.unlink_shared_queue(name[0])
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
