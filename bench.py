#!/usr/bin/env python
"""bench.py -- GP cells/sec (fit + predict) on the synthetic 25 km pan-Arctic day (BASELINE.json).

A step = one pass of the hot path (neighbour gather -> lockstep CG fit -> posterior) over one batch
of cells: ``2 x --gpus N`` stripes of the day (stripe s = every 16th ice cell starting at s, so each
stripe has the day's n-histogram; ~1195 cells per stripe, 2 stripes = 1/8 day per GPU and step,
per-GPU work fixed => weak scaling; at N=8 one step is the whole day sharded over the 8 GPUs =
BASELINE.json configs[2]).  At N=1 the whole day (configs[1], ~2.5 min) is additionally run ONCE through
the same ABI call and reported under "full_day"; it is too long to be the repeated step.  Cells of a step are sharded over ranks by LPT on n^3
(optimalinterpolation_b200/shard.py); the only collective is the final gather of the result rows.

  value : cells/s with observations + cell coordinates already resident in HBM (timed: gather +
          fit + predict + result gather), CUDA events on the launching stream, max over ranks
  e2e   : the same through the single C-ABI call oi_gpr_day with pinned HOST buffers
          (H2D of the inputs and D2H of the results inside the timed region)
  --impl reference : the reference's CPU path (oracle port, see oracle/) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before torch initialises CUDA: one hardware queue per stream group

METRIC = "GP cells/sec (fit+predict), 25km Arctic day"
N_STRIPES = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--stripes-per-gpu", type=int, default=2)
    ap.add_argument("--max-active", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-day", action="store_true", help="skip the single whole-day pass (N=1 only, ~2.5 min)")
    ap.add_argument("--no-family-pass", action="store_true", help="skip the single-stream pass that times each kernel family")
    ap.add_argument("--groups", type=int, default=0, help="lockstep stream groups (0 = library default)")
    ap.add_argument("--cpu-frac", type=float, default=0.2, help="cheapest fraction of cells the CPU sample is drawn from")
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


def step_cells(day, n_stripes_in_step):
    nc = len(day.X)
    return np.sort(np.concatenate([np.arange(s, nc, N_STRIPES) for s in range(min(n_stripes_in_step, N_STRIPES))]))


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    import warnings
    warnings.simplefilter("ignore")
    from optimalinterpolation_b200.synthetic import make_day
    from oracle import cpu_baseline
    from scipy.spatial import cKDTree
    day = make_day()
    cells = step_cells(day, args.gpus * args.stripes_per_gpu)
    counts = np.asarray(cKDTree(np.c_[day.x_train, day.y_train]).query_ball_point(
        day.X, r=day.radius_km * 1000.0, return_length=True))
    cores = os.cpu_count() or 1
    # evaluations per cell: full scipy-CG fits of the cheapest cells (nfev hardly depends on n: 158 +- 40 over n = 169..1100
    # in tests/golden/day_fit_sample_large.npz), untimed
    rf = cpu_baseline.run_sample(day, counts, cells, cores=cores, frac=0.02)
    nfev_mean = rf["nfev_mean"]
    vals, wall = [], []
    for s in range(args.warmup + args.steps):
        r = cpu_baseline.run_eval_sample(day, counts, cells, nfev=nfev_mean, cores=cores)
        if s >= args.warmup:
            vals.append(r["value"]); wall.append(r["wall_s"]); last = r
    v = float(np.mean(vals))
    sample = (f"per step one SMLII evaluation timed on {last['n_sample']} cells at evenly spaced quantiles of the step's n-distribution "
              f"(n {last['n_min']}..{last['n_max']}), one process per core, 1 BLAS thread each; t(n) ~ n^{last['exponent']:.2f}; cost of the "
              f"step = sum over its {len(cells)} cells of (nfev + 1/3) * t(n), nfev {nfev_mean:.0f} = mean of full scipy-CG fits of the "
              f"{rf['n_sample']} cheapest cells; {last['core_hours']:.1f} core-hours for the step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(wall)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.gpus * args.stripes_per_gpu}/16 stripes of the synthetic 25 km pan-Arctic day "
                               f"({len(cells)} cells); CPU arm times a bounded sample per step"},
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    from optimalinterpolation_b200.synthetic import make_day
    from optimalinterpolation_b200.shard import lpt_partition, gather_results, imbalance

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    day = make_day()
    cells = step_cells(day, world * args.stripes_per_gpu)
    Xstep = day.X[cells]
    h = oi.Handle(local_rank)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)

    # pinned host staging of the inputs (e2e path copies from these every step)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    px, py, pt, pz = map(pinned, (day.x_train, day.y_train, day.t_train, day.z))

    # planning (untimed, deterministic): counts of the step's cells -> LPT shard
    h.set_observations(px, py, pt, pz)
    h.set_cells(Xstep)
    counts_step = h.gather_neighbours(day.radius_km * 1000.0).copy()
    parts = lpt_partition(counts_step, world)
    mine = parts[rank]
    Xmine = pinned(Xstep[mine])
    params = h.make_params(day.radius_km * 1000.0, day.T_mid, day.mean, day.x0, mode=0, max_active=args.max_active,
                           n_groups=args.groups)

    def step_resident():
        h.gather_neighbours(day.radius_km * 1000.0)
        h.run(params)
        res = h.get_results()
        st = h.stats()
        full = gather_results(res["out"], mine, len(cells), parts)
        return res, st, full

    def step_e2e():
        res = h.gpr_day(px, py, pt, pz, Xmine, params)
        st = h.stats()
        full = gather_results(res["out"], mine, len(cells), parts)
        return res, st, full

    # ---------------- device-resident arm ----------------
    h.set_cells(Xmine)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    agg = {}
    for _ in range(args.steps):
        res, st, full = step_resident()
        for k, v in st.items():
            if not isinstance(v, list):
                agg[k] = agg.get(k, 0) + v
    e1.record(stream); barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    tt = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
    launches = torch.tensor([agg["n_launches"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms, wall_ms = float(tt[0]), float(tt[1])
    value = len(cells) * args.steps / (ms * 1e-3)

    # ---------------- end-to-end arm (host buffers through the one ABI call) ----------------
    step_e2e()
    barrier()
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(args.steps):
        res_e, st_e, full_e = step_e2e()
    e1.record(stream); barrier()
    tt = torch.tensor([e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = len(cells) * args.steps / (max(float(tt[0]), float(tt[1])) * 1e-3)
    h2d = 4 * day.z.size * 8 + Xmine.size * 8
    d2h = len(mine) * (64 + 12)
    same = bool(np.array_equal(full, full_e, equal_nan=True))

    # ---------------- single-stream pass: device time of each kernel family (diagnostic, untimed) ----------------
    # With several stream groups the families of different groups overlap, so their per-stream times do not
    # add up; one extra step with n_groups=1 gives the per-family numbers that the ncu launch list is compared to.
    fam1 = None
    if rank == 0 and not args.no_family_pass:
        p1 = h.make_params(day.radius_km * 1000.0, day.T_mid, day.mean, day.x0, mode=0, max_active=args.max_active, n_groups=1)
        h.set_cells(Xmine); h.gather_neighbours(day.radius_km * 1000.0); h.run(p1)
        fam1 = h.stats()
    # ---------------- the whole day in one call on one GPU (BASELINE.json configs[1]) ----------------
    full_day = None
    if world == 1 and not args.no_full_day:
        Xday = pinned(day.X)
        t0 = time.perf_counter(); e0.record(stream)
        res_d = h.gpr_day(px, py, pt, pz, Xday, params)
        e1.record(stream); torch.cuda.synchronize()
        sec = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
        st_d = h.stats()
        full_day = {"cells": int(len(day.X)), "seconds": sec, "cells_per_s": len(day.X) / sec,
                    "tflops": st_d["flops"] / sec * 1e-12, "evals": int(st_d["n_evals"]), "iterations": int(st_d["n_iterations"]),
                    "finite_frac": float(np.isfinite(res_d["out"][:, 0]).mean()),
                    "how": "one oi_gpr_day call with host buffers (H2D + gather + fit + predict + D2H), single pass, not part of the timed steps"}
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
        # k_scale_rows, k_trtri, k_lauum_trace).  achieved = their algorithmic flops / the device time of the whole
        # lockstep region of the timed steps (a lower bound: build, substitutions, finalize and host gaps included).
        ach = agg["flops_factor"] / agg["ms_total"] * 1e-9
        dmma_launches = agg["launches_chol"] + agg["launches_trtri"] + agg["launches_lauum"]
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("dmma_tile_kernels")
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "FP64 DMMA tile kernels (k_chol_update+k_chol_panel+k_scale_rows+k_trtri+k_lauum_trace)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA issue-rate microbenchmark: 37.1 TFLOP/s, profiles/r01_fp64_peak_microbench.txt)",
                    "flops_per_launch": agg["flops_factor"] / max(dmma_launches, 1),
                    "ms_per_launch": agg["ms_total"] / max(dmma_launches, 1),
                    "how": f"{int(st['n_groups'])} stream groups overlap their kernels, so the time base is the device time of the "
                           "whole lockstep region of the timed steps (CUDA events on the launching stream), not a sum of launches",
                    "whole_step_tflops": agg["flops"] / (ms * 1e-3) * 1e-12 if world == 1 else None}
        if fam1 is not None:
            names = {"chol": "k_chol_update+k_chol_panel+k_scale_rows", "trtri": "k_trtri", "lauum": "k_lauum_trace"}
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
            "config": {"workload": f"{world * args.stripes_per_gpu}/16 stripes of the synthetic 25 km pan-Arctic day "
                                   f"(SURVEY.md 8d: 320x320 lattice, 19109 ice cells, 38144 obs, r=300 km, 9 days)",
                       "cells_per_step": int(len(cells)), "n_obs": int(day.z.size),
                       "n_min_median_max": [int(counts_step.min()), int(np.median(counts_step)), int(counts_step.max())],
                       "optimiser": "scipy-CG restatement (reference gradient convention), x0 as GPR_CS2S3.py:217",
                       "sharding": f"LPT on n^3 over {world} ranks, imbalance {imbalance(counts_step, parts):.4f}",
                       "cache": "per-iteration working set (sum of n_pad^2*8 B over active cells, GBs) is far larger than the 126 MB L2; no L2 flush needed"},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "identical_to_resident_arm": same},
            "gpu_launches": int(launches.item()),
            "wall_ms_per_step": wall_ms / args.steps,
            "nfev_mean": float(res["nfev"].mean()), "evals_per_step": agg["n_evals"] / args.steps,
            "iterations_per_step": agg["n_iterations"] / args.steps,
            "roofline": roofline,
        }
        if full_day is not None:
            line["full_day"] = full_day
        if world == 1 and not args.no_cpu_baseline:
            import warnings
            warnings.simplefilter("ignore")
            from oracle import cpu_baseline
            cores = os.cpu_count() or 1
            cf = counts_full(day, h, cells, counts_step)
            nf = np.zeros(len(cells)); nf[mine] = res["nfev"]            # evaluations per cell as measured in the GPU run
            r = cpu_baseline.run_eval_sample(day, cf, cells, nfev=nf, cores=cores)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": "cells/s", "cores": cores, "kind": "port",
                "sample": (f"one SMLII evaluation (the reference spends >99% of its time there) timed on {r['n_sample']} cells at evenly "
                           f"spaced quantiles of the step's n-distribution (n {r['n_min']}..{r['n_max']}), one process per core with 1 BLAS "
                           f"thread, {r['wall_s']:.1f} s wall; power-law fit t(n) ~ n^{r['exponent']:.2f} ({r['t_eval_median_n'] * 1e3:.0f} ms at the "
                           f"median n); cost of the step = sum over its {len(cells)} cells of (nfev + 1/3) * t(n) with the measured mean "
                           f"nfev {r['nfev_mean']:.0f} = {r['core_hours']:.1f} core-hours")}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def counts_full(day, h, cells, counts_step):
    """neighbour counts indexed by the day's cell index (only the step's cells are filled)."""
    c = np.zeros(len(day.X), dtype=np.int64)
    c[cells] = counts_step
    return c


if __name__ == "__main__":
    main()
