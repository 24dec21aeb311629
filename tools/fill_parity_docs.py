"""Fill the parity placeholders of DESIGN.md / README.md from the reports tests/test_gpu_fit_parity.py and
tests/test_gpu_parity.py::test_fit_matches_oracle wrote into gpurun_out/ (and copy the reports into profiles/)."""
import json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = lambda n: json.load(open(os.path.join(ROOT, "gpurun_out", n)))
p1k, pl, pc, ps = g("parity_1k.json"), g("parity_lbfgs_1k.json"), g("parity_cfg5.json"), g("parity_small_day.json")
for n in ("parity_1k.json", "parity_lbfgs_1k.json", "parity_cfg5.json", "parity_small_day.json"):
    shutil.copy(os.path.join(ROOT, "gpurun_out", n), os.path.join(ROOT, "profiles", "r02_" + n))


def row(name, fl, gp):
    f = lambda s: (f"gate met by {100 * s['frac_gate']:.2f} % of {s['cells']} cells (abs Δfs ≤ 1 mm: {100 * s['frac_fs_1mm']:.2f} %, NLML rule: "
                   f"{100 * s['frac_nlml']:.2f} %; NaN only on one side: {s['nan_only_candidate'] + s['nan_only_reference']}, NaN on both: {s['nan_both']}; "
                   f"abs Δfs median {s['dfs_mm_median']:.1e} mm, p99 {s['dfs_mm_p99']:.2g} mm)")
    return f"| {name} | {f(fl)} | {f(gp)} |"


table = "\n".join([
    row(f"full day, {p1k['cells']} cells, n = {p1k['n_min']}…{p1k['n_max']} ({100 * p1k['frac_cells_n_gt_1100']:.0f} % with n > 1100)", p1k["reference_sorted_vs_tree"], p1k["gpu_vs_tree"]),
    row(f"small day, {ps['cells']} cells, n = 40…300", ps["reference_sorted_vs_tree"], ps["gpu_vs_tree"]),
    f"| config 5 (12.5 km / 500 km), {pc['cells']} cells, n = {min(pc['n'])}…{max(pc['n'])} | (one order only) | gate met by {pc['cells'] - int(round((1 - pc['frac_gate']) * pc['cells']))} of {pc['cells']}; abs Δfs max {pc['dfs_mm_max']:.1e} mm, likelihood within {max(abs(v) for v in pc['rel_lZ'] if v is not None):.0e} relative |"])
fl, gp = p1k["reference_sorted_vs_tree"], p1k["gpu_vs_tree"]
numbers = (f"`profiles/r02_parity_*.json`; the reference misses its own gate on {100 * (1 - fl['frac_gate']):.1f} % of the full-day cells, the CUDA path "
           f"misses the reference on {100 * (1 - gp['frac_gate']):.1f} %")
q = p1k["nonfinite_by_n_quartile"]
nan_text = ("all n classes on both sides — per quartile of n (" + ", ".join(f"{x['n_lo']}…{x['n_hi']}" for x in q) + "): CUDA path "
            + " / ".join(f"{100 * x['gpu']:.1f}" for x in q) + " %, reference " + " / ".join(f"{100 * x['ref_tree']:.1f}" for x in q)
            + " % (tree order) and " + " / ".join(f"{100 * x['ref_sorted']:.1f}" for x in q) + " % (sorted order); over the whole fixture "
            + f"{100 * (gp['nan_only_candidate'] + gp['nan_both']) / gp['cells']:.1f} % (CUDA) vs {100 * (gp['nan_only_reference'] + gp['nan_both']) / gp['cells']:.1f} % (reference): "
              "the NaN rate grows with n up to n ≈ 1300, which round 1's n ≤ 1100 sample could not show, and the day product's 4.0 % of holes "
              "(`full_day.status_hist`) is the reference's own rate.")
fast = (f"{100 * pl['frac_lZ_ge_ref']:.1f} % of the {pl['cells']} cells (strictly higher by > 1e-6 in {100 * pl['frac_lZ_strictly_higher_1e_6']:.0f} %), with "
        f"{pl['nfev_mean_fast']:.0f} evaluations per cell instead of {pl['nfev_mean_ref']:.0f} (max {pl['nfev_max_fast']}), {100 * pl['finite_fast']:.1f} % finite cells "
        f"(reference {100 * pl['finite_ref']:.1f} %), |Δfs| median {pl['dfs_mm_median']:.1e} mm but only {100 * pl['frac_fs_1mm']:.0f} % within 1 mm "
        f"(it does not stop where the reference's line searches give up); {pl['cells_per_s_fast']:.0f} vs {pl['cells_per_s_cg']:.0f} cells/s on that sample")
d = open(os.path.join(ROOT, "DESIGN.md")).read()
d = d.replace("PARITY_NUMBERS", numbers).replace("PARITY_TABLE", table).replace("NANRATE_TEXT", nan_text).replace("FASTMODE_NUMBERS", fast)
open(os.path.join(ROOT, "DESIGN.md"), "w").write(d)
r = open(os.path.join(ROOT, "README.md")).read()
r = r.replace("README_PARITY", f"the CUDA path meets the gate (|Δfs| ≤ 1 mm and NLML ≤ ref·(1+1e-6)) on {100 * gp['frac_gate']:.1f} % of {gp['cells']} cells; the reference "
              f"meets it against ITSELF (same cells, inputs permuted) on {100 * fl['frac_gate']:.1f} % — its stopping points are chaotic (DESIGN §2); median |Δfs| {gp['dfs_mm_median']:.0e} mm")
open(os.path.join(ROOT, "README.md"), "w").write(r)
print(table); print(numbers); print(nan_text); print(fast)
