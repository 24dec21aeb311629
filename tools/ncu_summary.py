"""Summarise an .ncu-rep (raw page) into one line per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_inst%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("launch__registers_per_thread", "regs"), ("launch__waves_per_multiprocessor", "waves"),
        ("smsp__cycles_active.avg", "smsp_act"), ("sm__cycles_elapsed.max", "cyc"),
        ("SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "dmma_cyc")]
print(",".join(c[1] for c in cols))
for r in rows[2:]:
    out = []
    for name, _ in cols:
        v = r[idx[name]] if name in idx else ""
        if name == "Kernel Name":
            v = v.split("(")[0]
        if name == "Grid Size":
            v = v.replace(",", "x").replace(" ", "")
        u = units[idx[name]] if name in idx else ""
        out.append(f"{v}{'' if u in ('', '%', 'cycle', 'block', 'register/thread') else ' ' + u}")
    print(",".join(out))
