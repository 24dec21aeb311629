"""Summarise an ncu --csv log holding gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum per launch
(one NLML+gradient evaluation of a stripe, tools/eval_bench.py): per kernel launches / time / DRAM bytes, and the
per-launch average of the FP64 DMMA tile kernels that bench.py reports as roofline.traffic."""
import csv, collections, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
hdr_units = {}
acc = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").split("<")[0]
    metric, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    acc[name][metric] += val * scale
    if metric == "gpu__time_duration.sum":
        acc[name]["launches"] += 1
fam = ("k_chol_update", "k_chol_panel", "k_trtri", "k_lauum_trace")
out = {"per_kernel": {}, "source": sys.argv[1]}
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    out["per_kernel"][k] = {"launches": int(v["launches"]), "us": v["gpu__time_duration.sum"],
                            "dram_read_bytes": v["dram__bytes_read.sum"], "dram_write_bytes": v["dram__bytes_write.sum"]}
L = sum(acc[k]["launches"] for k in fam)
B = sum(acc[k]["dram__bytes_read.sum"] + acc[k]["dram__bytes_write.sum"] for k in fam)
T = sum(acc[k]["gpu__time_duration.sum"] for k in fam)
out["dmma_tile_kernels"] = B / max(L, 1)
out["dmma_tile_kernels_detail"] = {"launches": int(L), "dram_bytes_total": B, "us_total": T, "dram_GBps": B / max(T, 1e-9) * 1e-3}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["dmma_tile_kernels_detail"]))
for k, v in out["per_kernel"].items():
    print(f"{k:16s} {v['launches']:4d} launches {v['us']:10.1f} us  rd {v['dram_read_bytes'] / 1e9:7.3f} GB  wr {v['dram_write_bytes'] / 1e9:7.3f} GB")
