"""Aggregate warp-stall samples per reason and list the hottest SASS instructions of one kernel in an .ncu-rep."""
import csv, collections, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(r)
b = blocks[which]; hdr = b["rows"][0]; data = [r for r in b["rows"][1:] if len(r) >= len(hdr)]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s_ in stalls:
        try: tot[s_] += float(r[idx[s_]])
        except ValueError: pass
T = sum(tot.values())
print(b["name"], "block", which, "of", len(blocks), "total samples", T)
for k, v in tot.most_common(8): print(f"  {k:26s} {v:9.0f} {v / T:.3f}")
top = sorted(data, key=lambda r: -float(r[idx["# Samples"]] or 0))[:int(sys.argv[4]) if len(sys.argv) > 4 else 14]
for r in top:
    st = {s_: float(r[idx[s_]]) for s_ in stalls if float(r[idx[s_]] or 0) > 0}
    big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[idx["# Samples"]].rjust(7), r[idx["Source"]][:58].ljust(58), big)
