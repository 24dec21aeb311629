"""Timeline of an OI_TRACE file (lockstep engine): bulk/express active cells, iterations and algorithmic TFLOP/s per time bucket."""
import sys
import numpy as np, pandas as pd
d = pd.read_csv(sys.argv[1]); d = d[d.iter != 'iter'].astype(float).sort_values('t_ms')
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 30
G = int(d.group.max()) + 1; nx = max(1, G // 4) if G >= 4 else 0
T = d.t_ms.max(); bins = np.linspace(0, T, nb + 1); d['b'] = np.digitize(d.t_ms, bins)
print('total ms', round(T), 'group iterations', len(d), 'groups', G)
for b, gp in d.groupby('b'):
    act = gp.groupby('group').active.mean(); fl = gp.flops_factor.sum(); ex = gp[gp.group >= G - nx]
    print(f"{int(bins[min(int(b) - 1, nb - 1)]):7d} ms  bulk active {int(act[act.index < G - nx].sum()):5d}  express active {int(act[act.index >= G - nx].sum()):3d}"
          f"  express its {len(ex):5d}  bulk its {len(gp) - len(ex):5d}  Nmax mean {gp.Nmax.mean():5.1f}  TFLOP/s {fl / (bins[1] - bins[0]) / 1e9:5.1f}")
