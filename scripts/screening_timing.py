"""Times the screening loop of config C4 (or a reduced N): per-threshold wall time through the host-pointer C ABI
(dcb200_screening = host bookkeeping + GPU pair scan), like Density::main's loop (density_clustering.cpp:806-816)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200 import density
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
n_thr = int(sys.argv[3]) if len(sys.argv) > 3 else 12
x = config_data(name, n)
t0 = time.perf_counter()
pops = density.calculate_populations(x, np.asarray(cfg["radii"][:1], np.float32))[0]
fe = density.calculate_free_energies(pops)
ni, nd, hi, hd = density.nearest_neighbors(x, fe)
t1 = time.perf_counter()
print(json.dumps(dict(workload=name, n=n, pops_fe_nn_ms=(t1 - t0) * 1e3, max_fe=float(fe.max()))), flush=True)
prev = None
t = np.float32(0.1)
step = np.float32(0.1)
k = 0
total = 0.0
while t < fe.max() + 0.1 and k < n_thr:
    a = time.perf_counter()
    lab = density.screening(fe, nd, t, x, prev)
    b = time.perf_counter()
    total += b - a
    print(json.dumps(dict(t=float(t), ms=(b - a) * 1e3, below=int((fe <= t).sum()), clusters=int(lab.max()))), flush=True)
    prev = lab
    t = np.float32(t + step)
    k += 1
print(json.dumps(dict(thresholds=k, total_ms=total * 1e3)))
