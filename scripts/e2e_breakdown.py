"""Wall-clock breakdown of the host-pointer C ABI calls (the e2e arm of bench.py) on one workload."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200 import density
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
x = config_data(name, n)
radii = np.asarray(cfg["radii"], np.float32)
for it in range(4):
    t0 = time.perf_counter(); pops = density.calculate_populations(x, radii)
    t1 = time.perf_counter(); fe = density.calculate_free_energies(pops[0])
    t2 = time.perf_counter(); nn = density.nearest_neighbors(x, fe)
    t3 = time.perf_counter()
    print(json.dumps(dict(workload=name, n=n, it=it, pops_ms=(t1 - t0) * 1e3, fe_ms=(t2 - t1) * 1e3, nn_ms=(t3 - t2) * 1e3)), flush=True)
