"""Runs each pair-scan kernel once on a reduced workload (for `ncu --set full` captures)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
cfg = CONFIGS[name]
x = config_data(name, n)
s = Session(0)
s.set_coords(x)
radii = np.asarray(cfg["radii"], np.float32)
for _ in range(2):
    pops = s.to_frame_order(s.populations(radii))
    fe = s.free_energies(pops[0].contiguous())
    nn = s.nearest_neighbors(fe)
s.sync()
st = s.stats()
print(name, n, st, int(pops.max()), 'evaluated/scheduled = %.4f' % (st['pairs_evaluated'] / max(1, st['pairs_scheduled'])))
