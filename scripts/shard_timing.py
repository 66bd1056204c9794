"""Times the block-cyclic shards of a multi-GPU density run on ONE GPU: the scans of a shard involve no communication, so
rank g of W runs exactly these launches.  Prints per-shard population and neighbour-scan times next to the unsharded scan.

    python scripts/shard_timing.py C3 8 [shards to time, default 0 and W/2]
"""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
which = [int(v) for v in sys.argv[3:]] or sorted({0, W // 2})
cfg = CONFIGS[name]
x = config_data(name)
n, d = x.shape
radii = np.asarray(cfg["radii"], np.float32)
s = Session(0)
stream = s.torch_stream()
xd = torch.from_numpy(x).cuda()
torch.cuda.synchronize()


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = fn()
        e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return r, best


s.set_coords(xd)
pp, t_full = timed(lambda: s.populations(radii))
pops = s.to_frame_order(pp)
fe = s.free_energies(pops[cfg.get("fe_radius_index", 0)].contiguous())
s.nn_prepare(fe)
_, t_nn_full = timed(lambda: s.nn_scan())
out = dict(workload=name, n=n, d=d, shards=W, pops_full_ms=t_full, nn_full_ms=t_nn_full, ideal_pops_ms=t_full / W, ideal_nn_ms=t_nn_full / W, per_shard=[])
for g in which:
    _, tp = timed(lambda: s.populations_shard(radii, g, W))
    s.nn_prepare(fe)
    _, tn = timed(lambda: s.nn_scan_shard(g, W))
    out["per_shard"].append(dict(shard=g, pops_ms=tp, nn_ms=tn, pops_eff=t_full / W / tp, nn_eff=t_nn_full / W / tn))
print(json.dumps(out), flush=True)
