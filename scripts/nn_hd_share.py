"""How much of the neighbour scan is the search for the neighbour with LOWER free energy?  Times the scan with the real free
energies and with a constant free energy (no frame has a lower one: the scan only looks for nearest neighbours -- what the
lumping-radius pre-pass runs), and reports the share of the pair matrix each evaluates.

    python scripts/nn_hd_share.py C3
"""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = CONFIGS[name]
x = config_data(name)
n, d = x.shape
radii = np.asarray(cfg["radii"], np.float32)
s = Session(0)
stream = s.torch_stream()
xd = torch.from_numpy(x).cuda()
torch.cuda.synchronize()
s.set_coords(xd)
pops = s.to_frame_order(s.populations(radii))
fe = s.free_energies(pops[cfg.get("fe_radius_index", 0)].contiguous())


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


out = dict(workload=name)
for label, f in (("real_fe", fe), ("constant_fe", torch.zeros_like(fe))):
    s.nn_prepare(f)
    s.stats(reset=True)
    keys, t = timed(lambda: s.nn_scan())
    st = s.stats(reset=True)
    nn = s.nn_finish(keys)
    s.sync()
    has_hd = int((nn[2].to(torch.int64) <= n).sum().item())
    out[label] = dict(nn_ms=round(t, 2), eval_frac=round(st["pairs_evaluated"] / 3 / (float(n) * n), 4), exact=st["exact_pairs"] // 3,
                      rows_with_lower_fe_neighbour=has_hd)
# how far the lower-free-energy neighbours are, relative to the nearest neighbours (real free energies)
s.nn_prepare(fe)
nn = s.nn_finish(s.nn_scan())
s.sync()
nd, hd = nn[1].cpu().numpy(), nn[3].cpu().numpy()
ok = hd < 3e38
ratio = np.sqrt(hd[ok] / np.maximum(nd[ok], 1e-30))
out["hd_over_nn_distance"] = {f"p{q}": round(float(np.percentile(ratio, q)), 2) for q in (50, 90, 99, 99.9)}
out["hd_distance"] = {f"p{q}": round(float(np.percentile(np.sqrt(hd[ok]), q)), 3) for q in (50, 90, 99, 99.9, 100)}
out["nn_distance"] = {f"p{q}": round(float(np.percentile(np.sqrt(nd), q)), 3) for q in (50, 90, 99, 100)}
print(json.dumps(out), flush=True)
