"""Times the population and neighbour scans of a workload under a list of environment knob settings (read by the library
at every launch), in one process: the layout is built once, each setting is timed as the best of three launches.

    python scripts/sweep_knobs.py C3 "DCB200_BIN_STEAL=0" "DCB200_BIN_DENSE_LANES=4" "DCB200_ITEMS_PER_CTA=96,DCB200_BIN_STEAL=1"
"""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1]
settings = [""] + sys.argv[2:]
cfg = CONFIGS[name]
x = config_data(name)
n, d = x.shape
radii = np.asarray(cfg["radii"], np.float32)
s = Session(0)
stream = s.torch_stream()
xd = torch.from_numpy(x).cuda()
torch.cuda.synchronize()
s.set_coords(xd)


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


ref = None
for st in settings:
    kv = dict(p.split("=") for p in st.split(",") if p)
    for k, v in kv.items():
        os.environ[k] = v
    s.stats(reset=True)
    pp, tp = timed(lambda: s.populations(radii))
    sp = s.stats(reset=True)
    pops = s.to_frame_order(pp)
    if ref is None:
        ref = pops.clone()
    same = bool(torch.equal(ref, pops))
    fe = s.free_energies(pops[cfg.get("fe_radius_index", 0)].contiguous())
    s.nn_prepare(fe)
    _, tn = timed(lambda: s.nn_scan())
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("DCB200_LIB", "libdcb200.so")), setting=st or "default", workload=name,
                          pops_ms=round(tp, 2), nn_ms=round(tn, 2), pops_eval_frac=round(sp["pairs_evaluated"] / 3 / (float(n) * n), 4),
                          pops_exact=sp["exact_pairs"] // 3, same_pops_as_default=same)), flush=True)
    for k in kv:
        del os.environ[k]
