"""Runs each pair-scan kernel on a (possibly reduced) workload and prints per-stage CUDA-event times.
Used for `ncu --set full` captures and for quick per-config timings on the GPU box.

    python scripts/profile_kernels.py C3 [n_frames] [repeats]
"""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x = config_data(name, n)
d = x.shape[1]
s = Session(0)
stream = s.torch_stream()
radii = np.asarray(cfg["radii"], np.float32)
xd = torch.from_numpy(x).cuda()
torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    r = fn()
    e1.record(stream)
    e1.synchronize()
    return r, e0.elapsed_time(e1)


for it in range(reps):
    s.stats(reset=True)
    _, t_layout = timed(lambda: s.set_coords(xd))
    pp, t_pops = timed(lambda: s.populations(radii))
    st_p = s.stats(reset=True)
    pops = s.to_frame_order(pp)
    fe = s.free_energies(pops[cfg.get("fe_radius_index", 0)].contiguous())
    _, t_prep = timed(lambda: s.nn_prepare(fe))
    keys, t_nn = timed(lambda: s.nn_scan())
    st_n = s.stats(reset=True)
    nn = s.nn_finish(keys)
    s.sync()
    pairs = float(n) * n
    out = dict(workload=name, n=n, d=d, radii=len(radii), rep=it, layout_ms=t_layout, pops_ms=t_pops, nn_prepare_ms=t_prep, nn_ms=t_nn,
               pops_eval_frac=st_p["pairs_evaluated"] / pairs, nn_eval_frac=st_n["pairs_evaluated"] / pairs,
               pops_slow=st_p["slow_pairs"], pops_exact=st_p["exact_pairs"], nn_slow=st_n["slow_pairs"], nn_exact=st_n["exact_pairs"],
               pops_eff_gpd=pairs * d / t_pops / 1e6, nn_eff_gpd=pairs * d / t_nn / 1e6,
               pops_exec_tflops=2 * st_p["pairs_evaluated"] * d / t_pops / 1e9, nn_exec_tflops=2 * st_n["pairs_evaluated"] * d / t_nn / 1e9,
               max_pop=int(pops.max()))
    print(json.dumps(out), flush=True)
