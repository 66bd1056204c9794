"""Times the screening of config C4 (or a reduced N): the whole threshold loop of Density::main (density_clustering.cpp:
806-816) through ONE screening run (dcb200_screening_begin / _next / _end), next to the call-per-threshold form
(dcb200_screening with the previous labels) for a few thresholds.  DCB200_TRACE=1 adds the per-phase breakdown on stderr.

    python scripts/screening_timing.py [C4] [n_frames] [max_thresholds]
"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200 import density
from clustering_b200.synth import CONFIGS, config_data

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
n_thr = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
x = config_data(name, n)
r = density.density_run(x, np.asarray(cfg["radii"][:1], np.float32), 0)          # warm-up (contexts, buffers)
t0 = time.perf_counter()
r = density.density_run(x, np.asarray(cfg["radii"][:1], np.float32), 0)
t1 = time.perf_counter()
fe, nd = r["fe"], r["nn"][1]
print(json.dumps(dict(workload=name, n=n, density_run_ms=(t1 - t0) * 1e3, max_fe=float(fe.max()))), flush=True)

a = time.perf_counter()
run = density.ScreeningRun(fe, nd, x)
b = time.perf_counter()
print(json.dumps(dict(begin_ms=(b - a) * 1e3, what="sort of the free energies + gather + upload + layout, once per run")), flush=True)
t, step, k, per = np.float32(0.1), np.float32(0.1), 0, []
t_to = float(fe.max())
lab = None
while (t < np.float32(t_to - 0.01 + 0.1)) and not (np.float32(t_to + 0.01 + 0.1) < t) and k < n_thr:
    a = time.perf_counter()
    lab = run.next(t)
    b = time.perf_counter()
    per.append((float(t), (b - a) * 1e3, int((fe <= t).sum()), int(lab.max())))
    t = np.float32(t + step)
    k += 1
run.close()
for q in per[:3] + per[len(per) // 2:len(per) // 2 + 2] + per[-3:]:
    print(json.dumps(dict(t=q[0], ms=q[1], below=q[2], clusters=q[3])), flush=True)
tot = sum(q[1] for q in per)
print(json.dumps(dict(thresholds=k, loop_ms=tot, mean_ms=tot / max(k, 1), max_ms=max(q[1] for q in per),
                      labels_checksum=int((lab.astype(np.int64) * (np.arange(n) % 1000003 + 1)).sum()))), flush=True)

# the call-per-threshold form (what the reference's signature forces): every call sorts, gathers and uploads again
prev, t = None, np.float32(0.1)
for q in range(3):
    a = time.perf_counter()
    prev = density.screening(fe, nd, t, x, prev)
    b = time.perf_counter()
    print(json.dumps(dict(form="call per threshold", t=float(t), ms=(b - a) * 1e3)), flush=True)
    t = np.float32(t + step)
