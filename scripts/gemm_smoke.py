"""First-light check of the GEMM-form path: tiny inputs against the CPU oracle, with per-step prints."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _oracle import Oracle
from clustering_b200 import density
from clustering_b200.synth import gaussian_mixture, contact_like

o = Oracle()
for (n, d) in ((100, 32), (300, 32), (1300, 64), (1300, 128), (3000, 40)):
    x = contact_like(n, d, k=3, seed=n + d)
    dd = ((x[:200, None, :] - x[None, :200, :]) ** 2).sum(-1)
    r = float(np.sqrt(np.percentile(dd[dd > 0], 20)))
    radii = np.array([r, 0.8 * r], np.float32)
    t0 = time.time()
    pg = density.calculate_populations(x, radii)
    po = o.populations(x, radii)
    bad = int((pg != po).sum())
    print(f"n={n} d={d} r={r:.4f} pops mismatches={bad} of {po.size} max_pop={po.max()} ({time.time()-t0:.2f}s)", flush=True)
    if bad:
        idx = np.argwhere(pg != po)[:8]
        for a, b in idx:
            print("   radius", a, "frame", b, "got", pg[a, b], "want", po[a, b])
    fe = o.free_energies(po[0])
    a, b = o.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe)
    badn = [int((u.view(np.uint32) != v.view(np.uint32)).sum()) for u, v in zip(a, b)]
    print(f"   nn mismatches (idx, d2, hd idx, hd d2) = {badn}", flush=True)
