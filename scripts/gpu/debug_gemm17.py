"""debug: d = 17 multi-radius populations on the GEMM-form path vs the oracle, under several knobs"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
if len(sys.argv) > 1:
    from _oracle import Oracle
    from clustering_b200 import density
    from clustering_b200.synth import gaussian_mixture
    d, n = 17, 1300
    x = gaussian_mixture(n, d, k=4, seed=700 + d)
    x[17] = x[3]; x[n - 1] = x[3]
    dd = ((x[:300, None, :] - x[None, :300, :]) ** 2).sum(-1)
    r_in = float(np.sqrt(np.percentile(dd[dd > 0], 20)))
    radii = np.linspace(0.3 * r_in, 3.0 * r_in, 9).astype(np.float32)
    o = Oracle()
    po = o.populations(x, radii)
    pg = density.calculate_populations(x, radii)
    diff = np.argwhere(po != pg)
    print(sys.argv[1], "differing entries:", len(diff), [(int(r), int(i), int(po[r, i]), int(pg[r, i])) for r, i in diff[:12]])
    for k in range(len(radii)):
        p1 = density.calculate_populations(x, radii[k:k + 1])
        if not np.array_equal(p1[0], po[k]):
            print("   single radius", k, "differs in", int((p1[0] != po[k]).sum()))
else:
    for name, env in (("default", {}), ("order6", {"DCB200_ORDER_DIMS": "6"}), ("ffma", {"DCB200_GEMM": "0"}), ("ra1", {"DCB200_GEMM_RA": "1"}),
                      ("order6+ra1", {"DCB200_ORDER_DIMS": "6", "DCB200_GEMM_RA": "1"})):
        subprocess.run([sys.executable, __file__, name], env=dict(os.environ, **env))
