"""A small pass over every FFMA scan kernel (count, table-driven, histogram fallback, neighbours, screening edges) for
`compute-sanitizer --tool memcheck|synccheck|racecheck`: sizes of a few thousand frames so that the instrumented run ends in
minutes.  Results are compared with the oracle as usual.

    compute-sanitizer --tool synccheck python scripts/sanitize_small.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _oracle import Oracle
from clustering_b200 import density
from clustering_b200.synth import gaussian_mixture

o = Oracle()
for d, n, radii in ((5, 3000, [0.3]), (10, 2600, [0.2 * i for i in range(1, 21)]), (3, 2500, [0.1, 0.2, 0.3, 0.4, 0.5]),
                    (10, 2100, [1.0, 1.0000001, 1.0000002, 0.5])):
    x = gaussian_mixture(n, d, seed=77 + d)
    radii = np.asarray(radii, np.float32)
    r = density.density_run(x, radii, 0)
    assert np.array_equal(r["pops"], o.populations(x, radii)), (d, "pops")
    fe = o.free_energies(r["pops"][0])
    for a, b in zip(o.nearest_neighbors(x, fe), r["nn"]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (d, "nn")
    prev = None
    for t in (np.float32(0.6), np.float32(1.8)):
        lab = density.screening(fe, r["nn"][1], t, x, prev)
        prev_o = o.screening(fe, r["nn"][1], t, x, prev)
        assert np.array_equal(lab, prev_o.astype(np.uint32)), (d, "screening", float(t))
        prev = lab
    print("ok", d, n, len(radii), flush=True)
