"""Writes tests/golden/io/* with the UNMODIFIED reference writers and host functions (oracle/_ref/libdcref.so).
Run in the build container:  python tests/golden/make_golden_io.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _oracle import Ref  # noqa: E402
import test_cli_io as t  # noqa: E402
from clustering_b200.synth import gaussian_mixture  # noqa: E402


def main():
    ref = Ref()
    ref.set_threads(1)
    out = os.path.join(HERE, "io")
    os.makedirs(out, exist_ok=True)
    t.write_all(ref, out)
    x = gaussian_mixture(2500, 3, k=5, seed=19)
    pops = ref.populations(x, np.array([0.3], np.float32))[0]
    fe = ref.free_energies(pops)
    ni, nd, hi, hd = ref.nearest_neighbors(x, fe)
    lab = ref.screening(fe, ni, nd, np.float32(1.0), x, None)
    assigned = ref.assign_low_density_frames(lab, hi, hd, fe)
    named = ref.sorted_cluster_names(assigned)
    ties = np.repeat(np.arange(1, 41), 5).astype(np.uint32)
    np.random.default_rng(1).shuffle(ties)
    np.savez_compressed(os.path.join(out, "microstates.npz"), initial=lab.astype(np.uint32), hd_idx=hi.astype(np.uint32), fe=fe,
                        assigned=assigned.astype(np.uint32), named=named.astype(np.uint32), ties=ties,
                        ties_named=ref.sorted_cluster_names(ties).astype(np.uint32))
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
