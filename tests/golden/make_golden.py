"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libdcref.so, compiled
from /root/reference/src by oracle/Makefile).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these fixtures -- outputs of
the reference's own CPU path with OMP_NUM_THREADS=1 -- are what pins the oracle and the CUDA path
on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _oracle import Ref  # noqa: E402
from clustering_b200.synth import gaussian_mixture, contact_like  # noqa: E402

CASES = {
    # name: (coords factory, radii, screening radius index, n thresholds)
    "d1_n600": (lambda: gaussian_mixture(600, 1, k=4, seed=11), [0.05, 0.2], 0),
    "d2_n1200": (lambda: gaussian_mixture(1200, 2, k=6, seed=12), [0.1, 0.3, 0.2], 0),
    "d3_n1500": (lambda: gaussian_mixture(1500, 3, k=8, seed=13), [0.15, 0.4], 0),
    "d5_n2000": (lambda: gaussian_mixture(2000, 5, k=12, seed=14), [0.1, 0.2, 0.3, 0.4, 0.5], 2),
    "d7_n1000": (lambda: gaussian_mixture(1000, 7, k=5, seed=15), [0.6, 0.9], 0),
    "d10_n1500": (lambda: gaussian_mixture(1500, 10, k=12, seed=16), [round(0.1 * i, 1) for i in range(1, 21)], 7),
    "d128_n600": (lambda: contact_like(600, 128, k=6, seed=17), [1.0, 0.8], 0),
}


def with_duplicates(x):
    x = x.copy()
    x[7] = x[3]
    x[len(x) // 2] = x[3]
    x[len(x) - 1] = x[0]
    return x


def main():
    ref = Ref()
    ref.set_threads(1)  # latent std::map race in the reference's parallel region (density_clustering.cpp:158,166)
    for name, (make, radii, r_scr) in CASES.items():
        x = with_duplicates(make())
        radii = np.asarray(radii, np.float32)
        pops = ref.populations(x, radii)
        fe_all = np.stack([ref.free_energies(p) for p in pops])
        fe = fe_all[r_scr]
        ni, nd, hi, hd = ref.nearest_neighbors(x, fe)
        order = ref.sorted_free_energies(fe)
        sigma2 = ref.sigma2(ni, nd)
        thresholds, labels = [], []
        t, step, prev = np.float32(0.1), np.float32(0.1), None
        t_to = fe.max()
        while (t < t_to - step / np.float32(10) + step) and not (t_to + step / np.float32(10) + step < t):
            lab = ref.screening(fe, ni, nd, t, x, prev)
            thresholds.append(t); labels.append(lab.astype(np.uint32)); prev = lab
            t = np.float32(t + step)
        final = ref.sorted_cluster_names(ref.assign_low_density_frames(labels[len(labels) // 2].astype(np.uint64), hi, hd, fe))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), coords=x, radii=radii, pops=pops.astype(np.uint32), fe=fe_all,
            r_scr=np.int64(r_scr), nn_idx=ni.astype(np.uint32), nn_d2=nd, hd_idx=hi.astype(np.uint32), hd_d2=hd,
            order=order.astype(np.uint32), sigma2=np.float64(sigma2), thresholds=np.asarray(thresholds, np.float32),
            labels=np.stack(labels), microstates_from_mid=final.astype(np.uint32))
        print(name, x.shape, "thresholds", len(thresholds), "clusters", int(labels[-1].max()))


if __name__ == "__main__":
    main()
