"""Mirror of the reference's density operators on numpy arrays, bound to libdcb200.so.

Same names, argument meaning and result conventions as Clustering::Density (reference
density_clustering.hpp / density_clustering_cuda.hpp), so that the parity tests read like tests of the
reference: populations are keyed by radius, neighbourhoods carry (index, squared distance), the
"no neighbour" entry is (n_rows+1, FLT_MAX), labels are 0 for unassigned frames.
"""
import ctypes as C

import numpy as np

from . import lib


def _coords(coords):
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    if coords.ndim != 2:
        raise ValueError("coords must be [n_rows][n_cols]")
    return coords


def calculate_populations(coords, radii):
    """-> uint32 [n_radii][n_rows], row r for radii[r]  (reference: density_clustering.cpp:126-195)."""
    coords = _coords(coords)
    radii = np.ascontiguousarray(np.atleast_1d(radii), dtype=np.float32)
    n, d = coords.shape
    pops = np.empty((radii.size, n), np.uint32)
    lib.check(lib.load().dcb200_populations(coords, n, d, radii, radii.size, pops))
    return pops


def calculate_free_energies(pops):
    """-> float32 [n_rows]  (reference: density_clustering.cpp:197-212)."""
    pops = np.ascontiguousarray(pops, dtype=np.uint32)
    fe = np.empty(pops.size, np.float32)
    lib.check(lib.load().dcb200_free_energies(pops, pops.size, fe))
    return fe


def nearest_neighbors(coords, free_energy):
    """-> (nn_idx, nn_d2, hd_idx, hd_d2)  (reference: density_clustering.cpp:230-288)."""
    coords = _coords(coords)
    fe = np.ascontiguousarray(free_energy, dtype=np.float32)
    n, d = coords.shape
    if fe.size != n:
        raise ValueError("free_energy must have n_rows entries")
    ni = np.empty(n, np.uint32); nd = np.empty(n, np.float32)
    hi = np.empty(n, np.uint32); hd = np.empty(n, np.float32)
    lib.check(lib.load().dcb200_nearest_neighbors(coords, n, d, fe, ni, nd, hi, hd))
    return ni, nd, hi, hd


def sorted_free_energies(free_energy):
    """-> order[k] = frame at sorted position k  (reference: density_clustering.cpp:214-228)."""
    fe = np.ascontiguousarray(free_energy, dtype=np.float32)
    order = np.empty(fe.size, np.uint32)
    lib.check(lib.load().dcb200_sorted_free_energies(fe, fe.size, order))
    return order


def compute_sigma2(nn_d2):
    nn_d2 = np.ascontiguousarray(nn_d2, dtype=np.float32)
    s = C.c_double(0.0)
    lib.check(lib.load().dcb200_sigma2(nn_d2, nn_d2.size, C.byref(s)))
    return s.value


def screening(free_energy, nn_d2, free_energy_threshold, coords, initial_clusters=None):
    """One screening threshold -> uint32 labels [n_rows]  (reference: density_clustering_common.cpp:37-134)."""
    coords = _coords(coords)
    fe = np.ascontiguousarray(free_energy, dtype=np.float32)
    nn_d2 = np.ascontiguousarray(nn_d2, dtype=np.float32)
    n, d = coords.shape
    labels = np.empty(n, np.uint32)
    init = None
    if initial_clusters is not None and len(initial_clusters) == n:
        init_arr = np.ascontiguousarray(initial_clusters, dtype=np.uint32)
        init = init_arr.ctypes.data
    lib.check(lib.load().dcb200_screening(fe, nn_d2, np.float32(free_energy_threshold), coords, n, d, init, labels))
    return labels


def assign_low_density_frames(initial_clustering, hd_idx, free_energy):
    """-> uint32 states [n_rows]  (reference: density_clustering.cpp:345-360)."""
    init = np.ascontiguousarray(initial_clustering, dtype=np.uint32)
    hd = np.ascontiguousarray(hd_idx, dtype=np.uint32)
    fe = np.ascontiguousarray(free_energy, dtype=np.float32)
    out = np.empty(init.size, np.uint32)
    lib.check(lib.load().dcb200_assign_low_density_frames(init, hd, fe, init.size, out))
    return out


def sorted_cluster_names(clustering):
    """-> uint32 [n_rows], states renamed 1..K by decreasing population  (reference: density_clustering.cpp:458-493)."""
    st = np.ascontiguousarray(clustering, dtype=np.uint32)
    out = np.empty(st.size, np.uint32)
    lib.check(lib.load().dcb200_sorted_cluster_names(st, st.size, out))
    return out


def screening_step(sorted_coords, m_prev, m_new, max_dist2, comp):
    sorted_coords = _coords(sorted_coords)
    comp = np.ascontiguousarray(comp, dtype=np.uint32)
    lib.check(lib.load().dcb200_screening_step(sorted_coords, sorted_coords.shape[1], m_prev, m_new,
                                                np.float32(max_dist2), comp))
    return comp


def density_run(coords, radii, fe_radius_index=0, neighbors=True, all_free_energies=False, out=None):
    """One density run through dcb200_density_run (one upload, one layout build): populations for all radii, the free
    energies of radii[fe_radius_index] and the neighbour search on them.
    -> dict(pops [R][n], fe [n], fe_all [R][n] or None, nn = (nn_idx, nn_d2, hd_idx, hd_d2) or None).
    out: optional dict of preallocated arrays with the same keys (e.g. views of pinned memory)."""
    coords = _coords(coords)
    radii = np.ascontiguousarray(np.atleast_1d(radii), dtype=np.float32)
    n, d = coords.shape
    out = out or {}
    pops = out.get("pops")
    if pops is None:
        pops = np.empty((radii.size, n), np.uint32)
    fe = out.get("fe")
    if fe is None:
        fe = np.empty(n, np.float32)
    fe_all = out.get("fe_all")
    if fe_all is None and all_free_energies:
        fe_all = np.empty((radii.size, n), np.float32)
    nn = out.get("nn")
    if nn is None and neighbors:
        nn = (np.empty(n, np.uint32), np.empty(n, np.float32), np.empty(n, np.uint32), np.empty(n, np.float32))
    p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    nnp = [p(a) for a in nn] if nn is not None else [None] * 4
    lib.check(lib.load().dcb200_density_run(C.c_void_p(coords.ctypes.data), n, d, radii, radii.size, int(fe_radius_index), p(pops),
                                            p(fe_all), p(fe), *nnp))
    return dict(pops=pops, fe=fe, fe_all=fe_all, nn=nn)


class ScreeningRun:
    """All thresholds of one screening run (dcb200_screening_begin / _next / _end): the free energies are sorted once and
    the sorted coordinates stay on the device(s); thresholds must not decrease."""

    def __init__(self, free_energy, nn_d2, coords):
        coords = _coords(coords)
        self.n, d = coords.shape
        fe = np.ascontiguousarray(free_energy, dtype=np.float32)
        nn_d2 = np.ascontiguousarray(nn_d2, dtype=np.float32)
        self.h = C.c_void_p()
        lib.check(lib.load().dcb200_screening_begin(fe, nn_d2, coords, self.n, d, C.byref(self.h)))

    def next(self, threshold, out=None):
        """labels (frame order) at `threshold`; out: optional uint32 [n] array to fill (e.g. pinned memory, reused over the
        thresholds like the command-line driver reuses its label vector -- a fresh 20 MB array per call costs more in page
        faults than the whole threshold on the device)."""
        labels = np.empty(self.n, np.uint32) if out is None else out
        assert labels.dtype == np.uint32 and labels.size == self.n and labels.flags["C_CONTIGUOUS"]
        lib.check(lib.load().dcb200_screening_next(self.h, np.float32(threshold), labels))
        return labels

    def close(self):
        if self.h:
            lib.load().dcb200_screening_end(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
