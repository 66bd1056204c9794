"""ctypes bindings of the parity checkers (TEST INFRASTRUCTURE, see oracle/dc_oracle.c).

  Oracle  -> oracle/libdcoracle.so       plain-C restatement, available everywhere (GPU box too)
  Ref     -> oracle/_ref/libdcref.so     the unmodified reference, compiled from /root/reference/src
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libdcoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdcref.so")
REFCUDA_SO = os.path.join(ROOT, "oracle", "_ref", "libdcrefcuda.so")

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_sz = C.c_uint64


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref():
    """Returns the path of libdcref.so, building it when the reference sources are present."""
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    def __init__(self):
        L = C.CDLL(build_oracle())
        L.dco_dist2.restype = C.c_float
        L.dco_dist2.argtypes = [_f, _f, _sz]
        L.dco_populations.argtypes = [_f, _sz, _sz, _f, _sz, _u32]
        L.dco_free_energies.argtypes = [_u32, _sz, _f]
        L.dco_nearest_neighbors.argtypes = [_f, _sz, _sz, _f, _u32, _f, _u32, _f]
        L.dco_sigma2.restype = C.c_double
        L.dco_sigma2.argtypes = [_f, _sz]
        L.dco_sorted_free_energies.argtypes = [_f, _sz, _u64]
        L.dco_screening.argtypes = [_f, _f, C.c_float, _f, _sz, _sz, _u64, C.c_void_p, _u64]
        L.dco_assign_low_density_frames.argtypes = [_u64, _u32, _u64, _sz, _u64]
        self.L = L

    def dist2(self, x, y):
        x = _c(x, np.float32); y = _c(y, np.float32)
        return np.float32(self.L.dco_dist2(x, y, x.size))

    def populations(self, coords, radii):
        coords = _c(coords, np.float32); n, d = coords.shape
        radii = _c(radii, np.float32)
        out = np.empty((radii.size, n), np.uint32)
        self.L.dco_populations(coords, n, d, radii, radii.size, out)
        return out

    def free_energies(self, pops):
        pops = _c(pops, np.uint32)
        fe = np.empty(pops.size, np.float32)
        self.L.dco_free_energies(pops, pops.size, fe)
        return fe

    def nearest_neighbors(self, coords, fe):
        coords = _c(coords, np.float32); n, d = coords.shape
        fe = _c(fe, np.float32)
        ni = np.empty(n, np.uint32); nd = np.empty(n, np.float32)
        hi = np.empty(n, np.uint32); hd = np.empty(n, np.float32)
        self.L.dco_nearest_neighbors(coords, n, d, fe, ni, nd, hi, hd)
        return ni, nd, hi, hd

    def sigma2(self, nn_d2):
        nn_d2 = _c(nn_d2, np.float32)
        return self.L.dco_sigma2(nn_d2, nn_d2.size)

    def sorted_free_energies(self, fe):
        fe = _c(fe, np.float32)
        order = np.empty(fe.size, np.uint64)
        self.L.dco_sorted_free_energies(fe, fe.size, order)
        return order

    def screening(self, fe, nn_d2, threshold, coords, initial=None):
        coords = _c(coords, np.float32); n, d = coords.shape
        fe = _c(fe, np.float32); nn_d2 = _c(nn_d2, np.float32)
        order = self.sorted_free_energies(fe)
        out = np.empty(n, np.uint64)
        init = None
        if initial is not None:
            init_arr = _c(initial, np.uint64)
            init = init_arr.ctypes.data
        self.L.dco_screening(fe, nn_d2, np.float32(threshold), coords, n, d, order, init, out)
        return out

    def assign_low_density_frames(self, initial, hd_idx, fe):
        initial = _c(initial, np.uint64); hd_idx = _c(hd_idx, np.uint32)
        order = self.sorted_free_energies(fe)
        out = np.empty(initial.size, np.uint64)
        self.L.dco_assign_low_density_frames(initial, hd_idx, order, initial.size, out)
        return out


class Ref:
    """The unmodified reference CPU path (only where oracle/_ref/libdcref.so exists)."""

    def __init__(self):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libdcref.so is not built (reference sources absent)")
        L = C.CDLL(so)
        L.dcref_set_threads.argtypes = [C.c_int]
        L.dcref_max_threads.restype = C.c_int
        L.dcref_populations.argtypes = [_f, _sz, _sz, _f, _sz, _u64]
        L.dcref_free_energies.argtypes = [_u64, _sz, _f]
        L.dcref_sorted_free_energies.argtypes = [_f, _sz, _u64]
        L.dcref_nearest_neighbors.argtypes = [_f, _sz, _sz, _f, _u64, _f, _u64, _f]
        L.dcref_sigma2.restype = C.c_double
        L.dcref_sigma2.argtypes = [_u64, _f, _sz]
        L.dcref_screening.argtypes = [_f, _u64, _f, C.c_float, _f, _sz, _sz, C.c_void_p, _u64]
        L.dcref_assign_low_density_frames.argtypes = [_u64, _u64, _f, _f, _sz, _u64]
        L.dcref_sorted_cluster_names.argtypes = [_u64, _sz, _u64]
        _kv = [C.c_char_p, C.POINTER(C.c_char_p), _f, C.c_int]
        L.dcref_write_pops.argtypes = [C.c_char_p, _u64, _sz] + _kv
        L.dcref_write_fes.argtypes = [C.c_char_p, _f, _sz] + _kv
        L.dcref_write_clustered_trajectory.argtypes = [C.c_char_p, _u64, _sz] + _kv
        L.dcref_write_neighborhood.argtypes = [C.c_char_p, _u64, _f, _u64, _f, _sz] + _kv
        L.dcref_read_coords.argtypes = [C.c_char_p, _f, _sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self.L = L

    # ---- the reference's own file writers / coordinate reader (tools.cpp, tools.hxx)
    @staticmethod
    def _kv(comments):
        keys = (C.c_char_p * len(comments))(*[k.encode() for k in comments])
        vals = np.array(list(comments.values()), np.float32)
        return keys, vals, len(comments)

    def write_pops(self, fname, pops, header, comments):
        pops = _c(pops, np.uint64)
        self.L.dcref_write_pops(fname.encode(), pops, pops.size, header.encode(), *self._kv(comments))

    def write_fes(self, fname, fe, header, comments):
        fe = _c(fe, np.float32)
        self.L.dcref_write_fes(fname.encode(), fe, fe.size, header.encode(), *self._kv(comments))

    def write_states(self, fname, states, header, comments):
        states = _c(states, np.uint64)
        self.L.dcref_write_clustered_trajectory(fname.encode(), states, states.size, header.encode(), *self._kv(comments))

    def write_neighborhood(self, fname, ni, nd, hi, hd, header, comments):
        ni = _c(ni, np.uint64); hi = _c(hi, np.uint64)
        self.L.dcref_write_neighborhood(fname.encode(), ni, _c(nd, np.float32), hi, _c(hd, np.float32), ni.size, header.encode(),
                                        *self._kv(comments))

    def read_coords(self, fname, cap=1 << 22):
        out = np.empty(cap, np.float32)
        r, k = C.c_uint64(0), C.c_uint64(0)
        rc = self.L.dcref_read_coords(fname.encode(), out, cap, C.byref(r), C.byref(k))
        assert rc == 0
        return out[: r.value * k.value].reshape(r.value, k.value).copy()

    def set_threads(self, n):
        self.L.dcref_set_threads(int(n))

    def max_threads(self):
        return self.L.dcref_max_threads()

    def populations(self, coords, radii):
        coords = _c(coords, np.float32); n, d = coords.shape
        radii = _c(radii, np.float32)
        out = np.empty((radii.size, n), np.uint64)
        self.L.dcref_populations(coords, n, d, radii, radii.size, out)
        return out

    def free_energies(self, pops):
        pops = _c(pops, np.uint64)
        fe = np.empty(pops.size, np.float32)
        self.L.dcref_free_energies(pops, pops.size, fe)
        return fe

    def sorted_free_energies(self, fe):
        fe = _c(fe, np.float32)
        order = np.empty(fe.size, np.uint64)
        self.L.dcref_sorted_free_energies(fe, fe.size, order)
        return order

    def nearest_neighbors(self, coords, fe):
        coords = _c(coords, np.float32); n, d = coords.shape
        fe = _c(fe, np.float32)
        ni = np.empty(n, np.uint64); nd = np.empty(n, np.float32)
        hi = np.empty(n, np.uint64); hd = np.empty(n, np.float32)
        self.L.dcref_nearest_neighbors(coords, n, d, fe, ni, nd, hi, hd)
        return ni, nd, hi, hd

    def sigma2(self, nn_idx, nn_d2):
        nn_idx = _c(nn_idx, np.uint64); nn_d2 = _c(nn_d2, np.float32)
        return self.L.dcref_sigma2(nn_idx, nn_d2, nn_d2.size)

    def screening(self, fe, nn_idx, nn_d2, threshold, coords, initial=None):
        coords = _c(coords, np.float32); n, d = coords.shape
        fe = _c(fe, np.float32); nn_idx = _c(nn_idx, np.uint64); nn_d2 = _c(nn_d2, np.float32)
        out = np.empty(n, np.uint64)
        init = None
        if initial is not None:
            init_arr = _c(initial, np.uint64)
            init = init_arr.ctypes.data
        self.L.dcref_screening(fe, nn_idx, nn_d2, np.float32(threshold), coords, n, d, init, out)
        return out

    def assign_low_density_frames(self, initial, hd_idx, hd_d2, fe):
        initial = _c(initial, np.uint64); hd_idx = _c(hd_idx, np.uint64)
        out = np.empty(initial.size, np.uint64)
        self.L.dcref_assign_low_density_frames(initial, hd_idx, _c(hd_d2, np.float32), _c(fe, np.float32), initial.size, out)
        return out

    def sorted_cluster_names(self, clustering):
        clustering = _c(clustering, np.uint64)
        out = np.empty(clustering.size, np.uint64)
        self.L.dcref_sorted_cluster_names(clustering, clustering.size, out)
        return out


class RefCuda:
    """The reference's own CUDA path (oracle/_ref/libdcrefcuda.so, built by `make -C oracle refcuda` from the unmodified
    reference .cu files for sm_100a).  Benchmark comparator only: its semantics differ from the CPU path."""

    def __init__(self):
        self.L = C.CDLL(REFCUDA_SO)
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
        self.L.dcrefcuda_populations.argtypes = [f32, C.c_uint64, C.c_uint64, f32, C.c_uint64, u64]
        self.L.dcrefcuda_nearest_neighbors.argtypes = [f32, C.c_uint64, C.c_uint64, f32, u64, f32, u64, f32]

    def num_gpus(self):
        return self.L.dcrefcuda_num_gpus()

    def populations(self, coords, radii):
        coords = _c(coords, np.float32); n, d = coords.shape
        radii = _c(radii, np.float32)
        out = np.empty((radii.size, n), np.uint64)
        self.L.dcrefcuda_populations(coords, n, d, radii, radii.size, out)
        return out

    def nearest_neighbors(self, coords, fe):
        coords = _c(coords, np.float32); n, d = coords.shape
        fe = _c(fe, np.float32)
        ni = np.empty(n, np.uint64); nd = np.empty(n, np.float32)
        hi = np.empty(n, np.uint64); hd = np.empty(n, np.float32)
        self.L.dcrefcuda_nearest_neighbors(coords, n, d, fe, ni, nd, hi, hd)
        return ni, nd, hi, hd
