"""ctypes binding of libdcb200.so (include/dcb200.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DCB200_LIB: alternative build of the same library (e.g. the -DDCB_GEMM_PROF diagnostics build)
SO_PATH = os.environ.get("DCB200_LIB") or os.path.join(HERE, "libdcb200.so")

# every symbol include/dcb200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "dcb200_device_count", "dcb200_set_gpus", "dcb200_last_error", "dcb200_version",
    "dcb200_populations", "dcb200_free_energies", "dcb200_nearest_neighbors", "dcb200_screening_step",
    "dcb200_sorted_free_energies", "dcb200_sigma2", "dcb200_screening",
    "dcb200_assign_low_density_frames", "dcb200_sorted_cluster_names",
    "dcb200_io_write_pops", "dcb200_io_write_fes", "dcb200_io_write_states", "dcb200_io_write_neighborhood",
    "dcb200_io_read_coords", "dcb200_io_read_column_float", "dcb200_io_read_column_uint", "dcb200_io_read_neighborhood",
    "dcb200_io_read_comment", "dcb200_io_write_states_record", "dcb200_io_read_states",
    "dcb200_ctx_create", "dcb200_ctx_destroy", "dcb200_ctx_stream", "dcb200_ctx_sync",
    "dcb200_ctx_set_coords", "dcb200_ctx_set_coords_device", "dcb200_ctx_set_coords_ex", "dcb200_ctx_order",
    "dcb200_ctx_to_frame_order", "dcb200_ctx_populations",
    "dcb200_ctx_free_energies", "dcb200_ctx_nn_prepare", "dcb200_ctx_nn_scan", "dcb200_ctx_nn_finish",
    "dcb200_ctx_screening_scan", "dcb200_ctx_screening_flatten", "dcb200_ctx_screening_merge", "dcb200_ctx_stats", "dcb200_ctx_ffma_peak",
    "dcb200_ctx_gemm_info", "dcb200_ctx_tf32_peak",
    "dcb200_density_run", "dcb200_screen_begin", "dcb200_screen_step", "dcb200_screen_end", "dcb200_screen_set_order", "dcb200_screen_labels",
    "dcb200_screening_begin", "dcb200_screening_next", "dcb200_screening_end",
    "dcb200_shard_capacity", "dcb200_ctx_shard_rows", "dcb200_ctx_populations_shard", "dcb200_ctx_nn_scan_shard",
    "dcb200_ctx_shards_to_frame_order", "dcb200_ctx_nn_finish_shards",
]

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_sz = C.c_size_t
_p = C.c_void_p

_lib = None


class Dcb200Error(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise Dcb200Error(
            f"{SO_PATH} is missing: build it with `make -C clustering_b200/csrc` (or __graft_entry__.build()); "
            "there is no Python/CPU fallback for the density kernels")
    L = C.CDLL(SO_PATH)
    L.dcb200_last_error.restype = C.c_char_p
    L.dcb200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.dcb200_set_gpus.argtypes = [C.c_int]
    L.dcb200_populations.argtypes = [_f, _sz, _sz, _f, _sz, _u32]
    L.dcb200_free_energies.argtypes = [_u32, _sz, _f]
    L.dcb200_nearest_neighbors.argtypes = [_f, _sz, _sz, _f, _u32, _f, _u32, _f]
    L.dcb200_screening_step.argtypes = [_f, _sz, _sz, _sz, C.c_float, _u32]
    L.dcb200_sorted_free_energies.argtypes = [_f, _sz, _u32]
    L.dcb200_sigma2.argtypes = [_f, _sz, C.POINTER(C.c_double)]
    L.dcb200_screening.argtypes = [_f, _f, C.c_float, _f, _sz, _sz, _p, _u32]
    L.dcb200_assign_low_density_frames.argtypes = [_u32, _u32, _f, _sz, _u32]
    L.dcb200_sorted_cluster_names.argtypes = [_u32, _sz, _u32]
    L.dcb200_ctx_create.argtypes = [C.c_int, C.POINTER(_p)]
    L.dcb200_ctx_destroy.argtypes = [_p]
    L.dcb200_ctx_stream.argtypes = [_p]
    L.dcb200_ctx_stream.restype = _p
    L.dcb200_ctx_sync.argtypes = [_p]
    L.dcb200_ctx_set_coords.argtypes = [_p, _p, _sz, _sz]
    L.dcb200_ctx_set_coords_device.argtypes = [_p, _p, _sz, _sz]
    L.dcb200_ctx_populations.argtypes = [_p, _f, _sz, _sz, _sz, _p]
    L.dcb200_ctx_free_energies.argtypes = [_p, _p, _sz, C.c_uint32, _p]
    L.dcb200_ctx_nn_prepare.argtypes = [_p, _p]
    L.dcb200_ctx_nn_scan.argtypes = [_p, _sz, _sz, _p, _p]
    L.dcb200_ctx_nn_finish.argtypes = [_p, _p, _p, _p, _p, _p, _p]
    L.dcb200_ctx_screening_scan.argtypes = [_p, _sz, _sz, _sz, _sz, C.c_float, _p]
    L.dcb200_ctx_screening_flatten.argtypes = [_p, _sz, _p]
    L.dcb200_ctx_screening_merge.argtypes = [_p, _sz, _p, _p]
    L.dcb200_ctx_stats.argtypes = [_p, C.POINTER(C.c_uint64), C.c_int]
    L.dcb200_ctx_set_coords_ex.argtypes = [_p, _p, _sz, _sz, C.c_int, C.c_int]
    L.dcb200_ctx_order.argtypes = [_p, _p]
    L.dcb200_ctx_to_frame_order.argtypes = [_p, _p, _sz, _p]
    L.dcb200_ctx_ffma_peak.argtypes = [_p, C.c_double, C.POINTER(C.c_double)]
    L.dcb200_ctx_tf32_peak.argtypes = [_p, C.c_double, C.POINTER(C.c_double)]
    L.dcb200_ctx_gemm_info.argtypes = [_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.dcb200_density_run.argtypes = [_p, _sz, _sz, _f, _sz, _sz, _p, _p, _p, _p, _p, _p, _p]
    L.dcb200_screen_begin.argtypes = [_f, _sz, _sz, C.POINTER(_p)]
    L.dcb200_screen_step.argtypes = [_p, _sz, C.c_float, _p, _u32]
    L.dcb200_screen_end.argtypes = [_p]
    L.dcb200_screening_begin.argtypes = [_f, _f, _f, _sz, _sz, C.POINTER(_p)]
    L.dcb200_screening_next.argtypes = [_p, C.c_float, _u32]
    L.dcb200_screening_end.argtypes = [_p]
    L.dcb200_shard_capacity.argtypes = [_sz, C.c_int]
    L.dcb200_shard_capacity.restype = _sz
    L.dcb200_ctx_shard_rows.argtypes = [_p, C.c_int, C.c_int, C.POINTER(_sz)]
    L.dcb200_ctx_populations_shard.argtypes = [_p, _f, _sz, C.c_int, C.c_int, _p]
    L.dcb200_ctx_nn_scan_shard.argtypes = [_p, C.c_int, C.c_int, _p, _p]
    L.dcb200_ctx_shards_to_frame_order.argtypes = [_p, _p, _sz, C.c_int, _p]
    L.dcb200_ctx_nn_finish_shards.argtypes = [_p, _p, _p, C.c_int, _sz, _p, _p, _p, _p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise Dcb200Error(load().dcb200_last_error().decode())


def device_count():
    n = C.c_int(0)
    check(load().dcb200_device_count(C.byref(n)))
    return n.value


def set_gpus(n):
    check(load().dcb200_set_gpus(int(n)))
