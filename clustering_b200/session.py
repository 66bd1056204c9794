"""Device-resident session over libdcb200.so's dcb200_ctx_* API.  torch is used only as the owner of
device memory (raw pointers are passed through the C ABI); all compute is in the library's CUDA kernels."""
import ctypes as C

import numpy as np
import torch

from . import lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _ordered(fn):
    """The library enqueues on the context's own NON-BLOCKING stream; torch tensors are produced and consumed on torch's
    current stream.  Every Session method that launches work therefore orders the context stream after the caller's
    current stream on entry and the caller's stream after the context stream on exit (two event waits, no host
    synchronisation): results can be used right away from whichever stream the caller is on, and inputs produced there
    are complete before the library reads them.  Inside `with torch.cuda.stream(session.torch_stream())` both are no-ops."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        cur = torch.cuda.current_stream(self.dev)
        mine = self.torch_stream()
        if cur.cuda_stream != mine.cuda_stream:
            mine.wait_stream(cur)
        try:
            return fn(self, *args, **kwargs)
        finally:
            if cur.cuda_stream != mine.cuda_stream:
                cur.wait_stream(mine)
    return wrapper


class Session:
    def __init__(self, device=0):
        self.L = lib.load()
        self.device = int(device)
        h = C.c_void_p()
        lib.check(self.L.dcb200_ctx_create(self.device, C.byref(h)))
        self.h = h
        self._stream = None
        self.dev = torch.device("cuda", self.device)
        self.n = self.d = 0

    def close(self):
        if self.h:
            self.L.dcb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream_handle(self):
        return self.L.dcb200_ctx_stream(self.h)

    def torch_stream(self):
        if self._stream is None:
            self._stream = torch.cuda.ExternalStream(self.stream_handle, device=self.dev)
        return self._stream

    def sync(self):
        lib.check(self.L.dcb200_ctx_sync(self.h))

    def stats(self, reset=False):
        s = (C.c_uint64 * 6)()
        lib.check(self.L.dcb200_ctx_stats(self.h, s, 1 if reset else 0))
        return dict(launches=int(s[0]), slow_pairs=int(s[1]), exact_pairs=int(s[2]), tiles_streamed=int(s[3]),
                    pairs_scheduled=int(s[4]), pairs_evaluated=int(s[3]) * int(s[5]))

    def ffma_peak(self, ms_target=200.0):
        t = C.c_double(0.0)
        lib.check(self.L.dcb200_ctx_ffma_peak(self.h, float(ms_target), C.byref(t)))
        return t.value

    def tf32_peak(self, ms_target=200.0):
        t = C.c_double(0.0)
        lib.check(self.L.dcb200_ctx_tf32_peak(self.h, float(ms_target), C.byref(t)))
        return t.value

    def gemm_info(self):
        """(active, check_ratio) of the GEMM-form tensor-core path (17 <= n_cols <= 256), see dcb200_ctx_gemm_info."""
        a, r = C.c_int(0), C.c_float(0.0)
        lib.check(self.L.dcb200_ctx_gemm_info(self.h, C.byref(a), C.byref(r)))
        return bool(a.value), float(r.value)

    # ---- coordinates
    @_ordered
    def set_coords(self, coords, keep_order=False):
        """coords: numpy [n][d] (host upload) or a CUDA torch tensor [n][d] (adopted from device memory).
        keep_order: positions = frame order (screening input); default: spatial order chosen by the library."""
        if isinstance(coords, torch.Tensor):
            assert coords.is_cuda and coords.dtype == torch.float32 and coords.is_contiguous()
            self.n, self.d = coords.shape
            lib.check(self.L.dcb200_ctx_set_coords_ex(self.h, _ptr(coords), self.n, self.d, 1, int(keep_order)))
        else:
            coords = np.ascontiguousarray(coords, np.float32)
            self.n, self.d = coords.shape
            lib.check(self.L.dcb200_ctx_set_coords_ex(self.h, C.c_void_p(coords.ctypes.data), self.n, self.d, 0, int(keep_order)))

    @_ordered
    def order(self):
        """int32 [n]: frame index at every position."""
        out = torch.empty(self.n, dtype=torch.int32, device=self.dev)
        lib.check(self.L.dcb200_ctx_order(self.h, _ptr(out)))
        return out

    @_ordered
    def to_frame_order(self, src, out=None):
        """src: 32-bit device tensor [k][n] in position order -> same shape in frame order."""
        assert src.is_cuda and src.is_contiguous() and src.element_size() == 4 and src.shape[-1] == self.n
        if out is None:
            out = torch.empty_like(src)
        lib.check(self.L.dcb200_ctx_to_frame_order(self.h, _ptr(src), src.numel() // self.n, _ptr(out)))
        return out

    # ---- populations / free energies
    @_ordered
    def populations(self, radii, row_begin=0, row_end=None, out=None):
        """positions [row_begin,row_end) -> int32 [n_radii][rows] in POSITION order (see to_frame_order)."""
        radii = np.ascontiguousarray(np.atleast_1d(radii), np.float32)
        row_end = self.n if row_end is None else row_end
        if out is None:
            out = torch.empty((radii.size, row_end - row_begin), dtype=torch.int32, device=self.dev)
        lib.check(self.L.dcb200_ctx_populations(self.h, radii, radii.size, row_begin, row_end, _ptr(out)))
        return out

    @_ordered
    def free_energies(self, pops, max_pop=0, out=None):
        assert pops.is_cuda and pops.dtype == torch.int32 and pops.is_contiguous()
        if out is None:
            out = torch.empty(pops.numel(), dtype=torch.float32, device=self.dev)
        lib.check(self.L.dcb200_ctx_free_energies(self.h, _ptr(pops), pops.numel(), int(max_pop), _ptr(out)))
        return out

    # ---- nearest neighbours
    @_ordered
    def nn_prepare(self, fe):
        assert fe.is_cuda and fe.dtype == torch.float32 and fe.is_contiguous() and fe.numel() == self.n
        lib.check(self.L.dcb200_ctx_nn_prepare(self.h, _ptr(fe)))

    @_ordered
    def nn_scan(self, pos_begin=0, pos_end=None, out=None):
        pos_end = self.n if pos_end is None else pos_end
        if out is None:
            out = torch.empty((2, pos_end - pos_begin), dtype=torch.int64, device=self.dev)
        lib.check(self.L.dcb200_ctx_nn_scan(self.h, pos_begin, pos_end, _ptr(out[0]), _ptr(out[1])))
        return out

    @_ordered
    def nn_finish(self, keys, out=None):
        """keys: int64 [2][n] in sorted-position order -> (nn_idx, nn_d2, hd_idx, hd_d2) in frame order."""
        assert keys.shape == (2, self.n) and keys.is_contiguous()
        if out is None:
            out = (torch.empty(self.n, dtype=torch.int32, device=self.dev), torch.empty(self.n, dtype=torch.float32, device=self.dev),
                   torch.empty(self.n, dtype=torch.int32, device=self.dev), torch.empty(self.n, dtype=torch.float32, device=self.dev))
        lib.check(self.L.dcb200_ctx_nn_finish(self.h, _ptr(keys[0]), _ptr(keys[1]), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                              _ptr(out[3])))
        return out

    def nearest_neighbors(self, fe):
        """fe: frame order -> (nn_idx, nn_d2, hd_idx, hd_d2) in frame order."""
        self.nn_prepare(fe)
        return self.nn_finish(self.nn_scan())

    # ---- shards (block-cyclic dealing of the positions to the ranks of a multi-GPU run, include/dcb200.h)
    def shard_capacity(self, n_shards):
        return int(self.L.dcb200_shard_capacity(self.n, int(n_shards)))

    def shard_rows(self, shard, n_shards):
        r = C.c_size_t(0)
        lib.check(self.L.dcb200_ctx_shard_rows(self.h, int(shard), int(n_shards), C.byref(r)))
        return int(r.value)

    @_ordered
    def populations_shard(self, radii, shard, n_shards, out=None):
        """-> int32 [n_radii][capacity] (the shard's rows first, the padding is not written).
        (torch.empty, not zeros: a fill kernel on torch's stream would race with the library's own non-blocking stream)"""
        radii = np.ascontiguousarray(np.atleast_1d(radii), np.float32)
        if out is None:
            out = torch.empty((radii.size, self.shard_capacity(n_shards)), dtype=torch.int32, device=self.dev)
        lib.check(self.L.dcb200_ctx_populations_shard(self.h, radii, radii.size, int(shard), int(n_shards), _ptr(out)))
        return out

    @_ordered
    def nn_scan_shard(self, shard, n_shards, out=None):
        """-> int64 [2][capacity] neighbour keys of the shard's rows."""
        if out is None:
            out = torch.empty((2, self.shard_capacity(n_shards)), dtype=torch.int64, device=self.dev)
        lib.check(self.L.dcb200_ctx_nn_scan_shard(self.h, int(shard), int(n_shards), _ptr(out[0]), _ptr(out[1])))
        return out

    @_ordered
    def shards_to_frame_order(self, gathered, n_arrays, n_shards, out=None):
        """gathered: 32-bit device tensor [n_shards][n_arrays][capacity] -> [n_arrays][n] in frame order."""
        assert gathered.is_cuda and gathered.is_contiguous() and gathered.element_size() == 4
        assert gathered.numel() == n_shards * n_arrays * self.shard_capacity(n_shards)
        if out is None:
            out = torch.empty((n_arrays, self.n), dtype=gathered.dtype, device=self.dev)
        lib.check(self.L.dcb200_ctx_shards_to_frame_order(self.h, _ptr(gathered), int(n_arrays), int(n_shards), _ptr(out)))
        return out

    @_ordered
    def nn_finish_shards(self, gathered, n_shards, out=None):
        """gathered: int64 [n_shards][2][capacity] -> (nn_idx, nn_d2, hd_idx, hd_d2) in frame order."""
        cap = self.shard_capacity(n_shards)
        assert gathered.is_cuda and gathered.is_contiguous() and gathered.numel() == n_shards * 2 * cap
        if out is None:
            out = (torch.empty(self.n, dtype=torch.int32, device=self.dev), torch.empty(self.n, dtype=torch.float32, device=self.dev),
                   torch.empty(self.n, dtype=torch.int32, device=self.dev), torch.empty(self.n, dtype=torch.float32, device=self.dev))
        base = gathered.data_ptr()            # shard s: nn keys at [s][0], hd keys at [s][1] -> stride 2 * capacity, no copies
        lib.check(self.L.dcb200_ctx_nn_finish_shards(self.h, C.c_void_p(base), C.c_void_p(base + 8 * cap), int(n_shards), 2 * cap,
                                                     _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3])))
        return out

    # ---- screening (coords must be the free-energy-sorted frames)
    @_ordered
    def screening_scan(self, m_prev, m_new, max_dist2, comp, row_begin=0, row_end=None):
        row_end = m_new if row_end is None else row_end
        assert comp.is_cuda and comp.dtype == torch.int32 and comp.numel() >= m_new
        lib.check(self.L.dcb200_ctx_screening_scan(self.h, m_prev, m_new, row_begin, row_end, float(max_dist2), _ptr(comp)))

    @_ordered
    def screening_flatten(self, m_new, comp):
        lib.check(self.L.dcb200_ctx_screening_flatten(self.h, m_new, _ptr(comp)))

    @_ordered
    def screening_merge(self, m_new, comp, other):
        lib.check(self.L.dcb200_ctx_screening_merge(self.h, m_new, _ptr(comp), _ptr(other)))
