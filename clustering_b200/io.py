"""File formats of `clustering density` through libdcb200.so's dcb200_io_* entry points (include/dcb200.h):
the reference's writers and readers (tools.cpp / tools.hxx) re-created byte for byte."""
import ctypes as C

import numpy as np

from . import lib

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_bound = False


def _L():
    global _bound
    L = lib.load()
    if not _bound:
        kv = [C.c_char_p, C.POINTER(C.c_char_p), _f, C.c_size_t]
        L.dcb200_io_write_pops.argtypes = [C.c_char_p, _u32, C.c_size_t] + kv
        L.dcb200_io_write_fes.argtypes = [C.c_char_p, _f, C.c_size_t] + kv
        L.dcb200_io_write_states.argtypes = [C.c_char_p, _u32, C.c_size_t] + kv
        L.dcb200_io_write_neighborhood.argtypes = [C.c_char_p, _u32, _f, _u32, _f, C.c_size_t] + kv
        L.dcb200_io_read_coords.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dcb200_io_read_column_float.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dcb200_io_read_column_uint.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dcb200_io_read_neighborhood.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                  C.POINTER(C.c_size_t)]
        L.dcb200_io_read_comment.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.POINTER(C.c_float)]
        L.dcb200_io_write_states_record.argtypes = [C.c_char_p, _u32, C.c_size_t, C.c_int]
        L.dcb200_io_read_states.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        _bound = True
    return L


def _kv(comments):
    keys = (C.c_char_p * max(1, len(comments)))(*[k.encode() for k in comments])
    vals = np.array(list(comments.values()) or [0.0], np.float32)
    return keys, vals, len(comments)


def write_pops(fname, pops, header="", comments=None):
    pops = np.ascontiguousarray(pops, np.uint32)
    lib.check(_L().dcb200_io_write_pops(fname.encode(), pops, pops.size, header.encode(), *_kv(comments or {})))


def write_fes(fname, fe, header="", comments=None):
    fe = np.ascontiguousarray(fe, np.float32)
    lib.check(_L().dcb200_io_write_fes(fname.encode(), fe, fe.size, header.encode(), *_kv(comments or {})))


def write_states(fname, states, header="", comments=None):
    states = np.ascontiguousarray(states, np.uint32)
    lib.check(_L().dcb200_io_write_states(fname.encode(), states, states.size, header.encode(), *_kv(comments or {})))


def write_neighborhood(fname, ni, nd, hi, hd, header="", comments=None):
    ni = np.ascontiguousarray(ni, np.uint32); hi = np.ascontiguousarray(hi, np.uint32)
    nd = np.ascontiguousarray(nd, np.float32); hd = np.ascontiguousarray(hd, np.float32)
    lib.check(_L().dcb200_io_write_neighborhood(fname.encode(), ni, nd, hi, hd, ni.size, header.encode(), *_kv(comments or {})))


def read_coords(fname):
    r, k = C.c_size_t(0), C.c_size_t(0)
    lib.check(_L().dcb200_io_read_coords(fname.encode(), None, 0, C.byref(r), C.byref(k)))
    out = np.empty((r.value, k.value), np.float32)
    lib.check(_L().dcb200_io_read_coords(fname.encode(), out.ctypes.data, out.size, C.byref(r), C.byref(k)))
    return out


def read_column(fname, kind=float):
    n = C.c_size_t(0)
    fn = _L().dcb200_io_read_column_float if kind is float else _L().dcb200_io_read_column_uint
    lib.check(fn(fname.encode(), None, 0, C.byref(n)))
    out = np.empty(n.value, np.float32 if kind is float else np.uint32)
    lib.check(fn(fname.encode(), out.ctypes.data, out.size, C.byref(n)))
    return out


def read_neighborhood(fname):
    n = C.c_size_t(0)
    lib.check(_L().dcb200_io_read_neighborhood(fname.encode(), None, None, None, None, 0, C.byref(n)))
    ni = np.empty(n.value, np.uint32); hi = np.empty(n.value, np.uint32)
    nd = np.empty(n.value, np.float32); hd = np.empty(n.value, np.float32)
    lib.check(_L().dcb200_io_read_neighborhood(fname.encode(), ni.ctypes.data, nd.ctypes.data, hi.ctypes.data, hd.ctypes.data, n.value,
                                               C.byref(n)))
    return ni, nd, hi, hd


def read_comment(fname, key, current=0.0):
    v = C.c_float(0.0)
    lib.check(_L().dcb200_io_read_comment(fname.encode(), key.encode(), float(current), C.byref(v)))
    return v.value


def write_states_record(text_fname, states, truncate=False):
    """binary side channel of a per-threshold label file: record of text_fname's labels in <out>.dcb200labels."""
    states = np.ascontiguousarray(states, np.uint32)
    lib.check(_L().dcb200_io_write_states_record(text_fname.encode(), states, states.size, 1 if truncate else 0))


def read_states(fname):
    """-> (labels uint32 [n], from_binary): what read_clustered_trajectory returns for fname, from the binary container when
    it holds a valid record for the file, else parsed from the ASCII file."""
    n, b = C.c_size_t(0), C.c_int(0)
    lib.check(_L().dcb200_io_read_states(fname.encode(), None, 0, C.byref(n), C.byref(b)))
    out = np.empty(n.value, np.uint32)
    lib.check(_L().dcb200_io_read_states(fname.encode(), out.ctypes.data, out.size, C.byref(n), C.byref(b)))
    return out, bool(b.value)
