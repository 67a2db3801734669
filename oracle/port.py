"""ctypes binding of oracle/_build/liboracle.so (oracle/hs_oracle.c, the plain-C
restatement).  TEST INFRASTRUCTURE ONLY.  Same call shapes as oracle/ref.py."""
import ctypes as C
import os
import subprocess

import numpy as np

from .ref import REC_DTYPE, _u8

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "hs_oracle.c")
OUT = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", OUT, SRC], check=True)
        L = C.CDLL(OUT)
        vp = C.c_void_p
        L.oracle_scan_collect.restype = C.c_long
        L.oracle_scan_collect.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t,
                                          C.POINTER(C.c_int)]
        L.oracle_scan_blocks_mt.restype = C.c_double
        L.oracle_scan_blocks_mt.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_uint, C.c_uint,
                                            C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.oracle_stream_collect.restype = C.c_long
        L.oracle_vector_collect.restype = C.c_long
        L.oracle_vector_collect.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_int)]
        L.oracle_stream_collect.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_int)]
        L.oracle_hwlm_exec.restype = C.c_long
        L.oracle_hwlm_exec.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_ulonglong, vp, C.c_size_t,
                                       C.c_size_t]
        _lib = L
    return _lib


def scan_collect(db_ptr, data, offsets, lengths, stop_after=0, cap=None):
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    cap = cap or (1 << 20)
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib().oracle_scan_collect(db_ptr, keep.ctypes.data, off.ctypes.data, ln.ctypes.data, off.size,
                                      out.ctypes.data, cap, stop_after, C.byref(err))
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def stream_collect(db_ptr, data, write_lengths, stop_after=0):
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    wl = np.ascontiguousarray(write_lengths, dtype=np.uint32)
    cap = 1 << 18
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib().oracle_stream_collect(db_ptr, keep.ctypes.data, wl.ctypes.data, wl.size, out.ctypes.data,
                                        cap, stop_after, C.byref(err))
        if n < 0:
            raise RuntimeError("oracle stream open failed")
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def vector_collect(db_ptr, data, buf_lengths, stop_after=0):
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    bl = np.ascontiguousarray(buf_lengths, dtype=np.uint32)
    cap = 1 << 18
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib().oracle_vector_collect(db_ptr, keep.ctypes.data, bl.ctypes.data, bl.size, out.ctypes.data,
                                        cap, stop_after, C.byref(err))
        if n < 0:
            raise RuntimeError("oracle: not a vectored pure-literal database")
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def scan_sorted(db_ptr, data, offsets, lengths):
    r, err = scan_collect(db_ptr, data, offsets, lengths)
    if err:
        raise RuntimeError("oracle scan error %d" % err)
    return np.sort(r, order=["block", "to", "id"])


def bench_blocks(db_ptr, data, offsets, lengths, threads, repeats):
    a = _u8(data)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    m = C.c_ulonglong()
    b = C.c_ulonglong()
    t = lib().oracle_scan_blocks_mt(db_ptr, a.ctypes.data, off.ctypes.data, ln.ctypes.data, off.size,
                                    threads, repeats, C.byref(m), C.byref(b))
    return t, int(m.value), int(b.value)


def hwlm_exec(hwlm_bytes, data, start=0, groups=0xFFFFFFFFFFFFFFFF, stop_after=0):
    raw = np.zeros(len(hwlm_bytes) + 64, dtype=np.uint8)
    o = (-raw.ctypes.data) % 64
    raw[o:o + len(hwlm_bytes)] = np.frombuffer(hwlm_bytes, dtype=np.uint8)
    a = _u8(data)
    buf = a if a.size else np.zeros(1, dtype=np.uint8)
    cap = 1 << 16
    out = np.zeros(cap, dtype=REC_DTYPE)
    n = lib().oracle_hwlm_exec(raw.ctypes.data + o, buf.ctypes.data, a.size, start, groups,
                               out.ctypes.data, cap, stop_after)
    return [(int(r["to"]), int(r["id"])) for r in out[:min(n, cap)]]
