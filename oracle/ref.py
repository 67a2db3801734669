"""ctypes binding of the unmodified reference runtime built into oracle/_ref/
(see oracle/ref/Makefile, oracle/ref/ref_driver.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REC_DTYPE = np.dtype([("id", "<u4"), ("block", "<u4"), ("to", "<u8")])

_libs = {}


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_isa():
    """Highest ISA level of the reference's fat runtime this host can run
    (reference: src/dispatcher.c:50-90)."""
    fl = cpu_flags()
    if {"avx512vbmi", "avx512bw", "avx512f"} <= fl:
        return "avx512vbmi"
    if {"avx512bw", "avx512f"} <= fl:
        return "avx512"
    if "avx2" in fl:
        return "avx2"
    return "corei7"


def available():
    return os.path.exists(os.path.join(REF_DIR, "libhsref_%s.so" % best_isa()))


def lib(isa=None):
    isa = isa or best_isa()
    if isa not in _libs:
        path = os.path.join(REF_DIR, "libhsref_%s.so" % isa)
        if not os.path.exists(path):
            raise RuntimeError("reference runtime not built: %s (run `make -C oracle/ref` where "
                               "/root/reference exists)" % path)
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_scan_collect.restype = C.c_long
        L.ref_scan_collect.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t,
                                       C.POINTER(C.c_int)]
        L.ref_stream_collect.restype = C.c_long
        L.ref_vector_collect.restype = C.c_long
        L.ref_vector_collect.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_int)]
        L.ref_stream_collect.argtypes = [vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.POINTER(C.c_int)]
        L.ref_scan_blocks_mt.restype = C.c_double
        L.ref_scan_blocks_mt.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_uint, C.c_uint,
                                         C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.ref_hwlm_exec.restype = C.c_long
        L.ref_hwlm_exec.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_ulonglong, vp, C.c_size_t,
                                    C.c_size_t]
        for f in (L.ref_shufti, L.ref_truffle):
            f.restype = C.c_long
            f.argtypes = [vp, vp, vp, C.c_size_t]
        L.ref_vermicelli.restype = C.c_long
        L.ref_vermicelli.argtypes = [C.c_ubyte, C.c_int, vp, C.c_size_t]
        L.ref_dvermicelli.restype = C.c_long
        L.ref_dvermicelli.argtypes = [C.c_ubyte, C.c_ubyte, C.c_int, vp, C.c_size_t]
        _libs[isa] = L
    return _libs[isa]


def _u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def scan_collect(db_ptr, data, offsets, lengths, stop_after=0, isa=None, cap=None):
    """Reference hs_scan() over blocks; records in delivery order.
    Returns (records, last_error)."""
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    cap = cap or (1 << 20)
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib(isa).ref_scan_collect(db_ptr, keep.ctypes.data, off.ctypes.data, ln.ctypes.data, off.size,
                                      out.ctypes.data, cap, stop_after, C.byref(err))
        if n < 0:
            raise RuntimeError("reference hs_alloc_scratch failed: %d" % n)
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def stream_collect(db_ptr, data, write_lengths, stop_after=0, isa=None):
    """Reference streaming scan of `data` cut into consecutive writes; records
    (id, write index, to = stream offset) in delivery order + last error."""
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    wl = np.ascontiguousarray(write_lengths, dtype=np.uint32)
    assert int(wl.sum()) == a.size
    cap = 1 << 18
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib(isa).ref_stream_collect(db_ptr, keep.ctypes.data, wl.ctypes.data, wl.size, out.ctypes.data,
                                        cap, stop_after, C.byref(err))
        if n < 0:
            raise RuntimeError("reference stream open failed: %d" % n)
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def vector_collect(db_ptr, data, buf_lengths, stop_after=0, isa=None):
    """Reference hs_scan_vector over `data` cut into consecutive buffers; records
    (id, 0, to counted from the first buffer) in delivery order + the call's
    return code."""
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    bl = np.ascontiguousarray(buf_lengths, dtype=np.uint32)
    assert int(bl.sum()) == a.size
    cap = 1 << 18
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        err = C.c_int()
        n = lib(isa).ref_vector_collect(db_ptr, keep.ctypes.data, bl.ctypes.data, bl.size, out.ctypes.data,
                                        cap, stop_after, C.byref(err))
        if n < 0:
            raise RuntimeError("reference scratch allocation failed: %d" % n)
        if n <= cap:
            return out[:n], err.value
        cap = int(n) + 16


def scan_sorted(db_ptr, data, offsets, lengths, isa=None):
    """Match multiset sorted by (block, to, id): what 'bit-exact' is defined on
    (SURVEY.md F8)."""
    r, err = scan_collect(db_ptr, data, offsets, lengths, isa=isa)
    if err:
        raise RuntimeError("reference hs_scan error %d" % err)
    return np.sort(r, order=["block", "to", "id"])


def physical_core_cpus():
    """CPUs of the affinity mask, one hardware thread per physical core first (then
    the SMT siblings), for pinning the bench threads 1:1 the way hsbench does."""
    cpus = sorted(os.sched_getaffinity(0))
    first, rest, seen = [], [], set()
    for c in cpus:
        try:
            with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c) as f:
                sib = f.read().strip()
        except OSError:
            sib = str(c)
        if sib in seen:
            rest.append(c)
        else:
            seen.add(sib)
            first.append(c)
    return first + rest


def pin_bench_threads(cpus):
    """Pin bench thread i to cpus[i % len(cpus)]; an empty list unpins."""
    arr = (C.c_int * max(1, len(cpus)))(*cpus)
    lib().ref_set_bench_cpus(arr, len(cpus))


def bench_blocks(db_ptr, data, offsets, lengths, threads, repeats, isa=None):
    """hsbench-style timing loop (tools/hsbench/main.cpp:503-527).  Returns
    (seconds, matches, bytes)."""
    a = _u8(data)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    m = C.c_ulonglong()
    b = C.c_ulonglong()
    t = lib(isa).ref_scan_blocks_mt(db_ptr, a.ctypes.data, off.ctypes.data, ln.ctypes.data, off.size,
                                    threads, repeats, C.byref(m), C.byref(b))
    if t < 0:
        raise RuntimeError("reference bench failed")
    return t, int(m.value), int(b.value)


def hwlm_exec(hwlm_bytes, data, start=0, groups=0xFFFFFFFFFFFFFFFF, stop_after=0, isa=None):
    """Reference hwlmExec() on a raw HWLM table (64-byte aligned copy)."""
    raw = np.zeros(len(hwlm_bytes) + 64, dtype=np.uint8)
    o = (-raw.ctypes.data) % 64
    raw[o:o + len(hwlm_bytes)] = np.frombuffer(hwlm_bytes, dtype=np.uint8)
    a = _u8(data)
    # the reference engines may read a few bytes around the buffer: pad it
    buf = np.zeros(a.size + 128, dtype=np.uint8)
    buf[64:64 + a.size] = a
    cap = 1 << 16
    out = np.zeros(cap, dtype=REC_DTYPE)
    n = lib(isa).ref_hwlm_exec(raw.ctypes.data + o, buf.ctypes.data + 64, a.size, start, groups,
                               out.ctypes.data, cap, stop_after)
    return [(int(r["to"]), int(r["id"])) for r in out[:min(n, cap)]]


def nfa_exec_blocks(nfa_bytes, data, offsets, lengths, isa=None, cap=1 << 20):
    """Reference nfaExecMcClellan8_B / 16_B / nfaExecSheng_B over every block (offset 0):
    the callbacks as records sorted by (block, to, id)."""
    a = _u8(data)
    keep = a if a.size else np.zeros(1, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint32)
    raw = np.zeros(len(nfa_bytes) + 64, dtype=np.uint8)     # struct NFA is cache-line aligned
    shift = (-raw.ctypes.data) % 64
    raw[shift:shift + len(nfa_bytes)] = np.frombuffer(nfa_bytes, dtype=np.uint8)
    L = lib(isa)
    L.ref_nfa_exec_blocks.restype = C.c_long
    L.ref_nfa_exec_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_size_t]
    while True:
        out = np.zeros(cap, dtype=REC_DTYPE)
        n = L.ref_nfa_exec_blocks(raw.ctypes.data + shift, keep.ctypes.data, off.ctypes.data, ln.ctypes.data,
                                  off.size, out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("reference: engine type not handled")
        if n <= cap:
            return np.sort(out[:n], order=["block", "to", "id"])
        cap = int(n) + 16
