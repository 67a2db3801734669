"""API behaviour of the scan / scratch entry points on the device build, after the
reference's own API tests: unit/hyperscan/arg_checks.cpp (NULL / bad-magic /
alignment errors), scratch_op.cpp (grow-only scratch shared by databases, clone,
size), scratch_in_use.cpp (re-entrancy guard), literals.cpp (random literal
sets x length bounds x flags)."""
import ctypes as C

import numpy as np
import pytest

from hyperscan_b200 import synth
import oracle.brute as brute

pytestmark = pytest.mark.gpu


def test_scan_arg_checks(hs):
    L = hs.lib()
    db = hs.compile_lit_multi([b"foobar"])
    scratch = hs.Scratch(db)
    buf = C.create_string_buffer(b"xxfoobarxx")
    cb = hs.MATCH_CB(lambda *a: 0)
    # arg_checks.cpp ScanBlockNoScratch / NoData / NoDatabase
    assert L.hs_scan(db.ptr, buf, 10, 0, None, cb, None) == hs.HS_INVALID
    assert L.hs_scan(db.ptr, None, 10, 0, scratch.ptr, cb, None) == hs.HS_INVALID
    assert L.hs_scan(None, buf, 10, 0, scratch.ptr, cb, None) == hs.HS_INVALID
    # bogus database bytes / bad version
    junk = C.create_string_buffer(4096)
    assert L.hs_scan(junk, buf, 10, 0, scratch.ptr, cb, None) == hs.HS_INVALID
    blob = bytearray(db.serialize())
    raw = C.create_string_buffer(len(blob) + 128)
    base = (C.addressof(raw) + 63) // 64 * 64
    assert L.hs_deserialize_database_at(bytes(blob), len(blob), C.c_void_p(base)) == 0
    C.memmove(base + 4, b"\x00\x00\x00\x01", 4)
    assert L.hs_scan(C.c_void_p(base), buf, 10, 0, scratch.ptr, cb, None) == hs.HS_DB_VERSION_ERROR
    # bogus scratch
    assert L.hs_scan(db.ptr, buf, 10, 0, junk, cb, None) == hs.HS_INVALID
    # NULL callback: matches are suppressed, scan succeeds (src/runtime.c:127)
    assert L.hs_scan(db.ptr, buf, 10, 0, scratch.ptr, hs.MATCH_CB(), None) == hs.HS_SUCCESS
    # hs_alloc_scratch argument checks (src/scratch.c:244-262)
    out = C.c_void_p()
    assert L.hs_alloc_scratch(None, C.byref(out)) == hs.HS_INVALID
    assert L.hs_alloc_scratch(db.ptr, None) == hs.HS_INVALID
    assert L.hs_alloc_scratch(junk, C.byref(out)) == hs.HS_INVALID
    bad = C.c_void_p(C.addressof(junk))
    assert L.hs_alloc_scratch(db.ptr, C.byref(bad)) == hs.HS_INVALID
    assert L.hs_free_scratch(None) == hs.HS_SUCCESS
    sz = C.c_size_t()
    assert L.hs_scratch_size(None, C.byref(sz)) == hs.HS_INVALID
    assert L.hs_scratch_size(scratch.ptr, None) == hs.HS_INVALID


def test_scratch_shared_by_databases_clone_and_size(hs, ref):
    dba = hs.compile_lit_multi([b"alpha", b"beta"], ids=[1, 2])
    lits, flags, ids = synth.literal_set(400, seed=77, alphabet=b"abcdefgh")
    dbb = hs.compile_lit_multi(lits, flags, ids)
    scratch = hs.Scratch(dba)
    s1 = scratch.size()
    scratch.add(dbb)                       # grow-only, same handle serves both
    assert scratch.size() >= s1
    data = b"..alpha..beta.." + lits[0] + b".."
    for db in (dba, dbb):
        rc, out = hs.scan(db, data, scratch)
        want = ref.scan_sorted(db.ptr, data, [0], [len(data)])
        assert rc == 0 and sorted(out, key=lambda x: (x[1], x[0])) == [(int(r["id"]), int(r["to"])) for r in want]
    clone = C.c_void_p()
    assert hs.lib().hs_clone_scratch(scratch.ptr, C.byref(clone)) == 0
    got = []
    cb = hs.MATCH_CB(lambda i, f, t, fl, ctx: got.append((i, t)) or 0)
    buf = C.create_string_buffer(data, len(data))
    assert hs.lib().hs_scan(dba.ptr, buf, len(data), 0, clone, cb, None) == 0
    assert got == [(1, 7), (2, 13)]
    assert hs.lib().hs_free_scratch(clone) == 0


def test_scratch_in_use_guard(hs):
    # unit/hyperscan/scratch_in_use.cpp: re-entering with the same scratch from
    # inside the match callback is refused
    db = hs.compile_lit_multi([b"x"])
    scratch = hs.Scratch(db)
    L = hs.lib()
    buf = C.create_string_buffer(b"axb")
    seen = []

    def cb(i, frm, to, flags, ctx):
        seen.append(L.hs_scan(db.ptr, buf, 3, 0, scratch.ptr, hs.MATCH_CB(), None))
        out = C.c_void_p(scratch.ptr.value)
        seen.append(L.hs_alloc_scratch(db.ptr, C.byref(out)))
        seen.append(L.hs_free_scratch(scratch.ptr))
        return 0

    assert L.hs_scan(db.ptr, buf, 3, 0, scratch.ptr, hs.MATCH_CB(cb), None) == 0
    assert seen == [hs.HS_SCRATCH_IN_USE] * 3
    assert L.hs_scan(db.ptr, buf, 3, 0, scratch.ptr, hs.MATCH_CB(), None) == 0   # released afterwards


@pytest.mark.parametrize("count", [1, 10, 100, 500])
@pytest.mark.parametrize("lo,hi", [(3, 10), (10, 100)])
@pytest.mark.parametrize("flag", [0, 8])
def test_random_literal_sets(hs, ref, count, lo, hi, flag):
    # unit/hyperscan/literals.cpp:160-260 (random [a-z] literals, block mode)
    lits, flags, ids = synth.literal_set(count, min_len=lo, max_len=hi, seed=29785643 % 1000 + count + lo,
                                         caseless_frac=0.0)
    flags = [flag] * count
    data, off, ln = synth.ragged_corpus([40000], lits, seed=count, plant_per_kb=2.0,
                                        alphabet=b"abcdefghijklmnopqrstuvwxyz")
    db = hs.compile_lit_multi(lits, flags, ids)
    scratch = hs.Scratch(db)
    rc, out = hs.scan(db, data, scratch)
    assert rc == 0
    tos = [t for _, t in out]
    assert tos == sorted(tos)              # unit/hyperscan/order.cpp: non-decreasing `to`
    want = ref.scan_sorted(db.ptr, data, off, ln)
    assert sorted(out, key=lambda x: (x[1], x[0])) == [(int(r["id"]), int(r["to"])) for r in want]
    assert np.array_equal(want, brute.scan_blocks(lits, flags, ids, data, off, ln))


def test_unaligned_and_gapped_host_layouts(hs, ref):
    # blocks at odd host offsets (the packed fallback) and with junk between blocks
    lits, flags, ids = synth.literal_set(50, min_len=3, max_len=9, seed=3, alphabet=b"abcdef")
    for align in (1, 3, 16, 64):
        data, off, ln = synth.ragged_corpus([100, 0, 7, 2500, 33, 1], lits, seed=align, plant_per_kb=20,
                                            align=align, alphabet=b"abcdefAB")
        db = hs.compile_lit_multi(lits, flags, ids)
        scratch = hs.Scratch(db)
        got = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
        assert np.array_equal(got, ref.scan_sorted(db.ptr, data, off, ln))


def test_scratch_is_usable_from_a_thread_with_another_current_device(hs, ref, real_gpu):
    """A scratch is bound to the device it was allocated on; a thread whose current
    CUDA device differs (a fresh thread starts on device 0) must still be able to
    scan with it -- the library switches devices inside the call and restores the
    caller's (round-1 advisor finding).  Needs two GPUs for the interesting case;
    on one GPU it still checks the cross-thread use."""
    import threading
    import numpy as np
    import torch
    from hyperscan_b200 import synth
    ndev = torch.cuda.device_count()
    dev = ndev - 1
    torch.cuda.set_device(dev)
    lits, flags, ids = synth.literal_set(200, seed=5)
    data, off, ln, _ = synth.block_corpus(64, 1024, lits, plant_per_kb=1.0, seed=6)
    db = hs.compile_lit_multi(lits, flags, ids)
    scratch = hs.Scratch(db)                      # lives on device `dev`
    want = ref.scan_sorted(db.ptr, data, off, ln)
    out = {}

    def worker():
        torch.cuda.set_device(0)                  # what a fresh thread has anyway
        out["recs"] = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
        rc, m = hs.scan(db, bytes(data[:1024]), scratch)
        out["rc"] = rc
        out["dev"] = torch.cuda.current_device()

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert np.array_equal(out["recs"], want)
    assert out["rc"] == hs.HS_SUCCESS and out["dev"] == 0
    scratch.free()
    torch.cuda.set_device(0)


def test_clone_scratch_after_the_database_was_freed(hs):
    """The reference's hs_clone_scratch never touches a database; here the clone rebuilds
    its device images from the source scratch's own copy of the bytes, so a database the
    application has already freed is not dereferenced (round-1 advisor finding)."""
    import ctypes as C
    db = hs.compile_lit_multi([b"abcdef", b"xyz"], [0, 0], [1, 2])
    keep = hs.compile_lit_multi([b"abcdef", b"xyz"], [0, 0], [1, 2])
    scratch = hs.Scratch(db)
    db.__del__()                                   # hs_free_database
    db.ptr = None
    clone = C.c_void_p()
    assert hs.lib().hs_clone_scratch(scratch.ptr, C.byref(clone)) == hs.HS_SUCCESS
    twin = hs.Scratch.__new__(hs.Scratch)
    twin.ptr = clone
    rc, out = hs.scan(keep, b"..abcdef..xyz", twin)
    assert rc == hs.HS_SUCCESS and out == [(1, 8), (2, 13)]
    assert hs.lib().hs_free_scratch(clone) == hs.HS_SUCCESS
    twin.ptr = None
    scratch.free()
