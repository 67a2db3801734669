"""Vectored mode (hs_scan_vector) for literal
databases: our compiler's HS_MODE_VECTORED databases drive the UNMODIFIED
reference hs_scan_vector() (src/runtime.c:1106-1175) to the matches the
definition demands over any cut of the data into buffers; the C restatement
reproduces the reference's callbacks, order and termination; on the GPU box
hs_scan_vector of this library does the same."""
import ctypes as C

import numpy as np
import pytest

from hyperscan_b200 import synth
import oracle.brute as brute
import oracle.port as port


def make(hs, nl, seed):
    lits, flags, ids = synth.literal_set(nl, min_len=1 if nl < 10 else 2, max_len=8, seed=seed,
                                         caseless_frac=0.2, alphabet=b"abcdef", singlematch_frac=0.15)
    ids = [i // 2 for i in ids]
    fm = {}
    for k in range(nl):
        fm.setdefault(ids[k], flags[k] & 8)
        flags[k] = (flags[k] & ~8) | fm[ids[k]]
    db = hs.compile_lit_multi(lits, flags, ids, mode=hs.HS_MODE_VECTORED)
    data, off, ln = synth.ragged_corpus([6000], lits, seed=seed + 1, plant_per_kb=15, alphabet=b"abcdefAB")
    return lits, flags, ids, db, data[:6000], off, ln


def cuts_of(n, seed, k=30):
    rng = np.random.default_rng(seed)
    cuts = sorted(rng.integers(0, n, size=k).tolist() + [0, 0, n, 1, 2, 3, n - 1])   # incl. an empty buffer
    return np.diff(np.array(cuts)).astype(np.uint32)


def pairs(recs):
    return [(int(r["id"]), int(r["to"])) for r in recs]


@pytest.mark.parametrize("nl", [1, 6, 40, 300, 1500])
def test_vectored_databases_on_reference_runtime(hs, ref, nl):
    lits, flags, ids, db, data, off, ln = make(hs, nl, nl + 3)
    want = sorted(pairs(brute.scan_blocks(lits, flags, ids, data, off, ln)))
    for seed in (1, 2):
        bl = cuts_of(data.size, seed)
        a, ea = ref.vector_collect(db.ptr, data, bl)
        assert ea == 0 and sorted(pairs(a)) == want
        b, eb = port.vector_collect(db.ptr, data, bl)
        assert eb == 0 and np.array_equal(a, b)          # same callbacks in the same order
        a, ea = ref.vector_collect(db.ptr, data, bl, stop_after=3)
        b, eb = port.vector_collect(db.ptr, data, bl, stop_after=3)
        assert ea == eb == hs.HS_SCAN_TERMINATED and np.array_equal(a, b)


def test_vectored_compile_rules_and_mode_errors(hs, ref):
    db = hs.compile_lit_multi([b"abcdefgh", b"xy"], mode=hs.HS_MODE_VECTORED)
    info = C.c_void_p()
    assert hs.lib().hs_database_info(db.ptr, C.byref(info)) == 0
    assert b"Mode: VECTORED" in C.string_at(info)
    sz = C.c_size_t()
    assert hs.lib().hs_stream_size(db.ptr, C.byref(sz)) == hs.HS_DB_MODE_ERROR   # src/runtime.c:1066-1068
    with pytest.raises(hs.HsError) as e:
        hs.compile_lit_multi([b"abcdefghi"], mode=hs.HS_MODE_VECTORED)
    assert "long literal" in e.value.message
    # the reference runtime agrees on which calls a database of each mode accepts
    blockdb = hs.compile_lit_multi([b"xy"])
    streamdb = hs.compile_lit_multi([b"xy"], mode=hs.HS_MODE_STREAM)
    buf = np.frombuffer(b"..xy..", dtype=np.uint8)
    one = np.array([6], dtype=np.uint32)
    for other in (blockdb, streamdb):
        _, err = ref.vector_collect(other.ptr, buf, one)
        assert err == hs.HS_DB_MODE_ERROR
    _, err = ref.scan_collect(db.ptr, buf, np.array([0], dtype=np.uint64), one)
    assert err == hs.HS_DB_MODE_ERROR
    r, err = ref.vector_collect(db.ptr, buf, one)
    assert err == 0 and pairs(r) == [(1, 4)]


@pytest.mark.gpu
@pytest.mark.parametrize("nl", [1, 6, 40, 300, 1500])
def test_device_scan_vector_equals_reference(hs, ref, nl):
    lits, flags, ids, db, data, off, ln = make(hs, nl, nl + 60)
    scratch = hs.Scratch(db)
    for seed in (1, 2):
        bl = cuts_of(data.size, seed)
        want, _ = ref.vector_collect(db.ptr, data, bl)
        bufs, pos = [], 0
        for n in bl:
            bufs.append(data[pos:pos + int(n)])
            pos += int(n)
        rc, got = hs.scan_vector(db, bufs, scratch)
        assert rc == 0
        # the device delivers each buffer's matches ordered by (to, id); the
        # reference orders equal ends by its literal-id order: same multiset per `to`
        assert sorted(got) == sorted(pairs(want))
        assert [t for _, t in got] == sorted(t for _, t in got)
        rc, part = hs.scan_vector(db, bufs, scratch, stop_after=3)
        if len(want) >= 3:
            assert rc == hs.HS_SCAN_TERMINATED and len(part) == 3
            assert [t for _, t in part] == [t for _, t in got[:3]]


@pytest.mark.gpu
def test_device_scan_vector_argument_checks(hs):
    vdb = hs.compile_lit_multi([b"needle"], mode=hs.HS_MODE_VECTORED)
    bdb = hs.compile_lit_multi([b"needle"])
    scratch = hs.Scratch(vdb)
    rc, got = hs.scan_vector(vdb, [b"..nee", b"", b"dle..needle"], scratch)
    assert rc == 0 and got == [(0, 8), (0, 16)]
    rc, got = hs.scan_vector(vdb, [], scratch)
    assert rc == 0 and got == []
    bs = hs.Scratch(bdb)
    assert hs.scan_vector(bdb, [b"needle"], bs)[0] == hs.HS_DB_MODE_ERROR
    assert hs.scan(vdb, b"needle", scratch)[0] == hs.HS_DB_MODE_ERROR
    L = hs.lib()
    lens = (C.c_uint * 1)(6)
    assert L.hs_scan_vector(vdb.ptr, None, lens, 1, 0, scratch.ptr, hs.MATCH_CB(), None) == hs.HS_INVALID
    ptrs = (C.c_void_p * 1)(None)
    assert L.hs_scan_vector(vdb.ptr, ptrs, lens, 1, 0, scratch.ptr, hs.MATCH_CB(), None) == hs.HS_INVALID
    assert L.hs_scan_vector(vdb.ptr, ptrs, lens, 1, 0, None, hs.MATCH_CB(), None) == hs.HS_INVALID
