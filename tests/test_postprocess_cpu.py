"""Host-side report rules (hs_b200_postprocess_matches): ordering, dedupe and
HS_FLAG_SINGLEMATCH / exhaustion, checked against the reference runtime's own
delivery on the same database (no GPU needed: raw records are synthesised by a
brute-force enumeration of every literal occurrence, duplicates included)."""
import numpy as np

from hyperscan_b200 import synth
import oracle.brute as brute


def raw_records(lits, flags, ids, data, off, ln):
    """What the device emits: one record per literal occurrence, no dedupe, no
    exhaustion, arbitrary order."""
    out = []
    for b, (o, n) in enumerate(zip(off, ln)):
        blk = bytes(data[int(o):int(o) + int(n)])
        up = brute.fold(blk)
        for lit, fl, rid in zip(lits, flags, ids):
            hay, needle = (up, brute.fold(lit)) if fl & 1 else (blk, bytes(lit))
            pos = hay.find(needle)
            while pos >= 0:
                out.append((rid, b, pos + len(needle)))
                pos = hay.find(needle, pos + 1)
    rng = np.random.default_rng(0)
    arr = np.array(out, dtype=[("id", "<u4"), ("block", "<u4"), ("to", "<u8")])
    return arr[rng.permutation(arr.size)]


def test_postprocess_equals_reference_delivery(hs, ref):
    lits = [b"abc", b"bc", b"ABC", b"cab", b"xyz", b"c"]
    flags = [8, 0, 8 | 1, 0, 8, 0]
    ids = [1, 2, 1, 2, 3, 4]       # shared ids: dedupe keys; id 1 and 3 single-match
    data, off, ln = synth.ragged_corpus([500, 0, 3000, 77], lits, seed=3, plant_per_kb=40,
                                        alphabet=b"abcxyzABC")
    db = hs.compile_lit_multi(lits, flags, ids)
    raw = raw_records(lits, flags, ids, data, off, ln)
    got = hs.postprocess_matches(db, raw)
    want = ref.scan_sorted(db.ptr, data, off, ln)
    assert raw.size > want.size          # there was something to dedupe / exhaust
    assert np.array_equal(got, want)
    # idempotent
    assert np.array_equal(hs.postprocess_matches(db, got), want)
    # empty input
    assert hs.postprocess_matches(db, raw[:0]).size == 0


def test_postprocess_noodle_singlematch(hs, ref):
    db = hs.compile_lit_multi([b"needle"], [8], [5])
    data = np.frombuffer(b"needle..needle...needle" * 3, dtype=np.uint8)
    off = np.array([0, 32], dtype=np.uint64)
    ln = np.array([30, 37], dtype=np.uint32)
    raw = raw_records([b"needle"], [8], [5], data, off, ln)
    got = hs.postprocess_matches(db, raw)
    assert np.array_equal(got, ref.scan_sorted(db.ptr, data, off, ln))
    assert got.size == 2
