"""Host literal compiler: the databases it emits are consumed by the UNMODIFIED
reference runtime (oracle/_ref) and yield exactly the matches the pattern
definition demands (oracle/brute.py); the plain-C restatement (oracle/port.py)
fires the same callbacks in the same order as the reference."""
import numpy as np
import pytest

from hyperscan_b200 import synth
import oracle.brute as brute
import oracle.port as port

F_CASELESS, F_SINGLE = 1, 8

CASES = [  # (nlits, forced engine, min_len, max_len)
    (1, -1, 1, 1), (1, -1, 2, 8), (1, -1, 9, 30), (2, -1, 1, 3), (8, -1, 1, 8), (8, 11, 3, 8),
    (20, 13, 2, 9), (40, 15, 3, 12), (40, 17, 4, 12), (48, 16, 3, 8), (48, 12, 1, 4),
    (6, 0, 1, 5), (100, -1, 3, 10), (300, -1, 2, 16), (1000, -1, 4, 8), (5000, -1, 4, 16),
]


@pytest.mark.parametrize("nlits,engine,lo,hi", CASES)
def test_reference_runtime_on_our_databases(hs, ref, nlits, engine, lo, hi):
    lits, flags, ids = synth.literal_set(nlits, min_len=lo, max_len=hi, seed=100 + nlits + lo,
                                         caseless_frac=0.25, alphabet=b"abcdeXYZ", singlematch_frac=0.1)
    ids = [i - (i % 3 == 2) for i in ids]  # some shared report ids -> dedupe keys
    fm = {}
    for k in range(nlits):
        fm.setdefault(ids[k], flags[k] & F_SINGLE)
        flags[k] = (flags[k] & ~F_SINGLE) | fm[ids[k]]
    if engine >= 0:
        hs.set_build_option("force_engine", engine)
    db = hs.compile_lit_multi(lits, flags, ids)
    info = db.info()
    if engine > 0:
        assert info.engine_id == engine
    if nlits == 1:
        assert info.hwlm_type == 16  # noodle
    data, off, ln = synth.ragged_corpus([0, 1, 2, 9, 33, 128, 1000, 4096, 30000], lits, seed=5,
                                        alphabet=b"abcdeXYZABCDExyz", plant_per_kb=8)
    want = brute.scan_blocks(lits, flags, ids, data, off, ln)
    got = ref.scan_sorted(db.ptr, data, off, ln)
    assert got.size == want.size
    assert np.array_equal(got, want)
    # the C restatement: same callbacks, same order, same termination
    a, ea = ref.scan_collect(db.ptr, data, off, ln)
    b, eb = port.scan_collect(db.ptr, data, off, ln)
    assert ea == eb == 0 and np.array_equal(a, b)
    a, ea = ref.scan_collect(db.ptr, data, off, ln, stop_after=3)
    b, eb = port.scan_collect(db.ptr, data, off, ln, stop_after=3)
    assert ea == eb and np.array_equal(a, b)


@pytest.mark.parametrize("domain,stride", [(9, 1), (10, 2), (12, 4), (13, 2), (15, 1)])
def test_fdr_domain_stride_variants(hs, ref, domain, stride):
    lits, flags, ids = synth.literal_set(200, min_len=4, max_len=10, seed=domain, caseless_frac=0.3,
                                         alphabet=b"abcdefgh")
    hs.set_build_option("force_engine", 0)
    hs.set_build_option("fdr_domain", domain)
    hs.set_build_option("fdr_stride", stride)
    db = hs.compile_lit_multi(lits, flags, ids)
    assert (db.info().fdr_domain, db.info().fdr_stride) == (domain, stride)
    data, off, ln = synth.ragged_corpus([5000, 17, 20000], lits, seed=6, alphabet=b"abcdefghABCDEFGH",
                                        plant_per_kb=10)
    want = brute.scan_blocks(lits, flags, ids, data, off, ln)
    assert np.array_equal(ref.scan_sorted(db.ptr, data, off, ln), want)
    assert np.array_equal(port.scan_sorted(db.ptr, data, off, ln), want)


def test_engine_choice_follows_reference_heuristics(hs):
    # src/fdr/fdr_engine_description.cpp:97-182, src/hwlm/hwlm_build.cpp:107-118
    def info(n, lo, hi):
        lits, flags, ids = synth.literal_set(n, min_len=lo, max_len=hi, seed=n)
        return hs.compile_lit_multi(lits, flags, ids).info()
    i = info(1, 6, 6)
    assert i.hwlm_type == 16
    i = info(30, 4, 8)
    assert i.hwlm_type == 12 and 11 <= i.engine_id <= 18          # Teddy, 8 buckets
    i = info(1000, 4, 8)
    assert (i.engine_id, i.fdr_domain, i.fdr_stride) == (0, 13, 2)  # BASELINE config 2
    i = info(50000, 4, 16)
    assert (i.engine_id, i.fdr_domain, i.fdr_stride) == (0, 15, 1)  # BASELINE config 5


def test_fat_teddy_needs_avx2_platform(hs, ref):
    if ref.best_isa() == "corei7":
        pytest.skip("host lacks AVX2")
    lits, flags, ids = synth.literal_set(90, min_len=3, max_len=8, seed=90, alphabet=b"abcdefgh")
    plat = hs.PlatformInfo(0, hs.HS_CPU_FEATURES_AVX2, 0, 0)
    import ctypes as C
    db = hs.compile_lit_multi(lits, flags, ids, platform=C.byref(plat))
    assert 3 <= db.info().engine_id <= 10
    data, off, ln = synth.ragged_corpus([20000, 100], lits, seed=2, alphabet=b"abcdefgh", plant_per_kb=4)
    want = brute.scan_blocks(lits, flags, ids, data, off, ln)
    assert np.array_equal(ref.scan_sorted(db.ptr, data, off, ln), want)
    assert np.array_equal(port.scan_sorted(db.ptr, data, off, ln), want)
