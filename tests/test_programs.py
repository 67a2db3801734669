"""Every stateless opcode of roseRunProgram_l (src/rose/program_runtime.c:3101-3522)
that the literal compiler here does not emit on its own -- CHECK_MASK, CHECK_BYTE,
CHECK_LONG_LIT(_NOCASE), CHECK_GROUPS, SQUASH_GROUPS, INCLUDED_JUMP, FINAL_REPORT,
multi-block programs -- reached through hand-assembled literal programs
(hs_b200_test_compile_programs; instruction layouts: src/rose/rose_program.h:214-724).

CPU half: the UNMODIFIED reference runtime scanning those databases reproduces the
reference's own ValidateMask known-answer formulas (unit/internal/rose_mask.cpp:
107-216) and the C restatement follows it.  GPU half (-m gpu): the device
interpreter (scan_kernels.cu runProgram / checkMask8 / checkByte / checkLiteral)
gives the same (id, to) sets from the same bytes."""
import struct

import numpy as np
import pytest

import oracle.port as port

END, CHECK_GROUPS, CHECK_MASK, CHECK_BYTE = 0, 3, 9, 11
PUSH_DELAYED, DEDUPE, REPORT_CHAIN, REPORT, REPORT_EXHAUST = 14, 28, 30, 33, 34
DEDUPE_AND_REPORT, FINAL_REPORT, CHECK_EXHAUSTED, SQUASH_GROUPS = 37, 38, 39, 43
CHECK_LONG_LIT, CHECK_LONG_LIT_NOCASE, CHECK_MED_LIT, CHECK_MED_LIT_NOCASE = 51, 52, 53, 54
INCLUDED_JUMP, SET_EXHAUST = 61, 65


def _pad(b):
    return b + b"\0" * (-len(b) % 8)


def i_end():
    return _pad(struct.pack("<B", END))


def i_check_groups(groups):
    return _pad(struct.pack("<B7xQ", CHECK_GROUPS, groups))


def i_squash_groups(groups):
    return _pad(struct.pack("<B7xQ", SQUASH_GROUPS, groups))


def i_check_mask(and_m, cmp_m, neg_m, offset, fail_jump):
    return _pad(struct.pack("<B7xQQQiI", CHECK_MASK, and_m, cmp_m, neg_m, offset, fail_jump))


def i_check_byte(and_m, cmp_m, neg, offset, fail_jump):
    return _pad(struct.pack("<BBBBiI", CHECK_BYTE, and_m, cmp_m, neg, offset, fail_jump))


def i_check_lit(code, lit_offset, lit_length, fail_jump):
    return _pad(struct.pack("<B3xIII", code, lit_offset, lit_length, fail_jump))


def i_report(onmatch, adj=0):
    return _pad(struct.pack("<B3xIi", REPORT, onmatch, adj))


def i_final_report(onmatch, adj=0):
    return _pad(struct.pack("<B3xIi", FINAL_REPORT, onmatch, adj))


def i_dedupe_and_report(dkey, onmatch, fail_jump, adj=0):
    return _pad(struct.pack("<BBxxIIiI", DEDUPE_AND_REPORT, 0, dkey, onmatch, adj, fail_jump))


def i_included_jump(squash, child_offset):
    return _pad(struct.pack("<BBxxI", INCLUDED_JUMP, squash, child_offset))


SZ_MASK, SZ_BYTE, SZ_LIT, SZ_REPORT = len(i_check_mask(0, 0, 0, 0, 0)), len(i_check_byte(0, 0, 0, 0, 0)), \
    len(i_check_lit(0, 0, 0, 0)), len(i_report(0))


def _scan_ref(ref, db, data):
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    r = ref.scan_sorted(db.ptr, a, np.array([0], np.uint64), np.array([a.size], np.uint32))
    return sorted((int(x["id"]), int(x["to"])) for x in r)


def _scan_port(db, data):
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    r = port.scan_sorted(db.ptr, a, np.array([0], np.uint64), np.array([a.size], np.uint32))
    return sorted((int(x["id"]), int(x["to"])) for x in r)


def _scan_dev(hs, db, data, scratch):
    rc, out = hs.scan(db, bytes(data), scratch)
    assert rc == hs.HS_SUCCESS
    return sorted(out)


# ---- ValidateMask known answers (unit/internal/rose_mask.cpp:44-62) ------------------------

TEST_BASIC = [
    (0x1234abcd4321dcba, 0xff09bbdd7f7ffeff, 0x1200abcd4561dcbb, 0xffff00ff),
    (0x56614c6944615465, 0xe0feffffdf7b5480, 0x40614c6946615400, 0xff0000ff000000),
    (0x4d41534b00, 0xfffffefebfdf002c, 0x5536344c0173002c, 0xffffff0000ff00ff),
    (0x464f6f3134666f6f, 0xdfdffffef8c0f000, 0x46466f3030406000, 0xff000000000000),
    (0x464f6f3134666f6f, 0xdfdffffef8c0f000, 0x44464f3034606f60, 0xffffff00ffffffff),
]
NEG_MASKS = [sum(0xff << (8 * i) for i in range(8) if (j >> i) & 1) for j in range(256)]


def _mask_db(hs, t):
    """literal "Z"; its program tries CHECK_MASK with all 256 neg masks and reports
    the index of each one that holds (window = the 8 bytes after the literal)."""
    data, and_m, cmp_m, _ = t
    prog = b""
    for j, nm in enumerate(NEG_MASKS):
        prog += i_check_mask(and_m, cmp_m, nm, 0, SZ_MASK + SZ_REPORT) + i_report(j)
    prog += i_end()
    return hs.compile_programs([b"Z"], [0], [0], prog)


def _mask_expect(t, k):
    """bytes in the future are not checked: valid_data_mask = the low k bytes
    (roseCheckMask, src/rose/program_runtime.c:644-726); the test's own formula:
    output = (truth table & vdm) == (neg_mask & vdm)  (rose_mask.cpp:164-186)"""
    vdm = (1 << (8 * k)) - 1
    return sorted((j, 1) for j, nm in enumerate(NEG_MASKS) if (t[3] & vdm) == (nm & vdm))


@pytest.mark.parametrize("ti", range(len(TEST_BASIC)))
def test_validate_mask_kats_reference_and_port(hs, ref, ti):
    t = TEST_BASIC[ti]
    db = _mask_db(hs, t)
    raw = struct.pack("<Q", t[0])
    for k in range(0, 9):
        buf = b"Z" + raw[:k]
        want = _mask_expect(t, k)
        assert _scan_ref(ref, db, buf) == want, (ti, k)
        assert _scan_port(db, buf) == want, (ti, k)


@pytest.mark.gpu
@pytest.mark.parametrize("ti", range(len(TEST_BASIC)))
def test_validate_mask_kats_device(hs, ref, ti):
    t = TEST_BASIC[ti]
    db = _mask_db(hs, t)
    scratch = hs.Scratch(db)
    raw = struct.pack("<Q", t[0])
    for k in range(0, 9):
        assert _scan_dev(hs, db, b"Z" + raw[:k], scratch) == _mask_expect(t, k), (ti, k)
    scratch.free()


# ---- seeded programs: device == reference == restatement ----------------------------------

STRINGS = [b"abcdefghijklmnop", b"QRSTUVWXYZABCDEFGHIJKLMNOPQRSTUVWXYZAB", b"xyzxyzabcab"]
# filler literals spread the set over several FDR buckets (longer literals get the lower bucket
# ids, src/fdr/fdr_compile.cpp:504-509), so that the parent "cab" is confirmed before its child "ab"
FILL = [bytes(b"mnopqrst"[(i * 7 + j * 3) % 8] for j in range(1 + i % 6)) + bytes([0x30 + i % 10, 0x41 + i % 26])
        for i in range(120)]
FILL = sorted(set(FILL))
LITS = [b"cab", b"ijklmnop", b"uvwxyzab", b"ab"] + FILL
NOCASE = [0, 0, 1, 0] + [0] * len(FILL)


def _emit(seed, base, str_abs, child_abs, squash):
    """Area = four programs, then the strings CHECK_*_LIT compare against.  Sizes
    do not depend on the values, so a first call with zeros yields the offsets."""
    rng = np.random.default_rng(seed)
    area = bytearray()
    offs = []

    def place(b):
        while len(area) % 8:
            area.append(0)
        off = len(area)
        area.extend(b)
        return off

    # literal 0 "cab" includes literal 3 "ab": INCLUDED_JUMP runs the child's program in place and
    # squashes the child's bucket at this position so that it is not confirmed a second time
    szd = len(i_dedupe_and_report(0, 0, 0))
    p = i_check_lit(CHECK_MED_LIT, str_abs[2], len(STRINGS[2]), SZ_LIT + szd + SZ_REPORT)
    p += i_dedupe_and_report(0, 400, szd) + i_report(401)
    p += i_squash_groups(0xffffffffffffffff) + i_included_jump(squash, child_abs) + i_end()
    offs.append(place(p))
    # literal 1: two blocks; the first ends in FINAL_REPORT, the second is reached by fail_jump
    p = i_check_lit(CHECK_LONG_LIT, str_abs[0], len(STRINGS[0]), SZ_LIT + SZ_REPORT) + i_final_report(200)
    offs.append(place(p + i_check_groups(1) + i_report(201) + i_end()))
    # literal 2 (caseless): group 2 is never switched on, 301 never fires
    p = i_check_lit(CHECK_LONG_LIT_NOCASE, str_abs[1], len(STRINGS[1]), SZ_LIT + SZ_REPORT) + i_report(300)
    offs.append(place(p + i_check_groups(2) + i_report(301) + i_end()))
    # literal 3 "ab": CHECK_BYTE / CHECK_MASK at assorted offsets; check j holds -> report 100 + j
    p = b""
    for j in range(24):
        if j % 2 == 0:
            and_m = int(rng.choice([0x01, 0x03, 0x20, 0xdf, 0xff]))
            cmp_m = int(rng.integers(0, 256)) & and_m
            p += i_check_byte(and_m, cmp_m, int(rng.integers(0, 2)), int(rng.integers(-6, 5)), SZ_BYTE + SZ_REPORT)
        else:
            and_m = sum((int(rng.choice([0x01, 0x20, 0x03])) if rng.integers(0, 2) else 0) << (8 * i)
                        for i in range(8))
            cmp_m = int(rng.integers(0, 1 << 62)) & and_m
            neg_m = sum(0xff << (8 * i) for i in range(8) if rng.integers(0, 3) == 0)
            p += i_check_mask(and_m, cmp_m, neg_m, int(rng.integers(-12, 4)), SZ_MASK + SZ_REPORT)
        p += i_report(100 + j)
    offs.append(place(p + i_end()))
    for i in range(len(FILL)):
        offs.append(place(i_report(1000 + i) + i_end()))
    strs = [base + place(x) for x in STRINGS]
    return bytes(area), offs, strs


def _child_bucket(hs, db, child_abs):
    """INCLUDED_JUMP squashes the CHILD literal's bucket bit in the confirm word
    (src/rose/program_runtime.c:3440-3456): find the bucket the builder gave it
    (walk FDRConfirm / LitInfo, src/fdr/fdr_confirm.h:36-94)."""
    bc = db.serialize()[32:]
    fm = struct.unpack_from("<I", bc, 96)[0]
    eng = fm + 192
    conf = eng + struct.unpack_from("<5I", bc, eng)[4]
    for b in range(8):
        cf = struct.unpack_from("<I", bc, conf + 4 * b)[0]
        if not cf:
            continue
        fc = conf + cf
        nbits = struct.unpack_from("<QQI", bc, fc)[2]
        for c in range(1 << nbits):
            st = struct.unpack_from("<I", bc, fc + 32 + 4 * c)[0]
            li = fc + st
            while st:
                v, msk, groups, lid, sz, fl, nxt = struct.unpack_from("<QQQIBBB", bc, li)
                if lid == child_abs:
                    return b
                if not nxt:
                    break
                li += 32
    raise AssertionError("child literal not found")


def _build(hs, seed):
    # INCLUDED_JUMP acts on FDR's confirm word; Teddy hands confWithBit a dummy
    # (src/fdr/teddy_runtime_common.h:436-439), so the FDR engine is forced
    hs.set_build_option("force_engine", 0)
    base = hs.test_program_base()
    _, offs, strs = _emit(seed, base, [0, 0, 0], 0, 0)
    child = base + offs[3]
    area, offs2, strs2 = _emit(seed, base, strs, child, 0)
    assert offs2 == offs and strs2 == strs
    probe = hs.compile_programs(LITS, NOCASE, offs, area, inv_dkey=[400])
    assert probe.info().engine_id == 0
    cb, pb = _child_bucket(hs, probe, child), _child_bucket(hs, probe, base + offs[0])
    assert cb > pb     # buckets are confirmed in ascending order: the squash must come before the child's turn
    squash = 1 << cb
    area, _, _ = _emit(seed, base, strs, child, squash)
    return hs.compile_programs(LITS, NOCASE, offs, area, inv_dkey=[400]), list(STRINGS)


def _corpus(seed, strings):
    rng = np.random.default_rng(seed + 1000)
    alpha = np.frombuffer(b"abcxyzABQ", dtype=np.uint8)
    data = alpha[rng.integers(0, alpha.size, size=6000)].copy()
    for s_ in strings + [b"qrstuvwxyzabcdefghijklmnopqrstuvwxyzAB", b"cab", b"abcdefghijklmnop"] + FILL[::9]:
        for _ in range(6):
            p = int(rng.integers(0, data.size - len(s_)))
            data[p:p + len(s_)] = np.frombuffer(s_, dtype=np.uint8)
    return data


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_assembled_programs_reference_and_port(hs, ref, seed):
    db, strings = _build(hs, seed)
    data = _corpus(seed, strings)
    want = _scan_ref(ref, db, data)
    ids = {i for i, _ in want}
    assert 200 in ids and 300 in ids and 400 in ids and 401 in ids and 301 not in ids
    assert len([i for i in ids if 100 <= i < 200]) >= 8          # a good share of the byte/mask checks hold
    assert _scan_port(db, data) == want
    for cut in (1, 2, 3, 9, 17, 100):                            # block edges: "too early" / "in the future"
        assert _scan_port(db, data[:cut]) == _scan_ref(ref, db, data[:cut])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_assembled_programs_device(hs, ref, seed):
    db, strings = _build(hs, seed)
    data = _corpus(seed, strings)
    scratch = hs.Scratch(db)
    assert _scan_dev(hs, db, data, scratch) == _scan_ref(ref, db, data)
    for cut in (1, 2, 3, 9, 17, 100, 2049):
        assert _scan_dev(hs, db, data[:cut], scratch) == _scan_ref(ref, db, data[:cut]), cut
    # many blocks at once through the batched entry
    n = 37
    off = (np.arange(n, dtype=np.uint64) * 160)
    ln = np.full(n, 150, dtype=np.uint32)
    got = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
    want = ref.scan_sorted(db.ptr, data, off, ln)
    assert np.array_equal(got, want)
    scratch.free()


def test_exhaustible_reports_found_in_later_program_blocks(hs, ref):
    """A program of several blocks, each ending in FINAL_REPORT and reached by
    fail_jump: the walk that looks for REPORT_EXHAUST must not stop at the first
    terminator (round-1 advisor finding on db_walk.cpp)."""
    import ctypes as C
    blk1 = i_check_byte(0xff, ord("x"), 0, -3, SZ_BYTE + SZ_REPORT) + i_final_report(7)
    rex = _pad(struct.pack("<B3xIiI", REPORT_EXHAUST, 9, 0, 0))
    cex = _pad(struct.pack("<B3xII", CHECK_EXHAUSTED, 0, len(_pad(b"x" * 12)) + len(rex)))
    prog = blk1 + cex + rex + i_end()
    db = hs.compile_programs([b"ab"], [0], [0], prog, ekey_count=1)
    recs = np.zeros(4, dtype=[("id", "<u4"), ("block", "<u4"), ("to", "<u8")])
    recs["id"] = [9, 9, 7, 9]
    recs["to"] = [5, 9, 3, 7]
    n = C.c_ulonglong()
    assert hs.lib().hs_b200_postprocess_matches(db.ptr, recs.ctypes.data, 4, C.byref(n)) == 0
    assert n.value == 2 and [(int(r["id"]), int(r["to"])) for r in recs[:2]] == [(7, 3), (9, 5)]
    data = b"yab..xab...ab..ab"
    assert _scan_ref(ref, db, data) == [(7, 8), (9, 3)]
    assert _scan_port(db, data) == [(7, 8), (9, 3)]


@pytest.mark.gpu
def test_exhaustible_later_block_device(hs, ref):
    blk1 = i_check_byte(0xff, ord("x"), 0, -3, SZ_BYTE + SZ_REPORT) + i_final_report(7)
    rex = _pad(struct.pack("<B3xIiI", REPORT_EXHAUST, 9, 0, 0))
    cex = _pad(struct.pack("<B3xII", CHECK_EXHAUSTED, 0, len(_pad(b"x" * 12)) + len(rex)))
    db = hs.compile_programs([b"ab"], [0], [0], blk1 + cex + rex + i_end(), ekey_count=1)
    scratch = hs.Scratch(db)
    assert _scan_dev(hs, db, b"yab..xab...ab..ab", scratch) == [(7, 8), (9, 3)]
    scratch.free()


@pytest.mark.gpu
@pytest.mark.parametrize("code,size", [(PUSH_DELAYED, 8), (REPORT_CHAIN, 24), (SET_EXHAUST, 8)])
def test_state_carrying_opcodes_are_refused_at_alloc(hs, code, size):
    """Delayed literals, chained reports and logical-combination state are not
    modelled on the device: such a database is refused when the scratch is
    allocated (HS_ARCH_ERROR), not in the middle of a scan."""
    prog = _pad(struct.pack("<B", code) + b"\0" * (size - 1)) + i_report(1) + i_end()
    db = hs.compile_programs([b"ab"], [0], [0], prog, ekey_count=1)
    with pytest.raises(hs.HsError) as e:
        hs.Scratch(db)
    assert e.value.code == hs.HS_ARCH_ERROR
