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
PUSH_DELAYED, DEDUPE, REPORT_CHAIN, REPORT, REPORT_EXHAUST = 18, 28, 30, 33, 34
CHECK_MASK_32, CHECK_MASK_64 = 10, 69
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


def i_check_mask_32(and_m, cmp_m, neg_m, offset, fail_jump):
    """ROSE_STRUCT_CHECK_MASK_32 (src/rose/rose_program.h:283-290): one negation BIT per byte"""
    return _pad(struct.pack("<B32s32s3xIiI", CHECK_MASK_32, bytes(and_m), bytes(cmp_m), neg_m, offset, fail_jump))


def i_check_mask_64(and_m, cmp_m, neg_m, offset, fail_jump):
    """ROSE_STRUCT_CHECK_MASK_64 (src/rose/rose_program.h:292-299)"""
    return _pad(struct.pack("<B64s64s7xQiI", CHECK_MASK_64, bytes(and_m), bytes(cmp_m), neg_m, offset, fail_jump))


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


# ---- ValidateMask32 / 64 known answers (unit/internal/rose_mask_32.cpp:57-131) ------------

# (index, data, and_mask, cmp_mask, negated) per line, as in testBasicIdx
TEST_BASIC_32 = [
    [(1, 0x34, 0xf8, 0x30, 0), (2, 0x34, 0xf8, 0x30, 0), (8, 0x23, 0xff, 0x23, 0), (9, 0x34, 0xf8, 0x30, 0),
     (10, 0x41, 0xdf, 0x41, 0), (11, 0x63, 0xdd, 0x41, 0), (12, 0x61, 0xdd, 0x41, 0), (13, 0x41, 0xdf, 0x41, 0),
     (14, 0x61, 0xdf, 0x41, 0), (15, 0x41, 0xdf, 0x41, 0), (16, 0x43, 0xdd, 0x41, 0), (17, 0x61, 0xdd, 0x41, 0),
     (23, 0x63, 0xdd, 0x41, 0), (24, 0x4f, 0xfc, 0x4c, 0), (25, 0x4d, 0xfc, 0x4c, 0), (26, 0x4d, 0xfc, 0x4c, 0)],
    [(11, 0, 0xff, 0x55, 1), (12, 0, 0xff, 0x36, 1), (13, 0, 0xfe, 0x34, 1), (14, 0x4d, 0xfe, 0x4c, 0),
     (15, 0x41, 0xbf, 0x01, 0), (16, 0x53, 0xdf, 0x73, 1), (17, 0x4b, 0, 0, 0), (18, 0, 0x2c, 0x2c, 1)],
    [(15, 0x46, 0xdf, 0x46, 0), (16, 0x4f, 0xdf, 0x46, 1), (17, 0x6f, 0xff, 0x6f, 0), (18, 0x31, 0xfe, 0x30, 0),
     (19, 0x34, 0xf8, 0x30, 0), (20, 0x66, 0xc0, 0x40, 0), (21, 0x6f, 0xf0, 0x60, 0), (22, 0x6f, 0, 0, 0),
     (23, 0x46, 0xdf, 0x44, 1), (24, 0x4f, 0xdf, 0x46, 1), (25, 0x6f, 0xff, 0x4f, 1), (26, 0x31, 0xfe, 0x30, 0),
     (27, 0x34, 0xf8, 0x34, 1), (28, 0x66, 0xc0, 0x60, 1), (29, 0x6f, 0xf0, 0x6f, 1), (30, 0x6f, 0, 0x60, 1)],
    [(31, 0x4a, 0x80, 0, 0)],
    [(12, 0x2b, 0x3d, 0x2d, 1), (13, 0x2b, 0x3d, 0x4c, 1), (23, 0x4a, 0x88, 0x0a, 1)],
]


def _wide_case(ti, width):
    """data / and / cmp arrays of `width` bytes and the truth mask (bit i = byte i differs);
    width 64: the 32-byte vectors twice, the second copy with the bytes reversed"""
    data, and_m, cmp_m = bytearray(width), bytearray(width), bytearray(width)
    for idx, d, a, c, _ in TEST_BASIC_32[ti]:
        for pos in ([idx] if width == 32 else [idx, 63 - idx]):
            data[pos], and_m[pos], cmp_m[pos] = d, a, c
    truth = sum(1 << i for i in range(width) if (data[i] & and_m[i]) != cmp_m[i])
    return data, and_m, cmp_m, truth


def _wide_negs(truth, width, seed):
    rng = np.random.default_rng(seed)
    full = (1 << width) - 1
    negs = [0, full, truth, truth ^ 1, truth ^ (1 << (width - 1))]
    negs += [truth ^ (1 << int(rng.integers(0, width))) for _ in range(8)]
    negs += [int(rng.integers(0, 1 << 32)) | (int(rng.integers(0, 1 << 32)) << 32 if width == 64 else 0)
             for _ in range(16)]
    return [n & full for n in negs]


def _wide_db(hs, ti, width, offset=0):
    data, and_m, cmp_m, truth = _wide_case(ti, width)
    negs = _wide_negs(truth, width, 451 + ti)
    ins = i_check_mask_32 if width == 32 else i_check_mask_64
    sz = len(ins(and_m, cmp_m, 0, 0, 0))
    prog = b"".join(ins(and_m, cmp_m, nm, offset, sz + SZ_REPORT) + i_report(j) for j, nm in enumerate(negs))
    return hs.compile_programs([b"Z"], [0], [0], prog + i_end()), data, truth, negs


def _wide_expect(truth, negs, k, to=1):
    """rose_mask_32.cpp:148-170: passes iff (cmp_result & valid) == (neg_mask & valid); in block
    mode the valid bytes are the first k of the window (the rest lie in the future)"""
    vdm = (1 << k) - 1
    return sorted((j, to) for j, nm in enumerate(negs) if (truth & vdm) == (nm & vdm))


def _isa_has_mask64(ref):
    return ref.best_isa().startswith("avx512")      # L_PROGRAM_CASE(CHECK_MASK_64) is HAVE_AVX512 only


@pytest.mark.parametrize("width", [32, 64])
@pytest.mark.parametrize("ti", range(len(TEST_BASIC_32)))
def test_validate_mask_wide_kats_reference_and_port(hs, ref, ti, width):
    db, data, truth, negs = _wide_db(hs, ti, width)
    assert truth == sum(neg << idx for idx, _, _, _, neg in TEST_BASIC_32[ti]) or width == 64   # testMask32_1
    for k in list(range(0, width + 1)) + [width + 5]:
        buf = b"Z" + bytes(data[:k]) + b"\x01" * max(0, k - width)
        want = _wide_expect(truth, negs, min(k, width))
        assert _scan_port(db, buf) == want, (ti, k)
        if width == 32 or _isa_has_mask64(ref):
            assert _scan_ref(ref, db, buf) == want, (ti, k)
    # a window that starts before the buffer fails whatever the masks say ("too early")
    db2, data, truth, negs = _wide_db(hs, ti, width, offset=-3)
    for buf in (b"Z" + bytes(data[2:]), b"xZ" + bytes(data[1:]), b"xyZ" + bytes(data), b"xyzZ" + bytes(data[1:])):
        want = _scan_port(db2, buf)
        z = buf.index(b"Z") + 1
        assert want == ([] if z < 3 else _wide_expect(sum(
            1 << i for i in range(min(width, len(buf) - z + 3)) if
            (buf[z - 3 + i] & _wide_case(ti, width)[1][i]) != _wide_case(ti, width)[2][i]), negs,
            min(width, len(buf) - z + 3), to=z))
        if width == 32 or _isa_has_mask64(ref):
            assert _scan_ref(ref, db2, buf) == want


@pytest.mark.gpu
@pytest.mark.parametrize("width", [32, 64])
@pytest.mark.parametrize("ti", range(len(TEST_BASIC_32)))
def test_validate_mask_wide_kats_device(hs, ref, ti, width):
    db, data, truth, negs = _wide_db(hs, ti, width)
    scratch = hs.Scratch(db)
    for k in list(range(0, width + 1)) + [width + 5]:
        buf = b"Z" + bytes(data[:k]) + b"\x01" * max(0, k - width)
        assert _scan_dev(hs, db, buf, scratch) == _wide_expect(truth, negs, min(k, width)), (ti, k)
    scratch.free()
    db2, data, truth, negs = _wide_db(hs, ti, width, offset=-3)
    scratch = hs.Scratch(db2)
    for buf in (b"Z" + bytes(data[2:]), b"xZ" + bytes(data[1:]), b"xyZ" + bytes(data), b"xyzZ" + bytes(data[1:])):
        assert _scan_dev(hs, db2, buf, scratch) == _scan_port(db2, buf)
    scratch.free()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_random_wide_mask_programs_device(hs, ref, seed):
    """seeded CHECK_MASK_32 / _64 at assorted offsets on a corpus with many hits: the device,
    the restatement and (where its build has the opcode) the reference agree"""
    rng = np.random.default_rng(seed)
    prog = b""
    for j in range(20):
        width = 32 if j % 2 == 0 else 64
        live = set(int(x) for x in rng.choice(width, size=3, replace=False))   # three constrained bytes
        and_m = bytes(int(rng.choice([0x01, 0x20, 0x03])) if i in live else 0 for i in range(width))
        cmp_m = bytes(int(rng.choice([0x61, 0x62, 0x41, 0x42])) & a for a in and_m)
        neg = sum(1 << i for i in live if rng.integers(0, 2))
        ins = i_check_mask_32 if width == 32 else i_check_mask_64
        sz = len(ins(and_m, cmp_m, 0, 0, 0))
        prog += ins(and_m, cmp_m, neg, int(rng.integers(-40, 8)), sz + SZ_REPORT) + i_report(100 + j)
    db = hs.compile_programs([b"ab"], [0], [0], prog + i_end())
    alpha = np.frombuffer(b"abAB", dtype=np.uint8)
    data = alpha[rng.integers(0, alpha.size, size=3000)].tobytes()
    scratch = hs.Scratch(db)
    want = _scan_port(db, data)
    assert len(want) > 50
    assert _scan_dev(hs, db, data, scratch) == want
    if _isa_has_mask64(ref):
        assert _scan_ref(ref, db, data) == want
    for cut in (1, 2, 3, 9, 33, 65, 100):
        assert _scan_dev(hs, db, data[:cut], scratch) == _scan_port(db, data[:cut]), cut
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
