/*
 * hs_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product (hyperscan_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * A plain scalar C restatement of the reference's block-mode literal scan path
 * (intel/hyperscan 5.4.2), operating on reference-format databases:
 *
 *   oracle_scan_collect   hs_scan()                    src/runtime.c:316-475
 *                         pureLiteralBlockExec         src/runtime.c:204-220
 *   hwlm_exec             hwlmExec                     src/hwlm/hwlm.c:172-199
 *   nood_exec             noodExec / final             src/hwlm/noodle_engine.c:114-141,374-442
 *   fdr_exec              fdr_engine_exec              src/fdr/fdr.c:157-364,694-790
 *   teddy_exec            fdr_exec_teddy_msks1..4      src/fdr/teddy.c:918-1064
 *                         (fat: src/fdr/teddy_avx2.c:395-447)
 *   conf_with_bit         confWithBit                  src/fdr/fdr_confirm_runtime.h:43-102
 *   rose_callback         roseCallback_i               src/rose/match.c:479-523
 *   run_program_l         roseRunProgram_l             src/rose/program_runtime.c:3101-3522
 *   dedupe                dedupeCatchup                src/report.h:55-119
 *   deliver_report        roseDeliverReport            src/report.h:301-337
 *
 * Parity pinning: tests/test_oracle_kat.py checks this file against the
 * reference's own known-answer tests (unit/internal/fdr.cpp, noodle.cpp --
 * SURVEY.md A.5) and, where oracle/_ref is built, against the unmodified
 * reference runtime on seeded random inputs (same callbacks in the same order).
 *
 * The SIMD first stages of the reference are restated as their scalar
 * meaning (one position at a time); zones, flood detection and acceleration are
 * optimisations with no effect on the callback sequence and are not restated.
 */
#define _GNU_SOURCE
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t s32;

#define API __attribute__((visibility("default")))
#define ROUNDUP8(x) (((x) + 7u) & ~7u)

struct rec16 {
    u32 id;
    u32 block;
    u64 to;
};

/* ---- layouts (all little-endian, natural alignment) ------------------------- */

struct db_header { /* struct hs_database, src/database.h:102-113 */
    u32 magic, version, length;
    u64 platform;
    u32 crc32, reserved0, reserved1, bytecode;
    u32 padding[16];
};

/* offsets inside struct RoseEngine (src/rose/rose_internal.h:330-496), pinned by
 * tests/golden/ref_layout.json */
enum {
    RE_runtimeImpl = 4, RE_canExhaust = 6, RE_mode = 12, RE_ekeyCount = 20, RE_dkeyCount = 44,
    RE_fmatcherOffset = 96, RE_minWidth = 200, RE_initialGroups = 240, RE_floating_group_mask = 248
};

static u32 rd32(const u8 *p) { u32 v; memcpy(&v, p, 4); return v; }
static u64 rd64(const u8 *p) { u64 v; memcpy(&v, p, 8); return v; }
static s32 rds32(const u8 *p) { s32 v; memcpy(&v, p, 4); return v; }

/* ---- scan context -------------------------------------------------------------- */

typedef int (*user_cb)(unsigned id, unsigned long long from, unsigned long long to, unsigned flags,
                       void *ctx);

struct scan {
    const u8 *rose;      /* bytecode; NULL for a bare HWLM run */
    const u8 *buf;
    size_t len;
    u64 groups;          /* "control": callback feedback (src/rose/match.c:516-517) */
    int terminated;
    /* rose state */
    u8 *evec;            /* exhaustion bits */
    u32 ekeyCount, dkeyCount;
    u8 *dlog[2];         /* dedupe logs, alternating by offset parity */
    u64 dedupe_offset;
    user_cb cb;
    void *cb_ctx;
    /* bare-HWLM recording */
    struct rec16 *out;
    size_t cap, n, stop_after;
    u32 block;
    u32 last_match;      /* FDR_LIT_FLAG_NOREPEAT state (fdr.c:737 last_match_id) */
    u64 base_offset;     /* streaming: stream offset of buf[0] (history included) */
};

/* ---- rose literal programs --------------------------------------------------------- */

enum {
    OP_END = 0, OP_CHECK_GROUPS = 3, OP_CHECK_MASK = 9, OP_CHECK_MASK_32 = 10, OP_CHECK_BYTE = 11, OP_CHECK_MASK_64 = 69, OP_DEDUPE = 28,
    OP_REPORT = 33, OP_REPORT_EXHAUST = 34, OP_DEDUPE_AND_REPORT = 37, OP_FINAL_REPORT = 38,
    OP_CHECK_EXHAUSTED = 39, OP_SQUASH_GROUPS = 43, OP_CHECK_LONG_LIT = 51, OP_CHECK_LONG_LIT_NOCASE = 52,
    OP_CHECK_MED_LIT = 53, OP_CHECK_MED_LIT_NOCASE = 54, OP_CLEAR_WORK_DONE = 55, OP_INCLUDED_JUMP = 61,
    OP_SET_EXHAUST = 65
};

static int bit_test_set(u8 *v, u32 k) {
    const int was = (v[k >> 3] >> (k & 7)) & 1;
    v[k >> 3] |= (u8)(1u << (k & 7));
    return was;
}
static int bit_test(const u8 *v, u32 k) { return (v[k >> 3] >> (k & 7)) & 1; }

static int all_exhausted(const struct scan *s) { /* isAllExhausted, src/report.h:131-138 */
    if (!s->rose[RE_canExhaust]) {
        return 0;
    }
    for (u32 k = 0; k < s->ekeyCount; k++) {
        if (!bit_test(s->evec, k)) {
            return 0;
        }
    }
    return 1;
}

/* dedupeCatchup (src/report.h:55-119), external reports without SOM. */
static int dedupe(struct scan *s, u64 offset, u64 to_offset, u32 dkey) {
    const size_t bytes = (s->dkeyCount + 7) / 8;
    if (offset != s->dedupe_offset) {
        if (offset == s->dedupe_offset + 1) {
            memset(s->dlog[offset % 2], 0, bytes);
        } else {
            memset(s->dlog[0], 0, bytes);
            memset(s->dlog[1], 0, bytes);
        }
        s->dedupe_offset = offset;
    }
    return bit_test_set(s->dlog[to_offset % 2], dkey); /* 1 = duplicate: skip */
}

/* roseDeliverReport + roseReport (src/report.h:301-337,
 * src/rose/program_runtime.c:464-481).  Returns 0 to halt. */
static int deliver_report(struct scan *s, u64 end, u32 onmatch, s32 adj, u32 ekey) {
    if (s->cb && s->cb(onmatch, 0, s->base_offset + end + adj, 0, s->cb_ctx)) {
        s->terminated = 1;
        return 0;
    }
    if (ekey != 0xffffffffu) {
        bit_test_set(s->evec, ekey);
        if (all_exhausted(s)) { /* roseHaltIfExhausted, src/rose/match.h:303 */
            return 0;
        }
    }
    return 1;
}

static u8 upper(u8 c) { return (c >= 'a' && c <= 'z') ? (u8)(c - 32) : c; }

/* roseCheckMediumLiteral / roseCheckLongLiteral in block mode
 * (src/rose/program_runtime.c:1883-2014). */
static int check_lit(const struct scan *s, u64 end, u32 off, u32 len, int nocase) {
    if (end < len) {
        return 0;
    }
    const u8 *lit = s->rose + off, *d = s->buf + end - len;
    for (u32 i = 0; i < len; i++) {
        const u8 c = nocase ? upper(d[i]) : d[i];
        if (c != lit[i]) {
            return 0;
        }
    }
    return 1;
}

/* roseCheckByte (src/rose/program_runtime.c:600-641), block mode. */
static int check_byte(const struct scan *s, u64 end, u8 and_mask, u8 cmp_mask, u8 neg, s32 off) {
    if (off < 0 && (u64)(0 - (long long)off) > end) {
        return 0;
    }
    const long long q = (long long)end + off;
    if (q >= (long long)s->len) {
        return 1;
    }
    return !(((and_mask & s->buf[q]) != cmp_mask) ^ (neg != 0));
}

/* roseCheckMask + validateMask (src/rose/program_runtime.c:644-726,
 * src/rose/validate_mask.h:83-103), block mode. */
static int check_mask(const struct scan *s, u64 end, u64 and_mask, u64 cmp_mask, u64 neg_mask, s32 off) {
    if (off < 0 && (u64)(0 - (long long)off) > end) {
        return 0;
    }
    const long long start = (long long)end + off;
    for (int i = 0; i < 8; i++) {
        const long long q = start + i;
        if (q < 0 || q >= (long long)s->len) {
            continue; /* lane outside the buffer: not validated */
        }
        const u8 r = (u8)((s->buf[q] & (u8)(and_mask >> (8 * i))) ^ (u8)(cmp_mask >> (8 * i)));
        const int neg = ((neg_mask >> (8 * i)) & 0xff) != 0;
        if ((r == 0) == neg) {
            return 0;
        }
    }
    return 1;
}

/* roseCheckMask32 / roseCheckMask64 + validateMask32 / 64 (src/rose/program_runtime.c:729-801,
 * :805-877; src/rose/validate_mask.h:106-153), block mode: n bytes at end + off, one
 * negation bit per byte, bytes past the buffer are masked out of the comparison. */
static int check_mask_wide(const struct scan *s, u64 end, const u8 *and_mask, const u8 *cmp_mask, u64 neg_mask,
                           s32 off, int n) {
    if (off < 0 && (u64)(0 - (long long)off) > end) {
        return 0; /* too early */
    }
    const long long start = (long long)end + off;
    u64 cmp_result = 0, valid = 0;
    for (int i = 0; i < n; i++) {
        const long long q = start + i;
        if (q < 0 || q >= (long long)s->len) {
            continue;
        }
        valid |= 1ull << i;
        if ((s->buf[q] & and_mask[i]) != cmp_mask[i]) {
            cmp_result |= 1ull << i;
        }
    }
    return (cmp_result & valid) == (neg_mask & valid);
}

/* roseRunProgram_l: the pure-literal interpreter.  `end` is the offset after
 * the literal's last byte.  Returns 0 to halt matching. */
static int run_program_l(struct scan *s, u32 prog, u64 end) {
    const u8 *pc = s->rose + prog;
    for (;;) {
        switch (*pc) {
        case OP_END:
            return 1;
        case OP_CHECK_GROUPS: /* {u8; u64 groups} */
            if (!(rd64(pc + 8) & s->groups)) {
                return 1;
            }
            pc += 16;
            break;
        case OP_CHECK_MASK: /* {u8; u64 and,cmp,neg; s32 offset; u32 fail_jump} */
            if (!check_mask(s, end, rd64(pc + 8), rd64(pc + 16), rd64(pc + 24), rds32(pc + 32))) {
                pc += rd32(pc + 36);
            } else {
                pc += 40;
            }
            break;
        case OP_CHECK_MASK_32: /* {u8; u8 and[32], cmp[32]; u32 neg; s32 offset; u32 fail_jump} */
            if (!check_mask_wide(s, end, pc + 1, pc + 33, rd32(pc + 68), rds32(pc + 72), 32)) {
                pc += rd32(pc + 76);
            } else {
                pc += 80;
            }
            break;
        case OP_CHECK_MASK_64: /* {u8; u8 and[64], cmp[64]; u64 neg; s32 offset; u32 fail_jump} */
            if (!check_mask_wide(s, end, pc + 1, pc + 65, rd64(pc + 136), rds32(pc + 144), 64)) {
                pc += rd32(pc + 148);
            } else {
                pc += 152;
            }
            break;
        case OP_CHECK_BYTE: /* {u8 code,and,cmp,neg; s32 offset; u32 fail_jump} */
            if (!check_byte(s, end, pc[1], pc[2], pc[3], rds32(pc + 4))) {
                pc += rd32(pc + 8);
            } else {
                pc += 16;
            }
            break;
        case OP_CHECK_LONG_LIT:
        case OP_CHECK_LONG_LIT_NOCASE:
        case OP_CHECK_MED_LIT:
        case OP_CHECK_MED_LIT_NOCASE: { /* {u8; u32 lit_offset, lit_length, fail_jump} */
            const int nc = *pc == OP_CHECK_LONG_LIT_NOCASE || *pc == OP_CHECK_MED_LIT_NOCASE;
            if (!check_lit(s, end, rd32(pc + 4), rd32(pc + 8), nc)) {
                pc += rd32(pc + 12);
            } else {
                pc += 16;
            }
            break;
        }
        case OP_CHECK_EXHAUSTED: /* {u8; u32 ekey; u32 fail_jump} */
            if (bit_test(s->evec, rd32(pc + 4))) {
                pc += rd32(pc + 8);
            } else {
                pc += 16;
            }
            break;
        case OP_DEDUPE: { /* {u8 code,quash_som; u32 dkey; s32 offset_adjust; u32 fail_jump} */
            const s32 adj = rds32(pc + 8);
            if (dedupe(s, end, end + adj, rd32(pc + 4))) {
                pc += rd32(pc + 12);
            } else {
                pc += 16;
            }
            break;
        }
        case OP_REPORT: /* {u8; u32 onmatch; s32 offset_adjust} */
            if (!deliver_report(s, end, rd32(pc + 4), rds32(pc + 8), 0xffffffffu)) {
                return 0;
            }
            pc += 16;
            break;
        case OP_REPORT_EXHAUST: /* {u8; u32 onmatch; s32 offset_adjust; u32 ekey} */
            if (!deliver_report(s, end, rd32(pc + 4), rds32(pc + 8), rd32(pc + 12))) {
                return 0;
            }
            pc += 16;
            break;
        case OP_DEDUPE_AND_REPORT: { /* {u8 code,quash; u32 dkey, onmatch; s32 adj; u32 fail_jump} */
            const s32 adj = rds32(pc + 12);
            if (dedupe(s, end, end + adj, rd32(pc + 4))) {
                pc += rd32(pc + 16);
                break;
            }
            if (!deliver_report(s, end, rd32(pc + 8), adj, 0xffffffffu)) {
                return 0;
            }
            pc += 24;
            break;
        }
        case OP_FINAL_REPORT:
            /* "One-shot specialisation: this instruction always terminates execution
             * of the program" -- of the PROGRAM, matching goes on
             * (src/rose/program_runtime.c:3349-3358) */
            return deliver_report(s, end, rd32(pc + 4), rds32(pc + 8), 0xffffffffu) ? 1 : 0;
        case OP_SQUASH_GROUPS: /* {u8; u64 groups}: groups &= mask */
            s->groups &= rd64(pc + 8);
            pc += 16;
            break;
        case OP_CLEAR_WORK_DONE:
            pc += 8;
            break;
        case OP_INCLUDED_JUMP: /* optimisation hint for the literal matcher */
            pc += 8;
            break;
        case OP_SET_EXHAUST:
            bit_test_set(s->evec, rd32(pc + 4));
            pc += 8;
            break;
        default:
            s->terminated = 2; /* unknown instruction */
            return 0;
        }
    }
}

/* The HWLM callback.  Rose: roseCallback_i (src/rose/match.c:479-523): run
 * the literal's program at real_end = end + 1; returns the new control
 * groups (0 = terminate).  Bare HWLM runs record (end, id). */
static u64 hwlm_cb(struct scan *s, size_t end, u32 id) {
    if (!s->rose) {
        if (s->n < s->cap) {
            s->out[s->n].id = id;
            s->out[s->n].block = s->block;
            s->out[s->n].to = end;
        }
        s->n++;
        if (s->stop_after && s->n >= s->stop_after) {
            return 0;
        }
        return s->groups;
    }
    if (!run_program_l(s, id, (u64)end + 1)) {
        return 0;
    }
    return s->groups;
}

/* ---- confirm ------------------------------------------------------------------------ */

/* confWithBit (src/fdr/fdr_confirm_runtime.h:43-102).  i = index of the last
 * byte; conf_key = LE u64 of buf[i-7..i] (bytes before the buffer read as 0:
 * block mode has no history). */
static void conf_with_bit(struct scan *s, const u8 *fdrc, size_t i, u64 conf_key) {
    const u64 andmsk = rd64(fdrc + 0), mult = rd64(fdrc + 8);
    const u32 nBits = rd32(fdrc + 16);
    const u32 c = (u32)(((conf_key & andmsk) * mult) >> (64 - nBits));
    const u32 start = rd32(fdrc + 32 + 4 * c);
    if (!start) {
        return;
    }
    const u8 *li = fdrc + start; /* struct LitInfo {v,msk,groups,id,size,flags,next} */
    u8 next;
    do {
        next = li[30];
        const u32 id = rd32(li + 24);
        if ((conf_key & rd64(li + 8)) != rd64(li + 0)) {
            goto out;
        }
        if (s->last_match == id && (li[29] & 1)) { /* FDR_LIT_FLAG_NOREPEAT */
            goto out;
        }
        if ((size_t)li[28] > i + 1) { /* literal would start before the buffer */
            goto out;
        }
        if (!(rd64(li + 16) & s->groups)) {
            goto out;
        }
        s->last_match = id;
        s->groups = hwlm_cb(s, i, id);
    out:
        li += 32;
    } while (next && s->groups);
}

static u64 conf_key_at(const struct scan *s, size_t i) {
    u64 v = 0;
    for (int z = 0; z < 8; z++) {
        const long long q = (long long)i - 7 + z;
        if (q >= 0) {
            v |= (u64)s->buf[q] << (8 * z);
        }
    }
    return v;
}

/* do_confirm_fdr / do_confWithBit_teddy: all candidate buckets at end position
 * i, lowest bucket first (src/fdr/fdr.c:330-364). */
static void confirm_position(struct scan *s, const u8 *confBase, size_t i, u32 buckets, u32 nbuckets) {
    for (u32 b = 0; b < nbuckets && s->groups; b++) {
        if (!(buckets & (1u << b))) {
            continue;
        }
        const u32 cf = rd32(confBase + 4 * b);
        if (!cf) {
            continue;
        }
        const u8 *fdrc = confBase + cf;
        if (!(rd64(fdrc + 24) & s->groups)) {
            continue;
        }
        conf_with_bit(s, fdrc, i, conf_key_at(s, i));
    }
}

/* ---- FDR ------------------------------------------------------------------------------- */

/* fdr_engine_exec as a scalar recurrence: `st` holds, per future end position
 * (byte lane) and bucket (bit), the OR of the table entries sampled so far;
 * a zero bit in lane 0 at position i is a candidate end at i
 * (src/fdr/fdr.c:157-327,694-723). */
static void fdr_exec(struct scan *s, const u8 *fdr, size_t start) {
    const u32 confOffset = rd32(fdr + 16);
    const u32 stride = fdr[24];
    const u32 dmask = (u32)fdr[26] | ((u32)fdr[27] << 8);
    const u8 *ft = fdr + 64; /* ROUNDUP_CL(sizeof(struct FDR)) */
    const u8 *confBase = fdr + confOffset;
    u64 st = rd64(fdr + 32); /* struct FDR.start: low half of the initial state */
    for (size_t i = 0; i < s->len && s->groups; i++) {
        if (i % stride == 0) {
            u32 h = s->buf[i];
            if (i + 1 < s->len) {
                h |= (u32)s->buf[i + 1] << 8;
            }
            st |= rd64(ft + 8 * (size_t)(h & dmask));
        }
        const u32 cand = (u32)(~st & 0xff);
        if (cand && i >= start) {
            confirm_position(s, confBase, i, cand, 8);
        }
        st >>= 8;
    }
}

/* ---- Teddy ----------------------------------------------------------------------------- */

/* prep_conf_teddy_mN / fat variants: candidate at end e for bucket b iff for
 * every mask m < numMasks the byte at e-m passes both nibble tables
 * (src/fdr/teddy.c:918-969; mask layout src/fdr/teddy_runtime_common.h:441,
 * src/fdr/teddy_compile.cpp:440-509).  Positions before the buffer start put
 * no constraint (confirm rejects literals that do not fit). */
static void teddy_exec(struct scan *s, const u8 *teddy, size_t start) {
    const u32 id = rd32(teddy + 0);
    const u32 confOffset = rd32(teddy + 16);
    const u32 nm = ((id - 3) % 8) / 2 + 1;
    const u32 oct = id <= 10 ? 2 : 1;
    const u8 *mb = teddy + 64;
    const u8 *confBase = teddy + confOffset;
    for (size_t e = 0; e < s->len && s->groups; e++) {
        u32 impossible = 0;
        for (u32 m = 0; m < nm && m <= e; m++) {
            const u8 c = s->buf[e - m];
            for (u32 o = 0; o < oct; o++) {
                const u8 *lo = mb + ((2 * m) * oct + o) * 16, *hi = mb + ((2 * m + 1) * oct + o) * 16;
                impossible |= (u32)(u8)(lo[c & 15] | hi[c >> 4]) << (8 * o);
            }
        }
        const u32 cand = ~impossible & (oct == 2 ? 0xffffu : 0xffu);
        if (cand && e >= start) {
            confirm_position(s, confBase, e, cand, 8 * oct);
        }
    }
}

/* ---- noodle ---------------------------------------------------------------------------- */

/* noodExec + final(): for every position pos with buf[pos] == key0 (and
 * buf[pos+1] == key1 unless single; caseless keys compare under 0xdf), the
 * msk_len bytes ending at pos + key_offset - 1 are compared under msk
 * (src/hwlm/noodle_engine.c:114-141,155-260,374-442). */
static void nood_exec(struct scan *s, const u8 *n, size_t start) {
    const u32 id = rd32(n + 0);
    const u64 msk = rd64(n + 8), cmp = rd64(n + 16);
    const u32 msk_len = n[24], key_offset = n[25];
    const int nocase = n[26], single = n[27];
    const u8 cm = nocase ? 0xdf : 0xff;
    const u8 k0 = n[28] & cm, k1 = n[29] & cm;
    for (size_t pos = 0; pos < s->len && s->groups; pos++) {
        if ((s->buf[pos] & cm) != k0) {
            continue;
        }
        if (!single && (pos + 1 >= s->len || (s->buf[pos + 1] & cm) != k1)) {
            continue;
        }
        const long long first = (long long)pos + key_offset - msk_len;
        const long long end = (long long)pos + key_offset - 1;
        if (first < 0 || end >= (long long)s->len || end < (long long)start) {
            continue;
        }
        u64 v = 0;
        for (u32 i = 0; i < msk_len; i++) {
            v |= (u64)s->buf[first + i] << (8 * i);
        }
        if ((v & msk) != cmp) {
            continue;
        }
        s->groups = hwlm_cb(s, (size_t)end, id);
    }
}

/* hwlmExec (src/hwlm/hwlm.c:172-199); acceleration is skipped (it only
 * advances `start` past bytes that cannot begin a match). */
static void hwlm_exec(struct scan *s, const u8 *hwlm, size_t start) {
    const u8 *eng = hwlm + 192; /* ROUNDUP_CL(sizeof(struct HWLM)) */
    s->last_match = 0xffffffffu; /* INVALID_MATCH_ID */
    if (hwlm[0] == 16) {
        nood_exec(s, eng, start);
    } else if (rd32(eng) == 0) {
        fdr_exec(s, eng, start);
    } else {
        teddy_exec(s, eng, start);
    }
}

API long oracle_hwlm_exec(const void *hwlm, const unsigned char *buf, size_t len, size_t start,
                          unsigned long long groups, struct rec16 *out, size_t cap,
                          size_t stop_after) {
    struct scan s;
    memset(&s, 0, sizeof(s));
    s.buf = buf;
    s.len = len;
    s.groups = groups;
    s.out = out;
    s.cap = cap;
    s.stop_after = stop_after;
    hwlm_exec(&s, (const u8 *)hwlm, start);
    return (long)s.n;
}

/* ---- hs_scan ----------------------------------------------------------------------------- */

struct collect {
    struct rec16 *out;
    size_t cap, n, stop_after;
    u32 block;
};

static int collect_cb(unsigned id, unsigned long long from, unsigned long long to, unsigned flags,
                      void *ctx) {
    (void)from;
    (void)flags;
    struct collect *c = (struct collect *)ctx;
    if (c->n < c->cap) {
        c->out[c->n].id = id;
        c->out[c->n].block = c->block;
        c->out[c->n].to = to;
    }
    c->n++;
    return c->stop_after && c->n >= c->stop_after;
}

/* hs_scan for ROSE_RUNTIME_PURE_LITERAL databases.  Returns an hs_error_t. */
static int scan_block(const void *db, const u8 *data, u32 len, user_cb cb, void *ctx, u8 *work,
                      size_t work_bytes) {
    const struct db_header *h = (const struct db_header *)db;
    if (!h || h->magic != 0xdbdbdbdbU) {
        return -1; /* HS_INVALID */
    }
    if (h->version != ((5u << 24) | (4u << 16) | (2u << 8))) {
        return -5; /* HS_DB_VERSION_ERROR */
    }
    const u8 *rose = (const u8 *)db + h->bytecode;
    if (rd32(rose + RE_mode) != 1) {
        return -7; /* HS_DB_MODE_ERROR */
    }
    if (rose[RE_runtimeImpl] != 1 || !rd32(rose + RE_fmatcherOffset)) {
        return -11; /* this restatement covers ROSE_RUNTIME_PURE_LITERAL only */
    }
    if (rd32(rose + RE_minWidth) > len) {
        return 0;
    }
    struct scan s;
    memset(&s, 0, sizeof(s));
    s.rose = rose;
    s.buf = data;
    s.len = len;
    s.cb = cb;
    s.cb_ctx = ctx;
    s.ekeyCount = rd32(rose + RE_ekeyCount);
    s.dkeyCount = rd32(rose + RE_dkeyCount);
    const size_t eb = (s.ekeyCount + 7) / 8 + 1, dbytes = (s.dkeyCount + 7) / 8 + 1;
    if (eb + 2 * dbytes > work_bytes) {
        return -2;
    }
    memset(work, 0, eb + 2 * dbytes); /* clearEvec, src/runtime.c:358 */
    s.evec = work;
    s.dlog[0] = work + eb;
    s.dlog[1] = work + eb + dbytes;
    s.dedupe_offset = ~0ULL;
    s.groups = rd64(rose + RE_initialGroups) & rd64(rose + RE_floating_group_mask);
    hwlm_exec(&s, rose + rd32(rose + RE_fmatcherOffset), 0);
    if (s.terminated == 2) {
        return -13; /* HS_UNKNOWN_ERROR */
    }
    return s.terminated ? -3 : 0; /* HS_SCAN_TERMINATED */
}

static size_t work_size(const void *db) {
    const struct db_header *h = (const struct db_header *)db;
    const u8 *rose = (const u8 *)db + h->bytecode;
    return (rd32(rose + RE_ekeyCount) + 7) / 8 + 2 * ((rd32(rose + RE_dkeyCount) + 7) / 8) + 16;
}

API long oracle_scan_collect(const void *db, const char *data, const unsigned long long *offsets,
                             const unsigned *lengths, size_t nblocks, struct rec16 *out, size_t cap,
                             size_t stop_after, int *last_err) {
    const size_t wb = work_size(db);
    u8 *work = (u8 *)malloc(wb);
    struct collect c = {out, cap, 0, stop_after, 0};
    int rv = 0;
    for (size_t i = 0; i < nblocks; i++) {
        c.block = (u32)i;
        rv = scan_block(db, (const u8 *)data + offsets[i], lengths[i], collect_cb, &c, work, wb);
        if (rv != 0) {
            break;
        }
    }
    if (last_err) {
        *last_err = rv;
    }
    free(work);
    return (long)c.n;
}

static int count_cb(unsigned id, unsigned long long from, unsigned long long to, unsigned flags,
                    void *ctx) {
    (void)id;
    (void)from;
    (void)to;
    (void)flags;
    (*(unsigned long long *)ctx)++;
    return 0;
}

/* Single-threaded hsbench-style loop (tools/hsbench/main.cpp:503-527) over
 * this scalar port; `nthreads` is accepted for signature parity and ignored. */
API double oracle_scan_blocks_mt(const void *db, const char *data, const unsigned long long *offsets,
                                 const unsigned *lengths, size_t nblocks, unsigned nthreads,
                                 unsigned repeats, unsigned long long *total_matches,
                                 unsigned long long *total_bytes) {
    (void)nthreads;
    const size_t wb = work_size(db);
    u8 *work = (u8 *)malloc(wb);
    unsigned long long m = 0, b = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned r = 0; r < repeats; r++) {
        for (size_t i = 0; i < nblocks; i++) {
            scan_block(db, (const u8 *)data + offsets[i], lengths[i], count_cb, &m, work, wb);
            b += lengths[i];
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(work);
    if (total_matches) {
        *total_matches = m;
    }
    if (total_bytes) {
        *total_bytes = b;
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- streaming ----------------------------------------------------------------------------
 * hs_open_stream / hs_scan_stream / hs_close_stream for pure-literal databases
 * (src/runtime.c:542-977): pureLiteralStreamExec = hwlmExecStreaming over the
 * write with the last historyRequired bytes of the stream as look-behind
 * (src/hwlm/hwlm.c:201-239, src/fdr/fdr.c:853-881); stream state = status byte,
 * history, exhaustion vector (src/runtime.c:478-508 maintainHistoryBuffer).
 * Restated as: scan (history ++ write), report ends inside the write only. */

enum { RE_historyRequired = 16 };

struct ostream {
    const void *db;
    u64 offset;
    u32 hlen, hreq;
    u8 hist[128];
    u8 status;           /* 1 terminated, 2 exhausted */
    u8 *work;            /* evec | dlog0 | dlog1 (evec persists across writes) */
    size_t work_bytes;
};

static void *open_stream_mode(const void *db, u32 mode) {
    const struct db_header *h = (const struct db_header *)db;
    if (!h || h->magic != 0xdbdbdbdbU) {
        return NULL;
    }
    const u8 *rose = (const u8 *)db + h->bytecode;
    if (rd32(rose + RE_mode) != mode || rose[RE_runtimeImpl] != 1) { /* HS_MODE_*, PURE_LITERAL */
        return NULL;
    }
    struct ostream *st = (struct ostream *)calloc(1, sizeof(*st));
    st->db = db;
    st->hreq = rd32(rose + RE_historyRequired);
    if (st->hreq > sizeof(st->hist)) {
        free(st);
        return NULL;
    }
    st->work_bytes = work_size(db);
    st->work = (u8 *)calloc(1, st->work_bytes);
    return st;
}

API void *oracle_open_stream(const void *db) {
    return open_stream_mode(db, 2); /* HS_MODE_STREAM */
}

API int oracle_scan_stream(void *stream, const char *data, unsigned len, user_cb cb, void *ctx) {
    struct ostream *st = (struct ostream *)stream;
    if (!st || !data) {
        return -1;
    }
    if (st->status & 1) {
        return -3; /* HS_SCAN_TERMINATED: the stream is broken (src/runtime.c:883-893) */
    }
    if ((st->status & 2) || len == 0) {
        return 0;
    }
    const struct db_header *h = (const struct db_header *)st->db;
    const u8 *rose = (const u8 *)st->db + h->bytecode;
    u8 *buf = (u8 *)malloc((size_t)st->hlen + len + 8);
    memcpy(buf, st->hist, st->hlen);
    memcpy(buf + st->hlen, data, len);
    struct scan s;
    memset(&s, 0, sizeof(s));
    s.rose = rose;
    s.buf = buf;
    s.len = (size_t)st->hlen + len;
    s.base_offset = st->offset - st->hlen;
    s.cb = cb;
    s.cb_ctx = ctx;
    s.ekeyCount = rd32(rose + RE_ekeyCount);
    s.dkeyCount = rd32(rose + RE_dkeyCount);
    const size_t eb = (s.ekeyCount + 7) / 8 + 1, dbytes = (s.dkeyCount + 7) / 8 + 1;
    s.evec = st->work;                      /* persists for the life of the stream */
    s.dlog[0] = st->work + eb;
    s.dlog[1] = st->work + eb + dbytes;
    memset(s.dlog[0], 0, 2 * dbytes);
    s.dedupe_offset = ~0ULL;
    s.groups = rd64(rose + RE_initialGroups) & rd64(rose + RE_floating_group_mask);
    hwlm_exec(&s, rose + rd32(rose + RE_fmatcherOffset), st->hlen);
    int rv = 0;
    if (s.terminated == 1) {
        st->status |= 1;
        rv = -3;
    } else if (s.terminated == 2) {
        rv = -13;
    } else {
        if (all_exhausted(&s)) {
            st->status |= 2;
        }
        /* maintainHistoryBuffer: keep the last hreq bytes of the stream */
        const size_t keep = s.len < st->hreq ? s.len : st->hreq;
        memmove(st->hist, buf + s.len - keep, keep);
        st->hlen = (u32)keep;
        st->offset += len;
    }
    free(buf);
    return rv;
}

API int oracle_close_stream(void *stream) {
    struct ostream *st = (struct ostream *)stream;
    if (st) {
        free(st->work);
        free(st);
    }
    return 0;
}

API long oracle_stream_collect(const void *db, const char *data, const unsigned *write_lengths,
                               size_t nwrites, struct rec16 *out, size_t cap, size_t stop_after,
                               int *last_err) {
    void *st = oracle_open_stream(db);
    if (!st) {
        return -1;
    }
    struct collect c = {out, cap, 0, stop_after, 0};
    int rv = 0;
    size_t pos = 0;
    for (size_t i = 0; i < nwrites; i++) {
        c.block = (u32)i;
        rv = oracle_scan_stream(st, data + pos, write_lengths[i], collect_cb, &c);
        pos += write_lengths[i];
        if (rv != 0) {
            break;
        }
    }
    oracle_close_stream(st);
    if (last_err) {
        *last_err = rv;
    }
    return (long)c.n;
}

/* hs_scan_vector (src/runtime.c:1106-1175): a temporary stream over a
 * HS_MODE_VECTORED database, one write per buffer, closed at the end. */
API long oracle_vector_collect(const void *db, const char *data, const unsigned *buf_lengths,
                               size_t nbufs, struct rec16 *out, size_t cap, size_t stop_after,
                               int *last_err) {
    void *st = open_stream_mode(db, 4); /* HS_MODE_VECTORED */
    if (!st) {
        return -1;
    }
    struct collect c = {out, cap, 0, stop_after, 0};
    int rv = 0;
    size_t pos = 0;
    for (size_t i = 0; i < nbufs; i++) {
        rv = oracle_scan_stream(st, data + pos, buf_lengths[i], collect_cb, &c);
        pos += buf_lengths[i];
        if (rv != 0) {
            break;
        }
    }
    oracle_close_stream(st);
    if (last_err) {
        *last_err = rv;
    }
    return (long)c.n;
}
