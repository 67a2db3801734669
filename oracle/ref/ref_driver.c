/*
 * ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin helpers linked next to the UNMODIFIED reference runtime objects inside
 * oracle/_ref/libhsref_<isa>.so.  They expose three things to the Python
 * tests / bench harness through ctypes:
 *
 *   ref_hwlm_exec        run the reference hwlmExec() (src/hwlm/hwlm.c:172) on a
 *                        raw HWLM table with a recording callback, the way
 *                        unit/internal/fdr.cpp:140-165 does.
 *   ref_scan_blocks_mt   hsbench-style block benchmark loop
 *                        (tools/hsbench/main.cpp:503-527): N threads, one
 *                        scratch each, `repeats` passes over a set of blocks,
 *                        counting callback; returns wall seconds.
 *   ref_scan_collect     scan blocks and collect (block,id,to) records.
 *   ref_stream_collect   one stream cut into writes, (id, write, to) records.
 *   ref_vector_collect   hs_scan_vector over consecutive buffers, (id, 0, to).
 *   ref_layout_dump      print sizeof/offsetof of every bytecode struct our
 *                        ref_layout.h restates (golden file for layout tests).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hs.h"
#include "scratch.h"
#include "database.h"
#include "hwlm/hwlm.h"
#include "hwlm/hwlm_internal.h"
#include "hwlm/noodle_internal.h"
#include "fdr/fdr_internal.h"
#include "fdr/fdr_confirm.h"
#include "fdr/teddy_internal.h"
#include "rose/rose_internal.h"
#include "rose/rose_program.h"
#include "nfa/nfa_internal.h"
#include "nfa/mcclellan_internal.h"
#include "nfa/nfa_api.h"
#include "nfa/mcclellan.h"
#include "nfa/sheng.h"
#include "nfa/callback.h"
#include "nfa/accel.h"
#include "nfa/shufti.h"
#include "nfa/truffle.h"
#include "nfa/vermicelli.h"

#define API __attribute__((visibility("default")))

struct rec16 {
    unsigned id;
    unsigned block;
    unsigned long long to;
};

/* ---- HWLM-level execution ------------------------------------------- */

struct hwlm_ctx {
    struct hs_scratch scratch; /* must be first: callback receives &scratch */
    struct rec16 *out;
    size_t cap;
    size_t n;
    size_t stop_after; /* terminate after this many matches (0 = never) */
};

static hwlmcb_rv_t record_cb(size_t end, u32 id, struct hs_scratch *scratch) {
    struct hwlm_ctx *c = (struct hwlm_ctx *)scratch;
    if (c->n < c->cap) {
        c->out[c->n].id = id;
        c->out[c->n].block = 0;
        c->out[c->n].to = end;
    }
    c->n++;
    if (c->stop_after && c->n >= c->stop_after) {
        return HWLM_TERMINATE_MATCHING;
    }
    return HWLM_CONTINUE_MATCHING;
}

API long ref_hwlm_exec(const void *hwlm, const unsigned char *buf, size_t len,
                       size_t start, unsigned long long groups,
                       struct rec16 *out, size_t cap, size_t stop_after) {
    struct hwlm_ctx *c = NULL;
    if (posix_memalign((void **)&c, 64, sizeof(*c))) {
        return -1;
    }
    memset(c, 0, sizeof(*c));
    c->out = out;
    c->cap = cap;
    c->stop_after = stop_after;
    hwlmExec((const struct HWLM *)hwlm, buf, len, start, record_cb,
             &c->scratch, groups);
    long n = (long)c->n;
    free(c);
    return n;
}

/* ---- accel primitives (reference: src/nfa/{shufti,truffle,vermicelli}) */

API long ref_shufti(const unsigned char lo[16], const unsigned char hi[16],
                    const unsigned char *buf, size_t len) {
    m128 l, h;
    memcpy(&l, lo, 16);
    memcpy(&h, hi, 16);
    return shuftiExec(l, h, buf, buf + len) - buf;
}

API long ref_truffle(const unsigned char m1[16], const unsigned char m2[16],
                     const unsigned char *buf, size_t len) {
    m128 a, b;
    memcpy(&a, m1, 16);
    memcpy(&b, m2, 16);
    return truffleExec(a, b, buf, buf + len) - buf;
}

API long ref_vermicelli(unsigned char c, int nocase, const unsigned char *buf,
                        size_t len) {
    return vermicelliExec((char)c, (char)nocase, buf, buf + len) - buf;
}

API long ref_dvermicelli(unsigned char c1, unsigned char c2, int nocase,
                         const unsigned char *buf, size_t len) {
    return vermicelliDoubleExec((char)c1, (char)c2, (char)nocase, buf,
                                buf + len) - buf;
}

/* ---- API-level collection ------------------------------------------- */

struct collect_ctx {
    struct rec16 *out;
    size_t cap;
    size_t n;
    unsigned block;
    size_t stop_after;
};

static int collect_cb(unsigned id, unsigned long long from,
                      unsigned long long to, unsigned flags, void *ctx) {
    (void)from;
    (void)flags;
    struct collect_ctx *c = (struct collect_ctx *)ctx;
    if (c->n < c->cap) {
        c->out[c->n].id = id;
        c->out[c->n].block = c->block;
        c->out[c->n].to = to;
    }
    c->n++;
    if (c->stop_after && c->n >= c->stop_after) {
        return 1;
    }
    return 0;
}

/* Scan nblocks blocks (data + offsets[i], lengths[i]) with the reference
 * hs_scan(); append (id, block, to) records in delivery order.  Returns the
 * total number of matches (may exceed cap) or a negative hs error. */
API long ref_scan_collect(const hs_database_t *db, const char *data,
                          const unsigned long long *offsets,
                          const unsigned *lengths, size_t nblocks,
                          struct rec16 *out, size_t cap, size_t stop_after,
                          int *last_err) {
    hs_scratch_t *scratch = NULL;
    hs_error_t err = hs_alloc_scratch(db, &scratch);
    if (err != HS_SUCCESS) {
        return (long)err;
    }
    struct collect_ctx c = {out, cap, 0, 0, stop_after};
    hs_error_t rv = HS_SUCCESS;
    for (size_t i = 0; i < nblocks; i++) {
        c.block = (unsigned)i;
        rv = hs_scan(db, data + offsets[i], lengths[i], 0, scratch, collect_cb,
                     &c);
        if (rv != HS_SUCCESS) {
            break;
        }
    }
    if (last_err) {
        *last_err = (int)rv;
    }
    hs_free_scratch(scratch);
    return (long)c.n;
}

/* Streaming: one stream, the data cut into `nwrites` consecutive writes
 * (hs_open_stream / hs_scan_stream / hs_close_stream, src/runtime.c:542-977).
 * Records (id, write index, to) in delivery order; `to` is the stream offset. */
API long ref_stream_collect(const hs_database_t *db, const char *data,
                            const unsigned *write_lengths, size_t nwrites,
                            struct rec16 *out, size_t cap, size_t stop_after,
                            int *last_err) {
    hs_scratch_t *scratch = NULL;
    hs_error_t err = hs_alloc_scratch(db, &scratch);
    if (err != HS_SUCCESS) {
        return (long)err;
    }
    hs_stream_t *stream = NULL;
    err = hs_open_stream(db, 0, &stream);
    if (err != HS_SUCCESS) {
        hs_free_scratch(scratch);
        return (long)err;
    }
    struct collect_ctx c = {out, cap, 0, 0, stop_after};
    hs_error_t rv = HS_SUCCESS;
    size_t pos = 0;
    for (size_t i = 0; i < nwrites; i++) {
        c.block = (unsigned)i;
        rv = hs_scan_stream(stream, data + pos, write_lengths[i], 0, scratch,
                            collect_cb, &c);
        pos += write_lengths[i];
        if (rv != HS_SUCCESS) {
            break;
        }
    }
    hs_error_t cv = hs_close_stream(stream, scratch, collect_cb, &c);
    if (rv == HS_SUCCESS) {
        rv = cv;
    }
    if (last_err) {
        *last_err = (int)rv;
    }
    hs_free_scratch(scratch);
    return (long)c.n;
}

/* Vectored mode: the data cut into `nbufs` consecutive buffers handed to the
 * reference hs_scan_vector() (src/runtime.c:1106-1175) in one call.  Records
 * (id, 0, to) in delivery order; `to` counts from the first buffer. */
API long ref_vector_collect(const hs_database_t *db, const char *data,
                            const unsigned *buf_lengths, size_t nbufs,
                            struct rec16 *out, size_t cap, size_t stop_after,
                            int *last_err) {
    hs_scratch_t *scratch = NULL;
    hs_error_t err = hs_alloc_scratch(db, &scratch);
    if (err != HS_SUCCESS) {
        return (long)err;
    }
    const char **ptrs = (const char **)malloc((nbufs + 1) * sizeof(*ptrs));
    size_t pos = 0;
    for (size_t i = 0; i < nbufs; i++) {
        ptrs[i] = data + pos;
        pos += buf_lengths[i];
    }
    struct collect_ctx c = {out, cap, 0, 0, stop_after};
    hs_error_t rv = hs_scan_vector(db, ptrs, buf_lengths, (unsigned)nbufs, 0,
                                   scratch, collect_cb, &c);
    if (last_err) {
        *last_err = (int)rv;
    }
    free(ptrs);
    hs_free_scratch(scratch);
    return (long)c.n;
}

/* ---- hsbench-style multi-threaded timing ------------------------------ */

/* CPUs the bench threads are pinned to, thread i -> cpu[i % n]; n == 0: unpinned.
 * hsbench pins its scan threads 1:1 (tools/hsbench/main.cpp:211-222 setAffinity). */
static int g_bench_cpus[1024];
static unsigned g_bench_ncpus;

API void ref_set_bench_cpus(const int *cpus, unsigned n) {
    g_bench_ncpus = n > 1024 ? 1024 : n;
    for (unsigned i = 0; i < g_bench_ncpus; i++) {
        g_bench_cpus[i] = cpus[i];
    }
}

struct bench_thread {
    unsigned index;
    pthread_t tid;
    const hs_database_t *db;
    const char *data;
    const unsigned long long *offsets;
    const unsigned *lengths;
    size_t first, last; /* block range [first,last) */
    unsigned repeats;
    unsigned long long matches;
    unsigned long long bytes;
    pthread_barrier_t *bar;
    int err;
};

static int count_cb(unsigned id, unsigned long long from,
                    unsigned long long to, unsigned flags, void *ctx) {
    (void)id;
    (void)from;
    (void)to;
    (void)flags;
    (*(unsigned long long *)ctx)++;
    return 0;
}

static void *bench_main(void *p) {
    struct bench_thread *t = (struct bench_thread *)p;
    if (g_bench_ncpus) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(g_bench_cpus[t->index % g_bench_ncpus], &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set); /* best effort */
    }
    hs_scratch_t *scratch = NULL;
    t->err = hs_alloc_scratch(t->db, &scratch);
    pthread_barrier_wait(t->bar);
    if (t->err == HS_SUCCESS) {
        for (unsigned r = 0; r < t->repeats; r++) {
            for (size_t i = t->first; i < t->last; i++) {
                hs_scan(t->db, t->data + t->offsets[i], t->lengths[i], 0,
                        scratch, count_cb, &t->matches);
                t->bytes += t->lengths[i];
            }
        }
    }
    pthread_barrier_wait(t->bar);
    if (scratch) {
        hs_free_scratch(scratch);
    }
    return NULL;
}

API double ref_scan_blocks_mt(const hs_database_t *db, const char *data,
                              const unsigned long long *offsets,
                              const unsigned *lengths, size_t nblocks,
                              unsigned nthreads, unsigned repeats,
                              unsigned long long *total_matches,
                              unsigned long long *total_bytes) {
    if (nthreads == 0) {
        nthreads = 1;
    }
    struct bench_thread *th = calloc(nthreads, sizeof(*th));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, nthreads + 1);
    for (unsigned i = 0; i < nthreads; i++) {
        th[i].index = i;
        th[i].db = db;
        th[i].data = data;
        th[i].offsets = offsets;
        th[i].lengths = lengths;
        th[i].first = nblocks * i / nthreads;
        th[i].last = nblocks * (i + 1) / nthreads;
        th[i].repeats = repeats;
        th[i].bar = &bar;
        pthread_create(&th[i].tid, NULL, bench_main, &th[i]);
    }
    struct timespec t0, t1;
    pthread_barrier_wait(&bar);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&bar);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    unsigned long long m = 0, b = 0;
    int err = 0;
    for (unsigned i = 0; i < nthreads; i++) {
        pthread_join(th[i].tid, NULL);
        m += th[i].matches;
        b += th[i].bytes;
        if (th[i].err) {
            err = th[i].err;
        }
    }
    pthread_barrier_destroy(&bar);
    free(th);
    if (total_matches) {
        *total_matches = m;
    }
    if (total_bytes) {
        *total_bytes = b;
    }
    if (err) {
        return -1.0;
    }
    return (double)(t1.tv_sec - t0.tv_sec) +
           1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- DFA engines in block mode ------------------------------------------- */

struct nfa_collect {
    struct rec16 *out;
    size_t cap, n;
    unsigned block;
};

static int nfa_cb(u64a start, u64a end, ReportID id, void *ctx) {
    (void)start;
    struct nfa_collect *c = (struct nfa_collect *)ctx;
    if (c->n < c->cap) {
        c->out[c->n].id = id;
        c->out[c->n].block = c->block;
        c->out[c->n].to = end;
    }
    c->n++;
    return MO_CONTINUE_MATCHING;
}

/* nfaExecMcClellan8_B / nfaExecMcClellan16_B / nfaExecSheng_B (src/nfa/mcclellan.c:
 * 937,963; src/nfa/sheng.c:706) over every block, offset 0 each, callbacks collected
 * as (report, block, end).  `nfa` must be 64-byte aligned.  Returns the number of
 * callbacks (may exceed cap) or -1 for an engine type not handled here. */
int ref_limex32_block(const struct NFA *n, const u8 *buf, size_t len, NfaCallback cb, void *ctx); /* ref_limex.c */

API long ref_nfa_exec_blocks(const void *nfa, const char *data, const unsigned long long *offsets,
                             const unsigned *lengths, size_t nblocks, struct rec16 *out,
                             size_t cap) {
    const struct NFA *n = (const struct NFA *)nfa;
    struct nfa_collect c = {out, cap, 0, 0};
    for (size_t b = 0; b < nblocks; b++) {
        const u8 *buf = (const u8 *)data + offsets[b];
        c.block = (unsigned)b;
        switch (n->type) {
        case MCCLELLAN_NFA_8:
            nfaExecMcClellan8_B(n, 0, buf, lengths[b], nfa_cb, &c);
            break;
        case MCCLELLAN_NFA_16:
            nfaExecMcClellan16_B(n, 0, buf, lengths[b], nfa_cb, &c);
            break;
        case SHENG_NFA:
            nfaExecSheng_B(n, 0, buf, lengths[b], nfa_cb, &c);
            break;
        default:
            if (!ref_limex32_block(n, buf, lengths[b], nfa_cb, &c)) {
                return -1;
            }
        }
    }
    return (long)c.n;
}

/* ---- layout dump -------------------------------------------------------- */

#define SZ(s) printf("  \"sizeof(%s)\": %zu,\n", #s, sizeof(struct s))
#define OFF(s, f)                                                              \
    printf("  \"%s.%s\": %zu,\n", #s, #f, offsetof(struct s, f))

void ref_layout_dump_sheng(void); /* ref_sheng_layout.c: sheng_internal.h redefines report_list */
void ref_layout_dump_limex(void); /* ref_limex.c */

API void ref_layout_dump(void) {
    printf("{\n");
    SZ(hs_database);
    OFF(hs_database, magic); OFF(hs_database, version);
    OFF(hs_database, length); OFF(hs_database, platform);
    OFF(hs_database, crc32); OFF(hs_database, reserved0);
    OFF(hs_database, reserved1); OFF(hs_database, bytecode);
    OFF(hs_database, padding); OFF(hs_database, bytes);

    SZ(RoseEngine);
    OFF(RoseEngine, pureLiteral); OFF(RoseEngine, noFloatingRoots);
    OFF(RoseEngine, requiresEodCheck); OFF(RoseEngine, hasOutfixesInSmallBlock);
    OFF(RoseEngine, runtimeImpl); OFF(RoseEngine, mpvTriggeredByLeaf);
    OFF(RoseEngine, canExhaust); OFF(RoseEngine, hasSom);
    OFF(RoseEngine, somHorizon); OFF(RoseEngine, mode);
    OFF(RoseEngine, historyRequired); OFF(RoseEngine, ekeyCount);
    OFF(RoseEngine, lkeyCount); OFF(RoseEngine, lopCount);
    OFF(RoseEngine, ckeyCount); OFF(RoseEngine, logicalTreeOffset);
    OFF(RoseEngine, combInfoMapOffset); OFF(RoseEngine, dkeyCount);
    OFF(RoseEngine, dkeyLogSize); OFF(RoseEngine, invDkeyOffset);
    OFF(RoseEngine, somLocationCount); OFF(RoseEngine, somLocationFatbitSize);
    OFF(RoseEngine, rolesWithStateCount); OFF(RoseEngine, stateSize);
    OFF(RoseEngine, anchorStateSize); OFF(RoseEngine, tStateSize);
    OFF(RoseEngine, scratchStateSize); OFF(RoseEngine, smallWriteOffset);
    OFF(RoseEngine, amatcherOffset); OFF(RoseEngine, ematcherOffset);
    OFF(RoseEngine, fmatcherOffset); OFF(RoseEngine, drmatcherOffset);
    OFF(RoseEngine, sbmatcherOffset); OFF(RoseEngine, longLitTableOffset);
    OFF(RoseEngine, amatcherMinWidth); OFF(RoseEngine, fmatcherMinWidth);
    OFF(RoseEngine, eodmatcherMinWidth);
    OFF(RoseEngine, amatcherMaxBiAnchoredWidth);
    OFF(RoseEngine, fmatcherMaxBiAnchoredWidth);
    OFF(RoseEngine, reportProgramOffset); OFF(RoseEngine, reportProgramCount);
    OFF(RoseEngine, delayProgramOffset); OFF(RoseEngine, anchoredProgramOffset);
    OFF(RoseEngine, activeArrayCount); OFF(RoseEngine, activeLeftCount);
    OFF(RoseEngine, queueCount); OFF(RoseEngine, activeQueueArraySize);
    OFF(RoseEngine, eagerIterOffset); OFF(RoseEngine, handledKeyCount);
    OFF(RoseEngine, handledKeyFatbitSize); OFF(RoseEngine, leftOffset);
    OFF(RoseEngine, roseCount); OFF(RoseEngine, eodProgramOffset);
    OFF(RoseEngine, flushCombProgramOffset);
    OFF(RoseEngine, lastFlushCombProgramOffset);
    OFF(RoseEngine, lastByteHistoryIterOffset); OFF(RoseEngine, minWidth);
    OFF(RoseEngine, minWidthExcludingBoundaries);
    OFF(RoseEngine, maxBiAnchoredWidth); OFF(RoseEngine, anchoredDistance);
    OFF(RoseEngine, anchoredMinDistance); OFF(RoseEngine, floatingDistance);
    OFF(RoseEngine, floatingMinDistance); OFF(RoseEngine, smallBlockDistance);
    OFF(RoseEngine, floatingMinLiteralMatchOffset);
    OFF(RoseEngine, nfaInfoOffset); OFF(RoseEngine, initialGroups);
    OFF(RoseEngine, floating_group_mask); OFF(RoseEngine, size);
    OFF(RoseEngine, delay_count); OFF(RoseEngine, delay_fatbit_size);
    OFF(RoseEngine, anchored_count); OFF(RoseEngine, anchored_fatbit_size);
    OFF(RoseEngine, maxFloatingDelayedMatch); OFF(RoseEngine, delayRebuildLength);
    OFF(RoseEngine, stateOffsets); OFF(RoseEngine, boundary);
    OFF(RoseEngine, totalNumLiterals); OFF(RoseEngine, asize);
    OFF(RoseEngine, outfixBeginQueue); OFF(RoseEngine, outfixEndQueue);
    OFF(RoseEngine, leftfixBeginQueue); OFF(RoseEngine, initMpvNfa);
    OFF(RoseEngine, rosePrefixCount); OFF(RoseEngine, activeLeftIterOffset);
    OFF(RoseEngine, ematcherRegionSize); OFF(RoseEngine, somRevCount);
    OFF(RoseEngine, somRevOffsetOffset); OFF(RoseEngine, longLitStreamState);
    OFF(RoseEngine, state_init);

    SZ(RoseStateOffsets);
    OFF(RoseStateOffsets, history); OFF(RoseStateOffsets, exhausted);
    OFF(RoseStateOffsets, exhausted_size); OFF(RoseStateOffsets, logicalVec);
    OFF(RoseStateOffsets, logicalVec_size); OFF(RoseStateOffsets, combVec);
    OFF(RoseStateOffsets, combVec_size); OFF(RoseStateOffsets, activeLeafArray);
    OFF(RoseStateOffsets, activeLeafArray_size);
    OFF(RoseStateOffsets, activeLeftArray);
    OFF(RoseStateOffsets, activeLeftArray_size);
    OFF(RoseStateOffsets, leftfixLagTable); OFF(RoseStateOffsets, anchorState);
    OFF(RoseStateOffsets, groups); OFF(RoseStateOffsets, groups_size);
    OFF(RoseStateOffsets, longLitState); OFF(RoseStateOffsets, longLitState_size);
    OFF(RoseStateOffsets, somLocation); OFF(RoseStateOffsets, somValid);
    OFF(RoseStateOffsets, somWritable); OFF(RoseStateOffsets, somMultibit_size);
    OFF(RoseStateOffsets, nfaStateBegin); OFF(RoseStateOffsets, end);

    SZ(RoseBoundaryReports);
    SZ(NfaInfo);
    OFF(NfaInfo, nfaOffset); OFF(NfaInfo, stateOffset);
    OFF(NfaInfo, fullStateOffset); OFF(NfaInfo, ekeyListOffset);
    OFF(NfaInfo, no_retrigger); OFF(NfaInfo, in_sbmatcher); OFF(NfaInfo, eod);

    SZ(HWLM);
    OFF(HWLM, type); OFF(HWLM, accel1_groups); OFF(HWLM, accel1);
    OFF(HWLM, accel0);
    printf("  \"sizeof(AccelAux)\": %zu,\n", sizeof(union AccelAux));

    SZ(noodTable);
    OFF(noodTable, id); OFF(noodTable, msk); OFF(noodTable, cmp);
    OFF(noodTable, msk_len); OFF(noodTable, key_offset);
    OFF(noodTable, nocase); OFF(noodTable, single); OFF(noodTable, key0);
    OFF(noodTable, key1);

    SZ(FDR);
    OFF(FDR, engineID); OFF(FDR, size); OFF(FDR, maxStringLen);
    OFF(FDR, numStrings); OFF(FDR, confOffset); OFF(FDR, floodOffset);
    OFF(FDR, stride); OFF(FDR, domain); OFF(FDR, domainMask);
    OFF(FDR, tabSize); OFF(FDR, start);

    SZ(Teddy);
    SZ(FDRFlood);
    OFF(FDRFlood, allGroups); OFF(FDRFlood, suffix); OFF(FDRFlood, idCount);
    OFF(FDRFlood, ids); OFF(FDRFlood, groups);

    SZ(FDRConfirm);
    OFF(FDRConfirm, andmsk); OFF(FDRConfirm, mult); OFF(FDRConfirm, nBits);
    OFF(FDRConfirm, groups);

    SZ(LitInfo);
    OFF(LitInfo, v); OFF(LitInfo, msk); OFF(LitInfo, groups); OFF(LitInfo, id);
    OFF(LitInfo, size); OFF(LitInfo, flags); OFF(LitInfo, next);

    SZ(NFA);
    OFF(NFA, flags); OFF(NFA, length); OFF(NFA, type); OFF(NFA, rAccelType);
    OFF(NFA, rAccelOffset); OFF(NFA, maxBiAnchoredWidth);
    OFF(NFA, rAccelData); OFF(NFA, queueIndex); OFF(NFA, nPositions);
    OFF(NFA, scratchStateSize); OFF(NFA, streamStateSize);
    OFF(NFA, maxWidth); OFF(NFA, minWidth); OFF(NFA, maxOffset);

    SZ(mcclellan);
    OFF(mcclellan, state_count); OFF(mcclellan, length);
    OFF(mcclellan, start_anchored); OFF(mcclellan, start_floating);
    OFF(mcclellan, aux_offset); OFF(mcclellan, sherman_offset);
    OFF(mcclellan, sherman_end); OFF(mcclellan, accel_limit_8);
    OFF(mcclellan, accept_limit_8); OFF(mcclellan, sherman_limit);
    OFF(mcclellan, wide_limit); OFF(mcclellan, alphaShift);
    OFF(mcclellan, flags); OFF(mcclellan, has_accel);
    OFF(mcclellan, has_wide); OFF(mcclellan, remap);
    OFF(mcclellan, arb_report); OFF(mcclellan, accel_offset);
    OFF(mcclellan, haig_offset); OFF(mcclellan, wide_offset);

    SZ(mstate_aux);
    OFF(mstate_aux, accept); OFF(mstate_aux, accept_eod);
    OFF(mstate_aux, top); OFF(mstate_aux, accel_offset);

ref_layout_dump_sheng();
    ref_layout_dump_limex();

    /* rose program instruction sizes (8-byte rounded stride is what matters) */
#define ISZ(n) printf("  \"sizeof(ROSE_STRUCT_%s)\": %zu,\n", #n, sizeof(struct ROSE_STRUCT_##n))
#define IOFF(n, f) printf("  \"ROSE_STRUCT_%s.%s\": %zu,\n", #n, #f, offsetof(struct ROSE_STRUCT_##n, f))
    ISZ(END); ISZ(CHECK_GROUPS); IOFF(CHECK_GROUPS, groups);
    ISZ(CHECK_MASK); IOFF(CHECK_MASK, and_mask); IOFF(CHECK_MASK, cmp_mask);
    IOFF(CHECK_MASK, neg_mask); IOFF(CHECK_MASK, offset);
    IOFF(CHECK_MASK, fail_jump);
    ISZ(CHECK_MASK_32); IOFF(CHECK_MASK_32, and_mask); IOFF(CHECK_MASK_32, cmp_mask);
    IOFF(CHECK_MASK_32, neg_mask); IOFF(CHECK_MASK_32, offset); IOFF(CHECK_MASK_32, fail_jump);
    ISZ(CHECK_MASK_64); IOFF(CHECK_MASK_64, and_mask); IOFF(CHECK_MASK_64, cmp_mask);
    IOFF(CHECK_MASK_64, neg_mask); IOFF(CHECK_MASK_64, offset); IOFF(CHECK_MASK_64, fail_jump);
    ISZ(CHECK_BOUNDS); IOFF(CHECK_BOUNDS, min_bound); IOFF(CHECK_BOUNDS, max_bound); IOFF(CHECK_BOUNDS, fail_jump);
    ISZ(CHECK_BYTE); IOFF(CHECK_BYTE, and_mask); IOFF(CHECK_BYTE, cmp_mask);
    IOFF(CHECK_BYTE, negation); IOFF(CHECK_BYTE, offset);
    IOFF(CHECK_BYTE, fail_jump);
    ISZ(DEDUPE); IOFF(DEDUPE, quash_som); IOFF(DEDUPE, dkey);
    IOFF(DEDUPE, offset_adjust); IOFF(DEDUPE, fail_jump);
    ISZ(REPORT); IOFF(REPORT, onmatch); IOFF(REPORT, offset_adjust);
    ISZ(REPORT_EXHAUST); IOFF(REPORT_EXHAUST, onmatch);
    IOFF(REPORT_EXHAUST, offset_adjust); IOFF(REPORT_EXHAUST, ekey);
    ISZ(DEDUPE_AND_REPORT); IOFF(DEDUPE_AND_REPORT, quash_som);
    IOFF(DEDUPE_AND_REPORT, dkey); IOFF(DEDUPE_AND_REPORT, onmatch);
    IOFF(DEDUPE_AND_REPORT, offset_adjust); IOFF(DEDUPE_AND_REPORT, fail_jump);
    ISZ(FINAL_REPORT); IOFF(FINAL_REPORT, onmatch);
    IOFF(FINAL_REPORT, offset_adjust);
    ISZ(CHECK_EXHAUSTED); IOFF(CHECK_EXHAUSTED, ekey);
    IOFF(CHECK_EXHAUSTED, fail_jump);
    ISZ(SQUASH_GROUPS); IOFF(SQUASH_GROUPS, groups);
    ISZ(CHECK_LONG_LIT); IOFF(CHECK_LONG_LIT, lit_offset);
    IOFF(CHECK_LONG_LIT, lit_length); IOFF(CHECK_LONG_LIT, fail_jump);
    ISZ(CHECK_MED_LIT); IOFF(CHECK_MED_LIT, lit_offset);
    IOFF(CHECK_MED_LIT, lit_length); IOFF(CHECK_MED_LIT, fail_jump);
    ISZ(INCLUDED_JUMP); IOFF(INCLUDED_JUMP, squash);
    IOFF(INCLUDED_JUMP, child_offset);
    ISZ(SET_EXHAUST); IOFF(SET_EXHAUST, ekey);
    printf("  \"ROSE_INSTR_CHECK_GROUPS\": %d,\n", ROSE_INSTR_CHECK_GROUPS);
    printf("  \"ROSE_INSTR_CHECK_MASK\": %d,\n", ROSE_INSTR_CHECK_MASK);
    printf("  \"ROSE_INSTR_CHECK_BOUNDS\": %d,\n", ROSE_INSTR_CHECK_BOUNDS);
    printf("  \"ROSE_INSTR_CHECK_BYTE\": %d,\n", ROSE_INSTR_CHECK_BYTE);
    printf("  \"ROSE_INSTR_CHECK_MASK_32\": %d,\n", ROSE_INSTR_CHECK_MASK_32);
    printf("  \"ROSE_INSTR_CHECK_MASK_64\": %d,\n", ROSE_INSTR_CHECK_MASK_64);
    printf("  \"ROSE_INSTR_DEDUPE\": %d,\n", ROSE_INSTR_DEDUPE);
    printf("  \"ROSE_INSTR_REPORT\": %d,\n", ROSE_INSTR_REPORT);
    printf("  \"ROSE_INSTR_REPORT_EXHAUST\": %d,\n", ROSE_INSTR_REPORT_EXHAUST);
    printf("  \"ROSE_INSTR_DEDUPE_AND_REPORT\": %d,\n",
           ROSE_INSTR_DEDUPE_AND_REPORT);
    printf("  \"ROSE_INSTR_FINAL_REPORT\": %d,\n", ROSE_INSTR_FINAL_REPORT);
    printf("  \"ROSE_INSTR_CHECK_EXHAUSTED\": %d,\n",
           ROSE_INSTR_CHECK_EXHAUSTED);
    printf("  \"ROSE_INSTR_SQUASH_GROUPS\": %d,\n", ROSE_INSTR_SQUASH_GROUPS);
    printf("  \"ROSE_INSTR_CHECK_LONG_LIT\": %d,\n", ROSE_INSTR_CHECK_LONG_LIT);
    printf("  \"ROSE_INSTR_CHECK_LONG_LIT_NOCASE\": %d,\n",
           ROSE_INSTR_CHECK_LONG_LIT_NOCASE);
    printf("  \"ROSE_INSTR_CHECK_MED_LIT\": %d,\n", ROSE_INSTR_CHECK_MED_LIT);
    printf("  \"ROSE_INSTR_CHECK_MED_LIT_NOCASE\": %d,\n",
           ROSE_INSTR_CHECK_MED_LIT_NOCASE);
    printf("  \"ROSE_INSTR_INCLUDED_JUMP\": %d,\n", ROSE_INSTR_INCLUDED_JUMP);
    printf("  \"ROSE_INSTR_SET_EXHAUST\": %d,\n", ROSE_INSTR_SET_EXHAUST);
    printf("  \"LAST_ROSE_INSTRUCTION\": %d,\n", LAST_ROSE_INSTRUCTION);
    printf("  \"MCCLELLAN_NFA_8\": %d,\n", MCCLELLAN_NFA_8);
    printf("  \"MCCLELLAN_NFA_16\": %d,\n", MCCLELLAN_NFA_16);
    printf("  \"SHENG_NFA\": %d,\n", SHENG_NFA);
    printf("  \"HS_DB_VERSION\": %u,\n", (unsigned)HS_DB_VERSION);
    printf("  \"hs_current_platform\": %llu,\n",
           (unsigned long long)hs_current_platform);
    printf("  \"SCRATCH_MAGIC\": %u\n", (unsigned)SCRATCH_MAGIC);
    printf("}\n");
}
