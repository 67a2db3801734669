/* ref_limex.c -- the reference's LimEx engines (32 ... 512 states) over blocks, the way Rose runs an outfix
 * in block mode (src/rose/block.c:218-262, src/rose/match.h:initQueue / pushQueue...):
 * a queue {MQE_START@0, MQE_TOP@0, MQE_END@len} through nfaExecLimEx<N>_Q, then
 * nfaExecLimEx<N>_testEOD; plus sizeof/offsetof of its structures for ref_layout_dump().
 * TEST INFRASTRUCTURE ONLY (part of oracle/_ref). */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ue2common.h"
#include "nfa/nfa_internal.h"
#include "nfa/nfa_api.h"
#include "nfa/nfa_api_queue.h"
#include "nfa/limex.h"
#include "nfa/limex_internal.h"

struct limex_collect {
    int (*cb)(u64a start, u64a end, ReportID id, void *ctx);
    void *ctx;
};

/* one block: returns 0 if the engine type is not one of the LimEx models (32 ... 512 states) */
#define LIMEX_CASES(X) X(32) X(64) X(128) X(256) X(384) X(512)
int ref_limex32_block(const struct NFA *n, const u8 *buf, size_t len, NfaCallback cb, void *ctx) {
    if (n->type > LIMEX_NFA_512) {
        return 0;
    }
    struct mq *q = (struct mq *)calloc(1, sizeof(struct mq));
    char *state = (char *)aligned_alloc(64, (n->scratchStateSize + 127) / 64 * 64);  /* m512 state: aligned loads */
    char *sstate = (char *)calloc(1, n->streamStateSize + 64);
    memset(state, 0, (n->scratchStateSize + 127) / 64 * 64);
    q->nfa = n;
    q->cur = q->end = 0;
    q->state = state;
    q->streamState = sstate;
    q->offset = 0;
    q->buffer = buf;
    q->length = len;
    q->history = NULL;
    q->hlength = 0;
    q->scratch = NULL;
    q->report_current = 0;
    q->cb = cb;
    q->context = ctx;
    switch (n->type) {
#define INIT(sz) case LIMEX_NFA_##sz: nfaExecLimEx##sz##_queueInitState(n, q); break;
        LIMEX_CASES(INIT)
    }
    pushQueue(q, MQE_START, 0);
    pushQueue(q, MQE_TOP, 0);
    pushQueue(q, MQE_END, (s64a)len);
    switch (n->type) {
#define RUN(sz) case LIMEX_NFA_##sz:                                                   \
        nfaExecLimEx##sz##_Q(n, q, (s64a)len);                                         \
        nfaExecLimEx##sz##_testEOD(n, q->state, q->streamState, len, cb, ctx);         \
        break;
        LIMEX_CASES(RUN)
    }
    free(sstate);
    free(state);
    free(q);
    return 1;
}

#define SZ(s) printf("  \"sizeof(%s)\": %zu,\n", #s, sizeof(struct s))
#define OFF(s, f) printf("  \"%s.%s\": %zu,\n", #s, #f, offsetof(struct s, f))

void ref_layout_dump_limex(void) {
    SZ(LimExNFA32);
    OFF(LimExNFA32, reachMap); OFF(LimExNFA32, reachSize); OFF(LimExNFA32, accelCount);
    OFF(LimExNFA32, accelTableOffset); OFF(LimExNFA32, accelAuxCount); OFF(LimExNFA32, accelAuxOffset);
    OFF(LimExNFA32, acceptCount); OFF(LimExNFA32, acceptOffset); OFF(LimExNFA32, acceptEodCount);
    OFF(LimExNFA32, acceptEodOffset); OFF(LimExNFA32, exceptionCount); OFF(LimExNFA32, exceptionOffset);
    OFF(LimExNFA32, repeatCount); OFF(LimExNFA32, repeatOffset); OFF(LimExNFA32, squashOffset);
    OFF(LimExNFA32, squashCount); OFF(LimExNFA32, topCount); OFF(LimExNFA32, topOffset);
    OFF(LimExNFA32, stateSize); OFF(LimExNFA32, flags); OFF(LimExNFA32, init); OFF(LimExNFA32, initDS);
    OFF(LimExNFA32, accept); OFF(LimExNFA32, acceptAtEOD); OFF(LimExNFA32, accel);
    OFF(LimExNFA32, accelPermute); OFF(LimExNFA32, accelCompare); OFF(LimExNFA32, accel_and_friends);
    OFF(LimExNFA32, compressMask); OFF(LimExNFA32, exceptionMask); OFF(LimExNFA32, repeatCyclicMask);
    OFF(LimExNFA32, zombieMask); OFF(LimExNFA32, shift); OFF(LimExNFA32, shiftCount);
    OFF(LimExNFA32, shiftAmount); OFF(LimExNFA32, exceptionShufMask); OFF(LimExNFA32, exceptionBitMask);
    OFF(LimExNFA32, exceptionAndMask);
    SZ(NFAException32);
    OFF(NFAException32, squash); OFF(NFAException32, successors); OFF(NFAException32, reports);
    OFF(NFAException32, repeatOffset); OFF(NFAException32, hasSquash); OFF(NFAException32, trigger);
    SZ(LimExNFA64);
    OFF(LimExNFA64, reachMap); OFF(LimExNFA64, reachSize); OFF(LimExNFA64, accelCount);
    OFF(LimExNFA64, accelTableOffset); OFF(LimExNFA64, accelAuxCount); OFF(LimExNFA64, accelAuxOffset);
    OFF(LimExNFA64, acceptCount); OFF(LimExNFA64, acceptOffset); OFF(LimExNFA64, acceptEodCount);
    OFF(LimExNFA64, acceptEodOffset); OFF(LimExNFA64, exceptionCount); OFF(LimExNFA64, exceptionOffset);
    OFF(LimExNFA64, repeatCount); OFF(LimExNFA64, repeatOffset); OFF(LimExNFA64, squashOffset);
    OFF(LimExNFA64, squashCount); OFF(LimExNFA64, topCount); OFF(LimExNFA64, topOffset);
    OFF(LimExNFA64, stateSize); OFF(LimExNFA64, flags); OFF(LimExNFA64, init); OFF(LimExNFA64, initDS);
    OFF(LimExNFA64, accept); OFF(LimExNFA64, acceptAtEOD); OFF(LimExNFA64, accel);
    OFF(LimExNFA64, accelPermute); OFF(LimExNFA64, accelCompare); OFF(LimExNFA64, accel_and_friends);
    OFF(LimExNFA64, compressMask); OFF(LimExNFA64, exceptionMask); OFF(LimExNFA64, repeatCyclicMask);
    OFF(LimExNFA64, zombieMask); OFF(LimExNFA64, shift); OFF(LimExNFA64, shiftCount);
    OFF(LimExNFA64, shiftAmount); OFF(LimExNFA64, exceptionShufMask); OFF(LimExNFA64, exceptionBitMask);
    OFF(LimExNFA64, exceptionAndMask);
    SZ(NFAException64);
    OFF(NFAException64, squash); OFF(NFAException64, successors); OFF(NFAException64, reports);
    OFF(NFAException64, repeatOffset); OFF(NFAException64, hasSquash); OFF(NFAException64, trigger);
#define DUMP_WIDE(L, E)                                                                                  \
    SZ(L); OFF(L, reachMap); OFF(L, reachSize); OFF(L, acceptCount); OFF(L, acceptOffset); OFF(L, acceptEodCount); \
    OFF(L, acceptEodOffset); OFF(L, exceptionCount); OFF(L, exceptionOffset); OFF(L, repeatCount); OFF(L, topOffset); OFF(L, stateSize); \
    OFF(L, flags); OFF(L, init); OFF(L, initDS); OFF(L, accept); OFF(L, acceptAtEOD); OFF(L, accel);         \
    OFF(L, compressMask); OFF(L, exceptionMask); OFF(L, repeatCyclicMask); OFF(L, zombieMask); OFF(L, shift); \
    OFF(L, shiftCount); OFF(L, shiftAmount); OFF(L, exceptionShufMask); OFF(L, exceptionBitMask);            \
    OFF(L, exceptionAndMask); SZ(E); OFF(E, squash); OFF(E, successors); OFF(E, reports); OFF(E, repeatOffset); \
    OFF(E, hasSquash); OFF(E, trigger);
    DUMP_WIDE(LimExNFA128, NFAException128)
    DUMP_WIDE(LimExNFA256, NFAException256)
    DUMP_WIDE(LimExNFA512, NFAException512)
    printf("  \"LIMEX_NFA_128\": %d,\n", (int)LIMEX_NFA_128);
    printf("  \"LIMEX_NFA_256\": %d,\n", (int)LIMEX_NFA_256);
    printf("  \"LIMEX_NFA_512\": %d,\n", (int)LIMEX_NFA_512);
    SZ(NFAAccept);
    OFF(NFAAccept, single_report); OFF(NFAAccept, reports); OFF(NFAAccept, squash);
    printf("  \"LIMEX_NFA_32\": %d,\n", (int)LIMEX_NFA_32);
    printf("  \"LIMEX_NFA_64\": %d,\n", (int)LIMEX_NFA_64);
    printf("  \"LIMEX_FLAG_CANNOT_DIE\": %d,\n", (int)LIMEX_FLAG_CANNOT_DIE);
    printf("  \"LIMEX_SQUASH_CYCLIC\": %d,\n", (int)LIMEX_SQUASH_CYCLIC);
    printf("  \"LIMEX_SQUASH_REPORT\": %d,\n", (int)LIMEX_SQUASH_REPORT);
}
