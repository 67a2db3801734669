/*
 * limex_build.h -- host-side emitter of the reference's LimEx NFA, 32- and 64-state
 * models, in the reference's own byte layout: `struct NFA` followed by `struct
 * LimExNFA32` / `LimExNFA64`, its reach table, accept / EOD-accept tables, exception table,
 * report lists (src/nfa/limex_internal.h:102-203; laid out the way
 * limex_compile.cpp's Factory::generateNfa does, src/nfa/limex_compile.cpp:2300-2480).
 *
 * The reference's compile side (parser -> Glushkov graph -> state numbering) stays
 * out of scope; what a test or a caller can give is a finished NFA of at most 64
 * states (`RawNfa`: per-state successor sets, per-byte reach, reports) or a
 * literal set, for which the position automaton is built here.  Transitions
 * i -> i + a for the (up to eight) most common forward distances a <= 16 become
 * the "limited" shift masks, everything else exception successors; bounded
 * repeats, tops beyond the single start and acceleration are not emitted.  The same
 * bytes run on the unmodified reference engines (nfaExecLimEx32_Q / 64_Q + _testEOD) --
 * the parity oracle -- and on the device kernel (device/dfa_kernels.cu).
 */
#ifndef HSB200_LIMEX_BUILD_H
#define HSB200_LIMEX_BUILD_H

#include <string>
#include <vector>

#include "../ref_layout.h"
#include "dfa_build.h"

namespace hsb {

struct RawNfa {
    u32 nstates = 0;                   /* <= 64: up to 32 states are emitted as LimEx-32, more as LimEx-64 */
    u64 reach[256] = {0};              /* states that may be ON after consuming the byte */
    u64 init = 0, initDS = 0;          /* switched on by a top at offset 0 / at a later offset */
    u32 mlStartState = 0;              /* regex_nfa.cpp: the shared "after a newline" state, 0 = none yet */
    u32 ctxWord = 0, ctxNonWord = 0;   /* regex_nfa.cpp: "the previous byte is / is not a word character" */
    std::vector<u64> succ;             /* [state] successor set */
    std::vector<u64> squashMask;       /* [state] kept states when the exception's squash applies */
    std::vector<u8> squashKind;        /* [state] LIMEX_SQUASH_NONE / _CYCLIC / _REPORT */
    std::vector<std::vector<u32>> reports;    /* raised while the state is on */
    std::vector<std::vector<u32>> reportsEod; /* raised if the data ends with the state on */
};
typedef RawNfa RawNfa32; /* the name the 32-state-only version had */

/* position automaton of a literal set: state 0 = floating start (always on), one state
 * per literal byte; throws if more than 63 positions are needed */
RawNfa nfaFromLiterals(const std::vector<DfaLiteral> &lits);

/* struct NFA + LimExNFA32 (nstates <= 32) or LimExNFA64 + tables */
std::vector<u8> emitLimEx(const RawNfa &n);
inline std::vector<u8> emitLimEx32(const RawNfa &n) { return emitLimEx(n); }

} // namespace hsb
#endif
