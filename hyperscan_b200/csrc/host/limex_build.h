/*
 * limex_build.h -- host-side emitter of the reference's LimEx NFA, 32- to 512-state
 * models, in the reference's own byte layout: `struct NFA` followed by `struct
 * LimExNFA32` / `64` / `128` / `256` / `512`, its reach table, accept / EOD-accept tables, exception table,
 * report lists (src/nfa/limex_internal.h:102-203; laid out the way
 * limex_compile.cpp's Factory::generateNfa does, src/nfa/limex_compile.cpp:2300-2480).
 *
 * The reference's compile side (parser -> Glushkov graph -> state numbering) stays
 * out of scope; what a test or a caller can give is a finished NFA of at most 512
 * states (`RawNfa`: per-state successor sets, per-byte reach, reports) or a
 * literal set, for which the position automaton is built here.  Transitions
 * i -> i + a for the (up to eight) most common forward distances a <= 16 become
 * the "limited" shift masks -- inside one 64-bit lane of the state only, as the wider models shift
 * lane by lane (isLimitedTransition, src/nfa/limex_compile.cpp:245-256) --, everything else exception successors; bounded
 * repeats, tops beyond the single start and acceleration are not emitted.  The same
 * bytes run on the unmodified reference engines (nfaExecLimEx32_Q ... 512_Q + _testEOD) --
 * the parity oracle -- and on the device kernel (device/dfa_kernels.cu).
 */
#ifndef HSB200_LIMEX_BUILD_H
#define HSB200_LIMEX_BUILD_H

#include <bitset>
#include <string>
#include <vector>

#include "../ref_layout.h"
#include "dfa_build.h"

namespace hsb {

/* a set of NFA states (or of Glushkov positions): bit i = state i */
static const u32 MAX_NFA_STATES = 512;
typedef std::bitset<MAX_NFA_STATES> StateSet;
inline StateSet stateBit(u32 i) {
    StateSet s;
    s.set(i);
    return s;
}
inline StateSet stateSetOf(u64 lowWord) { return StateSet(lowWord); }
inline StateSet allStates() { return StateSet().set(); }
/* bits 64 j .. 64 j + 63 */
inline u64 stateWord(const StateSet &s, u32 j) { return ((s >> (64 * j)) & StateSet(~0ull)).to_ullong(); }

struct RawNfa {
    u32 nstates = 0;                   /* <= 512; emitted as the smallest of the 32 / 64 / 128 / 256 / 512-state models that holds them */
    StateSet reach[256];               /* states that may be ON after consuming the byte */
    StateSet init, initDS;             /* switched on by a top at offset 0 / at a later offset */
    u32 mlStartState = 0;              /* regex_nfa.cpp: the shared "after a newline" state, 0 = none yet */
    u32 ctxWord = 0, ctxNonWord = 0;   /* regex_nfa.cpp: "the previous byte is / is not a word character" */
    std::vector<StateSet> succ;        /* [state] successor set */
    std::vector<StateSet> squashMask;  /* [state] kept states when the exception's squash applies */
    std::vector<u8> squashKind;        /* [state] LIMEX_SQUASH_NONE / _CYCLIC / _REPORT */
    std::vector<std::vector<u32>> reports;    /* raised while the state is on */
    std::vector<std::vector<u32>> reportsEod; /* raised if the data ends with the state on */
};
typedef RawNfa RawNfa32; /* the name the 32-state-only version had */

/* position automaton of a literal set: state 0 = floating start (always on), one state
 * per literal byte; throws if more than 511 positions are needed */
RawNfa nfaFromLiterals(const std::vector<DfaLiteral> &lits);

/* Subset construction: the DFA of an NFA without squashes (what regex_nfa.cpp builds), for the McClellan
 * emitters -- the reference also turns small NFA graphs into DFAs before it settles for LimEx
 * (src/nfagraph/ng_mcclellan.cpp, determinise with a state limit).  A DFA state's reports are those of its NFA
 * states: raised when the set is entered, i.e. at the offset after the byte, which is where the LimEx run raises
 * them (the exceptions of the next byte, or the accepts after the last one); EOD reports likewise.  Returns false
 * if more than maxStates subsets (the dead state included) are needed or the NFA has squashing exceptions. */
bool determinize(const RawNfa &n, size_t maxStates, RawDfa *out);

/* Merge the states of a DFA that no input tells apart (same reports now and after every byte string): Moore's
 * partition refinement; state 0 stays the dead state.  The reference minimises its DFAs too (minimize_hopcroft,
 * src/nfagraph/ng_mcclellan.cpp via util/determinise.h users). */
void minimizeDfa(RawDfa *dfa);

/* struct NFA + LimExNFA32 / 64 / 128 / 256 / 512 (the smallest that holds nstates) + tables */
std::vector<u8> emitLimEx(const RawNfa &n);
inline std::vector<u8> emitLimEx32(const RawNfa &n) { return emitLimEx(n); }

} // namespace hsb
#endif
