/*
 * dfa_build.h -- host-side emitters of the reference's DFA engines in the
 * reference's own byte layout: `struct NFA` (src/nfa/nfa_internal.h:84-126)
 * followed by McClellan with 8- or 16-bit states (`struct mcclellan`,
 * `struct mstate_aux`, report lists, Sherman states:
 * src/nfa/mcclellan_internal.h:36-106, laid out as mcclellanCompile8/16 do,
 * src/nfa/mcclellancompile.cpp:612-840, 909-1003) or Sheng (`struct sheng`,
 * `struct sstate_aux`: src/nfa/sheng_internal.h:56-79, shengcompile.cpp:630-700).
 *
 * The reference's compile side (parser -> NFA graph -> determinisation) stays
 * out of scope; what a test or a caller can give is a finished DFA (`RawDfa`,
 * the analogue of the reference's raw_dfa, src/nfa/rdfa.h) or a literal set, for
 * which the Aho-Corasick automaton is built here.  The same bytes run on the
 * unmodified reference engines (nfaExecMcClellan8_B / 16_B, nfaExecSheng_B) --
 * the parity oracle -- and on the device kernels (device/dfa_kernels.cu).
 */
#ifndef HSB200_DFA_BUILD_H
#define HSB200_DFA_BUILD_H

#include <array>
#include <string>
#include <vector>

#include "../ref_layout.h"

namespace hsb {

struct RawDfa {
    /* state 0 is the dead state (every transition of it leads to itself, no reports) */
    std::vector<std::array<u16, 256>> next;
    std::vector<std::vector<u32>> reports;    /* raised when the state is entered */
    std::vector<std::vector<u32>> reportsEod; /* raised if the block ends in the state */
    u16 startAnchored = 1, startFloating = 1;
    size_t size() const { return next.size(); }
};

struct DfaLiteral {
    std::string s;
    bool caseless = false;
    u32 report = 0;
};

/* Aho-Corasick automaton of a literal set as a complete DFA: floating (every
 * position may start a literal) or anchored (literals match at offset 0 only). */
RawDfa dfaFromLiterals(const std::vector<DfaLiteral> &lits, bool anchored);

enum DfaKind { DFA_AUTO = 0, DFA_MCCLELLAN8 = 1, DFA_MCCLELLAN16 = 2, DFA_SHENG = 3 };

/* Serialise.  sherman: McClellan16 only -- states that differ from another
 * ("daddy") state's row in at most 8 alphabet symbols are stored as 32-byte
 * exception lists instead of full rows (mcclellancompile.cpp find_better_daddy;
 * any valid choice gives an equivalent engine).  Throws std::runtime_error when
 * the automaton does not fit the kind (more than 256 / 16383 / 16 states). */
std::vector<u8> emitDfa(const RawDfa &d, DfaKind kind, bool sherman);

} // namespace hsb
#endif
