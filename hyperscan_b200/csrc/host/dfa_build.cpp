/* dfa_build.cpp -- see dfa_build.h */
#include "dfa_build.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>

#include "hwlm_build.h"

namespace hsb {

namespace {

struct Out {
    std::vector<u8> b;
    void grow(size_t n) {
        if (b.size() < n) {
            b.resize(n, 0);
        }
    }
    template <class T> void put(size_t off, const T &v) {
        grow(off + sizeof(T));
        memcpy(b.data() + off, &v, sizeof(T));
    }
};

u32 ceilLog2(u32 v) {
    u32 s = 0;
    while ((1u << s) < v) {
        s++;
    }
    return s;
}

/* bytes with identical columns share an alphabet symbol (the reference's
 * alpha_remap, src/nfa/rdfa.h; symbol count + 1: the extra TOP symbol) */
u32 alphabetOf(const RawDfa &d, u8 (&remap)[256]) {
    std::map<std::vector<u16>, u32> seen;
    for (u32 c = 0; c < 256; c++) {
        std::vector<u16> col(d.size());
        for (size_t s = 0; s < d.size(); s++) {
            col[s] = d.next[s][c];
        }
        auto it = seen.find(col);
        if (it == seen.end()) {
            it = seen.emplace(std::move(col), (u32)seen.size()).first;
        }
        remap[c] = (u8)it->second;
    }
    return (u32)seen.size();
}

/* report lists (struct report_list {u32 count; ReportID report[]}), identical lists shared */
struct ReportLists {
    std::map<std::vector<u32>, u32> offsetOf; /* relative to the start of the list area */
    std::vector<u8> bytes;
    u32 add(std::vector<u32> ids) {
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        auto it = offsetOf.find(ids);
        if (it != offsetOf.end()) {
            return it->second;
        }
        const u32 off = (u32)bytes.size();
        const u32 n = (u32)ids.size();
        bytes.resize(off + 4 + 4 * n);
        memcpy(bytes.data() + off, &n, 4);
        memcpy(bytes.data() + off + 4, ids.data(), 4 * n);
        offsetOf.emplace(std::move(ids), off);
        return off;
    }
};

/* "single": every accepting state raises the same one report and nothing fires at EOD only */
bool singleReport(const RawDfa &d, u32 *arb) {
    bool have = false, single = true;
    u32 r = 0;
    for (size_t s = 0; s < d.size(); s++) {
        for (const auto *lst : {&d.reports[s], &d.reportsEod[s]}) {
            for (u32 x : *lst) {
                if (!have) {
                    have = true;
                    r = x;
                } else if (x != r) {
                    single = false;
                }
            }
        }
        if (d.reports[s].size() > 1) {
            single = false;
        }
    }
    *arb = r;
    return have && single;
}

bool anyEod(const RawDfa &d) {
    for (const auto &l : d.reportsEod) {
        if (!l.empty()) {
            return true;
        }
    }
    return false;
}

void putNfaHeader(Out &o, u8 type, u32 total, const RawDfa &d, u32 stateBytes) {
    NFA n;
    memset(&n, 0, sizeof(n));
    n.flags = anyEod(d) ? NFA_ACCEPTS_EOD : 0;
    n.length = total;
    n.type = type;
    n.nPositions = (u32)d.size();
    n.scratchStateSize = stateBytes;
    n.streamStateSize = stateBytes;
    o.put(0, n);
}

std::vector<u8> emitMcClellan(const RawDfa &d, bool wide16, bool sherman) {
    const u32 n = (u32)d.size();
    if (n > (wide16 ? 16383u : 256u)) {
        throw std::runtime_error("too many states for this McClellan width");
    }
    u8 remap[256];
    const u32 alpha = alphabetOf(d, remap);
    const u32 alphaSize = alpha + 1; /* + TOP */
    const u32 as = ceilLog2(alphaSize);
    /* one representative byte per symbol */
    std::vector<u32> repByte(alpha, 0);
    for (u32 c = 256; c-- > 0;) {
        repByte[remap[c]] = c;
    }

    /* --- state numbering --- */
    std::vector<u16> impl(n, 0);
    std::vector<u32> daddy(n, 0);       /* sherman states: raw id of the row they are stored against */
    std::vector<bool> isSherman(n, false);
    u32 countReal = n, accelLimit8 = 0, acceptLimit8 = 0;
    if (wide16) {
        if (sherman && n > 2) {
            /* a state becomes a Sherman state when its row differs from the floating
             * start state's row in at most 8 symbols; the start state stays a full row */
            const u32 d0 = d.startFloating ? d.startFloating : 1;
            for (u32 s = 1; s < n; s++) {
                if (s == d0 || s == d.startAnchored) {
                    continue;
                }
                u32 diff = 0;
                for (u32 a = 0; a < alpha; a++) {
                    diff += d.next[s][repByte[a]] != d.next[d0][repByte[a]];
                }
                if (diff <= 8) {
                    isSherman[s] = true;
                    daddy[s] = d0;
                }
            }
        }
        u16 j = 1;
        for (u32 s = 1; s < n; s++) {
            if (!isSherman[s]) {
                impl[s] = j++;
            }
        }
        countReal = j;
        for (u32 s = 1; s < n; s++) {
            if (isSherman[s]) {
                impl[s] = j++;
            }
        }
    } else {
        /* 8 bit: dead, then the states without reports, then the accepting ones
         * (allocateFSN8, mcclellancompile.cpp:884-925; no accelerated states here) */
        u16 j = 1;
        for (u32 s = 1; s < n; s++) {
            if (d.reports[s].empty()) {
                impl[s] = j++;
            }
        }
        accelLimit8 = acceptLimit8 = j;
        for (u32 s = 1; s < n; s++) {
            if (!d.reports[s].empty()) {
                impl[s] = j++;
            }
        }
    }

    /* --- sizes and offsets (mcclellanCompile16 / 8) --- */
    const u32 stateBytes = wide16 ? 2 : 1;
    const size_t tranSize = ((size_t)1 << as) * stateBytes * (wide16 ? countReal : n);
    const size_t auxOffset = HSB_ROUNDUP(sizeof(NFA) + sizeof(McClellan) + tranSize, 16);
    const size_t auxSize = sizeof(MStateAux) * n;
    ReportLists rl;
    std::vector<u32> acc(n, 0xffffffffu), accEod(n, 0xffffffffu);
    for (u32 s = 0; s < n; s++) {
        if (!d.reports[s].empty()) {
            acc[s] = rl.add(d.reports[s]);
        }
        if (!d.reportsEod[s].empty()) {
            accEod[s] = rl.add(d.reportsEod[s]);
        }
    }
    const size_t rlOffset = auxOffset + auxSize;
    size_t accelOffset = HSB_ROUNDUP(rlOffset + rl.bytes.size(), 32);
    size_t total = accelOffset;
    size_t shermanOffset = 0, wideOffset = 0;
    u32 nSherman = n - countReal;
    if (wide16) {
        shermanOffset = HSB_ROUNDUP(accelOffset, 16);
        wideOffset = HSB_ROUNDUP(shermanOffset + (size_t)SHERMAN_FIXED_SIZE * nSherman, 16);
        total = wideOffset;
    }
    Out o;
    o.grow(total);
    putNfaHeader(o, wide16 ? NFA_MCCLELLAN_16 : NFA_MCCLELLAN_8, (u32)total, d, stateBytes);

    McClellan m;
    memset(&m, 0, sizeof(m));
    m.state_count = (u16)n;
    m.length = (u32)total;
    m.start_anchored = impl[d.startAnchored];
    m.start_floating = impl[d.startFloating];
    m.aux_offset = (u32)auxOffset;
    m.sherman_offset = (u32)shermanOffset;
    m.sherman_end = (u32)total;
    m.accel_limit_8 = (u16)accelLimit8;
    m.accept_limit_8 = (u16)acceptLimit8;
    m.sherman_limit = (u16)countReal;
    m.wide_limit = (u16)n;
    m.alphaShift = (u8)as;
    u32 arb = 0;
    m.flags = singleReport(d, &arb) ? MCCLELLAN_FLAG_SINGLE : 0;
    m.arb_report = arb;
    memcpy(m.remap, remap, 256);
    m.accel_offset = (u32)(accelOffset - sizeof(NFA));
    m.wide_offset = (u32)wideOffset;
    o.put(sizeof(NFA), m);

    memcpy(o.b.data() + rlOffset, rl.bytes.data(), rl.bytes.size());

    auto entry = [&](u32 target) -> u16 { /* successor as stored: id, ACCEPT_FLAG if it raises reports */
        u16 e = impl[target];
        if (wide16 && !d.reports[target].empty()) {
            e |= MCC_ACCEPT_FLAG;
        }
        return e;
    };
    const size_t succBase = sizeof(NFA) + sizeof(McClellan);
    for (u32 s = 0; s < n; s++) {
        MStateAux aux;
        memset(&aux, 0, sizeof(aux));
        aux.accept = acc[s] == 0xffffffffu ? 0 : (u32)(rlOffset + acc[s]);
        aux.accept_eod = accEod[s] == 0xffffffffu ? 0 : (u32)(rlOffset + accEod[s]);
        aux.top = s ? impl[s] : impl[d.startFloating]; /* no TOP events in block mode */
        o.put(auxOffset + sizeof(MStateAux) * impl[s], aux);
        if (isSherman[s]) {
            const size_t rec = shermanOffset + (size_t)SHERMAN_FIXED_SIZE * (impl[s] - countReal);
            u8 len = 0;
            u8 chars[9];
            u16 succs[9];
            for (u32 a = 0; a < alpha; a++) {
                if (d.next[s][repByte[a]] != d.next[daddy[s]][repByte[a]]) {
                    chars[len] = (u8)a;
                    succs[len] = entry(d.next[s][repByte[a]]);
                    len++;
                }
            }
            o.b[rec + SHERMAN_TYPE_OFFSET] = SHERMAN_STATE;
            o.b[rec + SHERMAN_LEN_OFFSET] = len;
            const u16 dd = impl[daddy[s]];
            memcpy(o.b.data() + rec + SHERMAN_DADDY_OFFSET, &dd, 2);
            memcpy(o.b.data() + rec + SHERMAN_CHARS_OFFSET, chars, len);
            memcpy(o.b.data() + rec + SHERMAN_CHARS_OFFSET + len, succs, 2 * (size_t)len);
            continue;
        }
        for (u32 a = 0; a < alphaSize; a++) {
            const u32 target = a < alpha ? d.next[s][repByte[a]] : (s ? s : d.startFloating); /* a == alpha: TOP */
            const size_t idx = ((size_t)impl[s] << as) + a;
            if (wide16) {
                const u16 e = entry(target);
                memcpy(o.b.data() + succBase + 2 * idx, &e, 2);
            } else {
                o.b[succBase + idx] = (u8)impl[target];
            }
        }
    }
    return o.b;
}

std::vector<u8> emitSheng(const RawDfa &d) {
    const u32 n = (u32)d.size();
    if (n > 16) {
        throw std::runtime_error("too many states for Sheng");
    }
    auto stateByte = [&](u32 s) -> u8 {
        u8 v = (u8)s;
        if (!d.reports[s].empty()) {
            v |= SHENG_STATE_ACCEPT;
        }
        bool dead = d.reports[s].empty() && d.reportsEod[s].empty();
        for (u32 c = 0; c < 256 && dead; c++) {
            dead = d.next[s][c] == s;
        }
        if (dead) {
            v |= SHENG_STATE_DEAD;
        }
        return v;
    };
    const size_t auxOffset = HSB_ROUNDUP(sizeof(NFA) + sizeof(Sheng), 16);
    const size_t reportOffset = auxOffset + sizeof(SstateAux) * n;
    ReportLists rl;
    std::vector<u32> acc(n, 0xffffffffu), accEod(n, 0xffffffffu);
    for (u32 s = 0; s < n; s++) {
        if (!d.reports[s].empty()) {
            acc[s] = rl.add(d.reports[s]);
        }
        if (!d.reportsEod[s].empty()) {
            accEod[s] = rl.add(d.reportsEod[s]);
        }
    }
    const size_t accelOffset = HSB_ROUNDUP(reportOffset + rl.bytes.size(), 16);
    const size_t total = HSB_ROUNDUP(accelOffset, 64);
    Out o;
    o.grow(total);
    putNfaHeader(o, NFA_SHENG, (u32)total, d, 1);
    Sheng sh;
    memset(&sh, 0, sizeof(sh));
    bool canDie = false;
    for (u32 c = 0; c < 256; c++) {
        for (u32 s = 0; s < n; s++) {
            const u8 nx = stateByte(d.next[s][c]);
            sh.shuffle_masks[c][s] = nx;
            canDie |= (nx & SHENG_STATE_DEAD) && !(stateByte(s) & SHENG_STATE_DEAD);
        }
    }
    sh.length = (u32)(total - sizeof(NFA));
    sh.aux_offset = (u32)auxOffset;
    sh.report_offset = (u32)reportOffset;
    sh.accel_offset = (u32)accelOffset;
    sh.n_states = (u8)n;
    sh.anchored = stateByte(d.startAnchored);
    sh.floating = stateByte(d.startFloating);
    u32 arb = 0;
    sh.flags = (canDie ? SHENG_FLAG_CAN_DIE : 0) | (singleReport(d, &arb) ? SHENG_FLAG_SINGLE_REPORT : 0);
    sh.report = arb;
    o.put(sizeof(NFA), sh);
    memcpy(o.b.data() + reportOffset, rl.bytes.data(), rl.bytes.size());
    for (u32 s = 0; s < n; s++) {
        SstateAux aux;
        memset(&aux, 0, sizeof(aux));
        aux.accept = acc[s] == 0xffffffffu ? 0 : (u32)(reportOffset + acc[s]);
        aux.accept_eod = accEod[s] == 0xffffffffu ? 0 : (u32)(reportOffset + accEod[s]);
        aux.top = stateByte(s ? s : d.startFloating);
        o.put(auxOffset + sizeof(SstateAux) * s, aux);
    }
    return o.b;
}

} // namespace

RawDfa dfaFromLiterals(const std::vector<DfaLiteral> &lits, bool anchored) {
    /* Position automaton of the set -- one chain of positions per literal, a caseless
     * letter admitting both cases, the start position looping on every byte unless the
     * set is anchored -- determinised by the subset construction over byte classes
     * (bytes no literal tells apart).  For literal sets this is the Aho-Corasick
     * automaton, without enumerating case variants. */
    struct Pos {
        u8 lo, hi;  /* the two bytes accepted (equal unless a caseless letter) */
        u32 next;   /* following position, 0 = the literal ends here */
        u32 report;
    };
    std::vector<Pos> pos(1); /* position 0 = start */
    std::vector<u32> first;  /* first position of every literal */
    for (const DfaLiteral &l : lits) {
        if (l.s.empty()) {
            throw std::runtime_error("empty literal");
        }
        first.push_back((u32)pos.size());
        for (size_t i = 0; i < l.s.size(); i++) {
            Pos p;
            const u8 c = (u8)l.s[i];
            const bool fold = l.caseless && isAsciiAlpha(c);
            p.lo = fold ? asciiLower(c) : c;
            p.hi = fold ? asciiUpper(c) : c;
            p.next = i + 1 < l.s.size() ? (u32)pos.size() + 1 : 0;
            p.report = l.report;
            pos.push_back(p);
        }
    }
    /* byte classes: bytes that occur in no literal behave alike */
    bool used[256] = {false};
    for (size_t i = 1; i < pos.size(); i++) {
        used[pos[i].lo] = used[pos[i].hi] = true;
    }
    std::vector<u8> reps;
    int other = -1;
    for (u32 c = 0; c < 256; c++) {
        if (used[c]) {
            reps.push_back((u8)c);
        } else if (other < 0) {
            other = (int)c;
            reps.push_back((u8)c);
        }
    }
    /* DFA state = sorted set of "literal positions about to be matched" (the start
     * position is implicit: always active when floating, active in the start state only
     * when anchored) */
    typedef std::vector<u32> Set;
    std::map<Set, u32> idOf;
    std::vector<Set> sets;
    RawDfa d;
    auto addState = [&](const Set &st, const std::vector<u32> &reports) -> u32 {
        /* accepting-ness is part of the state's identity: it is entered WITH these reports */
        Set key = st;
        key.push_back(0xffffffffu);
        key.insert(key.end(), reports.begin(), reports.end());
        auto it = idOf.find(key);
        if (it != idOf.end()) {
            return it->second;
        }
        const u32 id = (u32)sets.size();
        if (id >= 16383) {
            throw std::runtime_error("literal set too large for a 16-bit DFA");
        }
        idOf.emplace(std::move(key), id);
        sets.push_back(st);
        d.next.push_back(std::array<u16, 256>());
        d.next.back().fill(0);
        d.reports.push_back(reports);
        d.reportsEod.push_back({});
        return id;
    };
    /* 0 = dead: a row of its own, never looked up by set */
    sets.push_back(Set());
    d.next.push_back(std::array<u16, 256>());
    d.next.back().fill(0);
    d.reports.push_back({});
    d.reportsEod.push_back({});
    /* the literals' first positions are active in the start state and -- floating sets --
     * in every other state too, so they stay out of the sets: per byte, what stepping them
     * yields */
    std::vector<std::vector<u32>> firstNext(256), firstReports(256);
    for (u32 f : first) {
        for (u32 c : {(u32)pos[f].lo, (u32)pos[f].hi}) {
            if (pos[f].next) {
                firstNext[c].push_back(pos[f].next);
            } else {
                firstReports[c].push_back(pos[f].report);
            }
            if (pos[f].lo == pos[f].hi) {
                break;
            }
        }
    }
    Set startMark;
    if (anchored) {
        startMark.push_back(0); /* position 0 marks "at offset 0": only there do the literals begin */
    }
    const u32 start = addState(startMark, {});
    d.startAnchored = (u16)start;
    d.startFloating = anchored ? 0 : (u16)start;
    for (u32 sid = 1; sid < sets.size(); sid++) {
        const Set cur = sets[sid];
        for (u8 c : reps) {
            Set nxt;
            std::vector<u32> reports;
            bool starts = !anchored;
            for (u32 pi : cur) {
                if (pi == 0) {
                    starts = true;
                    continue;
                }
                const Pos &p = pos[pi];
                if (c == p.lo || c == p.hi) {
                    if (p.next) {
                        nxt.push_back(p.next);
                    } else {
                        reports.push_back(p.report);
                    }
                }
            }
            if (starts) {
                nxt.insert(nxt.end(), firstNext[c].begin(), firstNext[c].end());
                reports.insert(reports.end(), firstReports[c].begin(), firstReports[c].end());
            }
            std::sort(nxt.begin(), nxt.end());
            nxt.erase(std::unique(nxt.begin(), nxt.end()), nxt.end());
            std::sort(reports.begin(), reports.end());
            reports.erase(std::unique(reports.begin(), reports.end()), reports.end());
            /* floating: the empty set is the start state (it can always begin a literal);
             * anchored: nothing left to match = dead */
            const u32 tid = nxt.empty() && reports.empty() ? (anchored ? 0 : start) : addState(nxt, reports);
            d.next[sid][c] = (u16)tid;
        }
        /* bytes outside every literal share the transition of their representative */
        if (other >= 0) {
            for (u32 c = 0; c < 256; c++) {
                if (!used[c]) {
                    d.next[sid][c] = d.next[sid][other];
                }
            }
        }
    }
    return d;
}

std::vector<u8> emitDfa(const RawDfa &d, DfaKind kind, bool sherman) {
    if (d.size() < 2 || d.reports.size() != d.size() || d.reportsEod.size() != d.size() ||
        d.startAnchored >= d.size() || d.startFloating >= d.size()) {
        throw std::runtime_error("malformed DFA");
    }
    for (const auto &row : d.next) {
        for (u16 x : row) {
            if (x >= d.size()) {
                throw std::runtime_error("transition out of range");
            }
        }
    }
    if (kind == DFA_AUTO) {
        kind = d.size() <= 16 ? DFA_SHENG : d.size() <= 256 ? DFA_MCCLELLAN8 : DFA_MCCLELLAN16;
    }
    switch (kind) {
    case DFA_SHENG:
        return emitSheng(d);
    case DFA_MCCLELLAN8:
        return emitMcClellan(d, false, false);
    default:
        return emitMcClellan(d, true, sherman);
    }
}

} // namespace hsb
