/* limex_build.cpp -- see limex_build.h */
#include "limex_build.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>

namespace hsb {

RawNfa nfaFromLiterals(const std::vector<DfaLiteral> &lits) {
    RawNfa n;
    size_t total = 1;
    for (const DfaLiteral &l : lits) {
        if (l.s.empty()) {
            throw std::runtime_error("empty literal");
        }
        total += l.s.size();
    }
    if (total > MAX_NFA_STATES) {
        throw std::runtime_error("more than 512 NFA states");
    }
    n.nstates = (u32)total;
    n.succ.assign(total, StateSet());
    n.squashMask.assign(total, allStates());
    n.squashKind.assign(total, LIMEX_SQUASH_NONE);
    n.reports.resize(total);
    n.reportsEod.resize(total);
    n.init = n.initDS = stateBit(0);
    n.succ[0] = stateBit(0); /* the start state stays on (.* prefix) */
    for (u32 b = 0; b < 256; b++) {
        n.reach[b] = stateBit(0);
    }
    u32 next = 1;
    for (const DfaLiteral &l : lits) {
        u32 prev = 0;
        for (size_t i = 0; i < l.s.size(); i++) {
            const u32 st = next++;
            n.succ[prev].set(st);
            const u8 c = (u8)l.s[i];
            n.reach[c].set(st);
            if (l.caseless && ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) {
                n.reach[c ^ 0x20].set(st);
            }
            prev = st;
        }
        n.reports[prev].push_back(l.report);
    }
    return n;
}

namespace {

struct SetLess {
    bool operator()(const StateSet &a, const StateSet &b) const {
        for (u32 j = MAX_NFA_STATES / 64; j-- > 0;) {
            const u64 x = stateWord(a, j), y = stateWord(b, j);
            if (x != y) {
                return x < y;
            }
        }
        return false;
    }
};

} // namespace

bool determinize(const RawNfa &n, size_t maxStates, RawDfa *out) {
    for (u8 k : n.squashKind) {
        if (k != LIMEX_SQUASH_NONE) {
            return false;
        }
    }
    StateSet all;
    for (u32 i = 0; i < n.nstates; i++) {
        all.set(i);
    }
    /* bytes with the same reach mask behave alike */
    std::vector<StateSet> classes;
    u8 classOf[256];
    {
        std::map<StateSet, u32, SetLess> seen;
        for (u32 b = 0; b < 256; b++) {
            const StateSet m = n.reach[b] & all;
            auto it = seen.find(m);
            if (it == seen.end()) {
                it = seen.emplace(m, (u32)classes.size()).first;
                classes.push_back(m);
            }
            classOf[b] = (u8)it->second;
        }
    }
    RawDfa d;
    std::map<StateSet, u16, SetLess> ids;
    std::vector<StateSet> sets;
    auto add = [&](const StateSet &s) -> int {
        auto it = ids.find(s);
        if (it != ids.end()) {
            return it->second;
        }
        if (sets.size() >= maxStates || sets.size() >= 0xfffe) {
            return -1;
        }
        const u16 id = (u16)sets.size();
        ids.emplace(s, id);
        sets.push_back(s);
        d.next.emplace_back();
        d.next.back().fill(0);
        std::vector<u32> r, e;
        for (u32 q = 0; q < n.nstates; q++) {
            if (s.test(q)) {
                r.insert(r.end(), n.reports[q].begin(), n.reports[q].end());
                e.insert(e.end(), n.reportsEod[q].begin(), n.reportsEod[q].end());
            }
        }
        for (std::vector<u32> *v : {&r, &e}) {
            std::sort(v->begin(), v->end());
            v->erase(std::unique(v->begin(), v->end()), v->end());
        }
        d.reports.push_back(r);
        d.reportsEod.push_back(e);
        return id;
    };
    add(StateSet()); /* state 0: the dead state */
    const int start = add(n.init & all);
    if (start <= 0) {
        return false;
    }
    d.startAnchored = d.startFloating = (u16)start;
    for (size_t cur = 1; cur < sets.size(); cur++) {
        StateSet succ;
        const StateSet s = sets[cur];
        for (u32 q = 0; q < n.nstates; q++) {
            if (s.test(q)) {
                succ |= n.succ[q];
            }
        }
        std::vector<int> to(classes.size());
        for (size_t c = 0; c < classes.size(); c++) {
            to[c] = add(succ & classes[c]);
            if (to[c] < 0) {
                return false;
            }
        }
        for (u32 b = 0; b < 256; b++) {
            d.next[cur][b] = (u16)to[classOf[b]];
        }
    }
    *out = d;
    return true;
}

void minimizeDfa(RawDfa *dfa) {
    RawDfa &d = *dfa;
    const size_t n = d.size();
    /* bytes that behave alike in every state */
    std::vector<u32> byteClass(256, 0);
    u32 nclasses = 0;
    {
        std::map<std::vector<u16>, u32> cols;
        for (u32 b = 0; b < 256; b++) {
            std::vector<u16> col(n);
            for (size_t i = 0; i < n; i++) {
                col[i] = d.next[i][b];
            }
            auto it = cols.emplace(col, (u32)cols.size()).first;
            byteClass[b] = it->second;
        }
        nclasses = (u32)cols.size();
    }
    std::vector<u32> rep(nclasses, 0); /* one byte of every class */
    for (u32 b = 256; b-- > 0;) {
        rep[byteClass[b]] = b;
    }
    /* Moore's refinement: start from what a state reports, split by where the classes lead */
    std::vector<u32> part(n, 0);
    {
        std::map<std::pair<std::vector<u32>, std::vector<u32>>, u32> sig;
        for (size_t i = 0; i < n; i++) {
            part[i] = sig.emplace(std::make_pair(d.reports[i], d.reportsEod[i]), (u32)sig.size()).first->second;
        }
    }
    for (size_t count = 0;;) {
        std::map<std::vector<u32>, u32> sig;
        std::vector<u32> next(n);
        for (size_t i = 0; i < n; i++) {
            std::vector<u32> key(nclasses + 1);
            key[0] = part[i];
            for (u32 c = 0; c < nclasses; c++) {
                key[c + 1] = part[d.next[i][rep[c]]];
            }
            next[i] = sig.emplace(key, (u32)sig.size()).first->second;
        }
        part.swap(next);
        if (sig.size() == count) {
            break;
        }
        count = sig.size();
    }
    /* renumber: the dead state's class stays 0, the rest in order of first appearance */
    std::vector<int> id(n, -1);
    std::vector<size_t> first;
    id[part[0]] = 0;
    first.push_back(0);
    for (size_t i = 1; i < n; i++) {
        if (id[part[i]] < 0) {
            id[part[i]] = (int)first.size();
            first.push_back(i);
        }
    }
    if (first.size() == n) {
        return;
    }
    RawDfa m;
    m.next.resize(first.size());
    for (size_t k = 0; k < first.size(); k++) {
        for (u32 b = 0; b < 256; b++) {
            m.next[k][b] = (u16)id[part[d.next[first[k]][b]]];
        }
        m.reports.push_back(d.reports[first[k]]);
        m.reportsEod.push_back(d.reportsEod[first[k]]);
    }
    m.startAnchored = (u16)id[part[d.startAnchored]];
    m.startFloating = (u16)id[part[d.startFloating]];
    d = m;
}

namespace {

template <class T> void put(std::vector<u8> &b, size_t off, const T &v) {
    if (b.size() < off + sizeof(T)) {
        b.resize(off + sizeof(T), 0);
    }
    memcpy(b.data() + off, &v, sizeof(T));
}

size_t alignUp(std::vector<u8> &b, size_t a) {
    while (b.size() % a) {
        b.push_back(0);
    }
    return b.size();
}

} // namespace

/* a state set in an engine field: u32 / u64 take the low word, the wide models all their words */
template <class T> T fromSet(const StateSet &s) {
    T t;
    memset(&t, 0, sizeof(t));
    for (u32 j = 0; j < sizeof(T) / 8; j++) {
        const u64 w = stateWord(s, j);
        memcpy((u8 *)&t + 8 * j, &w, 8);
    }
    if (sizeof(T) < 8) {
        const u32 w = (u32)stateWord(s, 0);
        memcpy(&t, &w, sizeof(T));
    }
    return t;
}

template <class LX, class EX, class T> std::vector<u8> emitLimExT(const RawNfa &n, u8 nfaType) {
    const u32 WIDTH = (u32)sizeof(T) * 8;
    if (n.nstates == 0 || n.nstates > WIDTH || n.succ.size() != n.nstates || n.reports.size() != n.nstates ||
        n.reportsEod.size() != n.nstates || n.squashMask.size() != n.nstates || n.squashKind.size() != n.nstates) {
        throw std::runtime_error("bad NFA description");
    }
    StateSet all;
    for (u32 i = 0; i < n.nstates; i++) {
        all.set(i);
    }
    LX lx;
    memset(&lx, 0, sizeof(lx));

    /* reach classes: bytes with the same reach mask share an entry */
    std::vector<StateSet> reachTab;
    {
        std::map<StateSet, u32, SetLess> seen;
        for (u32 b = 0; b < 256; b++) {
            const StateSet m = n.reach[b] & all;
            auto it = seen.find(m);
            if (it == seen.end()) {
                it = seen.emplace(m, (u32)reachTab.size()).first;
                reachTab.push_back(m);
            }
            lx.reachMap[b] = (u8)it->second;
        }
    }
    lx.reachSize = (u32)reachTab.size();

    /* limited transitions: the most common forward distances become shifts; a shift never carries a
     * bit from one 64-bit lane of the state into the next ("can't jump over a bollard") */
    auto limited = [](u32 from, u32 to) { return (from & ~63u) == (to & ~63u); };
    u32 count[17] = {0};
    for (u32 i = 0; i < n.nstates; i++) {
        for (u32 j = i; j < n.nstates && j <= i + 16; j++) {
            if (n.succ[i].test(j) && limited(i, j)) {
                count[j - i]++;
            }
        }
    }
    std::vector<u32> amounts;
    for (u32 a = 0; a <= 16; a++) {
        if (count[a]) {
            amounts.push_back(a);
        }
    }
    std::sort(amounts.begin(), amounts.end(), [&](u32 x, u32 y) { return count[x] != count[y] ? count[x] > count[y] : x < y; });
    if (amounts.size() > 8) {
        amounts.resize(8);
    }
    std::sort(amounts.begin(), amounts.end());
    std::vector<StateSet> exceptional(n.nstates); /* successors not covered by a shift */
    for (u32 i = 0; i < n.nstates; i++) {
        exceptional[i] = n.succ[i] & all;
    }
    lx.shiftCount = std::max<u32>(1, (u32)amounts.size()); /* "should be always greater or equal to 1" */
    for (size_t k = 0; k < amounts.size(); k++) {
        const u32 a = amounts[k];
        lx.shiftAmount[k] = (u8)a;
        StateSet mask;
        for (u32 i = 0; i + a < n.nstates; i++) {
            if (n.succ[i].test(i + a) && limited(i, i + a)) {
                mask.set(i);
                exceptional[i].reset(i + a);
            }
        }
        lx.shift[k] = fromSet<T>(mask);
    }

    StateSet accept, acceptEod, excMask;
    for (u32 i = 0; i < n.nstates; i++) {
        if (!n.reports[i].empty()) {
            accept.set(i);
        }
        if (!n.reportsEod[i].empty()) {
            acceptEod.set(i);
        }
        if (exceptional[i].any() || !n.reports[i].empty() || n.squashKind[i] != LIMEX_SQUASH_NONE) {
            excMask.set(i);
        }
    }
    lx.accept = fromSet<T>(accept);
    lx.acceptAtEOD = fromSet<T>(acceptEod);
    lx.exceptionMask = fromSet<T>(excMask);
    lx.init = fromSet<T>(n.init & all);
    lx.initDS = fromSet<T>(n.initDS & all);
    lx.stateSize = (n.nstates + 7) / 8;
    lx.acceptCount = (u32)accept.count();
    lx.acceptEodCount = (u32)acceptEod.count();
    lx.exceptionCount = (u32)excMask.count();

    /* body after the struct: reach table, report lists, accept tables, exception table
     * (offsets relative to the LimExNFA) */
    std::vector<u8> body(sizeof(LX), 0);
    for (size_t i = 0; i < reachTab.size(); i++) {
        put(body, sizeof(LX) + sizeof(T) * i, fromSet<T>(reachTab[i]));
    }
    auto reportList = [&](const std::vector<u32> &r) -> u32 {
        const u32 off = (u32)alignUp(body, 4);
        for (u32 id : r) {
            put(body, body.size(), id);
        }
        put(body, body.size(), MO_INVALID_IDX);
        return off;
    };
    std::vector<u32> listOff(n.nstates, MO_INVALID_IDX), listOffEod(n.nstates, MO_INVALID_IDX);
    for (u32 i = 0; i < n.nstates; i++) {
        if (!n.reports[i].empty()) {
            listOff[i] = reportList(n.reports[i]);
        }
        if (!n.reportsEod[i].empty()) {
            listOffEod[i] = reportList(n.reportsEod[i]);
        }
    }
    auto acceptTable = [&](const StateSet &mask, const std::vector<std::vector<u32>> &reps, const std::vector<u32> &offs) -> u32 {
        const u32 off = (u32)alignUp(body, 4);
        for (u32 i = 0; i < n.nstates; i++) {
            if (!mask.test(i)) {
                continue;
            }
            NFAAccept a;
            memset(&a, 0, sizeof(a));
            a.single_report = reps[i].size() == 1;
            a.reports = a.single_report ? reps[i][0] : offs[i];
            a.squash = MO_INVALID_IDX;
            put(body, body.size(), a);
        }
        return off;
    };
    lx.acceptOffset = acceptTable(accept, n.reports, listOff);
    lx.acceptEodOffset = acceptTable(acceptEod, n.reportsEod, listOffEod);
    lx.exceptionOffset = (u32)alignUp(body, alignof(EX) > 16 ? alignof(EX) : 16);
    for (u32 i = 0; i < n.nstates; i++) {
        if (!excMask.test(i)) {
            continue;
        }
        EX e;
        memset(&e, 0, sizeof(e));
        e.squash = fromSet<T>(n.squashKind[i] != LIMEX_SQUASH_NONE ? (n.squashMask[i] & all) : all);
        e.successors = fromSet<T>(exceptional[i]);
        e.reports = listOff[i];
        e.repeatOffset = MO_INVALID_IDX;
        e.hasSquash = n.squashKind[i];
        e.trigger = LIMEX_TRIGGER_NONE;
        put(body, body.size(), e);
    }
    lx.accelTableOffset = lx.accelAuxOffset = lx.repeatOffset = lx.squashOffset = lx.topOffset = (u32)alignUp(body, 16);
    alignUp(body, 64);
    memcpy(body.data(), &lx, sizeof(lx));

    NFA hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.type = nfaType;
    hdr.length = (u32)(sizeof(NFA) + body.size());
    hdr.nPositions = n.nstates;
    hdr.scratchStateSize = (u32)std::max<size_t>(sizeof(T), nfaType == NFA_LIMEX_64 ? 16 : 0); /* NFATraits::scratch_state_size, limex_compile.cpp:2419-2428 */
    hdr.streamStateSize = lx.stateSize;
    hdr.flags = lx.acceptEodCount ? NFA_ACCEPTS_EOD : 0;
    std::vector<u8> out(sizeof(NFA));
    memcpy(out.data(), &hdr, sizeof(hdr));
    out.insert(out.end(), body.begin(), body.end());
    return out;
}

std::vector<u8> emitLimEx(const RawNfa &n) {
    if (n.nstates <= 32) {
        return emitLimExT<LimExNFA32, NFAException32, u32>(n, NFA_LIMEX_32);
    }
    if (n.nstates <= 64) {
        return emitLimExT<LimExNFA64, NFAException64, u64>(n, NFA_LIMEX_64);
    }
    if (n.nstates <= 128) {
        return emitLimExT<LimExNFA128, NFAException128, StateWord128>(n, NFA_LIMEX_128);
    }
    if (n.nstates <= 256) {
        return emitLimExT<LimExNFA256, NFAException256, StateWord256>(n, NFA_LIMEX_256);
    }
    return emitLimExT<LimExNFA512, NFAException512, StateWord512>(n, NFA_LIMEX_512);
}

} // namespace hsb
