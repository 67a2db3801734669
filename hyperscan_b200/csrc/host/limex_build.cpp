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
    if (total > 64) {
        throw std::runtime_error("more than 64 NFA states");
    }
    n.nstates = (u32)total;
    n.succ.assign(total, 0);
    n.squashMask.assign(total, ~0ull);
    n.squashKind.assign(total, LIMEX_SQUASH_NONE);
    n.reports.resize(total);
    n.reportsEod.resize(total);
    n.init = n.initDS = 1u;
    n.succ[0] = 1u; /* the start state stays on (.* prefix) */
    for (u32 b = 0; b < 256; b++) {
        n.reach[b] = 1u;
    }
    u32 next = 1;
    for (const DfaLiteral &l : lits) {
        u32 prev = 0;
        for (size_t i = 0; i < l.s.size(); i++) {
            const u32 st = next++;
            n.succ[prev] |= 1ull << st;
            const u8 c = (u8)l.s[i];
            n.reach[c] |= 1ull << st;
            if (l.caseless && ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) {
                n.reach[c ^ 0x20] |= 1ull << st;
            }
            prev = st;
        }
        n.reports[prev].push_back(l.report);
    }
    return n;
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

template <class LX, class EX, class T> std::vector<u8> emitLimExT(const RawNfa &n, u8 nfaType) {
    const u32 WIDTH = (u32)sizeof(T) * 8;
    if (n.nstates == 0 || n.nstates > WIDTH || n.succ.size() != n.nstates || n.reports.size() != n.nstates ||
        n.reportsEod.size() != n.nstates || n.squashMask.size() != n.nstates || n.squashKind.size() != n.nstates) {
        throw std::runtime_error("bad NFA description");
    }
    const T all = n.nstates == WIDTH ? (T)~(T)0 : (T)(((T)1 << n.nstates) - 1);
    LX lx;
    memset(&lx, 0, sizeof(lx));

    /* reach classes: bytes with the same reach mask share an entry */
    std::vector<T> reachTab;
    {
        std::map<T, u32> seen;
        for (u32 b = 0; b < 256; b++) {
            const T m = (T)n.reach[b] & all;
            auto it = seen.find(m);
            if (it == seen.end()) {
                it = seen.emplace(m, (u32)reachTab.size()).first;
                reachTab.push_back(m);
            }
            lx.reachMap[b] = (u8)it->second;
        }
    }
    lx.reachSize = (u32)reachTab.size();

    /* limited transitions: the most common forward distances become shifts */
    u32 count[17] = {0};
    for (u32 i = 0; i < n.nstates; i++) {
        for (u32 j = i; j < n.nstates && j <= i + 16; j++) {
            if ((n.succ[i] >> j) & 1) {
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
    std::vector<T> exceptional(n.nstates, 0); /* successors not covered by a shift */
    for (u32 i = 0; i < n.nstates; i++) {
        exceptional[i] = (T)n.succ[i] & all;
    }
    lx.shiftCount = std::max<u32>(1, (u32)amounts.size()); /* "should be always greater or equal to 1" */
    for (size_t k = 0; k < amounts.size(); k++) {
        const u32 a = amounts[k];
        lx.shiftAmount[k] = (u8)a;
        for (u32 i = 0; i + a < n.nstates; i++) {
            if ((n.succ[i] >> (i + a)) & 1) {
                lx.shift[k] |= (T)1 << i;
                exceptional[i] &= ~((T)1 << (i + a));
            }
        }
    }

    for (u32 i = 0; i < n.nstates; i++) {
        if (!n.reports[i].empty()) {
            lx.accept |= (T)1 << i;
        }
        if (!n.reportsEod[i].empty()) {
            lx.acceptAtEOD |= (T)1 << i;
        }
        if (exceptional[i] || !n.reports[i].empty() || n.squashKind[i] != LIMEX_SQUASH_NONE) {
            lx.exceptionMask |= (T)1 << i;
        }
    }
    lx.init = (T)n.init & all;
    lx.initDS = (T)n.initDS & all;
    lx.stateSize = (n.nstates + 7) / 8;
    lx.acceptCount = (u32)__builtin_popcountll(lx.accept);
    lx.acceptEodCount = (u32)__builtin_popcountll(lx.acceptAtEOD);
    lx.exceptionCount = (u32)__builtin_popcountll(lx.exceptionMask);

    /* body after the struct: reach table, report lists, accept tables, exception table
     * (offsets relative to the LimExNFA32) */
    std::vector<u8> body(sizeof(LX), 0);
    for (size_t i = 0; i < reachTab.size(); i++) {
        put(body, sizeof(LX) + sizeof(T) * i, reachTab[i]);
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
    auto acceptTable = [&](T mask, const std::vector<std::vector<u32>> &reps, const std::vector<u32> &offs) -> u32 {
        const u32 off = (u32)alignUp(body, 4);
        for (u32 i = 0; i < n.nstates; i++) {
            if (!((mask >> i) & 1)) {
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
    lx.acceptOffset = acceptTable(lx.accept, n.reports, listOff);
    lx.acceptEodOffset = acceptTable(lx.acceptAtEOD, n.reportsEod, listOffEod);
    lx.exceptionOffset = (u32)alignUp(body, 16);
    for (u32 i = 0; i < n.nstates; i++) {
        if (!((lx.exceptionMask >> i) & 1)) {
            continue;
        }
        EX e;
        memset(&e, 0, sizeof(e));
        e.squash = n.squashKind[i] != LIMEX_SQUASH_NONE ? ((T)n.squashMask[i] & all) : all;
        e.successors = exceptional[i];
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
    hdr.scratchStateSize = (u32)sizeof(T);
    hdr.streamStateSize = lx.stateSize;
    hdr.flags = lx.acceptEodCount ? NFA_ACCEPTS_EOD : 0;
    std::vector<u8> out(sizeof(NFA));
    memcpy(out.data(), &hdr, sizeof(hdr));
    out.insert(out.end(), body.begin(), body.end());
    return out;
}

std::vector<u8> emitLimEx(const RawNfa &n) {
    return n.nstates <= 32 ? emitLimExT<LimExNFA32, NFAException32, u32>(n, NFA_LIMEX_32)
                           : emitLimExT<LimExNFA64, NFAException64, u64>(n, NFA_LIMEX_64);
}

} // namespace hsb
