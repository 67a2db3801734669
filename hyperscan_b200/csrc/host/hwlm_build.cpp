/*
 * hwlm_build.cpp -- host-side builder for the literal matcher tables.
 *
 * A from-scratch restatement of what the reference's table builders produce
 * (we cannot build those here: they need Boost):
 *
 *   noodle table     src/hwlm/noodle_build.cpp:56-137
 *   FDR table        src/fdr/fdr_compile.cpp:130-212 (layout), :386-512
 *                    (bucket assignment), :527-632 (table fill);
 *                    engine choice src/fdr/fdr_engine_description.cpp:63-182
 *   Teddy masks      src/fdr/teddy_compile.cpp:151-316 (packing), :440-509
 *                    (nibble masks), :512-558 (reinforcement), :560-620
 *                    (layout); engine choice teddy_engine_description.cpp:55-199
 *   hash confirm     src/fdr/fdr_confirm_compile.cpp:73-339
 *   flood control    src/fdr/flood_compile.cpp:93-231
 *   HWLM header      src/hwlm/hwlm_build.cpp:120-163
 *
 * Output bytes follow the reference layouts exactly (ref_layout.h) so that the
 * unmodified reference engines accept them (tests/test_ref_crosscheck.py).
 */
#include "hwlm_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_set>

namespace hsb {

namespace {

const u32 CL = 64;
inline size_t roundCL(size_t v) { return HSB_ROUNDUP(v, CL); }

u32 floorLog2(u32 v) {
    u32 r = 0;
    while (v >>= 1) {
        r++;
    }
    return r;
}

/* ------------------------------------------------------------ noodle -- */

std::vector<u8> buildNoodle(const HwlmLit &lit) {
    const std::string &s = lit.s;
    const size_t len = s.size();
    /* key = first position whose successor differs (caseless compare when
     * the literal is nocase and the char is a letter) */
    size_t key = 0;
    for (size_t i = 0; i + 1 < len; i++) {
        u8 c = s[i], d = s[i + 1];
        bool diff = (lit.nocase && isAsciiAlpha(c))
                        ? asciiUpper(c) != asciiUpper(d)
                        : c != d;
        key = i;
        if (diff) {
            break;
        }
    }
    NoodTable n;
    memset(&n, 0, sizeof(n));
    u8 msk[8] = {0}, cmp[8] = {0};
    for (size_t i = 0; i < len; i++) { /* first char in the low byte */
        u8 c = s[i];
        u8 m = (lit.nocase && isAsciiAlpha(c)) ? 0xdf : 0xff;
        msk[i] = m;
        cmp[i] = c & m;
    }
    memcpy(&n.msk, msk, 8);
    memcpy(&n.cmp, cmp, 8);
    n.id = lit.id;
    n.msk_len = (u8)len;
    n.single = len == 1;
    n.key_offset = (u8)(len - key);
    n.nocase = lit.nocase;
    n.key0 = s[key];
    n.key1 = n.single ? 0 : s[key + 1];
    std::vector<u8> out(sizeof(n));
    memcpy(out.data(), &n, sizeof(n));
    return out;
}

/* ----------------------------------------------------- hash confirm -- */

typedef std::map<u32, std::vector<u32>> BucketMap; /* bucket -> literal idx */

std::vector<u8> buildConfirmOne(const std::vector<const HwlmLit *> &lits) {
    const size_t n = lits.size();
    std::vector<LitInfo> li(n);
    u64 andmsk = ~0ULL;
    for (size_t i = 0; i < n; i++) {
        const HwlmLit &l = *lits[i];
        LitInfo &x = li[i];
        memset(&x, 0, sizeof(x));
        x.id = l.id;
        x.flags = l.noruns ? FDR_LIT_FLAG_NOREPEAT : 0;
        x.size = (u8)l.s.size();
        x.groups = l.groups;
        u64 msk = ~0ULL, val = 0;
        for (u32 j = 0; j < 8; j++) { /* j-th char from the end -> lane 7-j */
            u32 sh = (7 - j) * 8;
            if (j >= l.s.size()) {
                msk &= ~(0xffULL << sh);
            } else {
                u8 c = l.s[l.s.size() - 1 - j];
                if (l.nocase && isAsciiAlpha(c)) {
                    msk &= ~(0x20ULL << sh);
                    val |= (u64)(c & 0xdf) << sh;
                } else {
                    val |= (u64)c << sh;
                }
            }
        }
        x.v = val;
        x.msk = msk;
        andmsk &= msk;
    }
    const u32 nBits = floorLog2((u32)n) + 4;
    std::map<u32, std::vector<u32>> chains;
    u64 gm = 0;
    for (size_t i = 0; i < n; i++) {
        u32 h = (u32)(((li[i].v & andmsk) * CONF_HASH_MULT) >> (64 - nBits));
        chains[h].push_back((u32)i);
        gm |= li[i].groups;
    }
    const size_t idxBytes = ((size_t)1 << nBits) * sizeof(u32);
    size_t litOff = HSB_ROUNDUP(sizeof(FDRConfirm) + idxBytes, 8);
    size_t total = HSB_ROUNDUP(litOff + n * sizeof(LitInfo), 8);
    std::vector<u8> out(total, 0);
    FDRConfirm hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.andmsk = andmsk;
    hdr.mult = CONF_HASH_MULT;
    hdr.nBits = nBits;
    hdr.groups = gm;
    memcpy(out.data(), &hdr, sizeof(hdr));
    u32 *idx = (u32 *)(out.data() + sizeof(FDRConfirm));
    size_t pos = litOff;
    for (const auto &c : chains) {
        idx[c.first] = (u32)pos;
        for (size_t k = 0; k < c.second.size(); k++) {
            LitInfo x = li[c.second[k]];
            x.next = (k + 1 == c.second.size()) ? 0 : 1;
            memcpy(out.data() + pos, &x, sizeof(x));
            pos += sizeof(LitInfo);
        }
    }
    return out;
}

std::vector<u8> buildConfirm(const std::vector<HwlmLit> &lits,
                             const BucketMap &b2l, u32 nBuckets) {
    std::map<u32, std::vector<u8>> parts;
    size_t body = 0;
    for (u32 b = 0; b < nBuckets; b++) {
        auto it = b2l.find(b);
        if (it == b2l.end() || it->second.empty()) {
            continue;
        }
        std::vector<const HwlmLit *> vl;
        for (u32 i : it->second) {
            vl.push_back(&lits[i]);
        }
        parts[b] = buildConfirmOne(vl);
        body += parts[b].size();
    }
    const size_t sw = roundCL(nBuckets * sizeof(u32));
    std::vector<u8> out(sw + body, 0);
    u32 *confBase = (u32 *)out.data();
    size_t pos = sw;
    for (auto &p : parts) {
        confBase[p.first] = (u32)pos;
        memcpy(out.data() + pos, p.second.data(), p.second.size());
        pos += p.second.size();
    }
    return out;
}

/* ---------------------------------------------------- flood control -- */

struct FloodCmp {
    bool operator()(const FDRFlood &a, const FDRFlood &b) const {
        return memcmp(&a, &b, sizeof(a)) < 0;
    }
};

std::vector<u8> buildFlood(const std::vector<HwlmLit> &lits, u32 defaultSuffix,
                           bool allowFlood) {
    std::vector<FDRFlood> fl(256);
    memset(fl.data(), 0, 256 * sizeof(FDRFlood));
    for (auto &f : fl) {
        f.suffix = defaultSuffix;
    }
    auto bump = [&](u8 c, u32 suffix) {
        fl[c].suffix = std::max(fl[c].suffix, suffix + 1);
    };
    auto add = [&](u8 c, const HwlmLit &l, u32 suffix) {
        FDRFlood &f = fl[c];
        f.suffix = std::max(f.suffix, suffix + 1);
        if (f.idCount < FDR_FLOOD_MAX_IDS) {
            f.ids[f.idCount] = l.id;
            f.allGroups |= l.groups;
            f.groups[f.idCount] = l.groups;
            f.idCount++;
        }
    };
    for (const HwlmLit &l : lits) {
        const u32 n = (u32)l.s.size();
        u8 c = l.s[n - 1];
        const bool nocase = isAsciiAlpha(c) ? l.nocase : false;
        /* length of the run of `c` (case-folded iff the literal is nocase)
         * at the tail of the literal; n if the literal is one long run */
        u32 run = n;
        for (u32 i = 0; i < n; i++) {
            u8 d = l.s[n - 1 - i];
            bool differs = l.nocase ? asciiLower(c) != asciiLower(d) : c != d;
            if (differs) {
                run = i;
                break;
            }
        }
        u8 up = nocase ? asciiUpper(c) : c;
        if (run != n) {
            bump(up, run);
            if (nocase) {
                bump(asciiLower(c), run);
            }
        } else {
            add(up, l, run);
            if (nocase) {
                add(asciiLower(c), l, run);
            }
        }
    }
    if (!allowFlood) {
        for (auto &f : fl) {
            f.idCount = FDR_FLOOD_MAX_IDS;
        }
    }
    std::map<FDRFlood, std::vector<u32>, FloodCmp> distinct;
    for (u32 c = 0; c < 256; c++) {
        distinct[fl[c]].push_back(c);
    }
    const size_t hdr = 256 * sizeof(u32);
    std::vector<u8> out(HSB_ROUNDUP(hdr + distinct.size() * sizeof(FDRFlood), 16), 0);
    u32 *index = (u32 *)out.data();
    u32 k = 0;
    for (const auto &d : distinct) {
        memcpy(out.data() + hdr + k * sizeof(FDRFlood), &d.first, sizeof(FDRFlood));
        for (u32 c : d.second) {
            index[c] = k;
        }
        k++;
    }
    return out;
}

/* -------------------------------------------------------------- FDR -- */

struct FdrParams {
    u32 domain;
    u32 stride;
};

u32 absdiff(u32 a, u32 b) { return a > b ? a - b : b - a; }

FdrParams chooseFdr(const std::vector<HwlmLit> &lits, const HwlmBuildOpts &o) {
    size_t msl = std::numeric_limits<size_t>::max(), mslCount = 0;
    for (const auto &l : lits) {
        if (l.s.size() < msl) {
            msl = l.s.size();
            mslCount = 1;
        } else if (l.s.size() == msl) {
            mslCount++;
        }
    }
    const size_t n = lits.size();
    u32 want = 1;
    if (msl > 1) {
        if (n < 250) {
            want = (u32)msl;
        } else if (n < 800) {
            want = (u32)msl - 1;
        } else if (n < 5000) {
            want = std::min<u32>((u32)msl - 1, 2);
        }
    }
    if (msl == 4 && want == 4 && mslCount > 2) {
        want = 2;
    }
    FdrParams best = {0, 0};
    u32 bestScore = 0;
    bool have = false;
    for (u32 domain = 9; domain <= 15; domain++) {
        for (u32 stride = 1; stride <= 4; stride *= 2) {
            if (domain > 13 && stride > 1) {
                continue;
            }
            if (msl < stride || (int)domain > o.maxDomain) {
                continue;
            }
            u32 score = 100 - absdiff(want, stride);
            if (stride <= want) {
                score += stride;
            }
            u32 ideal;
            if (n < 8) {
                ideal = stride == 1 ? 8 : 10;
            } else if (n < 20) {
                ideal = 10;
            } else if (n < 100) {
                ideal = 11;
            } else if (n < 1000) {
                ideal = 12;
            } else if (n < 10000) {
                ideal = 13;
            } else {
                ideal = 15;
            }
            if (stride > 1) {
                ideal++;
            }
            score -= absdiff(ideal, domain);
            if (!have || score > bestScore) {
                best = {domain, stride};
                bestScore = score;
                have = true;
            }
        }
    }
    if (!have) {
        throw std::runtime_error("no FDR engine fits this literal set");
    }
    if (o.forceDomain) {
        best.domain = (u32)o.forceDomain;
    }
    if (o.forceStride) {
        best.stride = (u32)o.forceStride;
    }
    return best;
}

double fdrScore(u32 len, u32 count) {
    if (len == 0) {
        return std::numeric_limits<double>::max();
    }
    return std::pow((double)count, 1.05) * std::pow((double)len, -3.0);
}

struct Chunk {
    u32 first, count, length;
};

/* Sort literals by (length, reversed text, nocase first), cut the sorted list
 * into at most 512 chunks, then split the chunk sequence into <= 8 contiguous
 * buckets minimising sum(score(len of first chunk, #lits)) by dynamic
 * programming; shortest literals end up in the highest bucket. */
BucketMap assignFdrBuckets(std::vector<HwlmLit> &lits, u32 nBuckets) {
    std::stable_sort(lits.begin(), lits.end(),
                     [](const HwlmLit &a, const HwlmLit &b) {
                         if (a.s.size() != b.s.size()) {
                             return a.s.size() < b.s.size();
                         }
                         for (size_t i = a.s.size(); i-- > 0;) {
                             if (a.s[i] != b.s[i]) {
                                 return (char)a.s[i] < (char)b.s[i];
                             }
                         }
                         return a.nocase > b.nocase;
                     });
    std::set<size_t> lens;
    for (const auto &l : lits) {
        lens.insert(l.s.size());
    }
    const u32 CHUNK_MAX = 512, MAX_LEN = 16;
    const u32 perChunk =
        (u32)(lits.size() / (CHUNK_MAX - std::min<size_t>(MAX_LEN, lens.size())) + 1);
    std::vector<Chunk> chunks;
    u32 curLen = 0, start = 0;
    const HwlmLit *lastNocase = nullptr;
    auto sameText = [](const HwlmLit &a, const HwlmLit &b, bool fold) {
        if (a.s.size() != b.s.size()) {
            return false;
        }
        for (size_t i = 0; i < a.s.size(); i++) {
            u8 x = a.s[i], y = b.s[i];
            if (fold ? asciiUpper(x) != asciiUpper(y) : x != y) {
                return false;
            }
        }
        return true;
    };
    for (u32 i = 0; i < lits.size() && chunks.size() < CHUNK_MAX - 1; i++) {
        const HwlmLit &l = lits[i];
        bool equiv = false;
        if (i != 0) {
            bool fold = lastNocase && sameText(l, *lastNocase, true);
            equiv = sameText(l, lits[i - 1], fold);
        }
        if (!equiv &&
            ((curLen < MAX_LEN && l.s.size() != curLen) ||
             (curLen != 1 && (i - start) >= perChunk))) {
            curLen = (u32)l.s.size();
            if (!chunks.empty()) {
                chunks.back().count = i - start;
            }
            start = i;
            chunks.push_back({i, 0, curLen});
        }
        if (l.nocase) {
            lastNocase = &l;
        }
    }
    chunks.back().count = (u32)lits.size() - start;
    chunks.push_back({(u32)lits.size(), 0, 0}); /* sentinel */

    const u32 nc = (u32)chunks.size();
    const double INF = std::numeric_limits<double>::max();
    /* t[j][i]: best cost of covering chunks j.. with i+1 buckets; link = first
     * chunk of the next bucket (0 = "rest in one bucket") */
    std::vector<std::vector<std::pair<double, u32>>> t(
        nc, std::vector<std::pair<double, u32>>(nBuckets, {0.0, 0}));
    for (u32 j = 0; j < nc; j++) {
        u32 cnt = 0;
        for (u32 k = j; k < nc; k++) {
            cnt += chunks[k].count;
        }
        t[j][0] = {fdrScore(chunks[j].length, cnt), 0};
    }
    for (u32 i = 1; i < nBuckets; i++) {
        for (u32 j = 0; j + 1 < nc; j++) {
            std::pair<double, u32> best = {INF, 0};
            u32 cnt = chunks[j].count;
            for (u32 k = j + 1; k + 1 < nc; k++) {
                double sc = fdrScore(chunks[j].length, cnt);
                if (sc > best.first) {
                    break;
                }
                sc += t[k][i - 1].first;
                if (sc < best.first) {
                    best = {sc, k};
                }
                cnt += chunks[k].count;
            }
            t[j][i] = best;
        }
        t[nc - 1][i] = {0.0, 0};
    }
    std::vector<std::vector<u32>> groups;
    for (u32 i = 0, n = nBuckets; n && i != nc - 1; n--) {
        u32 j = t[i][n - 1].second;
        if (j == 0) {
            j = nc - 1;
        }
        std::vector<u32> ids;
        for (u32 k = chunks[j].first; k-- > chunks[i].first;) {
            ids.push_back(k); /* longest first inside a bucket */
        }
        groups.push_back(ids);
        i = j;
    }
    BucketMap out;
    for (size_t i = 0; i < groups.size(); i++) {
        out[(u32)(groups.size() - 1 - i)] = groups[i];
    }
    return out;
}

std::vector<u8> buildFdr(std::vector<HwlmLit> lits, const HwlmBuildOpts &o,
                         HwlmBuildInfo *info) {
    const u32 nBuckets = 8, width = 8; /* 64-bit scheme: 8 buckets x 8 suffix positions */
    const FdrParams p = chooseFdr(lits, o);
    if (p.domain < 9 || p.domain > 15 || (p.stride != 1 && p.stride != 2 && p.stride != 4)) {
        throw std::runtime_error("bad FDR parameters");
    }
    BucketMap b2l = assignFdrBuckets(lits, nBuckets);
    const u32 entries = 1u << p.domain;
    const u32 dmask = entries - 1;
    std::vector<u64> tab(entries, ~0ULL);
    u64 defaultMask = ~0ULL;
    for (const auto &kv : b2l) {
        const u32 b = kv.first;
        for (u32 pos = 0; pos < width; pos++) {
            const u64 bit = 1ULL << (pos * nBuckets + b);
            /* (dontcare mask) -> set of required values */
            std::map<u32, std::unordered_set<u32>> want;
            bool everything = false;
            for (u32 li : kv.second) {
                const HwlmLit &l = lits[li];
                const u32 sz = (u32)l.s.size();
                u32 mask = 0, dc = 0;
                for (u32 cnt = 0; cnt < 2; cnt++) {
                    int np = (int)pos - (int)cnt;
                    u8 dcb = 0, mb = 0;
                    if (np < 0 || (u32)np >= sz) {
                        dcb = 0xff;
                    } else {
                        u8 c = l.s[sz - np - 1];
                        mb = c;
                        u32 rem = p.domain - cnt * 8;
                        if (rem < 8) {
                            u8 cm = (u8)((1u << rem) - 1);
                            mb &= cm;
                            dcb |= (u8)~cm;
                        }
                        if (l.nocase && isAsciiAlpha(c)) {
                            mb &= 0xdf;
                            dcb |= 0x20;
                        }
                    }
                    mask |= (u32)mb << (cnt * 8);
                    dc |= (u32)dcb << (cnt * 8);
                }
                mask &= dmask;
                dc &= dmask;
                if (dc == dmask) {
                    everything = true;
                    break;
                }
                want[dc].insert(mask & ~dc);
            }
            if (everything) {
                defaultMask &= ~bit;
                continue;
            }
            for (const auto &w : want) {
                const u32 dc = w.first;
                /* enumerate all subsets of the don't-care bits */
                u32 sub = 0;
                do {
                    for (u32 v : w.second) {
                        tab[v | sub] &= ~bit;
                    }
                    sub = (sub - dc) & dc;
                } while (sub != 0);
            }
        }
    }
    for (auto &e : tab) {
        e &= defaultMask;
    }

    std::vector<u8> conf = buildConfirm(lits, b2l, nBuckets);
    std::vector<u8> flood = buildFlood(lits, (64 + nBuckets - 1) / nBuckets + 1, o.allowFlood);
    const size_t tabBytes = (size_t)entries * 8;
    const size_t size = roundCL(sizeof(FDR)) + roundCL(tabBytes) + roundCL(conf.size()) + flood.size();
    std::vector<u8> out(size, 0);
    FDR h;
    memset(&h, 0, sizeof(h));
    h.engineID = 0;
    h.size = (u32)size;
    u32 maxLen = 0;
    for (const auto &l : lits) {
        maxLen = std::max<u32>(maxLen, (u32)l.s.size());
    }
    h.maxStringLen = maxLen;
    h.numStrings = (u32)lits.size();
    h.domain = (u8)p.domain;
    h.domainMask = (u16)dmask;
    h.tabSize = (u32)tabBytes;
    h.stride = (u8)p.stride;
    /* initial state: bucket b cannot match before its shortest literal fits */
    for (u32 b = 0; b < nBuckets; b++) {
        auto it = b2l.find(b);
        u32 minLen = ~0u;
        if (it != b2l.end()) {
            for (u32 li : it->second) {
                minLen = std::min<u32>(minLen, (u32)lits[li].s.size());
            }
        }
        for (u32 i = 0; i < width; i++) {
            if (i < minLen - 1) {
                u32 sb = i * nBuckets + b;
                h.start[sb / 8] |= (u8)(1u << (sb % 8));
            }
        }
    }
    size_t pos = roundCL(sizeof(FDR));
    memcpy(out.data() + pos, tab.data(), tabBytes);
    pos += roundCL(tabBytes);
    h.confOffset = (u32)pos;
    memcpy(out.data() + pos, conf.data(), conf.size());
    pos += roundCL(conf.size());
    h.floodOffset = (u32)pos;
    memcpy(out.data() + pos, flood.data(), flood.size());
    memcpy(out.data(), &h, sizeof(h));
    if (info) {
        info->engineID = 0;
        info->domain = p.domain;
        info->stride = p.stride;
        info->numBuckets = nBuckets;
    }
    return out;
}

/* ------------------------------------------------------------ Teddy -- */

const u32 TEDDY_BUCKET_LOAD = 6; /* src/fdr/teddy_engine_description.h:40 */

struct TeddyDef {
    u32 id, numMasks, numBuckets;
    bool packed;
};

const TeddyDef kTeddyDefs[] = {
    {3, 1, 16, false},  {4, 1, 16, true},  {5, 2, 16, false},  {6, 2, 16, true},
    {7, 3, 16, false},  {8, 3, 16, true},  {9, 4, 16, false},  {10, 4, 16, true},
    {11, 1, 8, false},  {12, 1, 8, true},  {13, 2, 8, false},  {14, 2, 8, true},
    {15, 3, 8, false},  {16, 3, 8, true},  {17, 4, 8, false},  {18, 4, 8, true},
};

bool teddyAllowed(const std::vector<HwlmLit> &lits, const TeddyDef &e,
                  size_t maxLen, const HwlmBuildOpts &o) {
    if (e.numBuckets == 16 && !o.allowFatTeddy) {
        return false;
    }
    if (e.numBuckets < lits.size() && !e.packed) {
        return false;
    }
    if (e.numBuckets * TEDDY_BUCKET_LOAD < lits.size()) {
        return false;
    }
    if (e.numMasks > maxLen) {
        return false;
    }
    if (lits.size() > 40) {
        u32 small = 0;
        for (const auto &l : lits) {
            if (l.s.size() < e.numMasks) {
                small++;
            }
        }
        if (small * 5 > lits.size()) {
            return false;
        }
    }
    return true;
}

const TeddyDef *chooseTeddy(const std::vector<HwlmLit> &lits,
                            const HwlmBuildOpts &o) {
    if (o.forceEngine >= 3) {
        for (const auto &d : kTeddyDefs) {
            if ((int)d.id == o.forceEngine) {
                return &d;
            }
        }
        return nullptr;
    }
    size_t maxLen = 0, maxTail = 0;
    for (const auto &l : lits) {
        maxLen = std::max(maxLen, l.s.size());
        size_t j = 1;
        for (; j < l.s.size(); j++) {
            if (l.s[l.s.size() - j - 1] != l.s[l.s.size() - 1]) {
                break;
            }
        }
        maxTail = std::max(maxTail, j);
    }
    const TeddyDef *best = nullptr;
    u32 bestScore = 0;
    for (const auto &e : kTeddyDefs) {
        if (!teddyAllowed(lits, e, maxLen, o)) {
            continue;
        }
        u32 score = 0;
        if (!e.packed) {
            score += 100;
        }
        if (lits.size() > 4 * e.numBuckets) {
            score += e.numMasks * 4;
        } else {
            score += 100;
        }
        if (e.numMasks > maxTail) {
            score += 50;
        }
        score += 6 / (std::abs(3 - (int)e.numMasks) + 1);
        score += 16 / e.numBuckets;
        if (!best || score > bestScore) {
            best = &e;
            bestScore = score;
        }
    }
    return best;
}

/* A candidate bucket while packing: per mask position the sets of low / high
 * nibbles it accepts (16-bit bitmaps) and the literal indices it holds. */
struct TeddySet {
    std::vector<u16> nib;
    std::vector<u32> ids;
    bool operator<(const TeddySet &o) const { return ids < o.ids; }
    u64 probability() const {
        u64 v = 1;
        for (u16 x : nib) {
            v *= (u64)__builtin_popcount(x);
        }
        return v;
    }
    u64 heuristic() const { return probability() * (2 + ids.size()); }
    bool runProne() const {
        u16 lo = 0xffff, hi = 0xffff;
        for (size_t i = 0; i < nib.size(); i += 2) {
            lo &= nib[i];
            hi &= nib[i + 1];
        }
        return lo && hi;
    }
};

TeddySet mergeSets(const TeddySet &a, const TeddySet &b) {
    TeddySet m = a;
    for (size_t i = 0; i < m.nib.size(); i++) {
        m.nib[i] |= b.nib[i];
    }
    m.ids.insert(m.ids.end(), b.ids.begin(), b.ids.end());
    std::sort(m.ids.begin(), m.ids.end());
    m.ids.erase(std::unique(m.ids.begin(), m.ids.end()), m.ids.end());
    return m;
}

bool packTeddy(const std::vector<HwlmLit> &lits, const TeddyDef &e, BucketMap &b2l) {
    if (lits.size() > e.numBuckets * TEDDY_BUCKET_LOAD) {
        return false;
    }
    std::set<TeddySet> sets;
    for (u32 i = 0; i < lits.size(); i++) {
        TeddySet ts;
        ts.nib.assign(e.numMasks * 2, 0);
        const std::string &s = lits[i].s;
        for (u32 m = 0; m < e.numMasks; m++) {
            if (m < s.size()) {
                u8 c = s[s.size() - 1 - m];
                u8 hi = c >> 4, lo = c & 0xf;
                ts.nib[m * 2] = (u16)(1u << lo);
                if (lits[i].nocase && isAsciiAlpha(c)) {
                    ts.nib[m * 2 + 1] = (u16)((1u << (hi & 0xd)) | (1u << (hi | 0x2)));
                } else {
                    ts.nib[m * 2 + 1] = (u16)(1u << hi);
                }
            } else {
                ts.nib[m * 2] = ts.nib[m * 2 + 1] = 0xffff;
            }
        }
        ts.ids.push_back(i);
        sets.insert(ts);
    }
    for (;;) {
        auto m1 = sets.end(), m2 = sets.end();
        u64 best = ~0ULL;
        for (auto i1 = sets.begin(); i1 != sets.end(); ++i1) {
            for (auto i2 = std::next(i1); i2 != sets.end(); ++i2) {
                if (sets.size() <= e.numBuckets && i1->nib != i2->nib) {
                    continue;
                }
                TeddySet t = mergeSets(*i1, *i2);
                u64 ns = t.heuristic(), os = i1->heuristic() + i2->heuristic();
                if (ns < os) { /* strictly better merged: take it, next i1 */
                    m1 = i1;
                    m2 = i2;
                    break;
                }
                u64 sc = ns - os;
                bool oldRun = i1->runProne() && i2->runProne();
                if (t.runProne() && !oldRun) {
                    continue;
                }
                if (sc < best) {
                    best = sc;
                    m1 = i1;
                    m2 = i2;
                }
            }
        }
        if (m1 == sets.end() || m2 == sets.end()) {
            break;
        }
        TeddySet t = mergeSets(*m1, *m2);
        sets.erase(m1);
        sets.erase(m2);
        sets.insert(t);
    }
    if (sets.size() > e.numBuckets) {
        return false;
    }
    u32 b = 0;
    for (const auto &s : sets) {
        b2l[b++] = s.ids;
    }
    return true;
}

/* Nibble masks: for mask m (distance from the last byte) and bucket b, byte
 * [lo table][nibble] has bit (b%8) CLEARED iff some literal in b accepts that
 * nibble there.  `dup` writes the Fat Teddy duplicate (32-byte rows). */
void fillNibbleMasks(const std::vector<HwlmLit> &lits, const BucketMap &b2l,
                     u32 numMasks, u32 maskWidth, bool dup, u8 *base, size_t len) {
    memset(base, 0xff, len);
    const u32 row = dup ? 32 : 16;
    const u32 mw = dup ? 2 : maskWidth;
    auto clr = [&](u32 mskId, u32 nibble, u8 bm) {
        base[mskId * row + nibble] &= (u8)~bm;
        if (dup) {
            base[mskId * row + 16 + nibble] &= (u8)~bm;
        }
    };
    for (const auto &kv : b2l) {
        const u32 b = kv.first;
        const u8 bm = (u8)(1u << (b % 8));
        for (u32 li : kv.second) {
            const HwlmLit &l = lits[li];
            const u32 sz = (u32)l.s.size();
            for (u32 j = 0; j < numMasks; j++) {
                const u32 lo = j * 2 * mw + b / 8, hi = (j * 2 + 1) * mw + b / 8;
                if (j >= sz) {
                    for (u32 n = 0; n < 16; n++) {
                        clr(lo, n, bm);
                        clr(hi, n, bm);
                    }
                    continue;
                }
                u8 c = l.s[sz - 1 - j];
                u32 nh = c >> 4, nl = c & 0xf;
                if (l.nocase && isAsciiAlpha(c)) {
                    clr(hi, nh & 0xd, bm);
                    clr(hi, nh | 0x2, bm);
                } else {
                    clr(hi, nh, bm);
                }
                clr(lo, nl, bm);
            }
        }
    }
}

const u32 RMSK_LEN = 8;
const u32 RTABLE_SIZE = (256 + 1) * RMSK_LEN;

/* "Reinforcement" table of the AVX2/AVX512 CPU variants: entry [c][j-1] has
 * the bucket bit cleared iff a literal of that bucket has byte c at distance
 * j (1..7) from its end (or is shorter). Entry 256 is all-zero. */
void fillReinforced(const std::vector<HwlmLit> &lits, const BucketMap &b2l, u8 *rt) {
    for (u32 c = 0; c < 256; c++) {
        u64 v = 0x00ffffffffffffffULL;
        memcpy(rt + c * RMSK_LEN, &v, 8);
    }
    auto clr = [&](u32 c, u32 j, u8 bm) { rt[c * RMSK_LEN + j - 1] &= (u8)~bm; };
    for (const auto &kv : b2l) {
        const u8 bm = (u8)(1u << (kv.first % 8));
        for (u32 li : kv.second) {
            const HwlmLit &l = lits[li];
            const u32 sz = (u32)l.s.size();
            for (u32 j = 1; j < RMSK_LEN; j++) {
                if (sz - 1 < j) {
                    for (u32 c = 0; c < 256; c++) {
                        clr(c, j, bm);
                    }
                } else {
                    u8 c = l.s[sz - 1 - j];
                    if (l.nocase && isAsciiAlpha(c)) {
                        clr(c & 0xdf, j, bm);
                        clr(c | 0x20, j, bm);
                    } else {
                        clr(c, j, bm);
                    }
                }
            }
        }
    }
    memset(rt + 256 * RMSK_LEN, 0, RMSK_LEN);
}

std::vector<u8> buildTeddy(const std::vector<HwlmLit> &lits, const TeddyDef &e,
                           const BucketMap &b2l, const HwlmBuildOpts &o,
                           HwlmBuildInfo *info) {
    const u32 maskWidth = e.numBuckets / 8;
    const size_t maskLen = (size_t)e.numMasks * 16 * 2 * maskWidth;
    const size_t extraLen = maskWidth == 2 ? maskLen * 2 : RTABLE_SIZE * maskWidth;
    std::vector<u8> conf = buildConfirm(lits, b2l, e.numBuckets);
    std::vector<u8> flood = buildFlood(lits, e.numMasks, o.allowFlood);
    const size_t size = roundCL(sizeof(Teddy)) + roundCL(maskLen) + roundCL(extraLen) +
                        roundCL(conf.size()) + flood.size();
    std::vector<u8> out(size, 0);
    Teddy h;
    memset(&h, 0, sizeof(h));
    h.engineID = e.id;
    h.size = (u32)size;
    for (const auto &l : lits) {
        h.maxStringLen = std::max<u32>(h.maxStringLen, (u32)l.s.size());
    }
    h.numStrings = (u32)lits.size();
    size_t pos = roundCL(sizeof(Teddy)) + roundCL(maskLen) + roundCL(extraLen);
    h.confOffset = (u32)pos;
    memcpy(out.data() + pos, conf.data(), conf.size());
    pos += roundCL(conf.size());
    h.floodOffset = (u32)pos;
    memcpy(out.data() + pos, flood.data(), flood.size());
    memcpy(out.data(), &h, sizeof(h));
    u8 *base = out.data() + roundCL(sizeof(Teddy));
    fillNibbleMasks(lits, b2l, e.numMasks, maskWidth, false, base, maskLen);
    u8 *extra = base + roundCL(maskLen);
    if (maskWidth == 1) {
        fillReinforced(lits, b2l, extra);
    } else {
        fillNibbleMasks(lits, b2l, e.numMasks, maskWidth, true, extra, extraLen);
    }
    if (info) {
        info->engineID = e.id;
        info->numBuckets = e.numBuckets;
        info->numMasks = e.numMasks;
    }
    return out;
}

} // namespace

/* ------------------------------------------------------------- HWLM -- */

std::vector<u8> buildHwlm(std::vector<HwlmLit> lits, const HwlmBuildOpts &opts,
                          HwlmBuildInfo *info) {
    if (lits.empty()) {
        throw std::runtime_error("no literals");
    }
    for (const auto &l : lits) {
        if (l.s.empty() || l.s.size() > 8) {
            throw std::runtime_error("HWLM literal length must be 1..8");
        }
        if (l.id == 0xffffffffu) {
            throw std::runtime_error("reserved literal id");
        }
        if (!l.groups) {
            throw std::runtime_error("literal without groups");
        }
    }
    std::vector<u8> eng;
    u32 type;
    HwlmBuildInfo local;
    if (lits.size() == 1 && opts.allowNoodle && opts.forceEngine < 0) {
        type = HWLM_ENGINE_NOOD;
        eng = buildNoodle(lits[0]);
    } else {
        type = HWLM_ENGINE_FDR;
        bool done = false;
        if (opts.allowTeddy && opts.forceEngine != 0) {
            const TeddyDef *e = chooseTeddy(lits, opts);
            BucketMap b2l;
            if (e && packTeddy(lits, *e, b2l)) {
                eng = buildTeddy(lits, *e, b2l, opts, &local);
                done = true;
            } else if (opts.forceEngine >= 3) {
                throw std::runtime_error("forced Teddy engine cannot hold this literal set");
            }
        }
        if (!done) {
            eng = buildFdr(lits, opts, &local);
        }
    }
    local.type = type;
    std::vector<u8> out(HWLM_ENGINE_OFFSET + eng.size(), 0);
    HWLM h;
    memset(&h, 0, sizeof(h));
    h.type = (u8)type; /* accel0/accel1 stay ACCEL_NONE */
    memcpy(out.data(), &h, sizeof(h));
    memcpy(out.data() + HWLM_ENGINE_OFFSET, eng.data(), eng.size());
    if (info) {
        *info = local;
    }
    return out;
}

/* CRC32C (Castagnoli, reflected 0x82F63B78), init as given, no final xor:
 * the convention of Crc32c_ComputeBuf(0, ...) (reference: src/crc32.c:516-522,
 * src/database.c:178-186). */
u32 crc32c(u32 crc, const void *buf, size_t len) {
    static u32 table[256];
    static bool init = false;
    if (!init) {
        for (u32 i = 0; i < 256; i++) {
            u32 c = i;
            for (int k = 0; k < 8; k++) {
                c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            }
            table[i] = c;
        }
        init = true;
    }
    const u8 *p = (const u8 *)buf;
    while (len--) {
        crc = table[(crc ^ *p++) & 0xff] ^ (crc >> 8);
    }
    return crc;
}

} // namespace hsb
