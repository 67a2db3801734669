/* pair_table.cpp -- see pair_table.h */
#include "pair_table.h"

#include <algorithm>
#include <cstring>
#include <map>

namespace hsb {

namespace {

const u32 GREEDY_LIMIT = 72; /* above this many classes a cheap pre-merge runs first */

struct Side {                       /* one partition of the byte values (rows: first byte, columns: second) */
    std::vector<std::vector<u8>> members;
    std::vector<double> w;          /* prior probability of the class in the corpus model */
};

struct Model {
    Side side[2];
    std::vector<u32> T;             /* [row class][column class]: bit 8 * slot + bucket SET = possible */
    u32 at(u32 r, u32 c) const { return T[(size_t)r * side[1].w.size() + c]; }
};

/* expected candidates per position if the slots were independent: for every
 * bucket the product over the four slots of P(sample admits the bucket) */
double objective(const double (&pass)[32]) {
    double total = 0;
    for (u32 k = 0; k < 8; k++) {
        total += pass[k] * pass[8 + k] * pass[16 + k] * pass[24 + k];
    }
    return total;
}

void mergeClasses(Model &m, int axis, u32 i, u32 j) { /* j into i, i < j */
    const u32 R = (u32)m.side[0].w.size(), C = (u32)m.side[1].w.size();
    std::vector<u32> nt;
    if (axis == 0) {
        nt.reserve((size_t)(R - 1) * C);
        for (u32 r = 0; r < R; r++) {
            if (r == j) {
                continue;
            }
            for (u32 c = 0; c < C; c++) {
                nt.push_back(r == i ? (m.at(i, c) | m.at(j, c)) : m.at(r, c));
            }
        }
    } else {
        nt.reserve((size_t)R * (C - 1));
        for (u32 r = 0; r < R; r++) {
            for (u32 c = 0; c < C; c++) {
                if (c == j) {
                    continue;
                }
                nt.push_back(c == i ? (m.at(r, i) | m.at(r, j)) : m.at(r, c));
            }
        }
    }
    Side &s = m.side[axis];
    s.members[i].insert(s.members[i].end(), s.members[j].begin(), s.members[j].end());
    s.members.erase(s.members.begin() + j);
    s.w[i] += s.w[j];
    s.w.erase(s.w.begin() + j);
    m.T.swap(nt);
}

/* masses[x][b] = sum over the other side's classes y of w[y] * bit b of T(x, y) */
void sideMasses(const Model &m, int axis, std::vector<double> *mass, double (&pass)[32]) {
    const u32 X = (u32)m.side[axis].w.size(), Y = (u32)m.side[1 - axis].w.size();
    mass->assign((size_t)X * 32, 0.0);
    for (u32 b = 0; b < 32; b++) {
        pass[b] = 0;
    }
    for (u32 x = 0; x < X; x++) {
        double *mx = mass->data() + (size_t)x * 32;
        for (u32 y = 0; y < Y; y++) {
            u32 t = axis == 0 ? m.at(x, y) : m.at(y, x);
            const double wy = m.side[1 - axis].w[y];
            while (t) {
                const u32 b = (u32)__builtin_ctz(t);
                t &= t - 1;
                mx[b] += wy;
            }
        }
        for (u32 b = 0; b < 32; b++) {
            pass[b] += m.side[axis].w[x] * mx[b];
        }
    }
}

/* best pair of classes to merge on `axis`: smallest objective afterwards */
double bestMerge(const Model &m, int axis, u32 *bi, u32 *bj) {
    const u32 X = (u32)m.side[axis].w.size(), Y = (u32)m.side[1 - axis].w.size();
    std::vector<double> mass;
    double pass[32];
    sideMasses(m, axis, &mass, pass);
    double best = -1;
    const std::vector<double> &wx = m.side[axis].w, &wy = m.side[1 - axis].w;
    for (u32 i = 0; i < X; i++) {
        for (u32 j = i + 1; j < X; j++) {
            double p[32];
            /* merged row admits T(i, y) | T(j, y): i gains what only j had and vice versa */
            for (u32 b = 0; b < 32; b++) {
                p[b] = pass[b];
            }
            for (u32 y = 0; y < Y; y++) {
                const u32 ti = axis == 0 ? m.at(i, y) : m.at(y, i);
                const u32 tj = axis == 0 ? m.at(j, y) : m.at(y, j);
                u32 gi = tj & ~ti, gj = ti & ~tj;
                while (gi) {
                    const u32 b = (u32)__builtin_ctz(gi);
                    gi &= gi - 1;
                    p[b] += wx[i] * wy[y];
                }
                while (gj) {
                    const u32 b = (u32)__builtin_ctz(gj);
                    gj &= gj - 1;
                    p[b] += wx[j] * wy[y];
                }
            }
            const double v = objective(p);
            if (best < 0 || v < best) {
                best = v;
                *bi = i;
                *bj = j;
            }
        }
    }
    return best;
}

} // namespace

void buildPairTables(const std::vector<LitTail> &tails, u32 slotBase, PairTables *out, u32 maxClass0,
                     u32 maxClass1) {
    /* 5 bits per byte of the sample; fewer classes of the second byte shrink the
     * pair table (4 KiB per class) in favour of a larger prefilter bitmap */
    const u32 maxClasses[2] = {std::max(1u, std::min(32u, maxClass0)), std::max(1u, std::min(32u, maxClass1))};
    /* poss[b0][b1]: bit 8 * i + bucket SET iff some literal of the bucket can have
     * byte b0 at suffix distance i + slotBase and b1 right after it (don't-care
     * bits of LitInfo.msk -- caseless letters -- admit both cases; distance 0 has
     * no following byte).  Same construction as setupTab
     * (src/fdr/fdr_compile.cpp:527-632) over byte pairs instead of hash values. */
    std::vector<u32> poss(256 * 256, 0);
    u32 dead = 0;
    for (const LitTail &t : tails) {
        for (u32 i = 0; i < 4; i++) {
            const u32 p = i + slotBase;
            const u32 bit = 1u << (8 * i + (t.bucket & 7));
            if (p >= t.size || p > 7) {
                dead |= bit; /* shorter literal: the slot cannot constrain its bucket */
                continue;
            }
            const u8 c0 = (u8)(t.v >> (8 * (7 - p))), m0 = (u8)(t.msk >> (8 * (7 - p)));
            u8 c1 = 0, m1 = 0;
            if (p > 0) {
                c1 = (u8)(t.v >> (8 * (8 - p)));
                m1 = (u8)(t.msk >> (8 * (8 - p)));
            }
            for (u32 b0 = 0; b0 < 256; b0++) {
                if ((b0 & m0) != c0) {
                    continue;
                }
                for (u32 b1 = 0; b1 < 256; b1++) {
                    if ((b1 & m1) == c1) {
                        poss[b0 * 256 + b1] |= bit;
                    }
                }
            }
        }
    }
    for (u32 &e : poss) {
        e |= dead;
    }

    /* corpus model for weighing merges: printable ASCII four times as likely as the rest */
    double wb[256], wsum = 0;
    for (u32 b = 0; b < 256; b++) {
        wb[b] = (b >= 0x20 && b < 0x7f) ? 1.0 : 0.25;
        wsum += wb[b];
    }
    for (u32 b = 0; b < 256; b++) {
        wb[b] /= wsum;
    }

    /* start from the lossless partitions: bytes with identical rows / columns */
    Model m;
    u32 cls[2][256];
    for (int axis = 0; axis < 2; axis++) {
        std::map<std::vector<u32>, u32> seen;
        for (u32 b = 0; b < 256; b++) {
            std::vector<u32> key(256);
            for (u32 o = 0; o < 256; o++) {
                key[o] = axis == 0 ? poss[b * 256 + o] : poss[o * 256 + b];
            }
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(std::move(key), (u32)m.side[axis].members.size()).first;
                m.side[axis].members.emplace_back();
                m.side[axis].w.push_back(0.0);
            }
            m.side[axis].members[it->second].push_back((u8)b);
            m.side[axis].w[it->second] += wb[b];
        }
    }
    {
        const u32 R = (u32)m.side[0].w.size(), C = (u32)m.side[1].w.size();
        m.T.resize((size_t)R * C);
        for (u32 r = 0; r < R; r++) {
            for (u32 c = 0; c < C; c++) {
                m.T[(size_t)r * C + c] = poss[m.side[0].members[r][0] * 256 + m.side[1].members[c][0]];
            }
        }
    }
    /* binary literal sets can start with hundreds of classes: fold the lightest
     * ones together first (cheap), the greedy search below does the rest */
    for (int axis = 0; axis < 2; axis++) {
        while (m.side[axis].w.size() > GREEDY_LIMIT) {
            const std::vector<double> &w = m.side[axis].w;
            u32 a = 0, b = 1;
            if (w[b] < w[a]) {
                std::swap(a, b);
            }
            for (u32 x = 2; x < w.size(); x++) {
                if (w[x] < w[a]) {
                    b = a;
                    a = x;
                } else if (w[x] < w[b]) {
                    b = x;
                }
            }
            mergeClasses(m, axis, std::min(a, b), std::max(a, b));
        }
    }
    while (m.side[0].w.size() > maxClasses[0] || m.side[1].w.size() > maxClasses[1]) {
        int axis = -1;
        u32 bi = 0, bj = 0;
        double best = -1;
        for (int a = 0; a < 2; a++) {
            if (m.side[a].w.size() <= maxClasses[a]) {
                continue;
            }
            u32 i = 0, j = 0;
            const double v = bestMerge(m, a, &i, &j);
            if (best < 0 || v < best) {
                best = v;
                axis = a;
                bi = i;
                bj = j;
            }
        }
        mergeClasses(m, axis, bi, bj);
    }

    for (int axis = 0; axis < 2; axis++) {
        for (u32 c = 0; c < m.side[axis].members.size(); c++) {
            for (u8 b : m.side[axis].members[c]) {
                cls[axis][b] = c;
            }
        }
    }
    for (u32 b = 0; b < 256; b++) {
        out->classWord[b] = (cls[0][b] << 7) | (cls[1][b] << 12);
    }
    for (u32 i = 0; i < 1024; i++) {
        out->pair[i] = 0xffffffffu;
    }
    const u32 R = (u32)m.side[0].w.size(), C = (u32)m.side[1].w.size();
    for (u32 r = 0; r < R; r++) {
        for (u32 c = 0; c < C; c++) {
            out->pair[(c << 5) | r] = ~m.at(r, c);
        }
    }
    out->nClass0 = R;
    out->nClass1 = C;
    std::vector<double> mass;
    double pass[32];
    sideMasses(m, 0, &mass, pass);
    out->modelRate = objective(pass);
}

} // namespace hsb
