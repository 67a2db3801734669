/* db_walk.cpp -- see db_walk.h */
#include "db_walk.h"

#include <algorithm>
#include <cstring>

#include "api_internal.h"

namespace hsb {

void collectProgramReports(const u8 *bc, u32 bcLen, u32 prog, std::unordered_set<u32> *ex) {
    u32 pc = prog;
    for (int guard = 0; guard < 4096 && pc < bcLen; guard++) {
        const u8 code = bc[pc];
        switch (code) {
        case OP_END:
        case OP_FINAL_REPORT:
            return;
        case OP_CHECK_GROUPS: pc += HSB_ROUNDUP(sizeof(InstrCheckGroups), 8); break;
        case OP_CHECK_MASK: pc += HSB_ROUNDUP(sizeof(InstrCheckMask), 8); break;
        case OP_CHECK_BYTE: pc += HSB_ROUNDUP(sizeof(InstrCheckByte), 8); break;
        case OP_CHECK_MED_LIT:
        case OP_CHECK_MED_LIT_NOCASE:
        case OP_CHECK_LONG_LIT:
        case OP_CHECK_LONG_LIT_NOCASE: pc += HSB_ROUNDUP(sizeof(InstrCheckLit), 8); break;
        case OP_CHECK_EXHAUSTED: pc += HSB_ROUNDUP(sizeof(InstrCheckExhausted), 8); break;
        case OP_DEDUPE: pc += HSB_ROUNDUP(sizeof(InstrDedupe), 8); break;
        case OP_REPORT: pc += HSB_ROUNDUP(sizeof(InstrReport), 8); break;
        case OP_REPORT_EXHAUST: {
            InstrReportExhaust in;
            memcpy(&in, bc + pc, sizeof(in));
            ex->insert(in.onmatch);
            pc += HSB_ROUNDUP(sizeof(InstrReportExhaust), 8);
            break;
        }
        case OP_DEDUPE_AND_REPORT: pc += HSB_ROUNDUP(sizeof(InstrDedupeAndReport), 8); break;
        case OP_SQUASH_GROUPS: pc += HSB_ROUNDUP(sizeof(InstrSquashGroups), 8); break;
        case OP_CLEAR_WORK_DONE: pc += 8; break;
        case OP_INCLUDED_JUMP: pc += HSB_ROUNDUP(sizeof(InstrIncludedJump), 8); break;
        case OP_SET_EXHAUST: pc += HSB_ROUNDUP(sizeof(InstrSetExhaust), 8); break;
        default:
            return;
        }
    }
}

/* Walk the hash-confirm structures to enumerate literal programs
 * (src/fdr/fdr_confirm.h:36-94). */
void walkConfirm(const u8 *bc, u32 bcLen, u32 confOff, u32 nBuckets,
                 std::unordered_set<u32> *ex, std::vector<LitTail> *tails) {
    const u8 *confBase = bc + confOff;
    for (u32 b = 0; b < nBuckets; b++) {
        u32 cf;
        memcpy(&cf, confBase + 4 * b, 4);
        if (!cf) {
            continue;
        }
        const u8 *fc = confBase + cf;
        FDRConfirm h;
        memcpy(&h, fc, sizeof(h));
        const u32 n = 1u << h.nBits;
        for (u32 c = 0; c < n; c++) {
            u32 start;
            memcpy(&start, fc + sizeof(FDRConfirm) + 4 * c, 4);
            if (!start) {
                continue;
            }
            const u8 *li = fc + start;
            for (;;) {
                LitInfo x;
                memcpy(&x, li, sizeof(x));
                collectProgramReports(bc, bcLen, x.id, ex);
                tails->push_back({x.v, x.msk, x.size, b});
                if (!x.next) {
                    break;
                }
                li += sizeof(LitInfo);
            }
        }
    }
}

hs_error_t collectExhaustible(const hs_database_t *db, std::unordered_set<u32> *ex) {
    const DbHeader *h = (const DbHeader *)db;
    if (!h || h->magic != DB_MAGIC) {
        return HS_INVALID;
    }
    const RoseEngine *r = dbRose(db);
    const u8 *bc = (const u8 *)r;
    if (r->runtimeImpl != RUNTIME_PURE_LITERAL || !r->fmatcherOffset) {
        return HS_ARCH_ERROR;
    }
    const HWLM *hw = (const HWLM *)(bc + r->fmatcherOffset);
    const u32 engOff = r->fmatcherOffset + HWLM_ENGINE_OFFSET;
    if (hw->type == HWLM_ENGINE_NOOD) {
        NoodTable n;
        memcpy(&n, bc + engOff, sizeof(n));
        collectProgramReports(bc, h->length, n.id, ex);
        return HS_SUCCESS;
    }
    FDR f;
    memcpy(&f, bc + engOff, sizeof(f));
    const u32 nb = f.engineID == 0 ? 8 : teddyNumBuckets(f.engineID);
    std::vector<LitTail> tails;
    walkConfirm(bc, h->length, engOff + f.confOffset, nb, ex, &tails);
    return HS_SUCCESS;
}

size_t postprocessRecords(const std::unordered_set<u32> &exhaustible, MatchRec *m, size_t n) {
    std::sort(m, m + n, [](const MatchRec &a, const MatchRec &b) {
        if (a.block != b.block) return a.block < b.block;
        if (a.to != b.to) return a.to < b.to;
        return a.id < b.id;
    });
    size_t w = 0;
    std::unordered_set<u32> seen;
    u32 curBlock = 0xffffffffu;
    const bool anyEx = !exhaustible.empty();
    for (size_t i = 0; i < n; i++) {
        if (w && m[w - 1].block == m[i].block && m[w - 1].to == m[i].to && m[w - 1].id == m[i].id) {
            continue;
        }
        if (anyEx) {
            if (m[i].block != curBlock) {
                curBlock = m[i].block;
                seen.clear();
            }
            if (exhaustible.count(m[i].id) && !seen.insert(m[i].id).second) {
                continue;
            }
        }
        m[w++] = m[i];
    }
    return w;
}

} // namespace hsb

extern "C" hs_error_t hs_b200_postprocess_matches(const hs_database_t *db, hs_b200_match_t *recs,
                                                  size_t n, unsigned long long *nout) {
    if (!db || (n && !recs) || !nout) {
        return HS_INVALID;
    }
    std::unordered_set<hsb::u32> ex;
    hs_error_t r = hsb::collectExhaustible(db, &ex);
    if (r != HS_SUCCESS) {
        return r;
    }
    static_assert(sizeof(hsb::MatchRec) == sizeof(hs_b200_match_t), "record layout");
    *nout = hsb::postprocessRecords(ex, (hsb::MatchRec *)recs, n);
    return HS_SUCCESS;
}
