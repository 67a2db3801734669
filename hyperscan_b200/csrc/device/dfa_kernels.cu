/*
 * dfa_kernels.cu -- the reference's DFA engines in block mode on sm_100a:
 * McClellan with 8- and 16-bit states (incl. Sherman states) and Sheng, run
 * over every block of a corpus, straight from the engines' own bytes.
 *
 * Replaces nfaExecMcClellan8_B / nfaExecMcClellan16_B (src/nfa/mcclellan.c:
 * 937-973; inner loops doNormal8 / doNormal16 :122-167,370-444, reports
 * doComplexReport :43-91, Sherman states mcclellan_common_impl.h:61-93) and
 * nfaExecSheng_B (src/nfa/sheng.c:706-739; loop sheng_impl.h:38-100, reports
 * fireReports :116-155) -- the entry points hs_scan uses for the anchored literal
 * table (src/rose/block.c:42-91) and the small-write engine
 * (src/runtime.c:285-315).
 *
 * A DFA run is one dependent table lookup per byte, so the parallelism is across
 * blocks: one thread per block (the configurations scan 10^6..10^7 blocks of
 * ~1 KiB), 16 corpus bytes per load.  The transition table sits in shared memory
 * when it fits (McClellan: remap + successor table; Sheng: the 4 KiB of shuffle
 * masks as a byte table), else it is read through L1/L2.  Acceleration schemes
 * (ACCEL_FLAG states, src/nfa/accel.h) are skip-ahead optimisations only and are
 * ignored; wide states (has_wide) are not handled -- the C ABI refuses them.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"

namespace hsb {

namespace {

__device__ __forceinline__ u32 g32(const u8 *p) { return __ldg(reinterpret_cast<const u32 *>(p)); }
__device__ __forceinline__ u16 g16(const u8 *p) { return __ldg(reinterpret_cast<const u16 *>(p)); }

__device__ __forceinline__ void emitDfaMatch(const DfaParams &p, u32 id, u32 block, u64 to) {
    const u32 i = atomicAdd(p.counters + CTR_MATCHES, 1u);
    if (i < p.outCap) {
        DevMatch m;
        m.id = id;
        m.block = block;
        m.to = to;
        *reinterpret_cast<uint4 *>(p.out + i) = *reinterpret_cast<const uint4 *>(&m);
    }
}

/* struct report_list {u32 count; ReportID report[]} at NFA offset `off` */
__device__ void emitReportList(const DfaParams &p, u32 off, u32 block, u64 to) {
    const u32 n = g32(p.nfa + off);
    for (u32 i = 0; i < n; i++) {
        emitDfaMatch(p, g32(p.nfa + off + 4 + 4 * i), block, to);
    }
}

struct BlockSpan {
    const u8 *base;
    u32 len;
};

/* 16 corpus bytes at q (16-byte aligned: blocks start aligned); bytes past the readable
 * end read as zero (they lie behind the block's end and are not consumed) */
__device__ __forceinline__ uint4 load16(const DfaParams &p, const u8 *q) {
    if (q + 16 <= p.corpus + p.readableEnd) {
        return __ldg(reinterpret_cast<const uint4 *>(q));
    }
    u32 w[4] = {0, 0, 0, 0};
    for (u32 i = 0; i < 16 && q + i < p.corpus + p.readableEnd; i++) {
        w[i >> 2] |= (u32)__ldg(q + i) << (8 * (i & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ BlockSpan blockSpan(const DfaParams &p, u32 b) {
    BlockSpan s;
    if (p.uniformPitch && p.uniformLen) {
        s.base = p.corpus + (u64)b * p.uniformPitch;
        s.len = p.uniformLen;
    } else {
        s.base = p.corpus + __ldg(p.blockOff + b);
        s.len = __ldg(p.blockLen + b);
    }
    return s;
}

/* ---- McClellan ---------------------------------------------------------------- */

/* doSherman16 (mcclellan_common_impl.h:61-93): a 32-byte record {type, len, daddy,
 * chars[len], states u16[len]}; a symbol not listed falls through to daddy's row */
__device__ u32 shermanNext(const u8 *nfa, u32 shermanOffset, u32 shermanLimit, u32 s, u32 cprime,
                           const u16 *succ, u32 as) {
    const u8 *rec = nfa + shermanOffset + SHERMAN_FIXED_SIZE * (s - shermanLimit);
    const u32 len = __ldg(rec + SHERMAN_LEN_OFFSET);
    for (u32 i = 0; i < len; i++) {
        if (__ldg(rec + SHERMAN_CHARS_OFFSET + i) == cprime) {
            const u8 *q = rec + SHERMAN_CHARS_OFFSET + len + 2 * i; /* unaligned u16 */
            return (u32)__ldg(q) | ((u32)__ldg(q + 1) << 8);
        }
    }
    const u32 daddy = g16(rec + SHERMAN_DADDY_OFFSET);
    return __ldg(succ + (daddy << as) + cprime);
}

template <int WIDE16, int SMEM_TABLE>
__global__ void __launch_bounds__(256) mcclellanKernel(const HSB_GRID_CONSTANT DfaParams p) {
    HSB_DYNAMIC_SMEM(smem);
    const u8 *m = p.nfa + sizeof(NFA); /* struct mcclellan */
    u8 *remapS = smem;
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        remapS[i] = __ldg(m + offsetof(McClellan, remap) + i);
    }
    const u8 *succG = m + sizeof(McClellan);
    if (SMEM_TABLE) {
        u32 *d = reinterpret_cast<u32 *>(smem + 256);
        for (u32 i = threadIdx.x; i < (p.tableBytes + 3) / 4; i += blockDim.x) {
            d[i] = __ldg(reinterpret_cast<const u32 *>(succG) + i);
        }
    }
    __syncthreads();
    const u32 as = __ldg(m + offsetof(McClellan, alphaShift));
    const u32 single = __ldg(m + offsetof(McClellan, flags)) & MCCLELLAN_FLAG_SINGLE;
    const u32 arb = g32(m + offsetof(McClellan, arb_report));
    const u32 start = g16(m + offsetof(McClellan, start_anchored));
    const u32 auxOffset = g32(m + offsetof(McClellan, aux_offset));
    const u32 shermanOffset = g32(m + offsetof(McClellan, sherman_offset));
    const u32 shermanLimit = WIDE16 ? g16(m + offsetof(McClellan, sherman_limit)) : 0xffffffffu;
    const u32 acceptLimit8 = g16(m + offsetof(McClellan, accept_limit_8));
    const u8 *succ8 = SMEM_TABLE ? smem + 256 : succG;
    const u16 *succ16 = reinterpret_cast<const u16 *>(succ8);

    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < p.nblocks; b += gridDim.x * blockDim.x) {
        const BlockSpan blk = blockSpan(p, b);
        u32 s = start;
        u32 i = 0;
        while (i < blk.len && s) {
            const uint4 v = load16(p, blk.base + i);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            const u32 n = blk.len - i < 16 ? blk.len - i : 16;
#pragma unroll
            for (u32 j = 0; j < 16; j++) {
                if (j < n && s) {
                    const u32 c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    const u32 cp = remapS[c];
                    u32 e;
                    bool accept;
                    if (WIDE16) {
                        if (s < shermanLimit) {
                            e = SMEM_TABLE ? succ16[(s << as) + cp] : __ldg(succ16 + (s << as) + cp);
                        } else {
                            e = shermanNext(p.nfa, shermanOffset, shermanLimit, s, cp,
                                            reinterpret_cast<const u16 *>(succG), as);
                        }
                        accept = (e & MCC_ACCEPT_FLAG) != 0;
                        e &= MCC_STATE_MASK;
                    } else {
                        e = SMEM_TABLE ? succ8[(s << as) + cp] : __ldg(succ8 + (s << as) + cp);
                        accept = e >= acceptLimit8;
                    }
                    s = e;
                    if (accept) {
                        if (single) {
                            emitDfaMatch(p, arb, b, (u64)i + j + 1);
                        } else {
                            emitReportList(p, g32(p.nfa + auxOffset + sizeof(MStateAux) * s), b, (u64)i + j + 1);
                        }
                    }
                }
            }
            i += 16;
        }
        /* nfaExecMcClellan*_Bi: reports of the final state that fire at end of data */
        const u32 eod = g32(p.nfa + auxOffset + sizeof(MStateAux) * s + offsetof(MStateAux, accept_eod));
        if (eod) {
            emitReportList(p, eod, b, blk.len);
        }
    }
}

/* ---- Sheng ---------------------------------------------------------------------- */

__global__ void __launch_bounds__(256) shengKernel(const HSB_GRID_CONSTANT DfaParams p) {
    HSB_DYNAMIC_SMEM(smem);
    const u8 *sh = p.nfa + sizeof(NFA); /* struct sheng: 256 x 16 successor bytes first */
    {
        u32 *d = reinterpret_cast<u32 *>(smem);
        for (u32 i = threadIdx.x; i < 1024; i += blockDim.x) {
            d[i] = __ldg(reinterpret_cast<const u32 *>(sh) + i);
        }
    }
    __syncthreads();
    const u32 start = __ldg(sh + offsetof(Sheng, anchored));
    const u32 single = __ldg(sh + offsetof(Sheng, flags)) & SHENG_FLAG_SINGLE_REPORT;
    const u32 report = g32(sh + offsetof(Sheng, report));
    const u32 auxOffset = g32(sh + offsetof(Sheng, aux_offset));
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < p.nblocks; b += gridDim.x * blockDim.x) {
        const BlockSpan blk = blockSpan(p, b);
        u32 s = start;
        u32 i = 0;
        while (i < blk.len && !(s & SHENG_STATE_DEAD)) {
            const uint4 v = load16(p, blk.base + i);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            const u32 n = blk.len - i < 16 ? blk.len - i : 16;
#pragma unroll
            for (u32 j = 0; j < 16; j++) {
                if (j < n) {
                    const u32 c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    s = smem[c * 16 + (s & SHENG_STATE_MASK)]; /* pshufb(masks[c], state) */
                    if (s & SHENG_STATE_ACCEPT) {
                        if (single) {
                            emitDfaMatch(p, report, b, (u64)i + j + 1);
                        } else {
                            emitReportList(p, g32(p.nfa + auxOffset + sizeof(SstateAux) * (s & SHENG_STATE_MASK)), b,
                                           (u64)i + j + 1);
                        }
                    }
                }
            }
            i += 16;
        }
        const u32 eod = g32(p.nfa + auxOffset + sizeof(SstateAux) * (s & SHENG_STATE_MASK) +
                            offsetof(SstateAux, accept_eod));
        if (eod) {
            emitReportList(p, eod, b, blk.len);
        }
    }
}

template <int WIDE16, int SMEM_TABLE>
cudaError_t launchMcClellan(const DfaParams &p, int grid, size_t smem, cudaStream_t stream) {
    void (*kern)(const DfaParams) = mcclellanKernel<WIDE16, SMEM_TABLE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        return e;
    }
    HSB_LAUNCH(kern, grid, 256, smem, stream, p);
    return cudaGetLastError();
}

} // namespace

cudaError_t launchDfa(const DfaParams &p, int smCount, int maxSmem, cudaStream_t stream) {
    if (!p.nblocks) {
        return cudaSuccess;
    }
    const u32 perSm = 8; /* 2048 threads per SM when the table is small */
    const int grid = (int)std::min<u64>((u64)smCount * perSm, ((u64)p.nblocks + 255) / 256);
    if (p.kind == NFA_SHENG) {
        HSB_LAUNCH(shengKernel, grid, 256, 4096, stream, p);
        return cudaGetLastError();
    }
    const bool inSmem = p.tableBytes && (size_t)p.tableBytes + 256 + 1024 <= (size_t)maxSmem;
    const size_t smem = 256 + (inSmem ? HSB_ROUNDUP((size_t)p.tableBytes, 16) : 0);
    /* a big table leaves room for one CTA per SM only: keep the grid at one wave of CTAs */
    const int g = inSmem && smem > 24 * 1024
                      ? (int)std::min<u64>((u64)grid, (u64)smCount * std::max(1, (int)(maxSmem / (int)smem)))
                      : grid;
    if (p.kind == NFA_MCCLELLAN_16) {
        return inSmem ? launchMcClellan<1, 1>(p, g, smem, stream) : launchMcClellan<1, 0>(p, g, smem, stream);
    }
    if (p.kind == NFA_MCCLELLAN_8) {
        return inSmem ? launchMcClellan<0, 1>(p, g, smem, stream) : launchMcClellan<0, 0>(p, g, smem, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace hsb
