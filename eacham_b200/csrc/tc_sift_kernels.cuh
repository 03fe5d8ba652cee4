// tc_sift_kernels.cuh -- the default engine for batched 128-d float (SIFT) pairs on sm_100a: bf16 tcgen05 scoring of BOTH
// directions as thread-local scans with threshold pruning, then the exact FP32 re-rank + certainty check of tc_match_kernels.cuh.
//
// Replaces, for N x 128 CV_32F descriptors, per unordered image pair:
//   knnMatch(k=2) both directions   /root/reference/modules/base/features/FeatureMatcherFlann.cpp:17  (exact NORM_L2 semantics)
//   ratio test                      /root/reference/modules/base/features/FeatureMatcherFlann.cpp:21-27
//   gates + mutual filter           /root/reference/apps/sfm/main.cpp:111-146
//
// What changed against the round-1 kernel (tc_match_kernels.cuh, kept behind EACHAM_CFG_SIFT_TC_V1 for A/B runs): that kernel was
// bound by the ALU pipe (68 % busy, tensor pipe 19 %): every score went through a key build and a 3-instruction top-2 update for
// its row, and through two warp-wide REDUX.MIN per column for the other direction. Here
//   * the score matrix of a (256 rows of `first`) x (128 rows of `second`) tile is computed TWICE by the tensor cores, once as
//     D1 = A_half . B^T (TMEM lanes = rows of `first`) and once transposed as D2 = B . A_block^T (TMEM lanes = rows of `second`,
//     one N = 256 MMA chain), from the same shared-memory operands. The tensor pipe has the headroom; in exchange BOTH directions
//     become the same thread-local scan (a thread owns one TMEM lane = one query row) and all cross-lane work disappears;
//   * the scan is pruned: a query's running second-best distance is a bar; eight scores are reduced with four 3-input integer
//     minima and compared with the bar once, and only if some lane of the warp sees a score under its bar (__any_sync) are the
//     eight keys built and inserted. The number of true insertions per query is O(log n); late in a sweep most groups of
//     eight are skipped at 5 ALU instructions instead of 32;
//   * the bar is the MERGED state of the query (all four column parts): each warp keeps tile-local candidates, hands them over
//     through shared memory, and one of the quadrant's four warps (rotating) merges them into the query's state -- shared
//     memory for the resident block of `first`, the per-CTA L2 scratch for `second` -- and publishes the new bar.
// Keys, composites, tie-breaking (lower index wins) and everything after the sweep are those of the round-1 kernel, so
// the candidates differ from it only where D1 and D2 round differently (never for bf16-exact, e.g. integer-valued, inputs), and
// the match sets are exact either way (rerank_ratio_checked).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"
#include "tc_common.cuh"
#include "tc_match_kernels.cuh"

#ifndef EACHAM_EXP
#define EACHAM_EXP 0
#endif

namespace eacham {
namespace tcs {

#if EACHAM_EXP & 16
__device__ long long g_strace[3][512][8];      // [0: epilogue warp 0, 1: epilogue warp 5, 2: MMA issuer][tile][event] (CTA 0, first pair only)
#define SIFT_TRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 512) g_strace[role][tile][ev] = clock64(); } while (0)
#else
#define SIFT_TRACE(role, tile, ev) do { } while (0)
#endif

using tcm::ImageDescTc;
using tcm::PairParamsTc;
using tcm::kEpiWarps;
using tcm::kEpiThreads;
using tcm::kColParts;
using tcm::kABlockRows;
using tcm::kEmptyKeyTc;
using tcm::kEmptyComp;
using tcm::epi_bar;

constexpr int kThreads = 64 + kEpiThreads;
constexpr int kBStages = 3;
constexpr int kAChunks = tc::kDataChunks + 2 * tc::kAugChunks;      // 20 K-chunks per row
constexpr uint32_t kAChunkStride = 2 * tc::kChunkStride;            // resident block: the two 128-row halves side by side per K-chunk (4,096 B)
constexpr int kAugChunkM = tc::kDataChunks;                         // augmentation of the M-side (negated) operand: chunks 16, 17
constexpr int kAugChunkN = tc::kDataChunks + tc::kAugChunks;        // augmentation of the N-side operand: chunks 18, 19

struct SmemSift {
    uint8_t a[kAChunks * kAChunkStride];          // 256 rows of `first`: chunk c, half h, row r at c * 4096 + h * 2048 + r * 16 (rows contiguous: N = 256 operand)
    uint8_t b[kBStages][tc::kBlockBytes];         // 128 rows of `second`, whole pre-tiled block (both augmentations)
    long long rowstate[kABlockRows][2];           // merged (best, second) composites of the resident rows
    int32_t rowbar[kABlockRows];                  // their bar: scores >= bar cannot enter the state
    uint2 slots1[kColParts][kABlockRows];         // tile-local candidates of D1, per column part
    uint2 slots2[kColParts][128];                 // tile-local candidates of D2, per column part
    uint64_t b_full[kBStages], b_empty[kBStages], a_full, a_empty, acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
    uint32_t red[2 * kEpiWarps + 8];
    unsigned long long base;
};

__device__ __forceinline__ void quad_bar(int q) { asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory"); }

// the bar that belongs to a state whose second-best composite has the value word `hi` (= key >> 8): a score can still enter
// (or tie, and win on the index) iff (score bits & ~0xFF) <= hi << 8, i.e. iff score bits < (hi + 1) << 8
__device__ __forceinline__ int32_t bar_of(int32_t hi) { return min((hi + 1) << 8, kEmptyKeyTc); }

__device__ __forceinline__ long long comp_of(uint32_t key, uint32_t base) {
    return ((long long)((int32_t)key >> 8) << 32) | (long long)(base + (key & 0xFFu));
}

// key = (score bits & mask) | index byte as ONE LOP3: mask in a register, index immediate
template <int kIdx>
__device__ __forceinline__ int32_t key_of(uint32_t v, uint32_t mask) {
    int32_t key;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(key) : "r"(v), "r"(mask), "n"(kIdx));
    return key;
}
__device__ __forceinline__ void insert(int32_t key, int32_t& c0, int32_t& c1) {
    c1 = min(c1, max(c0, key));
    c0 = min(c0, key);
}

// 8 scores v[kOff .. kOff + 8) of one query (bit patterns of D as signed integers), index byte kIdx0 + position. c0 <= c1: the
// tile-local candidates (keys); bar: scores >= bar are of no interest. Warp-uniform control flow.
template <int kIdx0, int kOff>
__device__ __forceinline__ void scan8(const uint32_t (&v)[16], uint32_t mask, int32_t& c0, int32_t& c1, int32_t& bar) {
    const int32_t x0 = (int32_t)v[kOff + 0], x1 = (int32_t)v[kOff + 1], x2 = (int32_t)v[kOff + 2], x3 = (int32_t)v[kOff + 3];
    const int32_t x4 = (int32_t)v[kOff + 4], x5 = (int32_t)v[kOff + 5], x6 = (int32_t)v[kOff + 6], x7 = (int32_t)v[kOff + 7];
    const int32_t mn = min(min(min(x0, x1), x2), min(min(min(x3, x4), x5), min(x6, x7)));
    if (__any_sync(0xffffffffu, mn < bar)) {
        insert(key_of<kIdx0 + 0>(v[kOff + 0], mask), c0, c1); insert(key_of<kIdx0 + 1>(v[kOff + 1], mask), c0, c1);
        insert(key_of<kIdx0 + 2>(v[kOff + 2], mask), c0, c1); insert(key_of<kIdx0 + 3>(v[kOff + 3], mask), c0, c1);
        insert(key_of<kIdx0 + 4>(v[kOff + 4], mask), c0, c1); insert(key_of<kIdx0 + 5>(v[kOff + 5], mask), c0, c1);
        insert(key_of<kIdx0 + 6>(v[kOff + 6], mask), c0, c1); insert(key_of<kIdx0 + 7>(v[kOff + 7], mask), c0, c1);
        bar = min(bar, (c1 + 255) & ~255);
    }
}
template <int kIdx0>
__device__ __forceinline__ void scan16(const uint32_t (&v)[16], uint32_t mask, int32_t& c0, int32_t& c1, int32_t& bar) {
    scan8<kIdx0, 0>(v, mask, c0, c1, bar);
    scan8<kIdx0 + 8, 8>(v, mask, c0, c1, bar);
}

// sum over the 32 lanes of 16 values per lane at once: afterwards lanes 2r and 2r + 1 hold the total of value r in p[0]. Same offsets
// (16, 8, 4, 2, 1) and the same operand pairs as the usual xor butterfly, so each total has exactly the bits of
//   for (o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
// -- at one shuffle + one add per value instead of five.
__device__ __forceinline__ void multi_reduce16(float (&p)[16], int lane) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {                   // w values stay per lane after this stage; lane bit (w << 1) picks the half
        const bool up = (lane & (w << 1)) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? p[i] : p[i + w], keep = up ? p[i + w] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, w << 1);
        }
    }
    p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
}

// Exact re-rank + ratio test (FeatureMatcherFlann.cpp:21-27) of a strided range of queries: query q_first + k * step (k < n_iter,
// q < q_end) has its two candidates in state[(s_first + k * step) * 2 ..]. Sixteen queries per warp pass: every lane accumulates
// its four dimensions of all sixteen (query, candidate) distances, one multi-value butterfly per candidate sums them, and lane
// 2r (and 2r + 1) then runs the certainty logic of tcm::rerank_ratio_checked for query r -- same arithmetic, same bits, but the
// scalar part once per sixteen queries instead of once per query, and twelve gathered rows in flight per warp instead of three
// (one query at a time was measured at ~2,400 cycles per query, a fifth of the kernel). The only difference: the query's own norm
// in the error bound is replaced by its image's maximum norm (still a bound). Not inlined: its registers must not leak into the
// scan loop's allocation.
__device__ __noinline__ void rerank_range(const long long* state, uint32_t s_first, uint32_t q_first, uint32_t step, uint32_t n_iter, uint32_t q_end,
                                          const float* __restrict__ Q, const float* __restrict__ T, uint32_t n_train, double ratio, int lane,
                                          bool exact_inputs, float own_max_norm, float other_max_norm, uint32_t* fallbacks, int32_t* dbg_idx, float* dbg_dist,
                                          uint32_t* out) {
    const float kInf = __int_as_float(0x7f800000);
    for (uint32_t k0 = 0; k0 < n_iter; k0 += 16) {
        float p0[16], p1[16];
#pragma unroll
        for (int g = 0; g < 16; g += 4) {                // four queries = twelve rows in flight
            float4 a4[4], b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t k = k0 + g + i, qi = q_first + k * step, si = s_first + k * step;
                a4[i] = b0[i] = b1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < n_iter && qi < q_end) {
                    const uint32_t j0 = (uint32_t)state[2 * (size_t)si], j1 = (uint32_t)state[2 * (size_t)si + 1];
                    a4[i] = __ldg(reinterpret_cast<const float4*>(Q + (size_t)qi * 128) + lane);
                    if (j0 < n_train) b0[i] = __ldg(reinterpret_cast<const float4*>(T + (size_t)j0 * 128) + lane);
                    if (j1 < n_train) b1[i] = __ldg(reinterpret_cast<const float4*>(T + (size_t)j1 * 128) + lane);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {                // the partial sums of tcm::exact_l2, before its butterfly
                float dx = a4[i].x - b0[i].x, dy = a4[i].y - b0[i].y, dz = a4[i].z - b0[i].z, dw = a4[i].w - b0[i].w;
                float t = dx * dx;
                t = fmaf(dy, dy, t); t = fmaf(dz, dz, t); t = fmaf(dw, dw, t);
                p0[g + i] = t;
                dx = a4[i].x - b1[i].x; dy = a4[i].y - b1[i].y; dz = a4[i].z - b1[i].z; dw = a4[i].w - b1[i].w;
                t = dx * dx;
                t = fmaf(dy, dy, t); t = fmaf(dz, dz, t); t = fmaf(dw, dw, t);
                p1[g + i] = t;
            }
        }
        multi_reduce16(p0, lane);
        multi_reduce16(p1, lane);
        // lane 2r / 2r + 1: query r of this pass
        const uint32_t k = k0 + (uint32_t)(lane >> 1), qi = q_first + k * step, si = s_first + k * step;
        const bool valid = k < n_iter && qi < q_end;
        uint32_t j0 = EACHAM_NONE, j1 = EACHAM_NONE;
        if (valid) { j0 = (uint32_t)state[2 * (size_t)si]; j1 = (uint32_t)state[2 * (size_t)si + 1]; }
        const bool v0 = j0 < n_train, v1 = j1 < n_train;
        float d0 = v0 ? __fsqrt_rn(p0[0]) : kInf, d1 = v1 ? __fsqrt_rn(p1[0]) : kInf;
        if (d1 < d0 || (d1 == d0 && j1 < j0)) { const float t = d0; d0 = d1; d1 = t; const uint32_t u = j0; j0 = j1; j1 = u; }
        uint32_t result = EACHAM_NONE;
        bool certain = true;
        if (v0 && v1) {                                  // fewer than two neighbours: the reference is UB, rejected
            float E = 0.5f * d1 * d1 * 3.0517578125e-5f;                                   // E_trunc
            if (!exact_inputs) {
                const float na = own_max_norm * 1.004f, nb = other_max_norm * 1.004f;
                const float delta = 1.953125e-3f * (na + nb) * 1.01f;
                E += 0.5f * delta * (2.f * d1 + delta) + 1.52587890625e-5f * 0.5f * (na * na + nb * nb);
            }
            const float X = d1 * d1 - 4.f * E * 1.001f;
            const bool lo_pass = (double)__fdiv_rn(d0, d1) < ratio;
            certain = false;
            if (X > 0.f) {
                const float sx = sqrtf(X) * 0.999999f;
                if ((double)(d0 / sx * 1.000001f) < ratio) { certain = true; result = j0; }
                else if (!lo_pass && (double)sx >= ratio * (double)d0 * 1.000001) certain = true;
            }
        }
        // the uncertain ones: exact scan over all train rows, one query at a time, the whole warp on each
        unsigned need = __ballot_sync(0xffffffffu, valid && !certain && (lane & 1) == 0);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint32_t qs = __shfl_sync(0xffffffffu, qi, src);
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(Q + (size_t)qs * 128) + lane);
            float e0, e1; uint32_t i0, i1;
            tcm::exact_scan_top2(a4, T, n_train, lane, e0, e1, i0, i1);
            if ((lane >> 1) == (src >> 1)) { d0 = e0; d1 = e1; j0 = i0; j1 = i1; result = ((double)__fdiv_rn(e0, e1) < ratio) ? i0 : EACHAM_NONE; }
            if (lane == 0 && fallbacks != nullptr) atomicAdd(fallbacks, 1u);
        }
        if (valid && (lane & 1) == 0) {
            out[qi] = result;
            if (dbg_idx != nullptr) {
                dbg_idx[2 * (size_t)qi] = v0 ? (int32_t)j0 : -1; dbg_idx[2 * (size_t)qi + 1] = v1 ? (int32_t)j1 : -1;
                dbg_dist[2 * (size_t)qi] = d0; dbg_dist[2 * (size_t)qi + 1] = d1;
            }
        }
    }
}

// =============================================================================================================
// The fused SIFT pair kernel: one persistent CTA per SM, pairs in the host's L2-blocked order. 576 threads:
//   warps 0-15 epilogue (4 per TMEM lane quadrant; column part cp = warp / 4)
//   warp 16   producer: 1-D bulk copies of pre-tiled bf16 blocks
//   warp 17   MMA issuer: per tile  D1: 2 halves x 9 x (M 128, N 128, K 16)   D2: 9 x (M 128, N 256, K 16)
// TMEM: two regions of 256 columns; MMA "steps" alternate between them (D1, D2, D1, ... or D1, D1, ... in single-direction
// mode), so the tensor core fills one region while the epilogue scans the other.
// =============================================================================================================
__global__ void __launch_bounds__(kThreads, 1) sift_tc_match_pairs_kernel(const PairParamsTc p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemSift& S = *reinterpret_cast<SmemSift*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { tc::mbar_init(&S.b_full[s], 1); tc::mbar_init(&S.b_empty[s], 1); }
        tc::mbar_init(&S.a_full, 1); tc::mbar_init(&S.a_empty, 1);
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&S.acc_full[s], 1); tc::mbar_init(&S.acc_empty[s], kEpiWarps); }
        tc::fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tc::tmem_alloc(&S.tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = S.tmem_slot;
    const bool both = !p.single_dir;

    if (warp == kEpiWarps) {
        // ===================================== producer =====================================
        uint32_t b_it = 0, a_it = 0;
        for (uint32_t wk = blockIdx.x; wk < p.n_pairs; wk += gridDim.x) {
            const uint32_t pi = p.single_dir ? wk : p.order[wk];
            const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
            const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
            if (A.rows == 0 || B.rows == 0) continue;
            const uint32_t na128 = (A.rows + 127) / 128, nbt = (B.rows + 127) / 128;
            const uint8_t* Ab = p.tc_arena + A.tc_offset;
            const uint8_t* Bb = p.tc_arena + B.tc_offset;
            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab, ++a_it) {
                const uint32_t blk0 = p.single_dir ? pi : ab * 2;
                const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2);
                tc::mbar_wait(&S.a_empty, (a_it & 1) ^ 1);
                if (tc::elect_one()) {
                    tc::mbar_expect_tx(&S.a_full, nh * tc::kBlockBytes);
                    for (uint32_t h = 0; h < nh; ++h) {
                        const uint8_t* src = Ab + (size_t)(blk0 + h) * tc::kBlockBytes;
                        for (int c = 0; c < kAChunks; ++c)
                            tc::bulk_g2s(S.a + c * kAChunkStride + h * tc::kChunkStride, src + c * tc::kChunkStride, tc::kChunkStride, &S.a_full);
                    }
                }
                __syncwarp();
                for (uint32_t bt = 0; bt < nbt; ++bt, ++b_it) {
                    const uint32_t st = b_it % kBStages;
                    tc::mbar_wait(&S.b_empty[st], ((b_it / kBStages) & 1) ^ 1);
                    if (tc::elect_one()) {
                        tc::mbar_expect_tx(&S.b_full[st], tc::kBlockBytes);
                        tc::bulk_g2s(S.b[st], Bb + (size_t)bt * tc::kBlockBytes, tc::kBlockBytes, &S.b_full[st]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===================================== MMA issuer =====================================
        const uint64_t dbase_a = tc::make_smem_desc_base(kAChunkStride, tc::kSBO);
        const uint64_t dbase_b = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
        const uint32_t idesc128 = tc::make_idesc_bf16_f32(128, 128, true), idesc256 = tc::make_idesc_bf16_f32(128, 256, true);
        const uint32_t a_addr = tc::smem_u32(S.a);
        uint32_t b_it = 0, a_it = 0, step_it = 0;
        for (uint32_t wk = blockIdx.x; wk < p.n_pairs; wk += gridDim.x) {
            const uint32_t pi = p.single_dir ? wk : p.order[wk];
            const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
            const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
            if (A.rows == 0 || B.rows == 0) continue;
            const uint32_t na128 = (A.rows + 127) / 128, nbt = (B.rows + 127) / 128;
            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab, ++a_it) {
                const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2);
                tc::mbar_wait(&S.a_full, a_it & 1);
                for (uint32_t bt = 0; bt < nbt; ++bt, ++b_it) {
                    const uint32_t st = b_it % kBStages;
                    const uint32_t b_addr = tc::smem_u32(S.b[st]);
                    tc::mbar_wait(&S.b_full[st], (b_it / kBStages) & 1);
                    SIFT_TRACE(2, b_it, 0);
                    {   // D1: rows of `first` in the TMEM lanes
                        const uint32_t reg = step_it & 1;
                        tc::mbar_wait(&S.acc_empty[reg], ((step_it >> 1) & 1) ^ 1);
                        tc::tc_fence_after();
                        SIFT_TRACE(2, b_it, 1);
                        if (tc::elect_one()) {
                            for (uint32_t h = 0; h < nh; ++h) {
                                const uint32_t d = tmem + reg * 256 + h * 128;
#pragma unroll
                                for (int ks = 0; ks < tc::kKSteps; ++ks) {
                                    const int ca = ks < 8 ? 2 * ks : kAugChunkM, cb = ks < 8 ? 2 * ks : kAugChunkN;
                                    const uint64_t da = tc::smem_desc(dbase_a, a_addr + ca * kAChunkStride + h * tc::kChunkStride);
                                    const uint64_t db = tc::smem_desc(dbase_b, b_addr + cb * tc::kChunkStride);
                                    tc::mma_bf16(d, da, db, idesc128, ks > 0);
                                }
                            }
                            if (!both) tc::mma_commit(&S.b_empty[st]);
                            tc::mma_commit(&S.acc_full[reg]);
                        }
                        __syncwarp();
                        ++step_it;
                    }
                    if (both) {   // D2: rows of `second` in the TMEM lanes, the resident block as one N = nh * 128 operand
                        const uint32_t reg = step_it & 1;
                        tc::mbar_wait(&S.acc_empty[reg], ((step_it >> 1) & 1) ^ 1);
                        tc::tc_fence_after();
                        SIFT_TRACE(2, b_it, 2);
                        if (tc::elect_one()) {
                            const uint32_t d = tmem + reg * 256;
#pragma unroll
                            for (int ks = 0; ks < tc::kKSteps; ++ks) {
                                const int cm = ks < 8 ? 2 * ks : kAugChunkM, cn = ks < 8 ? 2 * ks : kAugChunkN;
                                const uint64_t dm = tc::smem_desc(dbase_b, b_addr + cm * tc::kChunkStride);
                                const uint64_t dn = tc::smem_desc(dbase_a, a_addr + cn * kAChunkStride);
                                tc::mma_bf16(d, dm, dn, nh == 2 ? idesc256 : idesc128, ks > 0);
                            }
                            tc::mma_commit(&S.b_empty[st]);
                            tc::mma_commit(&S.acc_full[reg]);
                        }
                        __syncwarp();
                        ++step_it;
                    }
                }
                if (tc::elect_one()) tc::mma_commit(&S.a_empty);
                __syncwarp();
            }
        }
    } else {
        // ===================================== epilogue =====================================
        const int e = warp, q = warp & 3, cp = e >> 2;
        const int et = e * 32 + lane;
        uint8_t* my_scratch = p.scratch + (size_t)blockIdx.x * tcm::tc_scratch_bytes_per_cta(p.rows_cap, p.cols_cap);
        long long* colstate = reinterpret_cast<long long*>(my_scratch);                        // [cols_cap][2]
        uint32_t* m12 = reinterpret_cast<uint32_t*>(my_scratch + (size_t)p.cols_cap * 16);     // [rows_cap]
        uint32_t* m21 = m12 + p.rows_cap;                                                      // [cols_cap]
        uint32_t mask = 0xFFFFFF00u;
        asm volatile("" : "+r"(mask));                       // keep the mask in a register: key = (v & mask) | immediate is one LOP3
        uint32_t step_it = 0;
        for (int i = et; i < kABlockRows; i += kEpiThreads) { S.rowstate[i][0] = kEmptyComp; S.rowstate[i][1] = kEmptyComp; S.rowbar[i] = kEmptyKeyTc; }
        for (uint32_t wk = blockIdx.x; wk < p.n_pairs; wk += gridDim.x) {
            const uint32_t pi = p.single_dir ? wk : p.order[wk];
            const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
            const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
            const uint32_t N = A.rows, M = B.rows;
            if (N == 0 || M == 0) {
                if (et == 0) {
                    eacham_pair_result_t r;
                    r.n12 = 0; r.n21 = 0; r.n_mutual = 0; r.flags = (0u < p.min_dir) ? EACHAM_PAIR_GATED : 0u; r.offset = 0; r.count = 0;
                    p.results[pi] = r;
                }
                continue;
            }
            const float* Af = reinterpret_cast<const float*>(p.arena + A.offset);
            const float* Bf = reinterpret_cast<const float*>(p.arena + B.offset);
            const bool both_exact = A.bf16_exact != 0 && B.bf16_exact != 0;
            const uint32_t na128 = (N + 127) / 128, nbt = (M + 127) / 128;
            if (both)
                for (uint32_t j = et; j < nbt * 128; j += kEpiThreads) { colstate[2 * j] = kEmptyComp; colstate[2 * j + 1] = kEmptyComp; }
            epi_bar();

            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab) {
                const uint32_t blk0 = p.single_dir ? pi : ab * 2;
                const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2);
                for (uint32_t bt = 0; bt < nbt; ++bt) {
                    // second image's state of this tile's rows: issue the (L2) load early
                    longlong2 cs = make_longlong2(0, 0);
                    const uint32_t jrow = bt * 128 + q * 32 + lane;
                    if (both) cs = *reinterpret_cast<const longlong2*>(colstate + 2 * (size_t)jrow);
                    // ---------------- D1: each lane owns one row of each resident half; 32 columns per warp ----------------
                    {
                        const uint32_t reg = step_it & 1;
                        int32_t c0[2] = {kEmptyKeyTc, kEmptyKeyTc}, c1[2] = {kEmptyKeyTc, kEmptyKeyTc}, bar[2];
                        // padding rows (beyond N) never take part: a bar below every score
                        bar[0] = (blk0 * 128 + q * 32 + lane < N) ? S.rowbar[q * 32 + lane] : INT32_MIN;
                        bar[1] = (blk0 * 128 + 128 + q * 32 + lane < N) ? S.rowbar[128 + q * 32 + lane] : INT32_MIN;
                        // single-direction mode has one barrier per tile, so its hand-over slots alternate between the two halves of slots1
                        const int sb = both ? 0 : (int)(bt & 1u) * 128;
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 0);
                        tc::mbar_wait(&S.acc_full[reg], (step_it >> 1) & 1);
                        tc::tc_fence_after();
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 1);
                        const uint32_t taddr = tmem + reg * 256 + ((uint32_t)(q * 32) << 16) + cp * 32;
                        {
                            // 16-column loads, ping-pong: the next load is in flight while this one is scanned
                            uint32_t v[16], w[16];
                            tc::tmem_ld16(taddr, v);
                            tc::tmem_ld_wait();
                            tc::tmem_ld16(taddr + 16, w);
                            scan16<0>(v, mask, c0[0], c1[0], bar[0]);
                            tc::tmem_ld_wait();
                            if (nh == 2) tc::tmem_ld16(taddr + 128, v);
                            scan16<16>(w, mask, c0[0], c1[0], bar[0]);
                            if (nh == 2) {
                                tc::tmem_ld_wait();
                                tc::tmem_ld16(taddr + 144, w);
                                scan16<0>(v, mask, c0[1], c1[1], bar[1]);
                                tc::tmem_ld_wait();
                                scan16<16>(w, mask, c0[1], c1[1], bar[1]);
                            }
                        }
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&S.acc_empty[reg]);
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 2);
                        ++step_it;
                        S.slots1[cp][sb + q * 32 + lane] = make_uint2((uint32_t)c0[0], (uint32_t)c1[0]);
                        if (nh == 2) S.slots1[cp][128 + q * 32 + lane] = make_uint2((uint32_t)c0[1], (uint32_t)c1[1]);
                        quad_bar(q);
                        if (cp == (int)(bt & 3u)) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h < (int)nh) {
                                    const int row = h * 128 + q * 32 + lane;
                                    const int32_t bar0 = S.rowbar[row];
                                    const uint2 s0 = S.slots1[0][sb + row], s1 = S.slots1[1][sb + row], s2 = S.slots1[2][sb + row], s3 = S.slots1[3][sb + row];
                                    if (min(min((int32_t)s0.x, (int32_t)s1.x), min((int32_t)s2.x, (int32_t)s3.x)) < bar0) {
                                        long long g0 = S.rowstate[row][0], g1 = S.rowstate[row][1];
                                        const uint32_t base = bt * 128;
                                        if ((int32_t)s0.x < bar0) tcm::comp_merge(comp_of(s0.x, base), comp_of(s0.y, base), g0, g1);
                                        if ((int32_t)s1.x < bar0) tcm::comp_merge(comp_of(s1.x, base + 32), comp_of(s1.y, base + 32), g0, g1);
                                        if ((int32_t)s2.x < bar0) tcm::comp_merge(comp_of(s2.x, base + 64), comp_of(s2.y, base + 64), g0, g1);
                                        if ((int32_t)s3.x < bar0) tcm::comp_merge(comp_of(s3.x, base + 96), comp_of(s3.y, base + 96), g0, g1);
                                        S.rowstate[row][0] = g0; S.rowstate[row][1] = g1;
                                        S.rowbar[row] = bar_of((int32_t)(g1 >> 32));
                                    }
                                }
                            }
                        }
                    }
                    if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, (step_it - 1) >> 1, 3);
                    if (!both) continue;
                    // ---------------- D2: each lane owns one row of the streamed tile; 64 columns (resident rows) per warp ----------------
                    {
                        const uint32_t reg = step_it & 1;
                        const uint32_t ncols = nh * 128;
                        const int32_t bar0 = bar_of((int32_t)(cs.y >> 32));
                        int32_t c0 = kEmptyKeyTc, c1 = kEmptyKeyTc, bar = jrow < M ? bar0 : INT32_MIN;
                        tc::mbar_wait(&S.acc_full[reg], (step_it >> 1) & 1);
                        tc::tc_fence_after();
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 4);
                        const uint32_t taddr = tmem + reg * 256 + ((uint32_t)(q * 32) << 16) + cp * 64;
                        if ((uint32_t)cp * 64 < ncols) {
                            uint32_t v[16], w[16];                      // cp * 64 + 64 <= ncols too: ncols is 128 or 256
                            tc::tmem_ld16(taddr, v);
                            tc::tmem_ld_wait();
                            tc::tmem_ld16(taddr + 16, w);
                            scan16<0>(v, mask, c0, c1, bar);
                            tc::tmem_ld_wait();
                            tc::tmem_ld16(taddr + 32, v);
                            scan16<16>(w, mask, c0, c1, bar);
                            tc::tmem_ld_wait();
                            tc::tmem_ld16(taddr + 48, w);
                            scan16<32>(v, mask, c0, c1, bar);
                            tc::tmem_ld_wait();
                            scan16<48>(w, mask, c0, c1, bar);
                        }
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&S.acc_empty[reg]);
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 5);
                        ++step_it;
                        S.slots2[cp][q * 32 + lane] = make_uint2((uint32_t)c0, (uint32_t)c1);
                        quad_bar(q);
                        if (cp == (int)((bt + 2) & 3u)) {
                            const int r = q * 32 + lane;
                            const uint2 s0 = S.slots2[0][r], s1 = S.slots2[1][r], s2 = S.slots2[2][r], s3 = S.slots2[3][r];
                            if (min(min((int32_t)s0.x, (int32_t)s1.x), min((int32_t)s2.x, (int32_t)s3.x)) < bar0) {
                                long long g0 = cs.x, g1 = cs.y;
                                const uint32_t base = blk0 * 128;
                                if ((int32_t)s0.x < bar0) tcm::comp_merge(comp_of(s0.x, base), comp_of(s0.y, base), g0, g1);
                                if ((int32_t)s1.x < bar0) tcm::comp_merge(comp_of(s1.x, base + 64), comp_of(s1.y, base + 64), g0, g1);
                                if ((int32_t)s2.x < bar0) tcm::comp_merge(comp_of(s2.x, base + 128), comp_of(s2.y, base + 128), g0, g1);
                                if ((int32_t)s3.x < bar0) tcm::comp_merge(comp_of(s3.x, base + 192), comp_of(s3.y, base + 192), g0, g1);
                                *reinterpret_cast<longlong2*>(colstate + 2 * (size_t)jrow) = make_longlong2(g0, g1);
                            }
                        }
                        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, (step_it - 1) >> 1, 6);
                    }
                }
                // ---- rows of this block are complete: re-rank exactly, ratio test ----
                epi_bar();                                                  // every tile's merge is in rowstate
                rerank_range(&S.rowstate[0][0], e * (kABlockRows / kEpiWarps), blk0 * 128 + e * (kABlockRows / kEpiWarps), 1, kABlockRows / kEpiWarps,
                             min(N, (blk0 + nh) * 128), Af, Bf, M, p.ratio, lane, both_exact, __uint_as_float(A.max_norm_bits), __uint_as_float(B.max_norm_bits), p.exact_fallbacks,
                             p.dbg_idx12, p.dbg_dist12, p.single_dir ? p.single_out : m12);
                epi_bar();
                for (int i = et; i < kABlockRows; i += kEpiThreads) { S.rowstate[i][0] = kEmptyComp; S.rowstate[i][1] = kEmptyComp; S.rowbar[i] = kEmptyKeyTc; }
                epi_bar();
            }

            if (p.single_dir) continue;                       // one direction only
            // ---- rows of `second`: re-rank, ratio -> m21 ----
            __threadfence_block();
            rerank_range(colstate, e, e, kEpiWarps, (M + kEpiWarps - 1 - e) / kEpiWarps, M, Bf, Af, N, p.ratio, lane, both_exact, __uint_as_float(B.max_norm_bits), __uint_as_float(A.max_norm_bits),
                         p.exact_fallbacks, p.dbg_idx21, p.dbg_dist21, m21);
            __threadfence_block();
            epi_bar();

            tcm::gates_mutual_compact(S, p, pi, N, M, m12, m21, et, e, lane);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) tc::tmem_dealloc(tmem, 512);
#if EACHAM_EXP & 16
    if (blockIdx.x == 0 && tid == 0) {
        const long long t0 = g_strace[2][300][0];
        for (int t = 300; t < 324; ++t) {
            printf("tile %d mma: b_full %lld d1_issue %lld d2_issue %lld |", t, g_strace[2][t][0] - t0, g_strace[2][t][1] - t0, g_strace[2][t][2] - t0);
            for (int w = 0; w < 2; ++w)
                printf(" w%d: top %lld acc1 %lld scan1 %lld bar1 %lld acc2 %lld scan2 %lld bar2 %lld |", w ? 5 : 0, g_strace[w][t][0] - t0, g_strace[w][t][1] - t0,
                       g_strace[w][t][2] - t0, g_strace[w][t][3] - t0, g_strace[w][t][4] - t0, g_strace[w][t][5] - t0, g_strace[w][t][6] - t0);
            printf("\n");
        }
    }
#endif
}

}  // namespace tcs
}  // namespace eacham
