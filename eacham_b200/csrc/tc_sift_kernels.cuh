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
// its row, and through two warp-wide REDUX.MIN per column for the other direction; and a fifth of its time went into re-ranking
// one query at a time (a full memory round trip each). Here
//   * the score matrix of a (256 rows of `first`) x (128 rows of `second`) tile is computed TWICE by the tensor cores, once as
//     D1 = A_half . B^T (TMEM lanes = rows of `first`) and once transposed as D2 = B . A_block^T (TMEM lanes = rows of `second`,
//     one N = 256 MMA chain), from the same shared-memory operands. The tensor pipe has the headroom; in exchange BOTH directions
//     become the same thread-local scan (a thread owns one TMEM lane = one query row) and all cross-lane work disappears;
//   * the scan is pruned: a query's running second-best distance is a bar; eight scores are reduced with three 3-input integer
//     minima + one minimum and compared with the bar once, and only if some lane of the warp sees a score under its bar
//     (__any_sync) are the eight keys built and inserted. True insertions per query are O(log n); measured, 32 % of the groups
//     of eight take the slow path (that IS the insertion count: ~1.8 inserting lanes per slow group);
//   * no barrier inside the sweep: each of the four column parts (warps) of a lane quadrant keeps its own running (best, second)
//     keys -- registers for the resident rows, the per-CTA L2 scratch for the streamed ones -- publishes them (shared memory / the
//     scratch) and derives its bar from ALL four parts' published keys (second smallest of the eight), read without
//     synchronisation: a stale copy only means a higher bar, never a wrong result. The parts are merged once per row block /
//     once per pair;
//   * the exact re-rank handles sixteen queries per warp pass with twelve gathered rows in flight, a multi-value butterfly (one
//     shuffle + one add per distance instead of five, same operand pairs, same bits) and the certainty logic once per sixteen
//     queries (rerank_range).
// Keys, composites and tie-breaking (lower index wins) are those of the round-1 kernel, so the candidates differ from it only where
// D1 and D2 round differently (never for bf16-exact, e.g. integer-valued, inputs), and the match sets are exact either way
// (tcm::rerank_ratio_checked's bound; the query's own norm is replaced by its image's maximum norm, still a bound).
//   * ratio-aware pruning (low_bar / `skip`): scores that could only ever be the second of a non-match are not inserted at all.
// Measured (500 images x 8192, 124,750 pairs): 32.8k pairs/s against 22.4k; tensor pipe 69 %, ALU 51 %.
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
constexpr uint32_t kBarEvery = 4;                                   // D1: the other parts' keys are re-read every this many tiles (power of two)
constexpr int kBStages = 3;
constexpr int kAChunks = tc::kDataChunks + 2 * tc::kAugChunks;      // 20 K-chunks per row
constexpr uint32_t kAChunkStride = 2 * tc::kChunkStride;            // resident block: the two 128-row halves side by side per K-chunk (4,096 B)
constexpr int kAugChunkM = tc::kDataChunks;                         // augmentation of the M-side (negated) operand: chunks 16, 17
constexpr int kAugChunkN = tc::kDataChunks + tc::kAugChunks;        // augmentation of the N-side operand: chunks 18, 19

struct SmemSift {
    uint8_t a[kAChunks * kAChunkStride];          // 256 rows of `first`: chunk c, half h, row r at c * 4096 + h * 2048 + r * 16 (rows contiguous: N = 256 operand)
    uint8_t b[kBStages][tc::kBlockBytes];         // 128 rows of `second`, whole pre-tiled block (both augmentations)
    long long rowstate[kABlockRows][2];           // merged (best, second) composites of the resident rows, built at the end of a row block
    uint2 share1[kColParts][kABlockRows];         // D1: every column part's running (best, second) keys of the resident rows -- the other parts' bars
    uint2 tiles1[kColParts][kABlockRows];         // D1: their tile ids (t0 | t1 << 16) and their smallest passed-over score, published at the end of a row block
    int32_t rowskip[kABlockRows];                 // smallest passed-over score of the resident rows over all parts (end of a row block)
    uint64_t b_full[kBStages], b_empty[kBStages], a_full, a_empty, acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
    uint32_t red[2 * kEpiWarps + 8];
    unsigned long long base;
};

// What a query still has to look at.
//   (1) Only a score below its second best can enter its top two: bar <= second best (rounded up to the next multiple of 256, the
//       key granularity, so that ties go through and are settled by the index).
//   (2) The OUTPUT for a query is its best neighbour's index if best / second < ratio, else nothing. A score x >= ratio^2 * best
//       (scores are squared distances / 2) can never be the best of a MATCH: as a new best it would have the old best as a second
//       at ratio >= `ratio`; as a second it only decides match / no match, for which its VALUE is enough. So such scores are not
//       inserted at all (r2s = ratio^2 plus a margin); the smallest of them is kept per query (`skip`, a running minimum -- it
//       changes O(log n) times) and rerank_range folds it into the decision with the scorer's error bound, falling back to the
//       exact scan in the rare case where the decision would hinge on a passed-over score's exact value. Insertions drop from
//       every new best-or-second to the new bests of potential matches.
__device__ __forceinline__ int32_t low_bar(int32_t best_key, float r2s) {
    return (__float_as_int(__int_as_float(best_key & (int32_t)0xFFFFFF00) * r2s) + 255) & ~255;
}
__device__ __forceinline__ int32_t bar_of2(int32_t c0, int32_t c1, float r2s) { return min((c1 + 255) & ~255, low_bar(c0, r2s)); }
// from the running (best, second) keys of the query's four column parts: the second smallest of the eight keys and the smallest
// best. Any STALE copy of a part's keys gives a valid (higher) bar.
__device__ __forceinline__ int32_t bar_of8(uint2 a, uint2 b, uint2 c, uint2 d, float r2s) {
    const int32_t a0 = (int32_t)a.x, b0 = (int32_t)b.x, c0 = (int32_t)c.x, d0 = (int32_t)d.x;
    const int32_t lo01 = min(a0, b0), hi01 = max(a0, b0), lo23 = min(c0, d0), hi23 = max(c0, d0);
    const int32_t x = min(min(hi01, hi23), max(lo01, lo23));                                     // second smallest of the four bests
    const int32_t y = min(min((int32_t)a.y, (int32_t)b.y), min((int32_t)c.y, (int32_t)d.y));     // smallest of the four seconds
    return min((min(x, y) + 255) & ~255, low_bar(min(lo01, lo23), r2s));
}

__device__ __forceinline__ long long comp_of(uint32_t key, uint32_t base) {
    return ((long long)((int32_t)key >> 8) << 32) | (long long)(base + (key & 0xFFu));
}

// coarse index (tile / row block id) of the two running candidates after a tile: (o0, o1) the keys before it
__device__ __forceinline__ void track(int32_t m0, int32_t m1, int32_t o0, int32_t o1, uint32_t t, uint32_t& t0, uint32_t& t1) {
    if (m0 != o0) { t1 = (m1 == o0) ? t0 : t; t0 = t; }
    else if (m1 != o1) t1 = t;
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
// query's running candidates (keys); bar: scores >= bar are not inserted; skip: the smallest score passed over. Warp-uniform control flow.
template <int kIdx0, int kOff>
__device__ __forceinline__ void scan8(const uint32_t (&v)[16], uint32_t mask, float r2s, int32_t& c0, int32_t& c1, int32_t& bar, int32_t& skip) {
    const int32_t x0 = (int32_t)v[kOff + 0], x1 = (int32_t)v[kOff + 1], x2 = (int32_t)v[kOff + 2], x3 = (int32_t)v[kOff + 3];
    const int32_t x4 = (int32_t)v[kOff + 4], x5 = (int32_t)v[kOff + 5], x6 = (int32_t)v[kOff + 6], x7 = (int32_t)v[kOff + 7];
    const int32_t mn = min(min(min(x0, x1), x2), min(min(min(x3, x4), x5), min(x6, x7)));      // four instructions (three 3-input minima)
    if (__any_sync(0xffffffffu, mn < bar)) {
        // (measured alternatives, all slower or equal: a sort-2 / merge tree instead of this chain; narrowing to the sub-groups of the
        //  min tree first; one vote per sixteen scores -- the sweep is bound by dependent-issue latency with four warps per scheduler)
        insert(key_of<kIdx0 + 0>(v[kOff + 0], mask), c0, c1); insert(key_of<kIdx0 + 1>(v[kOff + 1], mask), c0, c1);
        insert(key_of<kIdx0 + 2>(v[kOff + 2], mask), c0, c1); insert(key_of<kIdx0 + 3>(v[kOff + 3], mask), c0, c1);
        insert(key_of<kIdx0 + 4>(v[kOff + 4], mask), c0, c1); insert(key_of<kIdx0 + 5>(v[kOff + 5], mask), c0, c1);
        insert(key_of<kIdx0 + 6>(v[kOff + 6], mask), c0, c1); insert(key_of<kIdx0 + 7>(v[kOff + 7], mask), c0, c1);
        bar = min(bar, bar_of2(c0, c1, r2s));
    } else {
        skip = min(skip, mn);                             // all eight passed over (by every lane)
    }
}
template <int kIdx0>
__device__ __forceinline__ void scan16(const uint32_t (&v)[16], uint32_t mask, float r2s, int32_t& c0, int32_t& c1, int32_t& bar, int32_t& skip) {
    scan8<kIdx0, 0>(v, mask, r2s, c0, c1, bar, skip);
    scan8<kIdx0 + 8, 8>(v, mask, r2s, c0, c1, bar, skip);
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
// q < q_end) has its two candidates in state[(s_first + k * step) * 2 ..] and its smallest passed-over score in skip[..]. Sixteen queries per warp pass: every lane accumulates
// its four dimensions of all sixteen (query, candidate) distances, one multi-value butterfly per candidate sums them, and lane
// 2r (and 2r + 1) then runs the certainty logic of tcm::rerank_ratio_checked for query r -- same arithmetic, same bits, but the
// scalar part once per sixteen queries instead of once per query, and twelve gathered rows in flight per warp instead of three
// (one query at a time was measured at ~2,400 cycles per query, a fifth of the kernel). The only difference: the query's own norm
// in the error bound is replaced by its image's maximum norm (still a bound). Not inlined: its registers must not leak into the
// scan loop's allocation.
__device__ __noinline__ void rerank_range(const long long* state, const int32_t* skip, uint32_t s_first, uint32_t q_first, uint32_t step, uint32_t n_iter, uint32_t q_end,
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
        // The decision (tcm::rerank_ratio_checked, extended by the passed-over scores of the pruned sweep). Every train row other than
        // the two candidates is either (a) a row the candidates pushed out or kept out: its score was >= score(j1), so its squared
        // distance is >= X = d1^2 - 4E; or (b) a row that was passed over: its score was >= S (the smallest such score of this query),
        // so its distance is >= ds_lo, and the row that had score S is at most ds_hi away. With L a lower bound on the distance of
        // every non-candidate and U an upper bound on the distance of SOME row other than j0:
        //   j0 is the match for certain    if d0 < L and d0 / min(d1, L) < ratio
        //   there is no match for certain  if d0 / min(d1, U) >= ratio and (L >= d0 or L / d0 >= ratio)   (second case: a non-candidate is the best)
        //   otherwise                      exact scan over all train rows
        const int32_t skey = valid ? skip[si] : kEmptyKeyTc;
        const float S = __int_as_float(skey);
        const bool has_skip = valid && S < 1.0e29f;       // (>= 1e29: only padding rows were passed over)
        uint32_t result = EACHAM_NONE;
        bool certain = true;
        if (v0 && (v1 || has_skip)) {                    // fewer than two neighbours: the reference is UB, rejected
            float na = 0.f, nb = 0.f, delta = 0.f;
            if (!exact_inputs) { na = own_max_norm * 1.004f; nb = other_max_norm * 1.004f; delta = 1.953125e-3f * (na + nb) * 1.01f; }
            float L = kInf, U = kInf;
            if (v1) {
                float E = 0.5f * d1 * d1 * 3.0517578125e-5f;                               // E_trunc
                if (!exact_inputs) E += 0.5f * delta * (2.f * d1 + delta) + 1.52587890625e-5f * 0.5f * (na * na + nb * nb);
                const float X = d1 * d1 - 4.f * E * 1.001f;
                L = X > 0.f ? sqrtf(X) * 0.999999f : 0.f;
            }
            if (has_skip) {
                const float ds = sqrtf(2.f * fmaxf(S, 0.f));
                const float Et = fabsf(S) * 3.0517578125e-5f * 1.01f;                      // slack of one key step on the upper side (S itself is a raw score, not a truncated key)
                const float En = exact_inputs ? 0.f : (0.5f * delta * (2.f * ds + 3.f * delta) + 1.52587890625e-5f * 0.5f * (na * na + nb * nb)) * 1.01f;
                L = fminf(L, sqrtf(2.f * fmaxf(S - En, 0.f)) * 0.999999f);
                U = sqrtf(2.f * fmaxf(S + Et + En, 0.f)) * 1.000001f;
            }
            const float lo2 = fminf(d1, L);
            const bool no_match = U < d1 ? (double)(d0 / U) >= ratio * 1.000001 : !((double)__fdiv_rn(d0, d1) < ratio);
            certain = false;
            if (d0 < L && (double)(d0 / lo2 * 1.000001f) < ratio) { certain = true; result = j0; }
            else if (no_match && (L >= d0 || (double)L >= ratio * (double)d0 * 1.000001)) certain = true;
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

// The sweep of one resident row block over all tiles of the second image, for one epilogue warp. Leaves the part's keys in S.share1 and
// its tile ids in S.tiles1; returns the advanced MMA step counter. Not inlined: the loop gets its own register allocation (inlined into
// the kernel it ran at the 96-register cap with spills inside the tile loop).
__device__ __noinline__ uint32_t scan_block(SmemSift& S, uint32_t tmem, int q, int cp, int lane, int e, uint32_t blk0, uint32_t nh, uint32_t nbt, uint32_t N, uint32_t M,
                                            uint32_t ab, bool both, float r2s, uint2* colkeys, uint2* colts, uint32_t step_it) {
    uint32_t mask = 0xFFFFFF00u;
    asm volatile("" : "+r"(mask));                       // keep the mask in a register: key = (v & mask) | immediate is one LOP3
    // this part's running candidates of its two resident rows (keys: score bits | column within the part), and their tiles
    int32_t m0[2] = {kEmptyKeyTc, kEmptyKeyTc}, m1[2] = {kEmptyKeyTc, kEmptyKeyTc};
    uint32_t t0[2] = {0xFFFFu, 0xFFFFu}, t1[2] = {0xFFFFu, 0xFFFFu};
    int32_t skip1[2] = {kEmptyKeyTc, kEmptyKeyTc};       // smallest score passed over (bar_of2 / bar_of8)
    const int r0 = q * 32 + lane, r1 = 128 + q * 32 + lane;
    const bool rv0 = blk0 * 128 + r0 < N, rv1 = nh == 2 && blk0 * 128 + r1 < N;
    S.share1[cp][r0] = make_uint2((uint32_t)kEmptyKeyTc, (uint32_t)kEmptyKeyTc);
    S.share1[cp][r1] = make_uint2((uint32_t)kEmptyKeyTc, (uint32_t)kEmptyKeyTc);
    epi_bar();                                   // nobody reads the previous block's keys as bars (also orders the column state init)
    int32_t bar1[2] = {rv0 ? kEmptyKeyTc : INT32_MIN, rv1 ? kEmptyKeyTc : INT32_MIN};      // D1 bars, carried from tile to tile
    for (uint32_t bt = 0; bt < nbt; ++bt) {
        const uint32_t jrow = bt * 128 + q * 32 + lane;
        // this tile's rows of the second image: their four parts' keys and this part's block ids, requested ahead of D2 (L2 latency)
        uint4 ka = make_uint4(0u, 0u, 0u, 0u), kb = ka;
        uint2 ts = make_uint2(0u, 0u);                   // this part's (row-block ids, smallest passed-over score) of the row
        if (both) {
            ka = reinterpret_cast<const uint4*>(colkeys)[2 * (size_t)jrow];
            kb = reinterpret_cast<const uint4*>(colkeys)[2 * (size_t)jrow + 1];
            ts = colts[4 * (size_t)jrow + cp];
        }
        // ---------------- D1: each lane owns one row of each resident half; 32 columns per warp ----------------
        {
            const uint32_t reg = step_it & 1;
            // bars from all four parts' published keys, every fourth tile (unsynchronised: stale only means a higher bar); in between
            // the bar only follows this part's own insertions. Padding rows never take part.
            if ((bt & (kBarEvery - 1u)) == 1u) {
                if (rv0) bar1[0] = min(bar1[0], bar_of8(S.share1[0][r0], S.share1[1][r0], S.share1[2][r0], S.share1[3][r0], r2s));
                if (rv1) bar1[1] = min(bar1[1], bar_of8(S.share1[0][r1], S.share1[1][r1], S.share1[2][r1], S.share1[3][r1], r2s));
            }
            const int32_t o00 = m0[0], o10 = m1[0], o01 = m0[1], o11 = m1[1];
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
                scan16<0>(v, mask, r2s, m0[0], m1[0], bar1[0], skip1[0]);
                tc::tmem_ld_wait();
                if (nh == 2) tc::tmem_ld16(taddr + 128, v);
                scan16<16>(w, mask, r2s, m0[0], m1[0], bar1[0], skip1[0]);
                if (nh == 2) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld16(taddr + 144, w);
                    scan16<0>(v, mask, r2s, m0[1], m1[1], bar1[1], skip1[1]);
                    tc::tmem_ld_wait();
                    scan16<16>(w, mask, r2s, m0[1], m1[1], bar1[1], skip1[1]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&S.acc_empty[reg]);
            if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 2);
            ++step_it;
            if (__any_sync(0xffffffffu, m0[0] != o00 || m1[0] != o10 || m0[1] != o01 || m1[1] != o11)) {     // some row of the warp changed
                track(m0[0], m1[0], o00, o10, bt, t0[0], t1[0]);
                track(m0[1], m1[1], o01, o11, bt, t0[1], t1[1]);
                S.share1[cp][r0] = make_uint2((uint32_t)m0[0], (uint32_t)m1[0]);
                if (nh == 2) S.share1[cp][r1] = make_uint2((uint32_t)m0[1], (uint32_t)m1[1]);
            }
        }
        if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, (step_it - 1) >> 1, 3);
        if (!both) continue;
        // ---------------- D2: each lane owns one row of the streamed tile; 64 columns (resident rows) per warp ----------------
        {
            const uint32_t reg = step_it & 1;
            const uint32_t ncols = nh * 128;
            const uint2 own = cp == 0 ? make_uint2(ka.x, ka.y) : cp == 1 ? make_uint2(ka.z, ka.w) : cp == 2 ? make_uint2(kb.x, kb.y) : make_uint2(kb.z, kb.w);
            int32_t c0 = (int32_t)own.x, c1 = (int32_t)own.y, skip = (int32_t)ts.y;
            int32_t bar = jrow < M ? bar_of8(make_uint2(ka.x, ka.y), make_uint2(ka.z, ka.w), make_uint2(kb.x, kb.y), make_uint2(kb.z, kb.w), r2s) : INT32_MIN;
            tc::mbar_wait(&S.acc_full[reg], (step_it >> 1) & 1);
            tc::tc_fence_after();
            if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 4);
            const uint32_t taddr = tmem + reg * 256 + ((uint32_t)(q * 32) << 16) + cp * 64;
            if ((uint32_t)cp * 64 < ncols) {
                uint32_t v[16], w[16];                      // cp * 64 + 64 <= ncols too: ncols is 128 or 256
                tc::tmem_ld16(taddr, v);
                tc::tmem_ld_wait();
                tc::tmem_ld16(taddr + 16, w);
                scan16<0>(v, mask, r2s, c0, c1, bar, skip);
                tc::tmem_ld_wait();
                tc::tmem_ld16(taddr + 32, v);
                scan16<16>(w, mask, r2s, c0, c1, bar, skip);
                tc::tmem_ld_wait();
                tc::tmem_ld16(taddr + 48, w);
                scan16<32>(v, mask, r2s, c0, c1, bar, skip);
                tc::tmem_ld_wait();
                scan16<48>(w, mask, r2s, c0, c1, bar, skip);
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&S.acc_empty[reg]);
            if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, step_it >> 1, 5);
            ++step_it;
            if (c0 != (int32_t)own.x || c1 != (int32_t)own.y || skip != (int32_t)ts.y) {      // all three are running minima: they change O(log n) times
                uint32_t u0 = ts.x & 0xFFFFu, u1 = ts.x >> 16;
                track(c0, c1, (int32_t)own.x, (int32_t)own.y, ab, u0, u1);
                colkeys[4 * (size_t)jrow + cp] = make_uint2((uint32_t)c0, (uint32_t)c1);
                colts[4 * (size_t)jrow + cp] = make_uint2(u0 | (u1 << 16), (uint32_t)skip);
            }
            if (lane == 0 && (e == 0 || e == 5)) SIFT_TRACE(e == 5, (step_it - 1) >> 1, 6);
        }
    }
    S.tiles1[cp][r0] = make_uint2(t0[0] | (t1[0] << 16), (uint32_t)skip1[0]);
    S.tiles1[cp][r1] = make_uint2(t0[1] | (t1[1] << 16), (uint32_t)skip1[1]);
    return step_it;
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
        // descriptors = a base built once + a constant per K-chunk (the address field counts 16-byte units; shared-memory addresses stay
        // below 2^18, so the 14-bit field never carries): rebuilding them from addresses per tile cost the issuing warp ~200 issue slots
        const uint32_t hi_a = (uint32_t)(dbase_a >> 32), hi_b = (uint32_t)(dbase_b >> 32);
        const uint32_t a_lo = (uint32_t)tc::smem_desc(dbase_a, tc::smem_u32(S.a));
        const uint32_t b_lo0 = (uint32_t)tc::smem_desc(dbase_b, tc::smem_u32(S.b[0])), b_lo1 = (uint32_t)tc::smem_desc(dbase_b, tc::smem_u32(S.b[1])),
                       b_lo2 = (uint32_t)tc::smem_desc(dbase_b, tc::smem_u32(S.b[2]));
        static_assert(kBStages == 3, "three B stages assumed by the descriptor selection");
        constexpr uint32_t kStepA = kAChunkStride >> 4, kStepB = tc::kChunkStride >> 4, kHalfA = tc::kChunkStride >> 4;      // per K-chunk / per resident half
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
                    const uint32_t b_lo = st == 0 ? b_lo0 : st == 1 ? b_lo1 : b_lo2;
                    tc::mbar_wait(&S.b_full[st], (b_it / kBStages) & 1);
                    SIFT_TRACE(2, b_it, 0);
                    {   // D1: rows of `first` in the TMEM lanes
                        const uint32_t reg = step_it & 1;
                        tc::mbar_wait(&S.acc_empty[reg], ((step_it >> 1) & 1) ^ 1);
                        tc::tc_fence_after();
                        SIFT_TRACE(2, b_it, 1);
                        if (tc::elect_one()) {
#pragma unroll
                            for (uint32_t h = 0; h < 2; ++h) {
                                if (h < nh) {
                                    const uint32_t d = tmem + reg * 256 + h * 128;
#pragma unroll
                                    for (int ks = 0; ks < tc::kKSteps; ++ks) {
                                        const int ca = ks < 8 ? 2 * ks : kAugChunkM, cb = ks < 8 ? 2 * ks : kAugChunkN;
                                        tc::mma_bf16(d, tc::desc_from(a_lo + ca * kStepA + h * kHalfA, hi_a), tc::desc_from(b_lo + cb * kStepB, hi_b), idesc128, ks > 0);
                                    }
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
                                tc::mma_bf16(d, tc::desc_from(b_lo + cm * kStepB, hi_b), tc::desc_from(a_lo + cn * kStepA, hi_a), nh == 2 ? idesc256 : idesc128, ks > 0);
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
        uint2* colkeys = reinterpret_cast<uint2*>(my_scratch);                                      // [cols_cap][4 parts]: running (best, second) keys
        uint2* colts = reinterpret_cast<uint2*>(my_scratch + (size_t)p.cols_cap * 32);              // [cols_cap][4 parts]: (row-block ids, smallest passed-over score)
        long long* colstate = reinterpret_cast<long long*>(my_scratch + (size_t)p.cols_cap * 64);   // [cols_cap][2]: merged composites (end of pair)
        int32_t* colskip = reinterpret_cast<int32_t*>(my_scratch + (size_t)p.cols_cap * 80);        // [cols_cap]: merged smallest passed-over score
        uint32_t* m12 = reinterpret_cast<uint32_t*>(my_scratch + (size_t)p.cols_cap * 96);          // [rows_cap]
        // scores >= r2s * best are passed over (see low_bar). The margin covers the scorer's error, so that a passed-over score is
        // practically never the best of a match (that case is still handled: exact scan). The debug view wants the true second
        // neighbour, so it switches the rule off.
        const float r2s = (p.dbg_idx12 != nullptr || p.dbg_idx21 != nullptr) ? 3.0e38f : (float)(p.ratio * p.ratio) * 1.0625f;
        uint32_t* m21 = m12 + p.rows_cap;                                                           // [cols_cap]
        uint32_t step_it = 0;
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
                for (uint32_t j = et; j < nbt * 128; j += kEpiThreads) {
                    const uint4 ek = make_uint4((uint32_t)kEmptyKeyTc, (uint32_t)kEmptyKeyTc, (uint32_t)kEmptyKeyTc, (uint32_t)kEmptyKeyTc);
                    reinterpret_cast<uint4*>(colkeys)[2 * (size_t)j] = ek;
                    reinterpret_cast<uint4*>(colkeys)[2 * (size_t)j + 1] = ek;
                    const uint4 et4 = make_uint4(0xFFFFFFFFu, (uint32_t)kEmptyKeyTc, 0xFFFFFFFFu, (uint32_t)kEmptyKeyTc);
                    reinterpret_cast<uint4*>(colts)[2 * (size_t)j] = et4;
                    reinterpret_cast<uint4*>(colts)[2 * (size_t)j + 1] = et4;
                }

            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab) {
                const uint32_t blk0 = p.single_dir ? pi : ab * 2;
                const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2);
                step_it = scan_block(S, tmem, q, cp, lane, e, blk0, nh, nbt, N, M, ab, both, r2s, colkeys, colts, step_it);
                // ---- rows of this block are complete: merge the four parts, re-rank exactly, ratio test ----
                epi_bar();
                if (et < kABlockRows) {
                    long long g0 = kEmptyComp, g1 = kEmptyComp;
                    int32_t sk = kEmptyKeyTc;
#pragma unroll
                    for (int c = 0; c < kColParts; ++c) {
                        const uint2 k = S.share1[c][et];
                        const uint2 tl = S.tiles1[c][et];
                        tcm::comp_merge(comp_of(k.x, (tl.x & 0xFFFFu) * 128 + c * 32), comp_of(k.y, (tl.x >> 16) * 128 + c * 32), g0, g1);
                        sk = min(sk, (int32_t)tl.y);
                    }
                    S.rowstate[et][0] = g0; S.rowstate[et][1] = g1;
                    S.rowskip[et] = sk;
                }
                epi_bar();
                rerank_range(&S.rowstate[0][0], S.rowskip, e * (kABlockRows / kEpiWarps), blk0 * 128 + e * (kABlockRows / kEpiWarps), 1, kABlockRows / kEpiWarps,
                             min(N, (blk0 + nh) * 128), Af, Bf, M, p.ratio, lane, both_exact, __uint_as_float(A.max_norm_bits), __uint_as_float(B.max_norm_bits), p.exact_fallbacks,
                             p.dbg_idx12, p.dbg_dist12, p.single_dir ? p.single_out : m12);
            }

            if (p.single_dir) continue;                       // one direction only
            // ---- rows of `second`: merge the four parts, re-rank, ratio -> m21 ----
            __threadfence_block();
            epi_bar();
            for (uint32_t j = et; j < M; j += kEpiThreads) {
                const uint4 ka = reinterpret_cast<const uint4*>(colkeys)[2 * (size_t)j], kb = reinterpret_cast<const uint4*>(colkeys)[2 * (size_t)j + 1];
                const uint4 ta = reinterpret_cast<const uint4*>(colts)[2 * (size_t)j], tb = reinterpret_cast<const uint4*>(colts)[2 * (size_t)j + 1];
                long long g0 = kEmptyComp, g1 = kEmptyComp;
                tcm::comp_merge(comp_of(ka.x, (ta.x & 0xFFFFu) * 256), comp_of(ka.y, (ta.x >> 16) * 256), g0, g1);
                tcm::comp_merge(comp_of(ka.z, (ta.z & 0xFFFFu) * 256 + 64), comp_of(ka.w, (ta.z >> 16) * 256 + 64), g0, g1);
                tcm::comp_merge(comp_of(kb.x, (tb.x & 0xFFFFu) * 256 + 128), comp_of(kb.y, (tb.x >> 16) * 256 + 128), g0, g1);
                tcm::comp_merge(comp_of(kb.z, (tb.z & 0xFFFFu) * 256 + 192), comp_of(kb.w, (tb.z >> 16) * 256 + 192), g0, g1);
                colstate[2 * (size_t)j] = g0; colstate[2 * (size_t)j + 1] = g1;
                colskip[j] = min(min((int32_t)ta.y, (int32_t)ta.w), min((int32_t)tb.y, (int32_t)tb.w));
            }
            __threadfence_block();
            epi_bar();
            rerank_range(colstate, colskip, e, e, kEpiWarps, (M + kEpiWarps - 1 - e) / kEpiWarps, M, Bf, Af, N, p.ratio, lane, both_exact, __uint_as_float(B.max_norm_bits), __uint_as_float(A.max_norm_bits),
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
