// tc_orb_kernels.cuh -- the default engine for batched 256-bit ORB pairs on sm_100a: bits through the tensor cores with
// F16 accumulators and an all-packed (two scores per register) top-2 epilogue.
//
// Replaces, for CV_8U 32-byte descriptors, per unordered image pair:
//   knnMatch(k=2) both directions   /root/reference/modules/base/features/FeatureMatcherFlann.cpp:17  (NORM_HAMMING, exact)
//   ratio test                      /root/reference/modules/base/features/FeatureMatcherFlann.cpp:21-27
//   gates + mutual filter           /root/reference/apps/sfm/main.cpp:111-146
//   distance                        /root/reference/modules/base/tools/Tools3d.h:46-63
//
// Scores: each descriptor bit is one e4m3 element (0.0 / 1.0, tc_match_kernels.cuh: orb_tc_prep_kernel); with negate-A and
// the popcount augmentation the kind::f8f6f4 MMA yields D = hamming(a, b) / 2. Every partial sum is a multiple of 1/2 with
// |x| <= 600, so the F16 accumulator is EXACT (tools/tc_peak_microbench.cu checks this on hardware) and tcgen05.ld
// .pack::16b hands the epilogue two scores per register (register i = column 2i | column 2i+1 << 16).
//
// Epilogue, per 256 x 128 tile and per warp (32 TMEM lanes x 2 row halves x 32 columns = 2,048 scores in 32 registers):
//   values only -- non-negative f16 bit patterns order like u16, so VIMNMX.U16x2 / VIMNMX3.U16x2 work on them directly;
//   rows (thread-local): the 16 registers of a half are reduced to a packed (best, second best) by a sort-2 / merge tree and
//     merged into the running state; the two 16-bit lanes are the even / odd columns of the warp's 32-column part;
//   columns (cross-lane): sort-2 across the two row halves, a transpose through a per-warp shared-memory buffer, a merge tree
//     over 16 lanes, one shuffle step, then a packed merge into the warp's OWN (per lane quadrant) column state in the per-CTA
//     L2 scratch -- no barrier between warps inside the tile loop; the four quadrant states are merged once per pair;
//   sort-2 can run on the FMA pipe instead of the ALU pipe:  d = relu(a - b) (HFMA2.RELU), lo = a - d, hi = b + d  (exact
//     on multiples of 1/2 below 1024), which takes a third of the epilogue off the binding pipe.
// Indices are not carried through the hot loop at all. A match exists only where the best distance is a strict unique minimum
// (5 * d0 < 4 * d1), so it is enough to remember WHERE the running best last decreased -- the 128-column tile for rows (tracked
// with three packed f16 ops per tile), the (row block, quadrant) for columns -- and to recover the exact index afterwards by
// re-evaluating XOR + POPC on those 16 / 64 candidates, for ratio-passing rows and columns only.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"
#include "tc_common.cuh"
#include "tc_match_kernels.cuh"

#ifndef EACHAM_EXP
#define EACHAM_EXP 0      // bit 0: skip rows, bit 1: skip the column transpose + merge, bit 2: skip the column state update (timing experiments only)
#endif

namespace eacham {
namespace tco {

#if EACHAM_EXP & 8
__device__ long long g_trace[2][512][4];       // [role: 0 epilogue warp 0, 1 MMA issuer][tile][event] (CTA 0 only)
#define EACHAM_TRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 512) g_trace[role][tile][ev] = clock64(); } while (0)
#else
#define EACHAM_TRACE(role, tile, ev) do { } while (0)
#endif

using tcm::ImageDescTc;
using tcm::PairParamsTc;
using tcm::kEpiWarps;
using tcm::kEpiThreads;
using tcm::kColParts;
using tcm::kColsPerWarp;
using tcm::kABlockRows;
using tcm::epi_bar;

constexpr int kThreads = 64 + kEpiThreads;
constexpr int kBStages = 2;
constexpr int kAccStages = 2;
constexpr int kRing = 4;                          // pair ids in flight between the producer and the other roles
constexpr uint32_t kSentX2 = 0x63D063D0u;         // f16 1000.0 | 1000.0: "nothing seen yet" (real D <= 128, padding 300..600)
constexpr uint32_t kInvalid = 0x5CB0u;            // f16 300.0: a value >= this is a padding row / column, not a neighbour
constexpr uint32_t kNoPair = 0xFFFFFFFFu;
constexpr int kXposeStride = 18;                  // uint2 per lane row: 16 (lo, hi) pairs + 2 padding (144 B, conflict-free)

struct SmemOrb {
    uint8_t a[2][tc::kAOperandBytes];
    uint8_t b[kBStages][tc::kAOperandBytes];
    union {
        uint2 xpose[kEpiWarps][32 * kXposeStride];            // per-warp transpose buffer of the column path
        uint4 rowmerge[kColParts - 1][kABlockRows];           // end of a row block: row state of column parts 1..3
    } u;
    uint64_t b_full[kBStages], b_empty[kBStages], a_full, a_empty, acc_full[kAccStages], acc_empty[kAccStages];
    uint64_t ring_full[kRing], ring_empty[kRing];
    uint32_t ring[kRing];
    uint32_t tmem_slot;
    uint32_t red[2 * kEpiWarps + 8];
    unsigned long long base;
};

// ---- packed 16-bit primitives --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t minu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t maxu2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t min3u2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

template <bool kFma>
__device__ __forceinline__ void sort2(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
    if (kFma) {
        uint32_t d;
        const uint32_t neg1 = 0xBC00BC00u;
        asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(neg1), "r"(a));      // relu(a - b)
        asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(d));
        asm("add.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(d));
    } else {
        lo = minu2(a, b);
        hi = maxu2(a, b);
    }
}
// top-2 of the union of two sorted pairs (a0 <= a1, b0 <= b1), per 16-bit lane: 3 ALU-pipe instructions
__device__ __forceinline__ void merge2(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t& lo, uint32_t& hi) {
    lo = minu2(a0, b0);
    hi = min3u2(maxu2(a0, b0), a1, b1);
}

// packed top-2 of 16 registers (32 scores of one row: 16 even + 16 odd columns)
template <bool kFma>
__device__ __forceinline__ void reduce16(const uint32_t (&v)[16], uint32_t& lo, uint32_t& hi) {
    uint32_t l[8], h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sort2<kFma>(v[2 * i], v[2 * i + 1], l[i], h[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) merge2(l[2 * i], h[2 * i], l[2 * i + 1], h[2 * i + 1], l[i], h[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) merge2(l[2 * i], h[2 * i], l[2 * i + 1], h[2 * i + 1], l[i], h[i]);
    merge2(l[0], h[0], l[1], h[1], lo, hi);
}

__device__ __forceinline__ void quad_bar(int q) { asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory"); }

__device__ __forceinline__ void tmem_ld16_pack(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint32_t f16_bits_to_hamming(uint32_t bits) {       // bits = f16(D), D = hamming / 2 exactly
    return (uint32_t)__float2int_rn(2.f * __half2float(__ushort_as_half((unsigned short)bits)));
}
__device__ __forceinline__ uint32_t hamming256(const uint4 a0, const uint4 a1, const uint4* __restrict__ brow) {
    const uint4 b0 = __ldg(brow), b1 = __ldg(brow + 1);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
// FeatureMatcherFlann.cpp:23: m[0].distance / m[1].distance < 0.8, float / float against a double; d1 of a padding row or
// column means "fewer than two neighbours" (the reference dereferences m[1] unguarded = UB; such queries are rejected here)
__device__ __forceinline__ bool ratio_pass_f16(uint32_t best, uint32_t second, double ratio, uint32_t& ham0) {
    if (second >= kInvalid) return false;
    ham0 = f16_bits_to_hamming(best);
    const float d0 = (float)ham0, d1 = (float)f16_bits_to_hamming(second);
    return (double)__fdiv_rn(d0, d1) < ratio;            // 0/0 -> NaN -> false
}

// =============================================================================================================
// The fused ORB pair kernel: one persistent CTA per SM, image pairs from a dynamic queue. 576 threads:
//   warps 0-15 epilogue (4 per TMEM lane quadrant, 32 columns of every tile each)
//   warp 16   producer: pair queue (atomic counter -> shared-memory ring), 1-D bulk copies of pre-tiled FP8 blocks
//   warp 17   MMA issuer: tcgen05.mma kind::f8f6f4, M = 128, N = 128, K = 32, D format F16; 9 K-steps x 2 row halves per tile
// (the two single-thread roles sit in the HIGHEST warps: the warp scheduler favours higher warp ids, and a starved MMA issuer
// or producer stalls everything else)
// A block of 256 rows of the first image stays in shared memory while the second streams through in 128-column tiles;
// accumulators: 2 stages x 2 halves x 128 columns = all 512 TMEM columns.
// =============================================================================================================
template <bool kFmaSort>
__global__ void __launch_bounds__(kThreads, 1) orb_tc_match_pairs_kernel(const PairParamsTc p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemOrb& S = *reinterpret_cast<SmemOrb*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { tc::mbar_init(&S.b_full[s], 1); tc::mbar_init(&S.b_empty[s], 1); }
        tc::mbar_init(&S.a_full, 1); tc::mbar_init(&S.a_empty, 1);
        for (int s = 0; s < kAccStages; ++s) { tc::mbar_init(&S.acc_full[s], 1); tc::mbar_init(&S.acc_empty[s], kEpiWarps); }
        for (int s = 0; s < kRing; ++s) { tc::mbar_init(&S.ring_full[s], 1); tc::mbar_init(&S.ring_empty[s], 2); }
        tc::fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tc::tmem_alloc(&S.tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = S.tmem_slot;

    if (warp == kEpiWarps) {
        // ===================================== producer =====================================
        // The whole warp walks the loop (warp-uniform control flow); one elected lane touches the queue and issues the copies.
        {
            uint32_t b_it = 0, a_it = 0;
            for (uint32_t k = 0;; ++k) {
                const uint32_t rs = k % kRing;
                tc::mbar_wait(&S.ring_empty[rs], ((k / kRing) & 1) ^ 1);
                uint32_t pi = 0;
                if (lane == 0) {
                    pi = atomicAdd(p.work_counter, 1u);
                    pi = pi < p.n_pairs ? (p.single_dir ? pi : p.order[pi]) : kNoPair;
                    S.ring[rs] = pi;
                    tc::mbar_arrive(&S.ring_full[rs]);
                }
                pi = __shfl_sync(0xffffffffu, pi, 0);
                if (pi == kNoPair) break;
                const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
                const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
                if (A.rows == 0 || B.rows == 0) continue;
                const uint32_t na128 = (A.rows + 127) / 128, nbt = (B.rows + 127) / 128;
                const uint8_t* Ab = p.tc_arena + A.tc_offset;
                const uint8_t* Bb = p.tc_arena + B.tc_offset;
                for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab, ++a_it) {
                    const uint32_t blk0 = p.single_dir ? pi : ab * 2;                     // first 128-row block of this work unit
                    const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2); (void)blk0;
                    tc::mbar_wait(&S.a_empty, (a_it & 1) ^ 1);
                    if (tc::elect_one()) {
                        tc::mbar_expect_tx(&S.a_full, nh * tc::kAOperandBytes);
                        for (uint32_t h = 0; h < nh; ++h)
                            tc::bulk_g2s(S.a[h], Ab + (size_t)(blk0 + h) * tc::kBlockBytes, tc::kAOperandBytes, &S.a_full);
                    }
                    __syncwarp();
                    for (uint32_t bt = 0; bt < nbt; ++bt, ++b_it) {
                        const uint32_t st = b_it % kBStages;
                        tc::mbar_wait(&S.b_empty[st], ((b_it / kBStages) & 1) ^ 1);
                        if (tc::elect_one()) {
                            tc::mbar_expect_tx(&S.b_full[st], tc::kAOperandBytes);
                            const uint8_t* src = Bb + (size_t)bt * tc::kBlockBytes;
                            tc::bulk_g2s(S.b[st], src, tc::kDataBytes, &S.b_full[st]);
                            tc::bulk_g2s(S.b[st] + tc::kDataBytes, src + tc::kDataBytes + tc::kAugBytes, tc::kAugBytes, &S.b_full[st]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===================================== MMA issuer =====================================
        // Whole warp in the loop, one elected lane issues (see tc::elect_one).
        {
            const uint64_t dbase = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
            const uint32_t idesc = tc::make_idesc_e4m3(128, 128, true, /*d_f32=*/false);
            const uint32_t desc_hi = (uint32_t)(dbase >> 32);
            const uint32_t a_lo0 = (uint32_t)tc::smem_desc(dbase, tc::smem_u32(S.a[0])), a_lo1 = (uint32_t)tc::smem_desc(dbase, tc::smem_u32(S.a[1]));
            const uint32_t b_lo0 = (uint32_t)tc::smem_desc(dbase, tc::smem_u32(S.b[0])), b_lo1 = (uint32_t)tc::smem_desc(dbase, tc::smem_u32(S.b[1]));
            constexpr uint32_t kDescStep = (2 * tc::kChunkStride) >> 4;      // one K-step = two 16-byte K-chunks, in the descriptor's 16-byte units
            static_assert(kBStages == 2, "two B stages assumed by the descriptor selection");
            uint32_t b_it = 0, a_it = 0, acc_it = 0;
            for (uint32_t k = 0;; ++k) {
                const uint32_t rs = k % kRing;
                tc::mbar_wait(&S.ring_full[rs], (k / kRing) & 1);
                const uint32_t pi = S.ring[rs];
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&S.ring_empty[rs]);
                if (pi == kNoPair) break;
                const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
                const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
                if (A.rows == 0 || B.rows == 0) continue;
                const uint32_t na128 = (A.rows + 127) / 128, nbt = (B.rows + 127) / 128;
                for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab, ++a_it) {
                    const uint32_t blk0 = p.single_dir ? pi : ab * 2;                     // first 128-row block of this work unit
                    const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2); (void)blk0;
                    tc::mbar_wait(&S.a_full, a_it & 1);
                    for (uint32_t bt = 0; bt < nbt; ++bt, ++b_it, ++acc_it) {
                        const uint32_t st = b_it % kBStages, as = acc_it % kAccStages;
                        if (lane == 0) EACHAM_TRACE(1, acc_it, 0);
                        tc::mbar_wait(&S.b_full[st], (b_it / kBStages) & 1);
                        if (lane == 0) EACHAM_TRACE(1, acc_it, 1);
                        tc::mbar_wait(&S.acc_empty[as], ((acc_it / kAccStages) & 1) ^ 1);
                        if (lane == 0) EACHAM_TRACE(1, acc_it, 2);
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            // descriptors = a base built once per kernel + a constant per K-step (the address field counts 16-byte units and
                            // shared-memory addresses stay below 2^18, so the 14-bit field never carries): the issuing warp shares its scheduler
                            // with four epilogue warps, and rebuilding 36 descriptors per tile from addresses cost ~135 of its issue slots
                            const uint32_t b_lo = st ? b_lo1 : b_lo0;
#pragma unroll
                            for (uint32_t h = 0; h < 2; ++h) {
                                if (h < nh) {
                                    const uint32_t a_lo = h ? a_lo1 : a_lo0;
                                    const uint32_t d = tmem + as * 256 + h * 128;
#pragma unroll
                                    for (int ks = 0; ks < tc::kKSteps; ++ks)
                                        tc::mma_f8(d, tc::desc_from(a_lo + ks * kDescStep, desc_hi), tc::desc_from(b_lo + ks * kDescStep, desc_hi), idesc, ks > 0);
                                }
                            }
                            tc::mma_commit(&S.b_empty[st]);      // B stage reusable once these MMAs have read it
                            tc::mma_commit(&S.acc_full[as]);     // accumulators ready for the epilogue
                        }
                        __syncwarp();
                        if (lane == 0) EACHAM_TRACE(1, acc_it, 3);
                    }
                    if (tc::elect_one()) tc::mma_commit(&S.a_empty);              // A block reusable
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================================== epilogue =====================================
        const int e = warp, q = warp & 3, cp = e >> 2;
        const int et = e * 32 + lane;                         // 0..511 within the epilogue group
        uint8_t* my_scratch = p.scratch + (size_t)blockIdx.x * tcm::tc_scratch_bytes_per_cta(p.rows_cap, p.cols_cap);
        // column state per TMEM lane quadrant and per column PAIR: {packed best, packed second best, packed f16 row block of the
        // last decrease of the best, -}: [4][cols_cap / 2] x 16 B
        uint4* colq = reinterpret_cast<uint4*>(my_scratch);
        const uint32_t colq_stride = p.cols_cap / 2;
        uint32_t* m12 = reinterpret_cast<uint32_t*>(my_scratch + (size_t)p.cols_cap * 32);     // [rows_cap]
        uint32_t* m21 = m12 + p.rows_cap;                                                      // [cols_cap]
        uint2* xp = S.u.xpose[e];
        const uint32_t rbar = (uint32_t)__half_as_ushort(__float2half_rn((float)p.ratio * 1.0625f)) * 0x10001u;      // packed f16: ratio plus a margin (row bars)
        uint32_t acc_it = 0;
        for (uint32_t k = 0;; ++k) {
            const uint32_t rs = k % kRing;
            tc::mbar_wait(&S.ring_full[rs], (k / kRing) & 1);
            const uint32_t pi = S.ring[rs];
            if (pi == kNoPair) break;
            const eacham_pair_t pr = p.pairs[p.single_dir ? 0u : pi];
            const ImageDescTc A = p.images[pr.first], B = p.images[pr.second];
            const uint32_t N = A.rows, M = B.rows;
            const uint32_t na128 = (N + 127) / 128, nbt = (M + 127) / 128;
            if (!p.single_dir)
            for (uint32_t j = et; j < nbt * 64 * 4; j += kEpiThreads) colq[(j / (nbt * 64)) * colq_stride + j % (nbt * 64)] = make_uint4(kSentX2, kSentX2, 0u, 0u);
            epi_bar();                                        // every epilogue thread has read the ring slot
            if (et == 0) tc::mbar_arrive(&S.ring_empty[rs]);
            if (N == 0 || M == 0) {
                if (et == 0) {
                    eacham_pair_result_t r;
                    r.n12 = 0; r.n21 = 0; r.n_mutual = 0; r.flags = (0u < p.min_dir) ? EACHAM_PAIR_GATED : 0u; r.offset = 0; r.count = 0;
                    p.results[pi] = r;
                }
                continue;
            }
            const uint4* Aq = reinterpret_cast<const uint4*>(p.arena + A.offset);
            const uint4* Bq = reinterpret_cast<const uint4*>(p.arena + B.offset);

            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab) {
                const uint32_t blk0 = p.single_dir ? pi : ab * 2;                     // first 128-row block of this work unit
                    const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2); (void)blk0;
                uint32_t m0[2] = {kSentX2, kSentX2}, m1[2] = {kSentX2, kSentX2};       // packed (even | odd columns) top-2 per row half
                uint32_t bar[2] = {kSentX2, kSentX2}, skip[2] = {kSentX2, kSentX2};   // scores >= bar are not merged; skip = the smallest of those
                uint32_t t0[2] = {0u, 0u};                                             // packed f16: tile where m0 last decreased
                uint32_t btx2 = 0u;                                                    // packed f16 (bt | bt)
                const uint32_t abx2 = (uint32_t)__half_as_ushort(__uint2half_rn(ab)) * 0x10001u;          // packed f16 (ab | ab)
                for (uint32_t bt = 0; bt < nbt; ++bt, ++acc_it) {
                    const uint32_t as = acc_it % kAccStages;
                    // this warp's column state of the tile (lane c < 16: column pair c of the part): issue the (L2) load early
                    uint4 cst = make_uint4(kSentX2, kSentX2, 0u, 0u);
                    uint4* cptr = colq + (size_t)q * colq_stride + bt * 64 + cp * (kColsPerWarp / 2) + (lane & 15);
                    const bool cols = !p.single_dir;
                    if (lane < 16 && cols && !(EACHAM_EXP & 4)) cst = *cptr;
                    if (tid == 0) EACHAM_TRACE(0, acc_it, 0);
                    tc::mbar_wait(&S.acc_full[as], (acc_it / kAccStages) & 1);
                    if (tid == 0) EACHAM_TRACE(0, acc_it, 1);
                    tc::tc_fence_after();
                    const uint32_t taddr = tmem + as * 256 + ((uint32_t)(q * 32) << 16) + cp * kColsPerWarp;
                    uint32_t v0[16], v1[16];
                    tmem_ld16_pack(taddr, v0);
                    if (nh == 2) tmem_ld16_pack(taddr + 128, v1);
                    tc::tmem_ld_wait();
                    if (tid == 0) EACHAM_TRACE(0, acc_it, 2);
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&S.acc_empty[as]);       // scores are in registers: TMEM stage free for the next MMA
                    if (nh != 2) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v1[i] = kSentX2;
                    }
                    // ---- columns: sort-2 across the row halves, transpose, merge over lanes ----
                    if (cols && !(EACHAM_EXP & 2))
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        uint32_t l0, h0, l1, h1;
                        sort2<kFmaSort>(v0[i], v1[i], l0, h0);
                        sort2<kFmaSort>(v0[i + 1], v1[i + 1], l1, h1);
                        *reinterpret_cast<uint4*>(&xp[lane * kXposeStride + i]) = make_uint4(l0, h0, l1, h1);
                    }
                    __syncwarp();
                    // ---- rows (independent of the transpose: fills the shared-memory latency) ----
                    // Pruned: what matters for a row is its best (with its tile) and whether anything else comes within 1 / ratio of
                    // it. A score >= ratio * best (x 1.0625) can never be the best of a match -- the old best would be its second at
                    // a ratio >= `ratio` -- so it is not merged; only the running minimum of such scores is kept (`skip`), and the
                    // finalisation takes second = min(second, skip). Both are exact integers (D = hamming / 2), so nothing changes
                    // in the result. One packed min tree + one compare per half; the merge runs only if some lane has a score
                    // under its bar.
                    if (!(EACHAM_EXP & 1)) {
                        uint32_t mn0 = min3u2(min3u2(v0[0], v0[1], v0[2]), min3u2(v0[3], v0[4], v0[5]), min3u2(v0[6], v0[7], v0[8]));
                        mn0 = min3u2(mn0, min3u2(v0[9], v0[10], v0[11]), min3u2(v0[12], v0[13], v0[14]));
                        mn0 = minu2(mn0, v0[15]);
                        uint32_t mn1 = min3u2(min3u2(v1[0], v1[1], v1[2]), min3u2(v1[3], v1[4], v1[5]), min3u2(v1[6], v1[7], v1[8]));
                        mn1 = min3u2(mn1, min3u2(v1[9], v1[10], v1[11]), min3u2(v1[12], v1[13], v1[14]));
                        mn1 = minu2(mn1, v1[15]);
                        const bool hit = minu2(mn0, bar[0]) != bar[0] || minu2(mn1, bar[1]) != bar[1];     // some 16-bit lane below its bar
                        if (__any_sync(0xffffffffu, hit)) {
                            const uint32_t o0 = m0[0], o1 = m0[1];
                            uint32_t lo, hi;
                            reduce16<kFmaSort>(v0, lo, hi);
                            merge2(m0[0], m1[0], lo, hi, m0[0], m1[0]);
                            reduce16<kFmaSort>(v1, lo, hi);
                            merge2(m0[1], m1[1], lo, hi, m0[1], m1[1]);
                            // tile of the running best: t0 = max(t0, changed ? bt : 0), per 16-bit lane, in f16 arithmetic
                            const uint32_t k1024 = 0x64006400u;
                            uint32_t d, c;
                            asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(o0), "r"(m0[0]));
                            asm("mul.f16x2 %0, %1, %2;" : "=r"(c) : "r"(d), "r"(k1024));      // 0 or >= 512 (inf is fine)
                            asm("min.f16x2 %0, %1, %2;" : "=r"(c) : "r"(c), "r"(btx2));
                            asm("max.f16x2 %0, %1, %2;" : "=r"(t0[0]) : "r"(t0[0]), "r"(c));
                            asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(o1), "r"(m0[1]));
                            asm("mul.f16x2 %0, %1, %2;" : "=r"(c) : "r"(d), "r"(k1024));
                            asm("min.f16x2 %0, %1, %2;" : "=r"(c) : "r"(c), "r"(btx2));
                            asm("max.f16x2 %0, %1, %2;" : "=r"(t0[1]) : "r"(t0[1]), "r"(c));
                            uint32_t tl;
                            asm("mul.f16x2 %0, %1, %2;" : "=r"(tl) : "r"(m0[0]), "r"(rbar));
                            bar[0] = minu2(m1[0], tl);
                            asm("mul.f16x2 %0, %1, %2;" : "=r"(tl) : "r"(m0[1]), "r"(rbar));
                            bar[1] = minu2(m1[1], tl);
                        } else {
                            skip[0] = minu2(skip[0], mn0);
                            skip[1] = minu2(skip[1], mn1);
                        }
                    }
                    {
                        const uint32_t one = 0x3C003C00u;
                        asm("add.f16x2 %0, %1, %2;" : "=r"(btx2) : "r"(btx2), "r"(one));
                    }
                    // ---- columns, continued: lane (c, H) merges rows H*16 .. H*16+15 of column pair c ----
                    uint32_t g0 = v0[0], g1 = v1[0];
                    if (cols && !(EACHAM_EXP & 2)) {
                        const int c = lane & 15, H = lane >> 4;
                        const uint2* src = xp + (H * 16) * kXposeStride + c;
                        uint32_t l[8], h[8];
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const uint2 pa = src[(2 * s) * kXposeStride], pb = src[(2 * s + 1) * kXposeStride];
                            merge2(pa.x, pa.y, pb.x, pb.y, l[s], h[s]);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) merge2(l[2 * i], h[2 * i], l[2 * i + 1], h[2 * i + 1], l[i], h[i]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) merge2(l[2 * i], h[2 * i], l[2 * i + 1], h[2 * i + 1], l[i], h[i]);
                        merge2(l[0], h[0], l[1], h[1], g0, g1);
                        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, g0, 16), r1 = __shfl_xor_sync(0xffffffffu, g1, 16);
                        merge2(g0, g1, r0, r1, g0, g1);
                        if (H == 0 && !(EACHAM_EXP & 4)) {       // merge into this quadrant's state of column pair c; no other warp touches it
                            const uint32_t k1024 = 0x64006400u;
                            uint32_t n0, n1, d, cc;
                            merge2(cst.x, cst.y, g0, g1, n0, n1);
                            asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(cst.x), "r"(n0));
                            asm("mul.f16x2 %0, %1, %2;" : "=r"(cc) : "r"(d), "r"(k1024));
                            asm("min.f16x2 %0, %1, %2;" : "=r"(cc) : "r"(cc), "r"(abx2));
                            asm("max.f16x2 %0, %1, %2;" : "=r"(cc) : "r"(cst.z), "r"(cc));
                            *cptr = make_uint4(n0, n1, cc, 0u);
                        }
                    }
                    if (EACHAM_EXP & 6) { if (g0 == 0x12345678u && g1 == 0x9abcdef0u) m12[0] = g0; }      // keep the work alive
                    __syncwarp();                                           // transpose buffer free for the next tile
                    if (tid == 0) EACHAM_TRACE(0, acc_it, 3);
                }
                // ---- rows of this block are complete: merge the four column parts, ratio test, recover the index ----
                // Only the four warps of a lane quadrant (they share the rows) meet: named barrier 2 + q, 128 threads. Their exchange
                // area is the transpose buffer of the quadrant's first warp (all four are past their tile loops).
                quad_bar(q);
                uint4* rm = reinterpret_cast<uint4*>(S.u.xpose[q]);
#pragma unroll
                for (int h = 0; h < 2; ++h) rm[(cp * 2 + h) * 32 + lane] = make_uint4(m0[h], m1[h], t0[h], skip[h]);
                quad_bar(q);
                {
                    // two threads per row: both merge the parts, each checks 8 of the 16 candidate columns
                    const int qt = cp * 32 + lane, rr = qt >> 1, half = qt & 1, h = rr >> 5, l = rr & 31;
                    const uint32_t row = blk0 * 128 + h * 128 + q * 32 + l;
                    uint32_t best = 0xFFFFu, second = 0xFFFFu, wt = 0, wid = 0;
#pragma unroll
                    for (int c = 0; c < kColParts; ++c) {
                        const uint4 o = rm[(c * 2 + h) * 32 + l];
#pragma unroll
                        for (int par = 0; par < 2; ++par) {
                            const uint32_t v = (o.x >> (16 * par)) & 0xFFFFu, w = (o.y >> (16 * par)) & 0xFFFFu;
                            if (v < best) { second = min(second, best); best = v; wt = (o.z >> (16 * par)) & 0xFFFFu; wid = c * 2 + par; }
                            else second = min(second, v);
                            second = min(second, min(w, (o.w >> (16 * par)) & 0xFFFFu));      // second best merged, or the smallest score passed over
                        }
                    }
                    uint32_t found = EACHAM_NONE, ham0 = 0;
                    if (row < N && h < (int)nh && ratio_pass_f16(best, second, p.ratio, ham0)) {
                        // the unique best lives in tile wt, column part wid / 2, column parity wid & 1: 16 candidates
                        const uint32_t tile = (uint32_t)__half2int_rn(__ushort_as_half((unsigned short)wt));
                        const uint32_t base = tile * 128 + (wid >> 1) * kColsPerWarp + (wid & 1) + half * 16;
                        const uint4 a0 = __ldg(Aq + 2 * (size_t)row), a1 = __ldg(Aq + 2 * (size_t)row + 1);
#pragma unroll 4
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t j = base + 2 * i;
                            if (j < M && hamming256(a0, a1, Bq + 2 * (size_t)j) == ham0) found = min(found, j);
                        }
                    }
                    found = min(found, __shfl_xor_sync(0xffffffffu, found, 1));
                    if (half == 0 && row < N && h < (int)nh) (p.single_dir ? p.single_out : m12)[row] = found;
                }
                quad_bar(q);                                                // exchange area read: the transpose buffer is free again
            }

            if (p.single_dir) continue;                       // one direction only: the rows of this block are the whole work item
            // ---- columns: merge the four quadrant states, ratio test, recover the row index among the 64 rows of (row block, quadrant) ----
            epi_bar();
            __threadfence_block();
            for (uint32_t jb = e * 32; jb < M; jb += kEpiThreads) {
                const uint32_t j = jb + lane;
                uint32_t best = 0xFFFFu, second = 0xFFFFu, wab = 0, wq = 0;
                if (j < M) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const uint4 st = colq[(size_t)qq * colq_stride + (j >> 1)];
                        const int sh = 16 * (j & 1);
                        const uint32_t v = (st.x >> sh) & 0xFFFFu, w = (st.y >> sh) & 0xFFFFu;
                        if (v < best) { second = min(second, best); best = v; wab = (st.z >> sh) & 0xFFFFu; wq = qq; }
                        else second = min(second, v);
                        second = min(second, w);
                    }
                }
                uint32_t ham0 = 0;
                const bool pass = j < M && ratio_pass_f16(best, second, p.ratio, ham0);
                // the unique best of a passing column lives in row block wab, lane quadrant wq: 64 candidate rows, checked by the
                // whole warp at once (two per lane)
                const uint32_t base = (uint32_t)__half2int_rn(__ushort_as_half((unsigned short)wab)) * kABlockRows + wq * 32;
                uint32_t found = EACHAM_NONE;
                for (uint32_t mask = __ballot_sync(0xffffffffu, pass); mask; mask &= mask - 1) {
                    const int src = __ffs(mask) - 1;
                    const uint32_t jj = __shfl_sync(0xffffffffu, j, src), bb = __shfl_sync(0xffffffffu, base, src), hh = __shfl_sync(0xffffffffu, ham0, src);
                    const uint4 b0 = __ldg(Bq + 2 * (size_t)jj), b1 = __ldg(Bq + 2 * (size_t)jj + 1);
                    const uint32_t r0 = bb + lane, r1 = bb + 128 + lane;
                    const bool e0 = r0 < N && hamming256(b0, b1, Aq + 2 * (size_t)r0) == hh;
                    const bool e1 = r1 < N && hamming256(b0, b1, Aq + 2 * (size_t)r1) == hh;
                    const uint32_t k0 = __ballot_sync(0xffffffffu, e0), k1 = __ballot_sync(0xffffffffu, e1);
                    const uint32_t r = k0 ? bb + (__ffs(k0) - 1) : (k1 ? bb + 128 + (__ffs(k1) - 1) : EACHAM_NONE);
                    if (lane == src) found = r;
                }
                if (j < M) m21[j] = found;
            }
            __threadfence_block();
            epi_bar();

            tcm::gates_mutual_compact(S, p, pi, N, M, m12, m21, et, e, lane);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) tc::tmem_dealloc(tmem, 512);
#if EACHAM_EXP & 8
    if (blockIdx.x == 0 && tid == 0) {
        const long long t0 = g_trace[1][64][0];
        for (int t = 64; t < 96; ++t)
            printf("tile %d  mma: wait_b %lld b_full %lld acc_empty %lld issued %lld | epi: top %lld acc_full %lld ld_done %lld end %lld\n", t,
                   g_trace[1][t][0] - t0, g_trace[1][t][1] - t0, g_trace[1][t][2] - t0, g_trace[1][t][3] - t0,
                   g_trace[0][t][0] - t0, g_trace[0][t][1] - t0, g_trace[0][t][2] - t0, g_trace[0][t][3] - t0);
    }
#endif
}

}  // namespace tco
}  // namespace eacham
