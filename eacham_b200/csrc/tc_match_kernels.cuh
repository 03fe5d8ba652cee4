// tc_match_kernels.cuh -- the tensor-core matching engine (tcgen05 / TMEM, sm_100a) for both descriptor kinds:
// 128-d float SIFT (bf16 scoring + exact FP32 re-rank) and 256-bit ORB (bits as FP8 0/1, exact).
//
// Replaces, for N x 128 CV_32F descriptors, per unordered image pair:
//   knnMatch(k=2) both directions   /root/reference/modules/base/features/FeatureMatcherFlann.cpp:17  (exact NORM_L2 semantics)
//   ratio test                      /root/reference/modules/base/features/FeatureMatcherFlann.cpp:21-27
//   gates + mutual filter           /root/reference/apps/sfm/main.cpp:111-146
// Candidate scoring runs as a bf16 GEMM with fp32 accumulation in tensor memory; the two best candidates per row and
// per column are then re-ranked with exact FP32 arithmetic (sqrtf(sum (a-b)^2), float accumulator) so that the
// distances the ratio test sees are the reference's.
// Exactness of the MATCH SET (see rerank_ratio_checked): the scorer's candidates minimise an approximate key whose distance
// from the true |a-b|^2/2 is bounded (bf16 rounding of the inputs -- zero for integer-valued rows, which is what cv::SIFT emits --
// plus the 2^-15 relative key truncation). From that bound every query gets an interval for the reference's ratio; only when
// the interval straddles the threshold (or the best itself could be a non-candidate) is the answer not yet certain, and those
// queries are re-done by an exact FP32 scan over ALL train rows. So the match sets equal the exact matcher's for any float
// input; there is no epsilon left in the result, only in how many queries take the slow path. See DESIGN.md "SIFT kernel".
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"
#include "tc_common.cuh"

namespace eacham {
namespace tcm {

// x = hi + mid + lo exactly (24-bit significand -> 3 x 8 bits)
__device__ __forceinline__ void split3(float x, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);
    lo = __float2bfloat16_rn(r2);
}

// fp32 rows [rows][128] -> pre-tiled bf16 blocks (tc_common.cuh layout). One warp per row; grid covers
// n_blocks * 128 rows (rows >= `rows` are padding: zero data, norm 1e30 so they never win a minimum).
__device__ __forceinline__ void sift_prep_row(const float* __restrict__ src, uint32_t rows, uint8_t* __restrict__ dst, uint32_t row, int lane,
                                              uint32_t* max_norm_bits = nullptr, uint32_t* bf16_exact = nullptr) {
    uint8_t* blk = dst + (size_t)(row / tc::kBlockRows) * tc::kBlockBytes;
    const uint32_t r = row % tc::kBlockRows;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * 128) + lane);
    const __nv_bfloat16 b0 = __float2bfloat16_rn(v.x), b1 = __float2bfloat16_rn(v.y), b2 = __float2bfloat16_rn(v.z),
                        b3 = __float2bfloat16_rn(v.w);
    const float f0 = __bfloat162float(b0), f1 = __bfloat162float(b1), f2 = __bfloat162float(b2), f3 = __bfloat162float(b3);
    float n = f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;       // norm of the ROUNDED row: D = |a-b|^2/2 stays consistent
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (max_norm_bits != nullptr && row < rows) {           // per-image facts the matcher's error bound needs
        const bool exact = __all_sync(0xffffffffu, f0 == v.x && f1 == v.y && f2 == v.z && f3 == v.w);
        if (lane == 0) {
            atomicMax(max_norm_bits, __float_as_uint(sqrtf(n)));          // n >= 0: the bit pattern orders like the value
            if (!exact) atomicAnd(bf16_exact, 0u);
        }
    }
    // lane l holds dims 4l..4l+3 -> chunk l/2, bytes (l&1)*8 .. +8 of the row's 16-byte slot
    __nv_bfloat162 p0 = __halves2bfloat162(b0, b1), p1 = __halves2bfloat162(b2, b3);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&p0);
    packed.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(blk + (size_t)(lane >> 1) * tc::kChunkStride + r * 16 + (lane & 1) * 8) = packed;
    if (lane < 4) {
        // augmentation: 16 bf16 as A operand (chunks 16,17), 16 bf16 as B operand (chunks 18,19); lanes 0..3 write
        // one 16-byte slot each
        const float x = (row < rows) ? -0.5f * n : -1e30f;
        __nv_bfloat16 hi, mid, lo;
        split3(x, hi, mid, lo);
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
        __nv_bfloat16 e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = zero;
        if (lane == 0) { e[0] = hi; e[1] = mid; e[2] = lo; e[3] = one; e[4] = one; e[5] = one; }       // A-role, first chunk
        if (lane == 2) { e[0] = one; e[1] = one; e[2] = one; e[3] = hi; e[4] = mid; e[5] = lo; }       // B-role, first chunk
        uint4 w;
        __nv_bfloat162 q0 = __halves2bfloat162(e[0], e[1]), q1 = __halves2bfloat162(e[2], e[3]), q2 = __halves2bfloat162(e[4], e[5]),
                       q3 = __halves2bfloat162(e[6], e[7]);
        w.x = *reinterpret_cast<uint32_t*>(&q0); w.y = *reinterpret_cast<uint32_t*>(&q1);
        w.z = *reinterpret_cast<uint32_t*>(&q2); w.w = *reinterpret_cast<uint32_t*>(&q3);
        *reinterpret_cast<uint4*>(blk + (size_t)(tc::kDataChunks + lane) * tc::kChunkStride + r * 16) = w;
    }
}

// 256-bit ORB rows [rows][32 bytes] -> the SAME pre-tiled block geometry, one FP8 (e4m3) element per bit: 0x00 = 0.0,
// 0x38 = 1.0. 256 elements = 256 bytes = 16 K-chunks; the augmentation chunks carry -n/2 (n = popcount of the row) split
// into three exactly representable e4m3 pieces (16*(n>>5), (n>>1)&15, (n&1)/2) and three ones. For 0/1 vectors
// |a-b|^2 = popcount(a xor b), so with negate-A the FP8 MMA yields D = hamming(a,b)/2 EXACTLY (all partial sums are small
// multiples of 1/2). This is the default engine for ORB pairs; EACHAM_CFG_ORB_POPC selects the XOR+POPC kernel (orb_kernels.cuh) instead.
__device__ __forceinline__ void orb_tc_prep_row(const uint8_t* __restrict__ src, uint32_t rows, uint8_t* __restrict__ dst, uint32_t row, int lane) {
    uint8_t* blk = dst + (size_t)(row / tc::kBlockRows) * tc::kBlockBytes;
    const uint32_t r = row % tc::kBlockRows;
    const uint32_t byte = (row < rows) ? (uint32_t)__ldg(src + (size_t)row * 32 + lane) : 0u;
    const uint32_t n = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(byte));
    uint2 out = make_uint2(0u, 0u);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (byte & (1u << b)) out.x |= 0x38u << (8 * b);
        if (byte & (16u << b)) out.y |= 0x38u << (8 * b);
    }
    *reinterpret_cast<uint2*>(blk + (size_t)(lane >> 1) * tc::kChunkStride + r * 16 + (lane & 1) * 8) = out;
    if (lane < 4) {
        uint32_t hi, mid, lo;
        if (row < rows) {
            hi = __nv_cvt_float_to_fp8(-16.f * (float)(n >> 5), __NV_SATFINITE, __NV_E4M3);
            mid = __nv_cvt_float_to_fp8(-(float)((n >> 1) & 15u), __NV_SATFINITE, __NV_E4M3);
            lo = __nv_cvt_float_to_fp8((n & 1u) ? -0.5f : 0.f, __NV_SATFINITE, __NV_E4M3);
        } else {
            hi = __nv_cvt_float_to_fp8(-256.f, __NV_SATFINITE, __NV_E4M3);     // padding rows: n/2 = 300, so they score
            mid = __nv_cvt_float_to_fp8(-32.f, __NV_SATFINITE, __NV_E4M3);     // D >= 300 > 128 (largest real D) and pad x pad
            lo = __nv_cvt_float_to_fp8(-12.f, __NV_SATFINITE, __NV_E4M3);      // = 600 still fits the packed 16-bit row keys
        }
        const uint32_t one = 0x38u;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) { w.x = hi | (mid << 8) | (lo << 16) | (one << 24); w.y = one | (one << 8); }           // A-role
        if (lane == 2) { w.x = one | (one << 8) | (one << 16) | (hi << 24); w.y = mid | (lo << 8); }           // B-role
        *reinterpret_cast<uint4*>(blk + (size_t)(tc::kDataChunks + lane) * tc::kChunkStride + r * 16) = w;
    }
}

// one image at a time (tools / unit tests); grid covers n_blocks * 128 rows, one warp per row
__global__ void __launch_bounds__(256) sift_prep_kernel(const float* __restrict__ src, uint32_t rows, uint8_t* __restrict__ dst,
                                                        uint32_t n_blocks, uint32_t* max_norm_bits = nullptr, uint32_t* bf16_exact = nullptr) {
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row < n_blocks * tc::kBlockRows) sift_prep_row(src, rows, dst, row, threadIdx.x & 31, max_norm_bits, bf16_exact);
}
__global__ void __launch_bounds__(256) orb_tc_prep_kernel(const uint8_t* __restrict__ src, uint32_t rows, uint8_t* __restrict__ dst,
                                                          uint32_t n_blocks) {
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row < n_blocks * tc::kBlockRows) orb_tc_prep_row(src, rows, dst, row, threadIdx.x & 31);
}

struct ImageDescTc {
    unsigned long long offset;      // raw rows in the arena (fp32 x 128 or 32 bytes)
    unsigned long long tc_offset;   // pre-tiled blocks in the tc arena
    uint32_t rows;
    uint32_t kind;
    uint32_t max_norm_bits;         // F32X128: float bits of the largest row norm (of the bf16-rounded rows); filled by the prep kernel
    uint32_t bf16_exact;            // F32X128: 1 if every value is exactly representable in bf16 (integer-valued SIFT): scoring is exact
};

// The whole image table in ONE launch: CTA b handles 8 rows of 128-row block b / 16; block_start[i] = first block of image i
// (prefix sums, n_images + 1 entries; images without a tensor-core copy have an empty range).
__global__ void __launch_bounds__(256) tc_prep_all_kernel(const uint8_t* __restrict__ arena, uint8_t* __restrict__ tc_arena,
                                                          ImageDescTc* images, const uint32_t* __restrict__ block_start,
                                                          uint32_t n_images) {
    const uint32_t blk = blockIdx.x / 16;
    uint32_t lo = 0, hi = n_images;                     // last image with block_start[i] <= blk
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(block_start + mid) <= blk) lo = mid; else hi = mid;
    }
    const ImageDescTc im = images[lo];
    const uint32_t row = (blk - __ldg(block_start + lo)) * tc::kBlockRows + (blockIdx.x % 16) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (im.kind == EACHAM_KIND_F32X128) sift_prep_row(reinterpret_cast<const float*>(arena + im.offset), im.rows, tc_arena + im.tc_offset, row, lane,
                                                      &images[lo].max_norm_bits, &images[lo].bf16_exact);
    else orb_tc_prep_row(arena + im.offset, im.rows, tc_arena + im.tc_offset, row, lane);
}

// =============================================================================================================
// The fused SIFT pair kernel.
//
// One persistent CTA per image pair (static round-robin over the pair list). 576 threads:
//   warps 0-15 epilogue: tcgen05.ld -> packed-key top-2 for rows (registers) and columns (REDUX + smem slots)
//   warp 16   producer: 1-D bulk copies (TMA engine) of pre-tiled bf16 blocks, mbarrier pipeline
//   warp 17   MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128, K=16) x 9 k-steps x 2 row halves per tile
// (the single-thread roles sit in the highest warps: the scheduler favours higher warp ids and they must never starve)
// A block of 256 rows of the first image stays in shared memory while the second image streams through in
// 128-column tiles (3 stages). Accumulators: 2 stages x 2 halves x 128 fp32 columns = all 512 TMEM columns, so the
// tensor core fills stage s+1 while the epilogue drains stage s.
// TMEM holds D = |a-b|^2/2 of the bf16-rounded rows (norms folded into the GEMM: tc_common.cuh). Keys: the low 8
// mantissa bits of D are replaced by an index byte (column within the thread's 64 columns / row within the 256-row
// block), so integer min/max on the bit pattern orders by (distance, index) and REDUX.MIN works on it. Coarse
// indices (tile id, block id) are tracked per tile, not per element.
// After a row block's sweep the two candidates of every row are re-ranked exactly in FP32 from the original
// fp32 rows, the ratio test is applied, and at the end of the pair the same happens for columns, followed by the
// reference's gates / mutual filter / compaction -- all inside the CTA.
// =============================================================================================================
struct PairParamsTc {
    const uint8_t* arena;           // fp32 descriptors (exact re-rank)
    const uint8_t* tc_arena;        // bf16 blocks (scoring)
    const ImageDescTc* images;
    const eacham_pair_t* pairs;
    uint32_t n_pairs;
    double ratio;
    uint32_t min_dir, min_mutual, cross_check, emit_all;
    eacham_pair_result_t* results;
    eacham_match_t* matches;
    unsigned long long matches_cap;
    unsigned long long* cursor;
    uint8_t* scratch;               // per CTA: colstate (16 B x cols_cap) + m12 (4 B x rows_cap) + m21 (4 B x cols_cap)
    uint32_t rows_cap, cols_cap;    // multiples of 128
    uint32_t* work_counter;         // dynamic pair queue (zeroed by the host before the launch)
    const uint32_t* order;          // processing order: work item k is pair order[k] (L2-blocked by the host); results stay in input order
    uint32_t* exact_fallbacks;      // F32X128: number of queries re-done by the exact scan (statistics)
    // debug (single-pair calls, F32X128): the kNN(k=2) the ratio test saw, per row of `first` / of `second`: idx[n][2], dist[n][2]
    int32_t* dbg_idx12; float* dbg_dist12; int32_t* dbg_idx21; float* dbg_dist21;
    // single-direction mode (the reference-shaped Match(): FeatureMatcherFlann.cpp:14-30). pairs[0] is THE pair; the n_pairs work
    // items are its 128-row blocks of `first`, one CTA each; only the row direction is evaluated and single_out[row] receives the
    // ratio-passing train index or EACHAM_NONE. Offsets in the image table are absolute device addresses then (arena = tc_arena = 0).
    uint32_t single_dir;
    uint32_t* single_out;
};

constexpr int kEpiWarps = 16;          // 4 per TMEM lane quadrant, 32 columns of every tile each
constexpr int kColParts = kEpiWarps / 4;
constexpr int kColsPerWarp = 128 / kColParts;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsTc = 64 + kEpiThreads;
constexpr int kBStages = 3;
constexpr int kAccStages = 2;
constexpr int kABlockRows = 256;
constexpr int32_t kEmptyKeyTc = 0x7F7FFF00;
constexpr long long kEmptyComp = ((long long)0x7F7FFF << 32) | 0xFFFFFFFFll;

__host__ __device__ inline size_t tc_scratch_bytes_per_cta(uint32_t rows_cap, uint32_t cols_cap) {
    return (size_t)cols_cap * 96 + (size_t)rows_cap * 4 + (size_t)cols_cap * 4;      // column state (<= 96 B per column: tc_sift_kernels.cuh) + m12 + m21
}

struct SmemTc {
    uint8_t a[2][tc::kAOperandBytes];
    uint8_t b[kBStages][tc::kAOperandBytes];
    uint2 slots[2][4][128];
    int4 rowkeys[kColParts - 1][kABlockRows];   // partial row state of column parts 1..: m0, m1, t0, t1
    uint2 rowcand[kABlockRows];         // merged candidates j0, j1
    uint64_t b_full[kBStages], b_empty[kBStages], a_full, a_empty, acc_full[kAccStages], acc_empty[kAccStages];
    uint32_t tmem_slot;
    uint32_t red[2 * kEpiWarps + 8];
    unsigned long long base;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__device__ __forceinline__ void comp_merge(long long b0, long long b1, long long& m0, long long& m1) {
    const long long lo = b0 < m0 ? b0 : m0;
    const long long mx = b0 < m0 ? m0 : b0;
    const long long mn = b1 < m1 ? b1 : m1;
    m0 = lo;
    m1 = mx < mn ? mx : mn;
}

// exact reference distance: sqrtf(sum_k (a_k - b_k)^2), float accumulator; all lanes return the same value
__device__ __forceinline__ float exact_l2(const float4 a4, const float* __restrict__ brow, int lane) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow) + lane);
    const float dx = a4.x - b4.x, dy = a4.y - b4.y, dz = a4.z - b4.z, dw = a4.w - b4.w;
    float s = dx * dx;
    s = fmaf(dy, dy, s); s = fmaf(dz, dz, s); s = fmaf(dw, dw, s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return __fsqrt_rn(s);
}

// kNN(k=2) of one query row by an exact scan over ALL train rows (OpenCV order: ascending distance, ties to the lower index),
// warp-cooperative with the same arithmetic as the re-rank (exact_l2). The slow path of rerank_ratio_checked.
__device__ __noinline__ void exact_scan_top2(const float4 a4, const float* __restrict__ tbase, uint32_t n_train, int lane,
                                                float& d0, float& d1, uint32_t& j0, uint32_t& j1) {
    d0 = d1 = __int_as_float(0x7f800000);
    j0 = j1 = EACHAM_NONE;
    uint32_t j = 0;
    for (; j + 2 <= n_train; j += 2) {                   // two train rows per step: independent loads
        const float da = exact_l2(a4, tbase + (size_t)j * 128, lane);
        const float db = exact_l2(a4, tbase + (size_t)(j + 1) * 128, lane);
        if (da < d1) { if (da < d0) { d1 = d0; j1 = j0; d0 = da; j0 = j; } else { d1 = da; j1 = j; } }
        if (db < d1) { if (db < d0) { d1 = d0; j1 = j0; d0 = db; j0 = j + 1; } else { d1 = db; j1 = j + 1; } }
    }
    if (j < n_train) {
        const float da = exact_l2(a4, tbase + (size_t)j * 128, lane);
        if (da < d1) { if (da < d0) { d1 = d0; j1 = j0; d0 = da; j0 = j; } else { d1 = da; j1 = j; } }
    }
}

// Re-rank the scorer's two candidates of one query row exactly and apply the ratio test (FeatureMatcherFlann.cpp:21-27); returns
// the train index or EACHAM_NONE. All lanes of the warp take part and return the same value.
//
// Certainty check. Let s(.) be the scorer's key order and D = |a-b|^2 / 2. The candidates j0, j1 minimise s, and
// |s(c) - D(c)| <= E for every train row c, so every NON-candidate c has D(c) >= D(j1) - 2E, i.e. d(c)^2 >= X := d1^2 - 4E.
//   E = E_round + E_trunc + E_acc
//   E_round = delta (2 d1 + delta) / 2,  delta = 2^-9 (|a| + max_c |b_c|)   bf16 rounding of both operands (0 if both images are bf16-exact)
//   E_trunc = 2^-15 D(j1)                                                  keys keep 16 mantissa bits of D
//   E_acc   = 2^-16 (|a|^2 + max|b|^2) / 2                                  fp32 accumulation of inexact products (0 if bf16-exact: integer sums < 2^24)
// The reference's ratio is d0/d1 if no non-candidate enters its top two, and lies in [d0/d1, d0/sqrt(X)] if one becomes second; a
// non-candidate can only be the BEST if X < d0^2. Hence:
//   d0/sqrt(X) < ratio                      -> match j0 for certain
//   d0/d1 >= ratio and sqrt(X) >= ratio d0  -> no match for certain
//   otherwise                               -> exact scan over all train rows (counted in *fallbacks)
// The gathered loads of one query's re-rank, separated from the arithmetic so that a warp can have several queries' rows in flight
// (one query at a time costs a full L2 / HBM round trip each: measured ~2,400 cycles per query).
struct RerankLoads {
    float4 a4, b0, b1;
    uint32_t j0, j1;
    bool v0, v1;
};
__device__ __forceinline__ RerankLoads rerank_load(const float* __restrict__ qrow, const float* __restrict__ tbase, uint32_t j0, uint32_t j1,
                                                   uint32_t n_train, int lane) {
    RerankLoads L;
    L.j0 = j0; L.j1 = j1; L.v0 = j0 < n_train; L.v1 = j1 < n_train;
    L.a4 = __ldg(reinterpret_cast<const float4*>(qrow) + lane);
    L.b0 = L.v0 ? __ldg(reinterpret_cast<const float4*>(tbase + (size_t)j0 * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    L.b1 = L.v1 ? __ldg(reinterpret_cast<const float4*>(tbase + (size_t)j1 * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    return L;
}
// same arithmetic as exact_l2, operands already in registers
__device__ __forceinline__ float exact_l2_regs(const float4 a4, const float4 b4) {
    const float dx = a4.x - b4.x, dy = a4.y - b4.y, dz = a4.z - b4.z, dw = a4.w - b4.w;
    float s = dx * dx;
    s = fmaf(dy, dy, s); s = fmaf(dz, dz, s); s = fmaf(dw, dw, s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return __fsqrt_rn(s);
}
__device__ __forceinline__ uint32_t rerank_finish(const RerankLoads& L, const float* __restrict__ tbase, uint32_t n_train, double ratio, int lane,
                                                  bool exact_inputs, float other_max_norm, uint32_t* fallbacks, int32_t* dbg_idx, float* dbg_dist) {
    const float4 a4 = L.a4;
    uint32_t j0 = L.j0, j1 = L.j1;
    const bool v0 = L.v0, v1 = L.v1;
    float d0 = v0 ? exact_l2_regs(a4, L.b0) : __int_as_float(0x7f800000);
    float d1 = v1 ? exact_l2_regs(a4, L.b1) : __int_as_float(0x7f800000);
    if (d1 < d0 || (d1 == d0 && j1 < j0)) { const float t = d0; d0 = d1; d1 = t; const uint32_t u = j0; j0 = j1; j1 = u; }
    uint32_t result = EACHAM_NONE;
    if (v0 && v1) {                                      // fewer than two neighbours: the reference is UB, rejected
        float E = 0.5f * d1 * d1 * 3.0517578125e-5f;                                   // E_trunc
        if (!exact_inputs) {
            float n = a4.x * a4.x + a4.y * a4.y + a4.z * a4.z + a4.w * a4.w;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
            const float na = sqrtf(n) * 1.004f, nb = other_max_norm * 1.004f;
            const float delta = 1.953125e-3f * (na + nb) * 1.01f;
            E += 0.5f * delta * (2.f * d1 + delta) + 1.52587890625e-5f * 0.5f * (na * na + nb * nb);
        }
        const float X = d1 * d1 - 4.f * E * 1.001f;
        const bool lo_pass = (double)__fdiv_rn(d0, d1) < ratio;
        bool certain = false;
        if (X > 0.f) {
            const float sx = sqrtf(X) * 0.999999f;
            if ((double)(d0 / sx * 1.000001f) < ratio) { certain = true; result = j0; }
            else if (!lo_pass && (double)sx >= ratio * (double)d0 * 1.000001) certain = true;
        }
        if (!certain) {
            exact_scan_top2(a4, tbase, n_train, lane, d0, d1, j0, j1);
            result = ((double)__fdiv_rn(d0, d1) < ratio) ? j0 : EACHAM_NONE;
            if (lane == 0 && fallbacks != nullptr) atomicAdd(fallbacks, 1u);
        }
    }
    if (dbg_idx != nullptr && lane == 0) {
        dbg_idx[0] = v0 ? (int32_t)j0 : -1; dbg_idx[1] = v1 ? (int32_t)j1 : -1;
        dbg_dist[0] = d0; dbg_dist[1] = d1;
    }
    return result;
}
__device__ __forceinline__ uint32_t rerank_ratio_checked(const float* __restrict__ qrow, const float* __restrict__ tbase, uint32_t j0, uint32_t j1,
                                                         uint32_t n_train, double ratio, int lane, bool exact_inputs, float other_max_norm,
                                                         uint32_t* fallbacks, int32_t* dbg_idx, float* dbg_dist) {
    const RerankLoads L = rerank_load(qrow, tbase, j0, j1, n_train, lane);
    return rerank_finish(L, tbase, n_train, ratio, lane, exact_inputs, other_max_norm, fallbacks, dbg_idx, dbg_dist);
}

// ORB engine: composites carry the exact Hamming distance (an integer) in their high word: ratio test straight from them.
__device__ __forceinline__ uint32_t comp_ratio(long long k0, long long k1, uint32_t n_train, double ratio) {
    const uint32_t j0 = (uint32_t)k0, j1 = (uint32_t)k1;
    if (!(j0 < n_train && j1 < n_train)) return EACHAM_NONE;          // fewer than two neighbours: rejected
    const float d0 = (float)(uint32_t)(k0 >> 32), d1 = (float)(uint32_t)(k1 >> 32);
    return ((double)__fdiv_rn(d0, d1) < ratio) ? j0 : EACHAM_NONE;    // 0/0 -> NaN -> rejected
}

// ORB epilogue. TMEM holds D = hamming/2 exactly (real pairs <= 128, padding 300..600). Keys are built on the FMA pipe with a
// magic-number FFMA: the low mantissa bits of fma(D, 2^s, 2^23 + idx) are (2D << (s-1)) | idx.
//   rows:    16-bit key (2D << 5) | column-in-part (5 bits); the keys of the thread's two rows are packed into one word
//            (PRMT) and the running top-2 of BOTH rows is updated with three VIMNMX.U16x2;
//   columns: 32-bit key (2D << 8) | row-in-block (8 bits) under the constant exponent bits, then min/max + two REDUX.MIN as
//            in the float version.
// 9 ALU-pipe instructions per column (two scores) instead of 15.
template <int NH>
__device__ __forceinline__ void epi_tile_orb(uint32_t acc_taddr, int cp, int q, int lane, uint32_t& m0x2, uint32_t& m1x2,
                                             uint2* __restrict__ slot_q) {
    const float cr0 = 8388608.f + (float)(q * 32 + lane), cr1 = 8388608.f + (float)(128 + q * 32 + lane);
#pragma unroll
    for (int ch = 0; ch < kColsPerWarp / 16; ++ch) {
        uint32_t v0[16], v1[16];
        tc::tmem_ld16(acc_taddr + cp * kColsPerWarp + ch * 16, v0);
        if (NH == 2) tc::tmem_ld16(acc_taddr + 128 + cp * kColsPerWarp + ch * 16, v1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = ch * 16 + k;
            const float cidx = 8388608.f + (float)idx;
            const float f0 = __uint_as_float(v0[k]);
            const uint32_t x0 = __float_as_uint(fmaf(f0, 64.f, cidx));
            uint32_t x1 = 0xFFFFFFFFu;
            uint32_t lo = __float_as_uint(fmaf(f0, 512.f, cr0)), hi = (uint32_t)kEmptyKeyTc;
            if (NH == 2) {
                const float f1 = __uint_as_float(v1[k]);
                x1 = __float_as_uint(fmaf(f1, 64.f, cidx));
                const uint32_t kc1 = __float_as_uint(fmaf(f1, 512.f, cr1));
                hi = max(lo, kc1);
                lo = min(lo, kc1);
            }
            const uint32_t kr = __byte_perm(x0, x1, 0x5410);          // low halves: row h=0 | row h=1 << 16
            m1x2 = __vminu2(m1x2, __vmaxu2(m0x2, kr));
            m0x2 = __vminu2(m0x2, kr);
            const uint32_t g0 = __reduce_min_sync(0xffffffffu, lo);
            const uint32_t x = (lo == g0) ? hi : lo;
            const uint32_t g1 = __reduce_min_sync(0xffffffffu, x);
            slot_q[cp * kColsPerWarp + idx] = make_uint2(g0, g1);
        }
    }
}

// key = (value bits & 0xFFFFFF00) | index byte, as ONE LOP3 ((a & b) | c, LUT 0xEA)
__device__ __forceinline__ int32_t make_key(uint32_t v, uint32_t mask, uint32_t idx) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(v), "r"(mask), "r"(idx));
    return (int32_t)r;
}

template <int NH, bool kCols>
__device__ __forceinline__ void epi_tile(uint32_t acc_taddr, int cp, int q, int lane, int32_t (&m0)[2], int32_t (&m1)[2],
                                         uint2* __restrict__ slot_q) {
    const uint32_t ridx0 = q * 32 + lane, ridx1 = 128 + q * 32 + lane;
    uint32_t mask = 0xFFFFFF00u;
    asm volatile("" : "+r"(mask));                       // keep the mask in a register so both keys are single LOP3s
#pragma unroll
    for (int ch = 0; ch < kColsPerWarp / 16; ++ch) {
        uint32_t v0[16], v1[16];
        tc::tmem_ld16(acc_taddr + cp * kColsPerWarp + ch * 16, v0);
        if (NH == 2) tc::tmem_ld16(acc_taddr + 128 + cp * kColsPerWarp + ch * 16, v1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t idx = ch * 16 + k;
            const int32_t kr0 = make_key(v0[k], mask, idx);
            m1[0] = min(m1[0], max(m0[0], kr0));
            m0[0] = min(m0[0], kr0);
            int32_t lo = kCols ? make_key(v0[k], mask, ridx0) : 0, hi = kEmptyKeyTc;
            if (NH == 2) {
                const int32_t kr1 = make_key(v1[k], mask, idx);
                m1[1] = min(m1[1], max(m0[1], kr1));
                m0[1] = min(m0[1], kr1);
                if (kCols) {
                    const int32_t kc1 = make_key(v1[k], mask, ridx1);
                    hi = max(lo, kc1);
                    lo = min(lo, kc1);
                }
            }
            if (kCols) {
                const int32_t g0 = __reduce_min_sync(0xffffffffu, lo);
                const int32_t x = (lo == g0) ? hi : lo;
                const int32_t g1 = __reduce_min_sync(0xffffffffu, x);
                slot_q[cp * kColsPerWarp + idx] = make_uint2((uint32_t)g0, (uint32_t)g1);   // warp-uniform value: all lanes store the same word
            }
        }
    }
}

// Gates, mutual filter and ordered compaction of one pair (/root/reference/apps/sfm/main.cpp:111-146), run by the 512
// epilogue threads of a CTA once m12[0..N) / m21[0..M) hold the ratio-filtered matches of both directions (EACHAM_NONE = no
// match). S needs `red[2 * kEpiWarps + 8]` and `base`. Ends with an epilogue barrier (scratch reusable).
template <class Smem>
__device__ __forceinline__ void gates_mutual_compact(Smem& S, const PairParamsTc& p, uint32_t pi, uint32_t N, uint32_t M,
                                                     const uint32_t* __restrict__ m12, const uint32_t* __restrict__ m21, int et, int e, int lane) {
    uint32_t c12 = 0, c21 = 0;
    for (uint32_t i = et; i < N; i += kEpiThreads) c12 += m12[i] != EACHAM_NONE;
    for (uint32_t j = et; j < M; j += kEpiThreads) c21 += m21[j] != EACHAM_NONE;
    c12 = __reduce_add_sync(0xffffffffu, c12);
    c21 = __reduce_add_sync(0xffffffffu, c21);
    if (lane == 0) { S.red[e] = c12; S.red[kEpiWarps + e] = c21; }
    epi_bar();
    uint32_t n12 = 0, n21 = 0;
#pragma unroll
    for (int w = 0; w < kEpiWarps; ++w) { n12 += S.red[w]; n21 += S.red[kEpiWarps + w]; }
    const bool gated = p.cross_check ? (n12 < p.min_dir || n21 < p.min_dir) : (n12 < p.min_dir);
    const uint32_t per = (N + kEpiThreads - 1) / kEpiThreads;
    const uint32_t lo = min(N, et * per), hi = min(N, lo + per);
    uint32_t mine = 0;
    if (!gated)
        for (uint32_t a = lo; a < hi; ++a) {
            const uint32_t b = m12[a];
            mine += (b != EACHAM_NONE) && (!p.cross_check || m21[b] == a);
        }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    epi_bar();
    if (lane == 31) S.red[e] = incl;
    epi_bar();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kEpiWarps; ++w) { if (w < e) woff += S.red[w]; total += S.red[w]; }
    const uint32_t excl = woff + incl - mine;
    const bool connected = !gated && total > p.min_mutual;
    const bool emit = !gated && (connected || p.emit_all) && total > 0;
    if (et == 0) {
        unsigned long long base = 0;
        if (emit) base = atomicAdd(p.cursor, (unsigned long long)total);
        S.base = base;
        eacham_pair_result_t r;
        r.n12 = n12; r.n21 = n21; r.n_mutual = gated ? 0u : total;
        r.flags = (gated ? EACHAM_PAIR_GATED : 0u) | (connected ? EACHAM_PAIR_CONNECTED : 0u);
        r.offset = base; r.count = emit ? total : 0u;
        p.results[pi] = r;
    }
    epi_bar();
    if (emit && S.base + total <= p.matches_cap) {
        unsigned long long o = S.base + excl;
        for (uint32_t a = lo; a < hi; ++a) {
            const uint32_t b = m12[a];
            if ((b != EACHAM_NONE) && (!p.cross_check || m21[b] == a)) {
                eacham_match_t mt; mt.query = a; mt.train = b;
                p.matches[o++] = mt;
            }
        }
    }
    epi_bar();
}

template <bool kOrb>
__global__ void __launch_bounds__(kThreadsTc, 1) tc_match_pairs_kernel(const PairParamsTc p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemTc& S = *reinterpret_cast<SmemTc*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { tc::mbar_init(&S.b_full[s], 1); tc::mbar_init(&S.b_empty[s], 1); }
        tc::mbar_init(&S.a_full, 1); tc::mbar_init(&S.a_empty, 1);
        for (int s = 0; s < kAccStages; ++s) { tc::mbar_init(&S.acc_full[s], 1); tc::mbar_init(&S.acc_empty[s], kEpiWarps); }
        tc::fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tc::tmem_alloc(&S.tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = S.tmem_slot;

    if (warp == kEpiWarps) {
        // ===================================== producer =====================================
        // the whole warp walks the loop (warp-uniform control flow); one elected lane issues the copies (tc::elect_one)
        {
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
        // whole warp in the loop, one elected lane issues: no ELECT / BRA.U.ANY waterfall around every tcgen05.mma
        {
            const uint64_t dbase = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
            const uint32_t idesc = kOrb ? tc::make_idesc_e4m3_f32(128, 128, true) : tc::make_idesc_bf16_f32(128, 128, true);
            uint32_t b_it = 0, a_it = 0, acc_it = 0;
            for (uint32_t wk = blockIdx.x; wk < p.n_pairs; wk += gridDim.x) {
                const uint32_t pi = p.single_dir ? wk : p.order[wk];
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
                        tc::mbar_wait(&S.b_full[st], (b_it / kBStages) & 1);
                        tc::mbar_wait(&S.acc_empty[as], ((acc_it / kAccStages) & 1) ^ 1);
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            const uint32_t b_addr = tc::smem_u32(S.b[st]);
                            for (uint32_t h = 0; h < nh; ++h) {
                                const uint32_t a_addr = tc::smem_u32(S.a[h]);
                                const uint32_t d = tmem + as * 256 + h * 128;
#pragma unroll
                                for (int ks = 0; ks < tc::kKSteps; ++ks) {
                                    const uint64_t da = tc::smem_desc(dbase, a_addr + ks * 2 * tc::kChunkStride);
                                    const uint64_t db = tc::smem_desc(dbase, b_addr + ks * 2 * tc::kChunkStride);
                                    if (kOrb) tc::mma_f8(d, da, db, idesc, ks > 0);
                                    else tc::mma_bf16(d, da, db, idesc, ks > 0);
                                }
                            }
                            tc::mma_commit(&S.b_empty[st]);      // B stage reusable once these MMAs have read it
                            tc::mma_commit(&S.acc_full[as]);     // accumulators ready for the epilogue
                        }
                        __syncwarp();
                    }
                    if (tc::elect_one()) tc::mma_commit(&S.a_empty);              // A block reusable
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================================== epilogue =====================================
        const int e = warp, q = warp & 3, cp = e >> 2;
        const int et = e * 32 + lane;                         // 0..255 within the epilogue group
        uint8_t* my_scratch = p.scratch + (size_t)blockIdx.x * tc_scratch_bytes_per_cta(p.rows_cap, p.cols_cap);
        long long* colstate = reinterpret_cast<long long*>(my_scratch);                       // [cols_cap][2]
        uint32_t* m12 = reinterpret_cast<uint32_t*>(my_scratch + (size_t)p.cols_cap * 16);     // [rows_cap]
        uint32_t* m21 = m12 + p.rows_cap;                                                      // [cols_cap]
        uint32_t acc_it = 0;
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
            const bool cols = !p.single_dir;
            if (cols)
            for (uint32_t j = et; j < nbt * 128; j += kEpiThreads) { colstate[2 * j] = kEmptyComp; colstate[2 * j + 1] = kEmptyComp; }
            epi_bar();

            for (uint32_t ab = p.single_dir ? pi : 0u, ab1 = p.single_dir ? pi + 1 : (na128 + 1) / 2; ab < ab1; ++ab) {
                const uint32_t blk0 = p.single_dir ? pi : ab * 2;                     // first 128-row block of this work unit
                    const uint32_t nh = p.single_dir ? 1u : min(2u, na128 - ab * 2); (void)blk0;
                int32_t m0[2] = {kEmptyKeyTc, kEmptyKeyTc}, m1[2] = {kEmptyKeyTc, kEmptyKeyTc};
                uint32_t m0x2 = 0xFFFFFFFFu, m1x2 = 0xFFFFFFFFu;           // ORB: packed 16-bit row keys of both rows
                uint32_t t0[2] = {0xFFFFu, 0xFFFFu}, t1[2] = {0xFFFFu, 0xFFFFu};
                for (uint32_t bt = 0; bt < nbt; ++bt, ++acc_it) {
                    const uint32_t as = acc_it % kAccStages;
                    // column state of this tile: issue the (L2) load early, consumed after the barrier
                    long long c0 = 0, c1 = 0;
                    if (cols && et < 128) { c0 = colstate[2 * (bt * 128 + et)]; c1 = colstate[2 * (bt * 128 + et) + 1]; }
                    if (kOrb) { m0[0] = (int32_t)(m0x2 & 0xFFFFu); m0[1] = (int32_t)(m0x2 >> 16); m1[0] = (int32_t)(m1x2 & 0xFFFFu); m1[1] = (int32_t)(m1x2 >> 16); }
                    const int32_t o00 = m0[0], o10 = m1[0], o01 = m0[1], o11 = m1[1];
                    tc::mbar_wait(&S.acc_full[as], (acc_it / kAccStages) & 1);
                    tc::tc_fence_after();
                    const uint32_t taddr = tmem + as * 256 + ((uint32_t)(q * 32) << 16);
                    uint2* slot_q = S.slots[bt & 1][q];
                    if (kOrb) {
                        if (nh == 2) epi_tile_orb<2>(taddr, cp, q, lane, m0x2, m1x2, slot_q);
                        else epi_tile_orb<1>(taddr, cp, q, lane, m0x2, m1x2, slot_q);
                        m0[0] = (int32_t)(m0x2 & 0xFFFFu); m0[1] = (int32_t)(m0x2 >> 16);
                        m1[0] = (int32_t)(m1x2 & 0xFFFFu); m1[1] = (int32_t)(m1x2 >> 16);
                    } else {
                        if (!cols) {               // single-direction mode: rows only
                            if (nh == 2) epi_tile<2, false>(taddr, cp, q, lane, m0, m1, slot_q);
                            else epi_tile<1, false>(taddr, cp, q, lane, m0, m1, slot_q);
                        } else if (nh == 2) epi_tile<2, true>(taddr, cp, q, lane, m0, m1, slot_q);
                        else epi_tile<1, true>(taddr, cp, q, lane, m0, m1, slot_q);
                    }
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&S.acc_empty[as]);       // TMEM stage free for the next MMA
                    // coarse (tile) index of the row candidates
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int32_t om0 = h ? o01 : o00, om1 = h ? o11 : o10;
                        if (m0[h] != om0) { t1[h] = (m1[h] == om0) ? t0[h] : bt; t0[h] = bt; }
                        else if (m1[h] != om1) t1[h] = bt;
                    }
                    if (cols) epi_bar();                                    // all 16 warps' slots of this tile are written
                    if (cols && et < 128) {
                        long long g0 = kEmptyComp, g1 = kEmptyComp;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const uint2 s = S.slots[bt & 1][qq][et];
                            // value part: float path = top 24 bits of D; ORB path = the integer 2D = hamming (below the exponent bits)
                            const long long v0 = kOrb ? (long long)((s.x >> 8) & 0x7FFFu) : (long long)((int32_t)s.x >> 8);
                            const long long v1 = kOrb ? (long long)((s.y >> 8) & 0x7FFFu) : (long long)((int32_t)s.y >> 8);
                            const long long k0 = (v0 << 32) | (long long)(ab * kABlockRows + (s.x & 0xFFu));
                            const long long k1 = (v1 << 32) | (long long)(ab * kABlockRows + (s.y & 0xFFu));
                            comp_merge(k0, k1, g0, g1);
                        }
                        comp_merge(c0, c1, g0, g1);
                        colstate[2 * (bt * 128 + et)] = g0;
                        colstate[2 * (bt * 128 + et) + 1] = g1;
                    }
                }
                // ---- rows of this block are complete: merge the two column halves, re-rank exactly, ratio test ----
                if (cp > 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) S.rowkeys[cp - 1][h * 128 + q * 32 + lane] = make_int4(m0[h], m1[h], (int)t0[h], (int)t1[h]);
                }
                epi_bar();
                if (cp == 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // composite = (value << 32) | full column. float path: key = value bits | index byte (8 index bits);
                        // ORB path: key = (hamming << 5) | column-in-part (5 index bits)
                        constexpr int kIdxBits = kOrb ? 5 : 8;
                        constexpr int kIdxMask = (1 << kIdxBits) - 1;
                        long long a0 = ((long long)(m0[h] >> kIdxBits) << 32) | (long long)(t0[h] * 128 + (m0[h] & kIdxMask));
                        long long a1 = ((long long)(m1[h] >> kIdxBits) << 32) | (long long)(t1[h] * 128 + (m1[h] & kIdxMask));
#pragma unroll
                        for (int c = 1; c < kColParts; ++c) {
                            const int4 o = S.rowkeys[c - 1][h * 128 + q * 32 + lane];
                            const long long b0 = ((long long)(o.x >> kIdxBits) << 32) | (long long)((uint32_t)o.z * 128 + c * kColsPerWarp + (o.x & kIdxMask));
                            const long long b1 = ((long long)(o.y >> kIdxBits) << 32) | (long long)((uint32_t)o.w * 128 + c * kColsPerWarp + (o.y & kIdxMask));
                            comp_merge(b0, b1, a0, a1);
                        }
                        if (kOrb) {                                   // exact distances are in the keys: no re-rank
                            const uint32_t row = blk0 * 128 + h * 128 + q * 32 + lane;
                            if (row < N && h < (int)nh) (p.single_dir ? p.single_out : m12)[row] = comp_ratio(a0, a1, M, p.ratio);
                        } else {
                            S.rowcand[h * 128 + q * 32 + lane] = make_uint2((uint32_t)a0, (uint32_t)a1);
                        }
                    }
                }
                epi_bar();
                if (!kOrb)
#pragma unroll 4                                                            // independent rows: overlap their gathered loads
                for (int r = 0; r < kABlockRows / kEpiWarps; ++r) {
                    const uint32_t lr = e * (kABlockRows / kEpiWarps) + r, row = blk0 * 128 + lr;
                    if (row < N && lr < nh * 128) {
                        const uint2 cand = S.rowcand[lr];
                        const uint32_t mm = rerank_ratio_checked(Af + (size_t)row * 128, Bf, cand.x, cand.y, M, p.ratio, lane, both_exact, __uint_as_float(B.max_norm_bits),
                                                                 p.exact_fallbacks, p.dbg_idx12 ? p.dbg_idx12 + 2 * (size_t)row : nullptr,
                                                                 p.dbg_dist12 ? p.dbg_dist12 + 2 * (size_t)row : nullptr);
                        if (lane == 0) (p.single_dir ? p.single_out : m12)[row] = mm;
                    }
                }
                epi_bar();                                                  // rowkeys / rowcand reusable
            }

            if (p.single_dir) continue;                       // one direction only
            // ---- columns: re-rank, ratio -> m21 ----
            __threadfence_block();
            if (kOrb) {
                for (uint32_t j = et; j < M; j += kEpiThreads) m21[j] = comp_ratio(colstate[2 * j], colstate[2 * j + 1], N, p.ratio);
            } else {
#pragma unroll 4
                for (uint32_t j = e; j < M; j += kEpiWarps) {
                    const long long k0 = colstate[2 * j], k1 = colstate[2 * j + 1];
                    const uint32_t mm = rerank_ratio_checked(Bf + (size_t)j * 128, Af, (uint32_t)k0, (uint32_t)k1, N, p.ratio, lane, both_exact, __uint_as_float(A.max_norm_bits),
                                                             p.exact_fallbacks, p.dbg_idx21 ? p.dbg_idx21 + 2 * (size_t)j : nullptr,
                                                             p.dbg_dist21 ? p.dbg_dist21 + 2 * (size_t)j : nullptr);
                    if (lane == 0) m21[j] = mm;
                }
            }
            __threadfence_block();
            epi_bar();

            gates_mutual_compact(S, p, pi, N, M, m12, m21, et, e, lane);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tcm
}  // namespace eacham
