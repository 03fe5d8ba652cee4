// orb_kernels.cuh -- 256-bit ORB Hamming matching for sm_100a: the XOR+POPC engine (EACHAM_CFG_ORB_POPC) and the
// single-direction kernels behind the reference-shaped Match()/knnMatch calls. (Batched ORB pairs default to the
// tensor-core engine in tc_match_kernels.cuh, which returns the same bytes.)
//
// Replaces, for CV_8U 32-byte descriptors, what the reference does per unordered image pair:
//   knnMatch(k=2) in both directions   /root/reference/modules/base/features/FeatureMatcherFlann.cpp:17
//   Lowe ratio test                    /root/reference/modules/base/features/FeatureMatcherFlann.cpp:21-27
//   per-direction gate, mutual filter, connect gate     /root/reference/apps/sfm/main.cpp:111-146
// Distance = popcount of XOR over 8 x 32-bit words       /root/reference/modules/base/tools/Tools3d.h:46-63
//
// Design (DESIGN.md section "ORB kernel"):  one persistent CTA owns one image pair at a time.
//   * the 512 threads hold up to 8 rows each of the FIRST image in registers (4096 rows per row block);
//   * the SECOND image streams through shared memory in 256-column chunks (8 KB, cp.async.bulk + mbarrier,
//     double buffered); every thread reads the same column at the same time (broadcast LDS.128 x2);
//   * each 256-bit distance is evaluated ONCE (8 XOR + 8 carry-save LOP3 + 4 POPC, see hamming256) and feeds both directions:
//       row direction    -> packed key (d << 16 | column) into a per-row top-2 kept in registers,
//       column direction -> packed key (d << 16 | row) into a per-thread top-2, reduced across the warp with
//                           two REDUX.MIN, across the 16 warps through a shared-memory slot table;
//     packed keys make "lowest index wins ties" (OpenCV batchDistance: strict <, ascending index) a plain min;
//   * ratio test, gates, mutual filter and ordered compaction run in the same CTA; only surviving matches
//     leave the chip.
// POPC issues at 16 lane-ops/clk/SM (profiles/r01_pipe_microbench.jsonl); with the carry-save popcount the ALU pipe
// (LOP3 + VIMNMX, 89 % busy) is the limiter: 44.4k pairs/s at 4k x 4k = 1.28 x the "8 POPC per distance" roofline.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"

namespace eacham {
namespace orb {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerThread = 8;
constexpr int kRowBlock = kThreads * kRowsPerThread;       // 4096 rows of the first image per sweep
constexpr int kChunkCols = 256;                            // columns of the second image per smem chunk
constexpr int kChunkBytes = kChunkCols * 32;
constexpr uint32_t kEmptyKey = 0xFFFF0000u;                // distance field 0xFFFF: "no neighbour"
constexpr uint32_t kPadBias = 512;                         // added to distances of padding rows (> 256 = max real)
constexpr uint32_t kMaxRowsFused = 16384;                  // smem budget: colstate 8 B/col + m12 2 B/row
constexpr uint32_t kNone16 = 0xFFFFu;

struct ImageDesc {
    unsigned long long offset;   // byte offset into the arena (128-byte aligned)
    uint32_t rows;
    uint32_t kind;
};

struct PairParams {
    const uint8_t* arena;
    const ImageDesc* images;
    const eacham_pair_t* pairs;
    uint32_t n_pairs;
    uint32_t* work_counter;
    const uint32_t* order;   // work item k is pair order[k] (L2-blocked processing order; results stay in input order)
    double ratio;
    uint32_t min_dir, min_mutual, cross_check, emit_all;
    eacham_pair_result_t* results;
    eacham_match_t* matches;
    unsigned long long matches_cap;
    unsigned long long* cursor;
    uint32_t smem_cols;   // capacity of colstate (>= max rows of any `second` image, multiple of 4)
    uint32_t smem_rows;   // capacity of m12 (>= max rows of any `first` image, multiple of 8)
};

__host__ __device__ inline size_t pair_smem_bytes(uint32_t smem_cols, uint32_t smem_rows) {
    return 2 * (size_t)kChunkBytes + (size_t)kWarps * kChunkCols * 8 + (size_t)smem_cols * 8 + (size_t)smem_rows * 2 + 256;
}

// ---------------------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done, addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 256-bit Hamming distance. The 8 XOR words are compressed with four carry-save adders (Harley-Seal step:
// sum = a^b^c, carry = maj(a,b,c), one LOP3 each) into 4 bit-planes of weight 1,1,2,4, so only 4 POPC are issued
// instead of 8:  d = popc(s2) + popc(x7) + 2*popc(s3) + 4*popc(c3).  Same integer result as Tools3d.h:46-63.
// POPC is the scarce pipe (16 lane-ops/clk/SM); LOP3 runs at 64/clk/SM and the adds go to the FMA pipe as IMAD,
// which is how this loop exceeds the "8 POPC per distance" roofline (tools/csa_microbench.cu).
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t hamming256(const uint32_t (&a)[8], const uint4& b0, const uint4& b1) {
    const uint32_t x0 = a[0] ^ b0.x, x1 = a[1] ^ b0.y, x2 = a[2] ^ b0.z, x3 = a[3] ^ b0.w;
    const uint32_t x4 = a[4] ^ b1.x, x5 = a[5] ^ b1.y, x6 = a[6] ^ b1.z, x7 = a[7] ^ b1.w;
    const uint32_t s0 = xor3(x0, x1, x2), c0 = maj3(x0, x1, x2);
    const uint32_t s1 = xor3(x3, x4, x5), c1 = maj3(x3, x4, x5);
    const uint32_t s2 = xor3(s0, s1, x6), c2 = maj3(s0, s1, x6);
    const uint32_t s3 = xor3(c0, c1, c2), c3 = maj3(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(s3) + 4 * __popc(c3);
}

// top-2 of the union of two sorted pairs (keys are unique apart from kEmptyKey-class fillers)
__device__ __forceinline__ uint2 merge2(uint2 a, uint2 b) {
    uint32_t lo = min(a.x, b.x);
    uint32_t hi = min(max(a.x, b.x), min(a.y, b.y));
    return make_uint2(lo, hi);
}

// FeatureMatcherFlann.cpp:23  `m[0].distance / m[1].distance < 0.8`  (float / float vs a double).
// Keys carry the integer distance in bits 16..; a distance field > 256 means "fewer than two neighbours"
// (the reference dereferences m[1] unguarded = UB; such rows are rejected here).
__device__ __forceinline__ bool ratio_pass(uint32_t k0, uint32_t k1, double ratio) {
    uint32_t d0 = k0 >> 16, d1 = k1 >> 16;
    if (d1 > 256u) return false;
    float r = __fdiv_rn((float)d0, (float)d1);   // 0/0 -> NaN -> compares false, as in the reference
    return (double)r < ratio;
}

// ---------------------------------------------------------------------------------------------------------
// column sweep over one chunk: NR live row slots per thread
// ---------------------------------------------------------------------------------------------------------
template <int NR>
__device__ __forceinline__ void sweep_chunk(const uint32_t (&a)[kRowsPerThread][8], uint32_t (&m0)[kRowsPerThread],
                                            uint32_t (&m1)[kRowsPerThread], const uint4* __restrict__ cols, int ncols,
                                            uint32_t jbase, uint32_t padbias, uint32_t rowbase,
                                            uint2* __restrict__ myslots, int lane) {
#pragma unroll 1
    for (int jj = 0; jj < ncols; ++jj) {
        const uint4 b0 = cols[2 * jj], b1 = cols[2 * jj + 1];
        const uint32_t j = jbase + jj;
        uint32_t c0 = kEmptyKey, c1 = kEmptyKey;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            uint32_t d = hamming256(a[r], b0, b1);
            if (r == NR - 1) d += padbias;                 // only the last live slot can hold padding rows
            const uint32_t kr = (d << 16) + j;             // row direction: tie -> lower column
            const uint32_t kc = (d << 16) + (uint32_t)(r * kThreads);   // column direction: tie -> lower row
            m1[r] = min(m1[r], max(m0[r], kr));
            m0[r] = min(m0[r], kr);
            c1 = min(c1, max(c0, kc));
            c0 = min(c0, kc);
        }
        c0 += rowbase;                                      // thread-local slot index -> row index in the image
        c1 += rowbase;
        const uint32_t g0 = __reduce_min_sync(0xffffffffu, c0);
        const uint32_t x = (c0 == g0) ? c1 : c0;
        const uint32_t g1 = __reduce_min_sync(0xffffffffu, x);
        if (lane == (jj & 31)) myslots[jj] = make_uint2(g0, g1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// the fused pair kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) orb_match_pairs_kernel(const PairParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* chunk = reinterpret_cast<uint4*>(smem);                                       // [2][kChunkCols*2]
    uint2* slots = reinterpret_cast<uint2*>(smem + 2 * kChunkBytes);                     // [kWarps][kChunkCols]
    uint2* colstate = slots + kWarps * kChunkCols;                                       // [smem_cols]
    uint16_t* m12 = reinterpret_cast<uint16_t*>(colstate + p.smem_cols);                 // [smem_rows]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(m12) + (size_t)p.smem_rows * 2);  // [2], 8B aligned
    uint32_t* s_u32 = reinterpret_cast<uint32_t*>(mbar + 2);                             // scratch: [0]=pair, [1..]=reductions

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    uint32_t seq = 0;   // chunks consumed so far by this CTA: buffer = seq & 1, parity = (seq >> 1) & 1

    for (;;) {
        if (tid == 0) { const uint32_t wk = atomicAdd(p.work_counter, 1u); s_u32[0] = wk < p.n_pairs ? p.order[wk] : 0xFFFFFFFFu; }
        __syncthreads();
        const uint32_t pi = s_u32[0];
        if (pi == 0xFFFFFFFFu) break;
        const eacham_pair_t pr = p.pairs[pi];
        const ImageDesc A = p.images[pr.first], B = p.images[pr.second];
        const uint32_t N = A.rows, M = B.rows;
        const uint4* __restrict__ Ap = reinterpret_cast<const uint4*>(p.arena + A.offset);
        const uint8_t* __restrict__ Bp = p.arena + B.offset;

        if (N == 0 || M == 0) {                      // nothing to match: both maps empty
            if (tid == 0) {
                eacham_pair_result_t r;
                r.n12 = 0; r.n21 = 0; r.n_mutual = 0;
                r.flags = (0u < p.min_dir) ? EACHAM_PAIR_GATED : 0u;
                r.offset = 0; r.count = 0;
                p.results[pi] = r;
            }
            __syncthreads();
            continue;
        }

        const uint32_t nchunks = (M + kChunkCols - 1) / kChunkCols;
        const uint32_t nrb = (N + kRowBlock - 1) / kRowBlock;

        for (uint32_t rb = 0; rb < nrb; ++rb) {
            const uint32_t row0 = rb * kRowBlock;
            const uint32_t rows_blk = min((uint32_t)kRowBlock, N - row0);
            const int nr = (int)((rows_blk + kThreads - 1) / kThreads);          // live row slots, 1..8
            const uint32_t rowbase = row0 + tid;

            // producer: first two chunks of this sweep
            if (tid == 0) {
#pragma unroll
                for (uint32_t c = 0; c < 2; ++c) {
                    if (c < nchunks) {
                        const uint32_t b = (seq + c) & 1u;
                        const uint32_t bytes = min((uint32_t)kChunkCols, M - c * kChunkCols) * 32u;
                        mbar_expect_tx(&mbar[b], bytes);
                        bulk_g2s(chunk + b * (kChunkCols * 2), Bp + (size_t)c * kChunkBytes, bytes, &mbar[b]);
                    }
                }
            }

            // rows of the first image -> registers (coalesced 2 x 128-bit loads per row)
            uint32_t a[kRowsPerThread][8];
            uint32_t m0[kRowsPerThread], m1[kRowsPerThread];
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                const uint32_t i = rowbase + r * kThreads;
                uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                if (i < N) { lo = __ldg(Ap + 2 * (size_t)i); hi = __ldg(Ap + 2 * (size_t)i + 1); }
                a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = lo.z; a[r][3] = lo.w;
                a[r][4] = hi.x; a[r][5] = hi.y; a[r][6] = hi.z; a[r][7] = hi.w;
                m0[r] = kEmptyKey; m1[r] = kEmptyKey;
            }
            const uint32_t padbias = (rowbase + (uint32_t)(nr - 1) * kThreads < N) ? 0u : kPadBias;

            for (uint32_t c = 0; c < nchunks; ++c) {
                const uint32_t b = seq & 1u;
                mbar_wait(&mbar[b], (seq >> 1) & 1u);
                const int ncols = (int)min((uint32_t)kChunkCols, M - c * kChunkCols);
                const uint4* cols = chunk + b * (kChunkCols * 2);
                uint2* myslots = slots + warp * kChunkCols;
                const uint32_t jbase = c * kChunkCols;
                switch (nr) {
                    case 8: sweep_chunk<8>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 7: sweep_chunk<7>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 6: sweep_chunk<6>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 5: sweep_chunk<5>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 4: sweep_chunk<4>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 3: sweep_chunk<3>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    case 2: sweep_chunk<2>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                    default: sweep_chunk<1>(a, m0, m1, cols, ncols, jbase, padbias, rowbase, myslots, lane); break;
                }
                __syncthreads();   // slots complete; everyone is done reading chunk buffer b
                if (tid == 0 && c + 2 < nchunks) {
                    const uint32_t bytes = min((uint32_t)kChunkCols, M - (c + 2) * kChunkCols) * 32u;
                    mbar_expect_tx(&mbar[b], bytes);
                    bulk_g2s(chunk + b * (kChunkCols * 2), Bp + (size_t)(c + 2) * kChunkBytes, bytes, &mbar[b]);
                }
                if (tid < ncols) {   // cross-warp merge of this chunk's columns (all rows of this row block seen)
                    uint2 g = make_uint2(kEmptyKey, kEmptyKey);
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) g = merge2(g, slots[w * kChunkCols + tid]);
                    const uint32_t j = jbase + tid;
                    if (rb > 0) g = merge2(g, colstate[j]);
                    colstate[j] = g;
                }
                __syncthreads();   // slot table free again
                ++seq;
            }

            // row direction is complete for this row block: ratio test -> m12
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                const uint32_t i = rowbase + r * kThreads;
                if (r < nr && i < N) m12[i] = ratio_pass(m0[r], m1[r], p.ratio) ? (uint16_t)(m0[r] & 0xFFFFu) : (uint16_t)kNone16;
            }
        }
        __syncthreads();

        // column direction: ratio test -> m21 (stored over colstate[j].x), count both maps
        uint32_t cnt12 = 0, cnt21 = 0;
        for (uint32_t j = tid; j < M; j += kThreads) {
            const uint2 s = colstate[j];
            const bool ok = ratio_pass(s.x, s.y, p.ratio);
            colstate[j].x = ok ? (s.x & 0xFFFFu) : EACHAM_NONE;
            cnt21 += ok;
        }
        for (uint32_t i = tid; i < N; i += kThreads) cnt12 += (m12[i] != kNone16);
        cnt12 = __reduce_add_sync(0xffffffffu, cnt12);
        cnt21 = __reduce_add_sync(0xffffffffu, cnt21);
        if (lane == 0) { s_u32[4 + warp] = cnt12; s_u32[4 + kWarps + warp] = cnt21; }
        __syncthreads();
        uint32_t n12 = 0, n21 = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { n12 += s_u32[4 + w]; n21 += s_u32[4 + kWarps + w]; }
        const bool gated = p.cross_check ? (n12 < p.min_dir || n21 < p.min_dir) : (n12 < p.min_dir);

        // mutual filter (main.cpp:133-140) + ordered compaction: thread t owns rows [t*per, (t+1)*per)
        const uint32_t per = (N + kThreads - 1) / kThreads;
        const uint32_t a_lo = min(N, tid * per), a_hi = min(N, a_lo + per);
        uint32_t mine = 0;
        if (!gated) {
            for (uint32_t ai = a_lo; ai < a_hi; ++ai) {
                const uint32_t bj = m12[ai];
                mine += (bj != kNone16) && (!p.cross_check || colstate[bj].x == ai);
            }
        }
        // block exclusive scan of `mine`
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();                       // previous readers of s_u32[4..] are done
        if (lane == 31) s_u32[4 + warp] = incl;
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t v = s_u32[4 + w];
            if (w < warp) warp_off += v;
            total += v;
        }
        const uint32_t excl = warp_off + incl - mine;
        const bool connected = !gated && total > p.min_mutual;
        const bool emit = !gated && (connected || p.emit_all) && total > 0;
        if (tid == 0) {
            unsigned long long base = 0;
            if (emit) base = atomicAdd(p.cursor, (unsigned long long)total);
            reinterpret_cast<unsigned long long*>(s_u32 + 40)[0] = base;
            eacham_pair_result_t r;
            r.n12 = n12; r.n21 = n21; r.n_mutual = gated ? 0u : total;
            r.flags = (gated ? EACHAM_PAIR_GATED : 0u) | (connected ? EACHAM_PAIR_CONNECTED : 0u);
            r.offset = base; r.count = emit ? total : 0u;
            p.results[pi] = r;
        }
        __syncthreads();
        if (emit) {
            const unsigned long long base = reinterpret_cast<unsigned long long*>(s_u32 + 40)[0];
            if (base + total <= p.matches_cap) {          // otherwise: overflow, the host sees cursor > cap
                unsigned long long o = base + excl;
                for (uint32_t ai = a_lo; ai < a_hi; ++ai) {
                    const uint32_t bj = m12[ai];
                    if ((bj != kNone16) && (!p.cross_check || colstate[bj].x == ai)) {
                        eacham_match_t m; m.query = ai; m.train = bj;
                        p.matches[o++] = m;
                    }
                }
            }
        }
        __syncthreads();   // colstate / m12 / s_u32 are reused by the next pair
    }
}

// ---------------------------------------------------------------------------------------------------------
// single-direction kNN(k=2) for the reference-shaped Match() call: grid (row blocks, column splits),
// partial packed keys to scratch, merged by knn2_finalize. Arbitrary sizes.
// ---------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr int kKnnRows = 4;                        // rows per thread
constexpr int kKnnRowBlock = kKnnThreads * kKnnRows;
constexpr int kKnnTile = 128;                      // columns per smem tile

__global__ void __launch_bounds__(kKnnThreads) orb_knn2_partial_kernel(const uint8_t* __restrict__ q, uint32_t nq,
                                                                       const uint8_t* __restrict__ t, uint32_t nt,
                                                                       uint32_t cols_per_split, uint2* __restrict__ partial) {
    __shared__ uint4 tile[kKnnTile * 2];
    const int tid = threadIdx.x;
    const uint32_t row0 = blockIdx.x * kKnnRowBlock + tid;
    const uint32_t c_lo = blockIdx.y * cols_per_split;
    const uint32_t c_hi = min(nt, c_lo + cols_per_split);
    const uint4* qp = reinterpret_cast<const uint4*>(q);
    const uint4* tp = reinterpret_cast<const uint4*>(t);
    uint32_t a[kKnnRows][8], m0[kKnnRows], m1[kKnnRows];
#pragma unroll
    for (int r = 0; r < kKnnRows; ++r) {
        const uint32_t i = row0 + r * kKnnThreads;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (i < nq) { lo = __ldg(qp + 2 * (size_t)i); hi = __ldg(qp + 2 * (size_t)i + 1); }
        a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = lo.z; a[r][3] = lo.w;
        a[r][4] = hi.x; a[r][5] = hi.y; a[r][6] = hi.z; a[r][7] = hi.w;
        m0[r] = kEmptyKey; m1[r] = kEmptyKey;
    }
    for (uint32_t c0 = c_lo; c0 < c_hi; c0 += kKnnTile) {
        const uint32_t n = min((uint32_t)kKnnTile, c_hi - c0);
        __syncthreads();
        for (uint32_t k = tid; k < n * 2; k += kKnnThreads) tile[k] = __ldg(tp + 2 * (size_t)c0 + k);
        __syncthreads();
        for (uint32_t jj = 0; jj < n; ++jj) {
            const uint4 b0 = tile[2 * jj], b1 = tile[2 * jj + 1];
#pragma unroll
            for (int r = 0; r < kKnnRows; ++r) {
                const uint32_t kr = (hamming256(a[r], b0, b1) << 16) + (c0 + jj);
                m1[r] = min(m1[r], max(m0[r], kr));
                m0[r] = min(m0[r], kr);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kKnnRows; ++r) {
        const uint32_t i = row0 + r * kKnnThreads;
        if (i < nq) partial[(size_t)blockIdx.y * nq + i] = make_uint2(m0[r], m1[r]);
    }
}

// merges the column splits; writes knn idx/dist (OpenCV DMatch shape) and/or the ratio-filtered match per query
__global__ void orb_knn2_finalize_kernel(const uint2* __restrict__ partial, uint32_t nq, uint32_t nsplit, double ratio,
                                         int32_t* __restrict__ idx, float* __restrict__ dist, uint32_t* __restrict__ match) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint2 g = make_uint2(kEmptyKey, kEmptyKey);
    for (uint32_t s = 0; s < nsplit; ++s) g = merge2(g, partial[(size_t)s * nq + i]);
    const bool v0 = (g.x >> 16) <= 256u, v1 = (g.y >> 16) <= 256u;
    if (idx) {
        idx[2 * i] = v0 ? (int32_t)(g.x & 0xFFFFu) : -1;
        idx[2 * i + 1] = v1 ? (int32_t)(g.y & 0xFFFFu) : -1;
        dist[2 * i] = v0 ? (float)(g.x >> 16) : __int_as_float(0x7f800000);
        dist[2 * i + 1] = v1 ? (float)(g.y >> 16) : __int_as_float(0x7f800000);
    }
    if (match) match[i] = ratio_pass(g.x, g.y, ratio) ? (g.x & 0xFFFFu) : EACHAM_NONE;
}

}  // namespace orb
}  // namespace eacham
