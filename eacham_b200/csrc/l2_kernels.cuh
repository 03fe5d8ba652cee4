// l2_kernels.cuh -- exact FP32 L2 kNN(k=2) for 128-d float descriptors (SIFT), sm_100a.
//
// Reference semantics: cv::BFMatcher(NORM_L2).knnMatch(k=2) as called from
// /root/reference/modules/base/features/FeatureMatcherFlann.cpp:17 on the N x 128 CV_32F matrices produced by
// /root/reference/modules/base/features/FeatureExtractorSift.cpp:14-26:
//   dist = sqrtf(sum_k (a_k - b_k)^2) with a float accumulator; ascending distance; ties -> lower train index.
// This file is the all-FP32 path (direct differences, no -2ab expansion, so no cancellation): it serves the
// reference-shaped Match()/knn2 calls and is the re-rank arithmetic the tensor-core scorer must agree with.
// Keys are 64-bit: (float bits of the distance << 32) | train index -- distances are >= 0 so the IEEE bit
// pattern orders like the value, and the index in the low word makes "lower index wins ties" a plain min.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"

namespace eacham {
namespace l2 {

constexpr int kDim = 128;
constexpr int kBM = 64, kBN = 64, kBK = 32, kThreads = 256;
constexpr unsigned long long kEmptyKey64 = 0x7f800000ffffffffull;   // +inf, idx 0xffffffff

__device__ __forceinline__ void top2_insert(unsigned long long k, unsigned long long& m0, unsigned long long& m1) {
    const unsigned long long hi = k > m0 ? k : m0;
    m0 = k < m0 ? k : m0;
    m1 = hi < m1 ? hi : m1;
}
__device__ __forceinline__ void top2_merge(unsigned long long b0, unsigned long long b1, unsigned long long& m0,
                                           unsigned long long& m1) {
    const unsigned long long lo = b0 < m0 ? b0 : m0;
    const unsigned long long mx = b0 < m0 ? m0 : b0;
    const unsigned long long mn = b1 < m1 ? b1 : m1;
    m0 = lo;
    m1 = mx < mn ? mx : mn;
}

// grid (ceil(nq/64), nsplit). Each CTA: 64 query rows x its column range, 4x4 outputs per thread.
__global__ void __launch_bounds__(kThreads) l2_knn2_partial_kernel(const float* __restrict__ q, uint32_t nq,
                                                                   const float* __restrict__ t, uint32_t nt,
                                                                   uint32_t cols_per_split,
                                                                   unsigned long long* __restrict__ partial) {
    __shared__ float As[kBK][kBM + 4];
    __shared__ float Bs[kBK][kBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t r0 = blockIdx.x * kBM;
    const uint32_t c_lo = blockIdx.y * cols_per_split, c_hi = min(nt, c_lo + cols_per_split);
    unsigned long long m0[4], m1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m0[i] = kEmptyKey64; m1[i] = kEmptyKey64; }

    for (uint32_t c0 = c_lo; c0 < c_hi; c0 += kBN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < kDim; k0 += kBK) {
            __syncthreads();
            // 64 rows x 32 dims per operand: 2048 floats, 8 per thread; thread -> (row = e / 32, k = e % 32)
#pragma unroll
            for (int e = tid; e < kBM * kBK; e += kThreads) {
                const int rr = e >> 5, kk = e & 31;
                const uint32_t qi = r0 + rr, tj = c0 + rr;
                As[kk][rr] = qi < nq ? __ldg(q + (size_t)qi * kDim + k0 + kk) : 0.f;
                Bs[kk][rr] = tj < c_hi ? __ldg(t + (size_t)tj * kDim + k0 + kk) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kBK; ++kk) {
                float av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float df = av[i] - bv[j];
                        acc[i][j] = fmaf(df, df, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t col = c0 + tx * 4 + j;
                if (col < c_hi) {
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(__fsqrt_rn(acc[i][j])) << 32) | col;
                    top2_insert(key, m0[i], m1[i]);
                }
            }
    }
    // reduce across the 16 threads (tx) that share a row: they are 16 consecutive lanes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, m0[i], o);
            const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, m1[i], o);
            top2_merge(b0, b1, m0[i], m1[i]);
        }
        const uint32_t row = r0 + ty * 4 + i;
        if (tx == 0 && row < nq) {
            partial[((size_t)blockIdx.y * nq + row) * 2] = m0[i];
            partial[((size_t)blockIdx.y * nq + row) * 2 + 1] = m1[i];
        }
    }
}

__global__ void l2_knn2_finalize_kernel(const unsigned long long* __restrict__ partial, uint32_t nq, uint32_t nsplit,
                                        double ratio, int32_t* __restrict__ idx, float* __restrict__ dist,
                                        uint32_t* __restrict__ match) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned long long m0 = kEmptyKey64, m1 = kEmptyKey64;
    for (uint32_t s = 0; s < nsplit; ++s)
        top2_merge(partial[((size_t)s * nq + i) * 2], partial[((size_t)s * nq + i) * 2 + 1], m0, m1);
    const uint32_t i0 = (uint32_t)m0, i1 = (uint32_t)m1;
    const float d0 = __uint_as_float((uint32_t)(m0 >> 32)), d1 = __uint_as_float((uint32_t)(m1 >> 32));
    if (idx) {
        idx[2 * i] = i0 == 0xffffffffu ? -1 : (int32_t)i0;
        idx[2 * i + 1] = i1 == 0xffffffffu ? -1 : (int32_t)i1;
        dist[2 * i] = d0;
        dist[2 * i + 1] = d1;
    }
    if (match) {
        // FeatureMatcherFlann.cpp:23: float / float compared with the double ratio; NaN (0/0) rejects.
        bool ok = (i1 != 0xffffffffu);
        if (ok) ok = (double)__fdiv_rn(d0, d1) < ratio;
        match[i] = ok ? i0 : EACHAM_NONE;
    }
}

}  // namespace l2

// ---------------------------------------------------------------------------------------------------------
// Generic pair finalisation from two ratio-filtered direction maps already in global memory
// (/root/reference/apps/sfm/main.cpp:111-146). One CTA per pair. Used by the paths that produce m12/m21
// with separate kNN kernels (SIFT exact path, ORB images larger than the fused kernel's limit).
// ---------------------------------------------------------------------------------------------------------
struct FinalizeParams {
    const uint32_t* m12; uint32_t n1;
    const uint32_t* m21; uint32_t n2;
    uint32_t min_dir, min_mutual, cross_check, emit_all;
    eacham_pair_result_t* result;
    eacham_match_t* matches; unsigned long long matches_cap; unsigned long long* cursor;
};

__global__ void __launch_bounds__(256) pair_finalize_kernel(const FinalizeParams p) {
    __shared__ uint32_t s_a[8], s_b[8];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t c12 = 0, c21 = 0;
    for (uint32_t i = tid; i < p.n1; i += 256) c12 += p.m12[i] != EACHAM_NONE;
    for (uint32_t j = tid; j < p.n2; j += 256) c21 += p.m21[j] != EACHAM_NONE;
    c12 = __reduce_add_sync(0xffffffffu, c12);
    c21 = __reduce_add_sync(0xffffffffu, c21);
    if (lane == 0) { s_a[warp] = c12; s_b[warp] = c21; }
    __syncthreads();
    uint32_t n12 = 0, n21 = 0;
    for (int w = 0; w < 8; ++w) { n12 += s_a[w]; n21 += s_b[w]; }
    const bool gated = p.cross_check ? (n12 < p.min_dir || n21 < p.min_dir) : (n12 < p.min_dir);
    const uint32_t per = (p.n1 + 255) / 256;
    const uint32_t lo = min(p.n1, tid * per), hi = min(p.n1, lo + per);
    uint32_t mine = 0;
    if (!gated)
        for (uint32_t a = lo; a < hi; ++a) {
            const uint32_t b = p.m12[a];
            mine += (b != EACHAM_NONE) && (!p.cross_check || p.m21[b] == a);
        }
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();
    if (lane == 31) s_a[warp] = incl;
    __syncthreads();
    uint32_t off = 0, total = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) off += s_a[w]; total += s_a[w]; }
    const uint32_t excl = off + incl - mine;
    const bool connected = !gated && total > p.min_mutual;
    const bool emit = !gated && (connected || p.emit_all) && total > 0;
    if (tid == 0) {
        unsigned long long base = 0;
        if (emit) base = atomicAdd(p.cursor, (unsigned long long)total);
        s_base = base;
        eacham_pair_result_t r;
        r.n12 = n12; r.n21 = n21; r.n_mutual = gated ? 0u : total;
        r.flags = (gated ? EACHAM_PAIR_GATED : 0u) | (connected ? EACHAM_PAIR_CONNECTED : 0u);
        r.offset = base; r.count = emit ? total : 0u;
        *p.result = r;
    }
    __syncthreads();
    if (emit && s_base + total <= p.matches_cap) {
        unsigned long long o = s_base + excl;
        for (uint32_t a = lo; a < hi; ++a) {
            const uint32_t b = p.m12[a];
            if ((b != EACHAM_NONE) && (!p.cross_check || p.m21[b] == a)) {
                eacham_match_t m; m.query = a; m.train = b;
                p.matches[o++] = m;
            }
        }
    }
}

}  // namespace eacham
