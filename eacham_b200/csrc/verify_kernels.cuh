// verify_kernels.cuh -- batched hypothesis scoring for two-view geometric verification, sm_100a.
//
// The step right after matching in the reference: ReconstructionManager::RecoverPoseTwoView
// (/root/reference/modules/sfm/reconstruction/ReconstructionManager.cpp:47-86) runs cv::findEssentialMat(..., cv::LMEDS, ...) and
// cv::findHomography(..., cv::LMEDS, ...) on the matched keypoints of a factor and counts the inliers of each (E_Inliers, H_Inliers);
// FindBestPair calls it twice per factor (/root/reference/modules/sfm/utils/Utils.h:24-68). Both OpenCV calls are
// hypothesise-and-score loops; what they spend their time on is scoring: the residual of every match under every hypothesis, the
// median of the residuals (least median of squares) and the inlier mask of the winner. This kernel does exactly that part for
// every pair of a batch at once, reading the match lists where the matching kernels left them (device memory) and the keypoints
// from a device table. The arithmetic restates OpenCV 4.x (the reference's un-vendored dependency, pinned opencv/4.5.5):
//   essential   modules/calib3d/src/five-point.cpp  EMEstimatorCallback::computeError: double precision on normalised coordinates
//               x = ((u - cx) / f, (v - cy) / f, 1):   err = (x2' E x1)^2 / ((E x1)_0^2 + (E x1)_1^2 + (E' x2)_0^2 + (E' x2)_1^2), stored as float
//   homography  modules/calib3d/src/fundam.cpp      HomographyEstimatorCallback::computeError: single precision forward transfer error
//               ww = 1 / (H6 u + H7 v + H8);  err = ((H0 u + H1 v + H2) ww - u')^2 + ((H3 u + H4 v + H5) ww - v')^2
//   LMedS       modules/calib3d/src/ptsetreg.cpp    median of the sorted errors (mean of the two middle ones for even n), least median
//               wins (first on ties), sigma = 2.5 * 1.4826 * (1 + 5 / (n - modelPoints)) * sqrt(median), at least 0.001; inlier iff err <= sigma^2
// Every floating-point operation is an explicit round-to-nearest intrinsic in a fixed order (no FMA contraction), so the NumPy
// oracle (oracle/verify_oracle.py) reproduces the residuals bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eacham_gpu.h"

namespace eacham {
namespace verify {

constexpr int kThreads = 256;
constexpr uint32_t kMaxMatches = 8192;          // per pair: errors + points live in shared memory

struct Params {
    const eacham_pair_t* pairs;                 // the batch's pair list (device)
    const eacham_pair_result_t* results;        // per pair: offset / count into `matches`
    const eacham_match_t* matches;
    const float2* keypoints;                    // all images' keypoints, image i at kp_offset[i]
    const unsigned long long* kp_offset;
    const double* hyps;                         // [n_pairs][n_hyp][9] row-major, or [n_hyp][9] when shared
    uint32_t n_pairs, n_hyp, shared, model;
    double focal, cx, cy;
    eacham_verify_result* out;                  // [n_pairs]
    float* medians;                             // [n_pairs][n_hyp] or null
    uint8_t* mask;                              // [matches of the batch] or null
    uint32_t cap;                               // shared-memory capacity in matches (power of two, <= kMaxMatches)
};

__device__ __forceinline__ float essential_error(const double* __restrict__ E, float u1, float v1, float u2, float v2, double f, double cx, double cy) {
    const double x1 = __ddiv_rn(__dsub_rn((double)u1, cx), f), y1 = __ddiv_rn(__dsub_rn((double)v1, cy), f);
    const double x2 = __ddiv_rn(__dsub_rn((double)u2, cx), f), y2 = __ddiv_rn(__dsub_rn((double)v2, cy), f);
    // E x1 (rows of E) and E' x2 (columns of E), left-to-right sums
    const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1), __dmul_rn(E[1], y1)), E[2]);
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], x1), __dmul_rn(E[4], y1)), E[5]);
    const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], x1), __dmul_rn(E[7], y1)), E[8]);
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x2), __dmul_rn(E[3], y2)), E[6]);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x2), __dmul_rn(E[4], y2)), E[7]);
    const double x2tEx1 = __dadd_rn(__dadd_rn(__dmul_rn(x2, a0), __dmul_rn(y2, a1)), a2);
    const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(b0, b0)), __dmul_rn(b1, b1));
    return (float)__ddiv_rn(__dmul_rn(x2tEx1, x2tEx1), den);
}

__device__ __forceinline__ float homography_error(const float* __restrict__ H, float u1, float v1, float u2, float v2) {
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(H[6], u1), __fmul_rn(H[7], v1)), H[8]));
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(H[0], u1), __fmul_rn(H[1], v1)), H[2]), ww), u2);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(H[3], u1), __fmul_rn(H[4], v1)), H[5]), ww), v2);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// One CTA per pair (grid-stride). Shared memory: points[cap] (float4: u1 v1 u2 v2), err[cap], sorted[cap].
__global__ void __launch_bounds__(kThreads) verify_pairs_kernel(const Params p) {
    extern __shared__ __align__(16) uint8_t smem[];
    float4* pts = reinterpret_cast<float4*>(smem);
    float* err = reinterpret_cast<float*>(pts + p.cap);
    float* srt = err + p.cap;
    __shared__ double s_hyp[9];
    __shared__ float s_hypf[9];
    __shared__ float s_best_median;
    __shared__ uint32_t s_best, s_count;
    const int tid = threadIdx.x;
    const uint32_t model_points = p.model == EACHAM_MODEL_ESSENTIAL ? 5u : 4u;

    for (uint32_t pi = blockIdx.x; pi < p.n_pairs; pi += gridDim.x) {
        const eacham_pair_result_t r = p.results[pi];
        const uint32_t n = (uint32_t)min((unsigned long long)r.count, (unsigned long long)p.cap);
        eacham_verify_result res;
        res.best = 0; res.n_inliers = 0; res.median = 0.f; res.sigma = 0.f;
        if (n <= model_points) {                         // nothing to verify (not connected, or too few matches for a model)
            if (tid == 0) p.out[pi] = res;
            if (p.medians) for (uint32_t hh = tid; hh < p.n_hyp; hh += kThreads) p.medians[(size_t)pi * p.n_hyp + hh] = 0.f;
            if (p.mask) for (uint32_t i = tid; i < n; i += kThreads) p.mask[r.offset + i] = 0;
            continue;
        }
        const eacham_pair_t pr = p.pairs[pi];
        const float2* k1 = p.keypoints + p.kp_offset[pr.first];
        const float2* k2 = p.keypoints + p.kp_offset[pr.second];
        for (uint32_t i = tid; i < n; i += kThreads) {
            const eacham_match_t m = p.matches[r.offset + i];
            const float2 a = k1[m.query], b = k2[m.train];
            pts[i] = make_float4(a.x, a.y, b.x, b.y);
        }
        uint32_t npad = 1;
        while (npad < n) npad <<= 1;
        if (tid == 0) { s_best = 0; s_best_median = __int_as_float(0x7f800000); }
        __syncthreads();

        for (uint32_t hh = 0; hh <= p.n_hyp; ++hh) {     // pass n_hyp re-evaluates the winner for the mask
            const uint32_t which = hh < p.n_hyp ? hh : s_best;
            const double* H = p.hyps + ((size_t)(p.shared ? 0 : pi) * p.n_hyp + which) * 9;
            if (tid < 9) { s_hyp[tid] = H[tid]; s_hypf[tid] = (float)H[tid]; }
            __syncthreads();
            for (uint32_t i = tid; i < npad; i += kThreads) {
                float e = __int_as_float(0x7f800000);     // padding sorts last
                if (i < n) {
                    const float4 q = pts[i];
                    e = p.model == EACHAM_MODEL_ESSENTIAL ? essential_error(s_hyp, q.x, q.y, q.z, q.w, p.focal, p.cx, p.cy)
                                                          : homography_error(s_hypf, q.x, q.y, q.z, q.w);
                    if (!(e == e)) e = __int_as_float(0x7f7fffff);       // NaN (degenerate hypothesis): as bad as it gets, but sortable
                    err[i] = e;
                }
                srt[i] = e;
            }
            __syncthreads();
            if (hh == p.n_hyp) break;
            // bitonic sort of srt[0..npad)
            for (uint32_t k = 2; k <= npad; k <<= 1)
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    for (uint32_t i = tid; i < npad; i += kThreads) {
                        const uint32_t l = i ^ j;
                        if (l > i) {
                            const float a = srt[i], b = srt[l];
                            const bool up = (i & k) == 0;
                            if ((a > b) == up) { srt[i] = b; srt[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            if (tid == 0) {
                const float med = (n & 1u) ? srt[n / 2] : __fmul_rn(__fadd_rn(srt[n / 2 - 1], srt[n / 2]), 0.5f);      // float + float, then halved (exact)
                if (p.medians) p.medians[(size_t)pi * p.n_hyp + hh] = med;
                if (med < s_best_median) { s_best_median = med; s_best = hh; }
            }
            __syncthreads();
        }
        // err[] now holds the winner's residuals
        const double sig0 = 2.5 * 1.4826 * (1.0 + 5.0 / (double)(n - model_points)) * sqrt((double)s_best_median);
        const double sigma = sig0 > 0.001 ? sig0 : 0.001;
        const float thr = (float)(sigma * sigma);          // findInliers: float t = (float)(thresh * thresh); err[i] <= t
        if (tid == 0) s_count = 0;
        __syncthreads();
        uint32_t mine = 0;
        for (uint32_t i = tid; i < n; i += kThreads) {
            const bool in = err[i] <= thr;
            mine += in;
            if (p.mask) p.mask[r.offset + i] = in ? 1 : 0;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((tid & 31) == 0 && mine) atomicAdd(&s_count, mine);
        __syncthreads();
        if (tid == 0) {
            res.best = s_best; res.n_inliers = s_count; res.median = s_best_median; res.sigma = (float)sigma;
            p.out[pi] = res;
        }
        __syncthreads();
    }
}

}  // namespace verify
}  // namespace eacham
