/*
 * eacham_gpu.h -- C ABI of the B200-native descriptor matcher (libeacham_gpu.so).
 *
 * This is the drop-in boundary for ONE path of fatlipp/eacham: the all-pairs descriptor matching phase.
 * Every entry point names the reference interface it replaces (paths relative to the reference root):
 *
 *   eacham_gpu_match        <- FeatureMatcherFlann::Match      modules/base/features/FeatureMatcherFlann.cpp:14-30
 *                              (IFeatureMatcher<T>::Match       modules/base/features/IFeatureMatcher.h:18-19)
 *   eacham_gpu_match_pairs  <- the pair loop + cross-check      apps/sfm/main.cpp:84-147
 *   eacham_gpu_set_descriptors(_batch) / eacham_gpu_commit
 *                           <- Node::GetDescriptors() feeding Match   modules/sfm/data/Node.h:136-139, apps/sfm/main.cpp:107-108
 *   eacham_match_t          <- one entry of match_t              modules/sfm/data/Types.h:34
 *
 * Conventions: plain pointers and sizes only; no allocation crosses the ABI (outputs go into caller buffers with
 * capacity in / count out); every function returns an eacham_status (0 = ok, <0 = error) and never throws;
 * eacham_gpu_last_error() returns a thread-local message for the last failure on the calling thread.
 * There is NO CPU fallback: without a usable CUDA device eacham_gpu_create fails with EACHAM_ERR_NO_DEVICE.
 *
 * Thread-safety: a handle may be used from several threads at once (the reference calls Match concurrently on one
 * matcher object from TBB workers, apps/sfm/main.cpp:98-109); calls on one handle are serialised internally.
 */
#ifndef EACHAM_GPU_H
#define EACHAM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EACHAM_GPU_ABI_VERSION 2

#if defined(__GNUC__)
#define EACHAM_API __attribute__((visibility("default")))
#else
#define EACHAM_API
#endif

typedef enum eacham_status {
    EACHAM_OK = 0,
    EACHAM_ERR_INVALID_ARG = -1,
    EACHAM_ERR_NO_DEVICE = -2,      /* no CUDA device / wrong architecture: there is no CPU fallback          */
    EACHAM_ERR_CUDA = -3,           /* a CUDA runtime call failed; see eacham_gpu_last_error()                 */
    EACHAM_ERR_OUT_OF_MEMORY = -4,
    EACHAM_ERR_BUFFER_TOO_SMALL = -5, /* caller buffer too small; the required count is reported back           */
    EACHAM_ERR_NOT_COMMITTED = -6,  /* match_pairs before commit, or an image id without descriptors            */
    EACHAM_ERR_TOO_LARGE = -7,      /* more descriptors per image than the kernels support (65535)              */
    EACHAM_ERR_KIND_MISMATCH = -8,  /* a pair mixes ORB and SIFT images                                          */
    EACHAM_ERR_NCCL = -9            /* multi-device: libnccl.so.2 missing or an NCCL call failed                  */
} eacham_status;

/* Descriptor kinds. ORB256: rows of 32 bytes (8 x int32, modules/base/tools/Tools3d.h:46-63), Hamming distance.
 * F32X128: rows of 128 float32 (cv::SIFT output, modules/base/features/FeatureExtractorSift.cpp:14-26), L2.  */
typedef enum eacham_kind {
    EACHAM_KIND_ORB256 = 0,
    EACHAM_KIND_F32X128 = 1
} eacham_kind;

#define EACHAM_NONE 0xFFFFFFFFu

typedef struct eacham_gpu_handle eacham_gpu_handle;

#define EACHAM_CFG_SIFT_EXACT_FP32 1u  /* SIFT pairs on the all-FP32 kernels (no tensor cores)                          */
#define EACHAM_CFG_ORB_POPC 2u         /* ORB pairs on the XOR+POPC kernel instead of the default tensor-core engine (bits as FP8 0/1, exact) */
#define EACHAM_CFG_ORB_TC_V1 4u        /* ORB pairs on the round-1 tensor-core kernel (F32 accumulators, 32-bit keys); for A/B measurements   */
#define EACHAM_CFG_ORB_TC_ALU_SORT 8u  /* default ORB engine with its sort-2 steps on the ALU pipe instead of the FMA pipe; for A/B measurements */
#define EACHAM_CFG_SIFT_TC_V1 128u     /* SIFT pairs on the round-1 tensor-core kernel (one MMA pass, REDUX column path); for A/B measurements */
#define EACHAM_CFG_MATCH_NO_CACHE 32u  /* eacham_gpu_match: do not keep descriptors on the device between calls (see eacham_gpu_match) */
#define EACHAM_CFG_MATCH_LEGACY 64u    /* eacham_gpu_match on the round-1 CUDA-core kernels (upload both images per call, one call at a time) */
#define EACHAM_CFG_MULTI_PARALLEL_H2D 16u /* eacham_gpu_create_multi: one H2D copy per device instead of H2D + NCCL broadcast (no NCCL needed) */

typedef struct eacham_gpu_config {
    int32_t device;               /* CUDA device ordinal                                                     */
    uint32_t max_images;          /* 0 = grow on demand                                                      */
    uint64_t match_buffer_entries;/* device-side capacity for compacted matches; 0 = sized per call          */
    uint32_t flags;               /* EACHAM_CFG_* bits                                                        */
} eacham_gpu_config;

/* {queryIdx -> trainIdx}: one entry of the reference's std::unordered_map<unsigned, unsigned>. */
typedef struct eacham_match_t {
    uint32_t query;
    uint32_t train;
} eacham_match_t;

/* An unordered image pair {first, second}; the reference enumerates both ordered pairs (main.cpp:84-92) and
 * recomputes the transposed distance matrix; here one entry covers both directions. */
typedef struct eacham_pair_t {
    uint32_t first;
    uint32_t second;
} eacham_pair_t;

typedef struct eacham_match_opts {
    double ratio;         /* Lowe ratio; the reference hard-codes the double literal 0.8 (FeatureMatcherFlann.cpp:23) */
    uint32_t min_dir;     /* per-direction gate: drop the pair if a direction has < min_dir matches (main.cpp:111) = 30 */
    uint32_t min_mutual;  /* connect iff mutual count > min_mutual (main.cpp:142, strict)                    = 30 */
    uint32_t cross_check; /* 1 = mutual filter of main.cpp:133-140; 0 = first->second direction only             */
    uint32_t emit_all;    /* 1 = also write matches of pairs that are not connected (tests); default 0           */
} eacham_match_opts;

#define EACHAM_PAIR_GATED 1u      /* a direction had fewer than min_dir ratio-passing matches (main.cpp:111-114) */
#define EACHAM_PAIR_CONNECTED 2u  /* mutual count > min_mutual: the reference calls Graph::Connect (main.cpp:142-146) */

typedef struct eacham_pair_result_t {
    uint32_t n12;       /* |matches12| after the ratio test, first -> second                                  */
    uint32_t n21;       /* |matches21| after the ratio test, second -> first                                  */
    uint32_t n_mutual;  /* |bestMatches12| (0 when gated)                                                     */
    uint32_t flags;     /* EACHAM_PAIR_*                                                                      */
    uint64_t offset;    /* index of this pair's first eacham_match_t in the output buffer                     */
    uint64_t count;     /* number of entries written for this pair (n_mutual if connected or emit_all, else 0) */
} eacham_pair_result_t;

/* Phase timings of the most recent match_pairs / commit on this handle, from CUDA events on the library's stream. */
typedef struct eacham_gpu_timing {
    float upload_ms;    /* descriptor H2D inside the last eacham_gpu_commit                                   */
    float pairs_h2d_ms; /* pair list H2D                                                                      */
    float kernel_ms;    /* matching kernels only (inputs resident in HBM)                                     */
    float d2h_ms;       /* results + compacted matches D2H                                                    */
    uint32_t kernel_launches; /* kernels launched by the last match_pairs                                     */
    float prep_ms;      /* tensor-core operand copy of the arena (bf16 / one e4m3 per bit), built once after a commit */
    uint32_t exact_fallbacks; /* F32X128 tensor engine: queries of the last batch re-done by the exact FP32 scan (see below) */
    float verify_ms;    /* eacham_gpu_verify_pairs: the scoring kernel                                              */
} eacham_gpu_timing;

EACHAM_API int eacham_gpu_abi_version(void);
EACHAM_API int eacham_gpu_device_count(void);
EACHAM_API const char* eacham_gpu_last_error(void);
EACHAM_API void eacham_gpu_default_opts(eacham_match_opts* opts);   /* ratio 0.8, min_dir 30, min_mutual 30, cross_check 1 */

EACHAM_API int eacham_gpu_create(const eacham_gpu_config* cfg, eacham_gpu_handle** out);
EACHAM_API void eacham_gpu_destroy(eacham_gpu_handle* h);

/* Stage one image's descriptors (copied; the host pointer is not retained). rows may be 0.
 * row_stride_bytes >= 32 (ORB256) or >= 512 (F32X128): cv::Mat::step of a possibly non-continuous Mat. */
EACHAM_API int eacham_gpu_set_descriptors(eacham_gpu_handle* h, uint32_t image_id, int kind, const void* data, uint32_t rows,
                               size_t row_stride_bytes);
/* The same for image ids first_id .. first_id + n - 1 in one call: data[i] / rows[i] / row_stride_bytes[i] describe image first_id + i
 * (row_stride_bytes may be NULL = dense rows). Arguments are validated before anything is staged; the copies into the pinned staging
 * buffer run on several host threads. What a loader that already holds every Node's descriptors (apps/sfm/main.cpp:72-79) should call. */
EACHAM_API int eacham_gpu_set_descriptors_batch(eacham_gpu_handle* h, uint32_t first_id, uint32_t n, int kind, const void* const* data,
                                     const uint32_t* rows, const size_t* row_stride_bytes);
/* Declare an image's shape without data (ranks that receive the arena by broadcast). */
EACHAM_API int eacham_gpu_reserve(eacham_gpu_handle* h, uint32_t image_id, int kind, uint32_t rows);
/* Lay the staged images out in the device arena and upload whatever data was staged (one H2D copy). */
EACHAM_API int eacham_gpu_commit(eacham_gpu_handle* h);
/* Drop all images (keeps device allocations for reuse). */
EACHAM_API int eacham_gpu_clear(eacham_gpu_handle* h);

/* Device arena after commit: the plumbing layer (torch.distributed / NCCL) broadcasts these bytes between ranks. */
EACHAM_API int eacham_gpu_arena(eacham_gpu_handle* h, void** device_ptr, size_t* bytes);
EACHAM_API int eacham_gpu_image_info(eacham_gpu_handle* h, uint32_t image_id, int* kind, uint32_t* rows, size_t* arena_offset);

/* One direction, one pair: the reference's Match(). Synchronous, thread-safe. out gets <= cap entries, *n_out the
 * number of ratio-passing matches (sorted by query index); EACHAM_ERR_BUFFER_TOO_SMALL if *n_out > cap.
 * Runs on the tensor-core engines, one CTA per 128 query rows, and up to 8 calls proceed concurrently on their own streams (the
 * reference calls Match from TBB workers, apps/sfm/main.cpp:98-109). Descriptor matrices stay resident on the device between
 * calls: the reference passes the same Node-owned cv::Mat for an image in every pair it takes part in (main.cpp:107-108), so an
 * image is uploaded and converted once, not once per pair. A cached copy is recognised by (pointer, rows, stride, kind) AND a hash
 * of 16 sampled rows; a caller that rewrites a descriptor matrix in place between calls without touching any sampled row must
 * create the handle with EACHAM_CFG_MATCH_NO_CACHE. */
EACHAM_API int eacham_gpu_match(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride_bytes,
                     const void* train, uint32_t t_rows, size_t t_stride_bytes, double ratio, eacham_match_t* out,
                     size_t cap, size_t* n_out);

/* kNN(k=2) only (what cv::DescriptorMatcher::knnMatch returns, FeatureMatcherFlann.cpp:17): idx[q_rows][2]
 * (-1 = absent) and dist[q_rows][2] (float, as DMatch.distance; +inf = absent). */
EACHAM_API int eacham_gpu_knn2(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride_bytes,
                    const void* train, uint32_t t_rows, size_t t_stride_bytes, int32_t* idx, float* dist);

/* Exactness of the batched path.
 * ORB256: bit-exact on every engine -- indices, counts, flags and match lists equal cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + the
 *   reference's ratio / gate / mutual code. The tensor-core engines carry no indices in their hot loop and rely on a match being a
 *   strict unique minimum (true for ratio <= 1); a call with ratio > 1 runs on the packed-key XOR+POPC kernels instead.
 * F32X128: candidates are scored as a bf16 GEMM (exact for integer-valued rows, which is what cv::SIFT emits) and re-ranked in exact
 *   FP32. With |score - |a-b|^2/2| <= E for every train row, where
 *       E = delta (2 d1 + delta) / 2 + 2^-15 d1^2 / 2 + 2^-16 (max|a|^2 + max|b|^2) / 2,   delta = 2^-9 (max|a| + max|b|)
 *   (maxima over the rows of the query's / the train image)
 *   (first and last term 0 when both images are bf16-exact), the reference's ratio lies in [d0/d1, d0/sqrt(d1^2 - 4E)]. Only if that
 *   interval straddles `ratio` (or the best itself could be a non-candidate) is the query re-done by an exact FP32 scan over all train
 *   rows (eacham_gpu_timing.exact_fallbacks counts them). The returned match sets are therefore the exact matcher's for ANY float
 *   input; epsilon (relative width 3e-5 of the ratio for integer-valued rows, about 1.4 % for unit-norm floats) only decides how
 *   many queries take the slow path. Distances: eacham_gpu_knn2 / eacham_gpu_debug_pair_knn2. */

/* The batched path: for every unordered pair both directions + ratio + gates + mutual filter on the GPU
 * (apps/sfm/main.cpp:84-147). res has n_pairs entries; matches of pair p are buf[res[p].offset .. +count),
 * sorted by query index, query = row in image `first`, train = row in image `second`.
 * *buf_used = entries needed; EACHAM_ERR_BUFFER_TOO_SMALL (res still valid, buf partially filled) if > buf_cap. */
EACHAM_API int eacham_gpu_match_pairs(eacham_gpu_handle* h, const eacham_pair_t* pairs, size_t n_pairs,
                           const eacham_match_opts* opts, eacham_pair_result_t* res, eacham_match_t* buf,
                           size_t buf_cap, size_t* buf_used);

/* Debug view of the batched tensor-core path for ONE F32X128 pair: the kNN(k=2) its ratio test saw, for the rows of `first`
 * against `second` (idx12[n1][2], dist12[n1][2]) and the other way round (idx21[n2][2], dist21[n2][2]); -1 / +inf = absent.
 * dist is the reference's float sqrt(sum (a-b)^2). Where the certainty check (below) sends a query to the exact scan the entries
 * are the exact matcher's; elsewhere entry 1 is the scorer's second candidate, whose distance can exceed the true second-best
 * by at most the stated bound -- never enough to change the ratio test's outcome. */
EACHAM_API int eacham_gpu_debug_pair_knn2(eacham_gpu_handle* h, uint32_t first, uint32_t second, const eacham_match_opts* opts,
                                          int32_t* idx12, float* dist12, int32_t* idx21, float* dist21);

/* Same computation with inputs AND outputs resident in HBM: enqueue, wait, no D2H of matches.
 * Used by bench.py for the device-resident throughput figure; results stay readable via _fetch. */
EACHAM_API int eacham_gpu_match_pairs_device(eacham_gpu_handle* h, const eacham_pair_t* pairs, size_t n_pairs,
                                  const eacham_match_opts* opts, size_t* total_matches);
EACHAM_API int eacham_gpu_fetch_results(eacham_gpu_handle* h, eacham_pair_result_t* res, size_t n_pairs, eacham_match_t* buf,
                             size_t buf_cap, size_t* buf_used);

/* Device addresses of the last batch's outputs (results[n_pairs], matches[n_matches]); valid until the next
 * match_pairs call on this handle. Lets the plumbing layer gather shards over NVLink without a host round trip. */
EACHAM_API int eacham_gpu_device_results(eacham_gpu_handle* h, void** results, void** matches, size_t* n_pairs, size_t* n_matches);

/* ---- several devices of one box in ONE process (the reference is a single-process C++ app, apps/sfm/main.cpp) ----------
 * The pair list shards with no data-path exchange: every device holds the whole arena, matches its share of
 * the pair list (whole 16 x 16 blocks of the image x image grid, so that its working set stays in L2) and copies its own shard of the results into its slice of the caller's buffers over its own PCIe link.
 * eacham_gpu_multi_commit = one H2D copy to devices[0] + ONE ncclBroadcast of the arena over NVLink (ncclCommInitAll, one stream
 * per device; libnccl.so.2 is bound at run time). Results are exactly those of the single-device calls, in input order:
 * buf holds device 0's matches, then device 1's, ...; res[k].offset indexes into buf. */
typedef struct eacham_gpu_multi eacham_gpu_multi;

typedef struct eacham_gpu_multi_timing {
    float upload_ms;      /* wall clock: staging -> devices[0] (layout + H2D) inside the last commit                   */
    float broadcast_ms;   /* wall clock: arena layout on the other devices + NCCL broadcast (or the per-device H2D)     */
    float match_ms;       /* wall clock: pair sharding + all devices' kernels (the slowest device)                      */
    float d2h_ms;         /* wall clock: all devices' result copies + re-interleaving                                   */
    float kernel_ms_max;  /* CUDA events: matching kernels, max over devices                                            */
    float prep_ms_max;    /* CUDA events: tensor-core operand copy, max over devices                                    */
    uint32_t kernel_launches; /* summed over devices                                                                    */
    uint32_t reserved;
} eacham_gpu_multi_timing;

EACHAM_API int eacham_gpu_create_multi(const int32_t* devices, uint32_t n_devices, const eacham_gpu_config* cfg, eacham_gpu_multi** out); /* cfg->device ignored */
EACHAM_API void eacham_gpu_destroy_multi(eacham_gpu_multi* m);
EACHAM_API uint32_t eacham_gpu_multi_device_count(eacham_gpu_multi* m);
EACHAM_API int eacham_gpu_multi_set_descriptors(eacham_gpu_multi* m, uint32_t image_id, int kind, const void* data, uint32_t rows,
                                                size_t row_stride_bytes);
EACHAM_API int eacham_gpu_multi_set_descriptors_batch(eacham_gpu_multi* m, uint32_t first_id, uint32_t n, int kind, const void* const* data,
                                           const uint32_t* rows, const size_t* row_stride_bytes);
EACHAM_API int eacham_gpu_multi_clear(eacham_gpu_multi* m);
EACHAM_API int eacham_gpu_multi_commit(eacham_gpu_multi* m);
EACHAM_API int eacham_gpu_multi_match_pairs(eacham_gpu_multi* m, const eacham_pair_t* pairs, size_t n_pairs, const eacham_match_opts* opts,
                                            eacham_pair_result_t* res, eacham_match_t* buf, size_t buf_cap, size_t* buf_used);
EACHAM_API int eacham_gpu_multi_last_timing(eacham_gpu_multi* m, eacham_gpu_multi_timing* t);
/* Page-locked host memory usable by every device (result buffers copied at full PCIe rate); NULL on failure. */
EACHAM_API void* eacham_gpu_host_alloc(size_t bytes);
EACHAM_API void eacham_gpu_host_free(void* p);

/* ---- next on the path: geometric verification of the matched pairs ------------------------------------
 * ReconstructionManager::RecoverPoseTwoView (modules/sfm/reconstruction/ReconstructionManager.cpp:47-86) runs
 * cv::findEssentialMat(LMEDS) and cv::findHomography(LMEDS) on every factor's matched keypoints and counts each model's inliers;
 * FindBestPair does so twice per factor (modules/sfm/utils/Utils.h:24-68). Both are hypothesise-and-score loops. This call does
 * the scoring -- residual of every match under every hypothesis, least median of squares, sigma and inlier mask exactly as
 * OpenCV's LMedS registrator defines them -- for ALL pairs of the last eacham_gpu_match_pairs* batch on this handle, straight from
 * the device-resident match lists and the keypoints set below. Hypotheses come from the caller (minimal solvers stay on the host). */
typedef enum eacham_model {
    EACHAM_MODEL_ESSENTIAL = 0,   /* 3x3 E on normalised coordinates ((u - cx) / focal, (v - cy) / focal): Sampson-type error of five-point.cpp, double */
    EACHAM_MODEL_HOMOGRAPHY = 1   /* 3x3 H on pixel coordinates: forward transfer error of fundam.cpp, float                                           */
} eacham_model;

typedef struct eacham_verify_opts {
    double focal, cx, cy;     /* ESSENTIAL: as cv::findEssentialMat(points1, points2, focal, pp, ...) (ReconstructionManager.cpp:58-61) */
    uint32_t n_hyp;           /* hypotheses per pair                                                                                     */
    uint32_t shared;          /* 1 = one set of n_hyp hypotheses for all pairs ([n_hyp][9]); 0 = per pair ([n_pairs][n_hyp][9])          */
} eacham_verify_opts;

typedef struct eacham_verify_result {
    uint32_t best;        /* index of the hypothesis with the least median (first on ties)                          */
    uint32_t n_inliers;   /* matches with err <= sigma^2 under it: the reference's E_Inliers / H_Inliers              */
    float median;         /* its median squared error                                                               */
    float sigma;          /* 2.5 * 1.4826 * (1 + 5 / (n - modelPoints)) * sqrt(median), at least 0.001                */
} eacham_verify_result;

/* Keypoints of one image: rows x (x, y) float32, row r belongs to descriptor row r (Node::GetKeyPoint, ReconstructionManager.cpp:15-30). */
EACHAM_API int eacham_gpu_set_keypoints(eacham_gpu_handle* h, uint32_t image_id, const float* xy, uint32_t rows, size_t row_stride_bytes);
/* hyps: row-major 3x3 doubles. res[n_pairs] (pairs of the last batch, input order). medians ([n_pairs][n_hyp]) and mask (one byte per
 * entry of the last batch's match buffer, 1 = inlier of the pair's best hypothesis) may be NULL. Pairs with no more matches than
 * the model's minimal sample (5 / 4) get zeros. At most 8192 matches per pair. */
EACHAM_API int eacham_gpu_verify_pairs(eacham_gpu_handle* h, int model, const double* hyps, const eacham_verify_opts* opts,
                                       eacham_verify_result* res, float* medians, uint8_t* mask);

EACHAM_API int eacham_gpu_last_timing(eacham_gpu_handle* h, eacham_gpu_timing* t);
/* Write `bytes` of device memory (L2 flush between timed benchmark iterations). */
EACHAM_API int eacham_gpu_flush_l2(eacham_gpu_handle* h, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* EACHAM_GPU_H */
