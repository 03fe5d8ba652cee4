/*
 * oracle/match_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's pairwise descriptor matching path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, link, load or call anything in this directory; the product (eacham_b200/) never does.
 *
 * What it restates (reference file:line, all under /root/reference):
 *   - knn (k=2) brute force: the call `mather->knnMatch(d1, d2, matches, 2)`
 *     modules/base/features/FeatureMatcherFlann.cpp:17.  The arithmetic lives in OpenCV
 *     (un-vendored dependency, pinned opencv/4.5.5 in conanfile.txt:3), whose exact matcher
 *     cv::BFMatcher::knnMatchImpl -> cv::batchDistance (modules/core/src/batch_distance.cpp) does, per
 *     query row: buf[j] = dist(q_i, t_j) for all j, then for j ascending
 *         if (buf[j] < dist[K-1]) { k = K-2; while (k >= 0 && buf[j] < dist[k]) shift; insert at k+1 }
 *     i.e. STRICT comparisons, ascending distance, ties keep the LOWER train index first, and a row
 *     emits min(K, n_train) neighbours.  (FlannBased, which the reference literally names, is
 *     approximate, non-deterministic and rejects CV_8U -- SURVEY.md 0.1 -- so the exact matcher is the
 *     oracle north_star asks for.)
 *   - 256-bit Hamming distance over 8 x int32 words: modules/base/tools/Tools3d.h:46-63
 *     (tools::BinaryDescriptorDist); distance is converted int -> float for DMatch.distance.
 *   - L2 distance for SIFT: OpenCV NORM_L2 = sqrtf(sum_k (a_k-b_k)^2) with a float accumulator.
 *   - Lowe ratio filter: modules/base/features/FeatureMatcherFlann.cpp:21-27
 *         if (m[0].distance / m[1].distance < 0.8) map[queryIdx] = trainIdx
 *     float / float, compared with the DOUBLE literal 0.8.  0/0 = NaN compares false (rejected).
 *     m[1] is unguarded in the reference (UB when train has one row); here such rows are rejected.
 *   - pair logic (both directions, gates, mutual filter): apps/sfm/main.cpp:107-146.
 *
 * Parity pinning: the reference ships NO tests, fixtures or golden vectors for this path
 * (SURVEY.md section 4), so this file is pinned against OpenCV itself -- cv2 4.13.0 BFMatcher outputs generated
 * in the build container by tests/golden/make_golden.py and committed under tests/golden/ -- plus
 * hand-constructed known-answer cases (tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_NONE 0xFFFFFFFFu

/* Tools3d.h:46-63 -- 8 x (int32 XOR, popcount).  __builtin_popcount replaces the SWAR bithack; same value. */
static inline int hamming256(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        dist += __builtin_popcount(wa ^ wb);
    }
    return dist;
}

/* The SWAR form exactly as written at Tools3d.h:55-58, kept to prove it equals the builtin (tests). */
int oracle_hamming256_swar(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

int oracle_hamming256(const uint8_t *a, const uint8_t *b) { return hamming256(a, b); }

/* batchDistance's K=2 insertion (strict <, j ascending).  idx = -1 / dist = +inf when fewer than 2 rows. */
static inline void insert2(float d, int j, float *dist, int32_t *idx)
{
    if (d < dist[1]) {
        if (d < dist[0]) {
            dist[1] = dist[0]; idx[1] = idx[0];
            dist[0] = d;       idx[0] = j;
        } else {
            dist[1] = d;       idx[1] = j;
        }
    }
}

/* FeatureMatcherFlann.cpp:17 with BFMatcher(NORM_HAMMING): out_idx/out_dist are [nq][2]. */
void oracle_knn2_hamming(const uint8_t *q, int nq, int q_stride, const uint8_t *t, int nt, int t_stride,
                         int32_t *out_idx, float *out_dist)
{
    for (int i = 0; i < nq; ++i) {
        float dist[2] = {INFINITY, INFINITY};
        int32_t idx[2] = {-1, -1};
        const uint8_t *qi = q + (size_t)i * q_stride;
        for (int j = 0; j < nt; ++j)
            insert2((float)hamming256(qi, t + (size_t)j * t_stride), j, dist, idx);
        out_idx[2 * i] = idx[0]; out_idx[2 * i + 1] = idx[1];
        out_dist[2 * i] = dist[0]; out_dist[2 * i + 1] = dist[1];
    }
}

/* FeatureMatcherFlann.cpp:17 with BFMatcher(NORM_L2): float accumulator, sqrtf at the end. */
void oracle_knn2_l2(const float *q, int nq, int q_stride, const float *t, int nt, int t_stride, int dim,
                    int32_t *out_idx, float *out_dist)
{
    for (int i = 0; i < nq; ++i) {
        float dist[2] = {INFINITY, INFINITY};
        int32_t idx[2] = {-1, -1};
        const float *qi = q + (size_t)i * q_stride;
        for (int j = 0; j < nt; ++j) {
            const float *tj = t + (size_t)j * t_stride;
            float s = 0.f;
            for (int k = 0; k < dim; ++k) {
                float df = qi[k] - tj[k];
                s += df * df;
            }
            insert2(sqrtf(s), j, dist, idx);
        }
        out_idx[2 * i] = idx[0]; out_idx[2 * i + 1] = idx[1];
        out_dist[2 * i] = dist[0]; out_dist[2 * i + 1] = dist[1];
    }
}

/* FeatureMatcherFlann.cpp:21-27.  match[i] = train index or ORACLE_NONE.  Returns |map|. */
int oracle_ratio_filter(const int32_t *idx, const float *dist, int nq, double ratio, uint32_t *match)
{
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        match[i] = ORACLE_NONE;
        if (idx[2 * i] < 0 || idx[2 * i + 1] < 0) continue;      /* <2 train rows: reference is UB; reject */
        float r = dist[2 * i] / dist[2 * i + 1];                  /* float / float */
        if ((double)r < ratio) { match[i] = (uint32_t)idx[2 * i]; ++n; }
    }
    return n;
}

/*
 * apps/sfm/main.cpp:107-146 for one unordered pair, given both ratio-filtered directions.
 * m12[a] = b (a in image 1, b in image 2) or NONE; m21[b] = a or NONE.
 * Writes the mutual matches (a ascending) into out_a/out_b (capacity n1) and returns:
 *   -1 if |m12| < min_dir (main.cpp:111) or |m21| < min_dir (same gate reached from the (j,i) ordered pair),
 *   otherwise the mutual count; the pair is CONNECTED iff count > min_mutual (main.cpp:142, strict).
 */
int oracle_mutual(const uint32_t *m12, int n1, const uint32_t *m21, int n2, int min_dir,
                  uint32_t *out_a, uint32_t *out_b, int *n12_out, int *n21_out)
{
    int n12 = 0, n21 = 0;
    for (int a = 0; a < n1; ++a) n12 += (m12[a] != ORACLE_NONE);
    for (int b = 0; b < n2; ++b) n21 += (m21[b] != ORACLE_NONE);
    if (n12_out) *n12_out = n12;
    if (n21_out) *n21_out = n21;
    if (n12 < min_dir || n21 < min_dir) return -1;
    int n = 0;
    for (int a = 0; a < n1; ++a) {
        uint32_t b = m12[a];
        if (b != ORACLE_NONE && m21[b] == (uint32_t)a) {           /* main.cpp:133-140 */
            out_a[n] = (uint32_t)a; out_b[n] = b; ++n;
        }
    }
    return n;
}

/*
 * Whole pair, ORB: both directions (main.cpp:107 called for (i,j) and (j,i)), ratio, gates, mutual.
 * Returns mutual count (or -1 if a direction gate failed); *connected = count > min_mutual.
 */
int oracle_match_pair_hamming(const uint8_t *d1, int n1, const uint8_t *d2, int n2, double ratio, int min_dir,
                              int min_mutual, uint32_t *out_a, uint32_t *out_b, int *n12, int *n21, int *connected)
{
    int nmax = n1 > n2 ? n1 : n2;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nmax);
    float *dist = (float *)malloc(sizeof(float) * 2 * (size_t)nmax);
    uint32_t *m12 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n1 > 0 ? n1 : 1));
    uint32_t *m21 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n2 > 0 ? n2 : 1));
    oracle_knn2_hamming(d1, n1, 32, d2, n2, 32, idx, dist);
    oracle_ratio_filter(idx, dist, n1, ratio, m12);
    oracle_knn2_hamming(d2, n2, 32, d1, n1, 32, idx, dist);
    oracle_ratio_filter(idx, dist, n2, ratio, m21);
    int n = oracle_mutual(m12, n1, m21, n2, min_dir, out_a, out_b, n12, n21);
    if (connected) *connected = (n > min_mutual);
    free(idx); free(dist); free(m12); free(m21);
    return n;
}

int oracle_match_pair_l2(const float *d1, int n1, const float *d2, int n2, int dim, double ratio, int min_dir,
                         int min_mutual, uint32_t *out_a, uint32_t *out_b, int *n12, int *n21, int *connected)
{
    int nmax = n1 > n2 ? n1 : n2;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nmax);
    float *dist = (float *)malloc(sizeof(float) * 2 * (size_t)nmax);
    uint32_t *m12 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n1 > 0 ? n1 : 1));
    uint32_t *m21 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n2 > 0 ? n2 : 1));
    oracle_knn2_l2(d1, n1, dim, d2, n2, dim, dim, idx, dist);
    oracle_ratio_filter(idx, dist, n1, ratio, m12);
    oracle_knn2_l2(d2, n2, dim, d1, n1, dim, dim, idx, dist);
    oracle_ratio_filter(idx, dist, n2, ratio, m21);
    int n = oracle_mutual(m12, n1, m21, n2, min_dir, out_a, out_b, n12, n21);
    if (connected) *connected = (n > min_mutual);
    free(idx); free(dist); free(m12); free(m21);
    return n;
}
