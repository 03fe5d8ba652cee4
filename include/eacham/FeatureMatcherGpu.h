// FeatureMatcherGpu.h -- header-only C++ shim over the C ABI (include/eacham_gpu.h).
//
// Keeps the shape of the reference's matcher so apps/sfm/main.cpp compiles with the type swapped:
//   class FeatureMatcherFlann            /root/reference/modules/base/features/FeatureMatcherFlann.h:11-24
//     ctor (const float inliersRatio)    :16
//     MatchType Match(const cv::Mat&, const cv::Mat&)   :19,   MatchType = std::unordered_map<unsigned, unsigned>  :14
//   IFeatureMatcher<T>::Match            /root/reference/modules/base/features/IFeatureMatcher.h:18-19
// and adds MatchPairs(), which subsumes the pair loop + cross-check of /root/reference/apps/sfm/main.cpp:84-147.
//
// `Mat` is any type with the cv::Mat members used here: `rows`, `cols`, `type()`, `step` (convertible to size_t,
// bytes per row) and `ptr<T>()`/`data`. With OpenCV present use cv::Mat directly; the unit test uses a stub.
// Errors from the C ABI become std::runtime_error (the reference propagates cv::Exception the same way).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../eacham_gpu.h"

namespace eacham
{

struct PairMatches
{
    unsigned first = 0, second = 0;
    unsigned n12 = 0, n21 = 0;   // |matches12|, |matches21| after the ratio test (main.cpp:107-114)
    bool gated = false;          // a direction had < 30 matches (main.cpp:111)
    bool connected = false;      // |bestMatches12| > 30 (main.cpp:142): the reference calls Graph::Connect both ways
    std::unordered_map<unsigned, unsigned> bestMatches12;   // first -> second (main.cpp:137)
    std::unordered_map<unsigned, unsigned> bestMatches21;   // second -> first (main.cpp:138)
};

class FeatureMatcherGpu
{
public:
    using MatchType = std::unordered_map<unsigned, unsigned>;

public:
    // inliersRatio is stored and, like the reference (FeatureMatcherFlann.cpp:23 hard-codes 0.8), not used.
    // flags: EACHAM_CFG_* bits (0 = defaults: tensor-core engine for ORB and SIFT; EACHAM_CFG_ORB_POPC selects the
    // XOR+POPC kernel for ORB pairs, EACHAM_CFG_SIFT_EXACT_FP32 the all-FP32 SIFT kernels). Results do not depend on it.
    explicit FeatureMatcherGpu(const float inliersRatio, const int device = 0, const unsigned flags = 0)
        : inliersRatio{inliersRatio}
    {
        eacham_gpu_config cfg{};
        cfg.device = device;
        cfg.flags = flags;
        Check(eacham_gpu_create(&cfg, &handle));
        eacham_gpu_default_opts(&opts);
    }

    ~FeatureMatcherGpu() { eacham_gpu_destroy(handle); }
    FeatureMatcherGpu(const FeatureMatcherGpu&) = delete;
    FeatureMatcherGpu& operator=(const FeatureMatcherGpu&) = delete;

public:
    // One direction, one pair; safe to call concurrently on one object (main.cpp:98-109 does).
    template <typename Mat>
    MatchType Match(const Mat& descriptor1, const Mat& descriptor2)
    {
        const int kind = KindOf(descriptor1);
        if (KindOf(descriptor2) != kind) throw std::runtime_error("FeatureMatcherGpu::Match: descriptor types differ");
        std::vector<eacham_match_t> out(static_cast<size_t>(descriptor1.rows > 0 ? descriptor1.rows : 1));
        size_t n = 0;
        Check(eacham_gpu_match(handle, kind, descriptor1.data, static_cast<uint32_t>(descriptor1.rows), Step(descriptor1),
                               descriptor2.data, static_cast<uint32_t>(descriptor2.rows), Step(descriptor2), opts.ratio,
                               out.data(), out.size(), &n));
        MatchType matchesPair;
        matchesPair.reserve(n);
        for (size_t i = 0; i < n; ++i) matchesPair.insert({out[i].query, out[i].train});
        return matchesPair;
    }

    // Replaces main.cpp:84-147: descriptors[k] belongs to image id k; `pairs` holds each unordered pair once.
    template <typename Mat>
    std::vector<PairMatches> MatchPairs(const std::vector<Mat>& descriptors,
                                        const std::vector<std::pair<unsigned, unsigned>>& pairs)
    {
        Check(eacham_gpu_clear(handle));
        for (size_t k = 0; k < descriptors.size(); ++k)
            Check(eacham_gpu_set_descriptors(handle, static_cast<uint32_t>(k), KindOf(descriptors[k]), descriptors[k].data,
                                             static_cast<uint32_t>(descriptors[k].rows), Step(descriptors[k])));
        Check(eacham_gpu_commit(handle));
        std::vector<eacham_pair_t> p(pairs.size());
        for (size_t k = 0; k < pairs.size(); ++k) { p[k].first = pairs[k].first; p[k].second = pairs[k].second; }
        std::vector<eacham_pair_result_t> res(pairs.size());
        std::vector<eacham_match_t> buf(pairs.size() * 192 + 1024);
        size_t used = 0;
        int rc = eacham_gpu_match_pairs(handle, p.data(), p.size(), &opts, res.data(), buf.data(), buf.size(), &used);
        if (rc == EACHAM_ERR_BUFFER_TOO_SMALL)
        {
            buf.resize(used);
            rc = eacham_gpu_fetch_results(handle, res.data(), res.size(), buf.data(), buf.size(), &used);
        }
        Check(rc);
        std::vector<PairMatches> out(pairs.size());
        for (size_t k = 0; k < pairs.size(); ++k)
        {
            PairMatches& m = out[k];
            m.first = pairs[k].first; m.second = pairs[k].second;
            m.n12 = res[k].n12; m.n21 = res[k].n21;
            m.gated = (res[k].flags & EACHAM_PAIR_GATED) != 0;
            m.connected = (res[k].flags & EACHAM_PAIR_CONNECTED) != 0;
            for (uint64_t e = 0; e < res[k].count; ++e)
            {
                const eacham_match_t& mm = buf[res[k].offset + e];
                m.bestMatches12[mm.query] = mm.train;
                m.bestMatches21[mm.train] = mm.query;
            }
        }
        return out;
    }

    // All unordered pairs of n images: one entry per two ordered pairs of main.cpp:84-92.
    static std::vector<std::pair<unsigned, unsigned>> ExhaustivePairs(const unsigned n)
    {
        std::vector<std::pair<unsigned, unsigned>> pairs;
        for (unsigned i = 0; i < n; ++i)
            for (unsigned j = i + 1; j < n; ++j) pairs.push_back({i, j});
        return pairs;
    }

    eacham_match_opts& Options() { return opts; }

private:
    template <typename Mat>
    static int KindOf(const Mat& m)
    {
        // cv::Mat::type(): CV_8U == 0 (ORB, 32 columns), CV_32F == 5 (SIFT, 128 columns)
        if (m.type() == 0 && (m.rows == 0 || m.cols == 32)) return EACHAM_KIND_ORB256;
        if (m.type() == 5 && (m.rows == 0 || m.cols == 128)) return EACHAM_KIND_F32X128;
        throw std::runtime_error("FeatureMatcherGpu: unsupported descriptor matrix (need CV_8U x32 or CV_32F x128)");
    }

    template <typename Mat>
    static size_t Step(const Mat& m) { return static_cast<size_t>(m.step); }

    static void Check(const int rc)
    {
        if (rc != EACHAM_OK) throw std::runtime_error(std::string("eacham_gpu: ") + eacham_gpu_last_error());
    }

private:
    float inliersRatio;
    eacham_gpu_handle* handle = nullptr;
    eacham_match_opts opts{};
};

}
