// FeatureMatcherGpu.h -- header-only C++ shim over the C ABI (include/eacham_gpu.h).
//
// Keeps the shape of the reference's matcher so apps/sfm/main.cpp compiles with ONLY the type swapped:
//   class FeatureMatcherFlann            /root/reference/modules/base/features/FeatureMatcherFlann.h:11-24
//     ctor (const float inliersRatio)    :16
//     MatchType Match(const cv::Mat&, const cv::Mat&)   :19,   MatchType = std::unordered_map<unsigned, unsigned>  :14
//   IFeatureMatcher<T>::Match            /root/reference/modules/base/features/IFeatureMatcher.h:18-19
// `Match` is a plain (non-template, non-overloaded) member, so `&FeatureMatcherGpu::Match` has one type and the reference's
//   std::async(std::launch::async, &FeatureMatcherFlann::Match, &matcher, d1, d2)      /root/reference/apps/sfm/main.cpp:107-108
// binds unchanged. MatchPairs() subsumes the pair loop + cross-check of /root/reference/apps/sfm/main.cpp:84-147 and, given more
// than one device, shards it over the GPUs of the box inside this process (eacham_gpu_multi_*).
//
// cv::Mat: with OpenCV on the include path the header pulls in <opencv2/core/mat.hpp> itself. Without OpenCV (this repo's test
// image) define EACHAM_HAVE_CV_MAT after declaring a `cv::Mat` with the members used here -- `rows`, `cols`, `type()`, `step`
// (convertible to size_t, bytes per row) and `data` -- which is what tests/cpp/shim_test.cpp does. MatchAny / MatchPairsAny are
// the same calls for any other Mat-like type. Errors from the C ABI become std::runtime_error (the reference propagates
// cv::Exception the same way).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#if !defined(EACHAM_HAVE_CV_MAT) && defined(__has_include)
#if __has_include(<opencv2/core/mat.hpp>)
#include <opencv2/core/mat.hpp>
#define EACHAM_HAVE_CV_MAT 1
#endif
#endif

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
    // flags: EACHAM_CFG_* bits (0 = defaults: tensor-core engines for ORB and SIFT pairs). ORB results do not depend on the
    // engine (bit-exact); SIFT results are bit-exact for integer-valued descriptors (what cv::SIFT emits) and agree within the
    // epsilon stated in include/eacham_gpu.h for general floats.
    explicit FeatureMatcherGpu(const float inliersRatio, const int device = 0, const unsigned flags = 0)
        : inliersRatio{inliersRatio}
    {
        eacham_gpu_config cfg{};
        cfg.device = device;
        cfg.flags = flags;
        Check(eacham_gpu_create(&cfg, &handle));
        eacham_gpu_default_opts(&opts);
    }

    // Several GPUs of one box in this process: Match() runs on devices[0]; MatchPairs() uploads once, broadcasts the arena
    // with NCCL and shards the pair list over all devices.
    FeatureMatcherGpu(const float inliersRatio, const std::vector<int>& devices, const unsigned flags = 0)
        : FeatureMatcherGpu(inliersRatio, devices.empty() ? 0 : devices[0], flags)
    {
        if (devices.size() > 1)
        {
            eacham_gpu_config cfg{};
            cfg.flags = flags;
            std::vector<int32_t> d(devices.begin(), devices.end());
            const int rc = eacham_gpu_create_multi(d.data(), static_cast<uint32_t>(d.size()), &cfg, &multi);
            if (rc != EACHAM_OK)
            {
                eacham_gpu_destroy(handle);
                Check(rc);
            }
        }
    }

    ~FeatureMatcherGpu()
    {
        eacham_gpu_destroy_multi(multi);
        eacham_gpu_destroy(handle);
    }
    FeatureMatcherGpu(const FeatureMatcherGpu&) = delete;
    FeatureMatcherGpu& operator=(const FeatureMatcherGpu&) = delete;

public:
#ifdef EACHAM_HAVE_CV_MAT
    // One direction, one pair: FeatureMatcherFlann::Match. Safe to call concurrently on one object (main.cpp:98-109 does).
    MatchType Match(const cv::Mat& descriptor1, const cv::Mat& descriptor2) { return MatchAny(descriptor1, descriptor2); }

    // Replaces main.cpp:84-147: descriptors[k] belongs to image id k; `pairs` holds each unordered pair once.
    std::vector<PairMatches> MatchPairs(const std::vector<cv::Mat>& descriptors, const std::vector<std::pair<unsigned, unsigned>>& pairs)
    {
        return MatchPairsAny(descriptors, pairs);
    }
#endif

    template <typename Mat>
    MatchType MatchAny(const Mat& descriptor1, const Mat& descriptor2)
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

    // all images through one batched staging call per run of equal descriptor kind (several host threads copy into pinned memory)
    template <typename Mat, typename Fn>
    static void Stage(const std::vector<Mat>& descriptors, Fn&& stage)
    {
        const size_t n = descriptors.size();
        std::vector<const void*> data(n);
        std::vector<uint32_t> rows(n);
        std::vector<size_t> steps(n);
        for (size_t k = 0; k < n; ++k)
        {
            data[k] = descriptors[k].data;
            rows[k] = static_cast<uint32_t>(descriptors[k].rows);
            steps[k] = Step(descriptors[k]);
        }
        for (size_t i = 0; i < n;)
        {
            size_t j = i;
            const int kind = KindOf(descriptors[i]);
            while (j < n && KindOf(descriptors[j]) == kind) ++j;
            Check(stage(static_cast<uint32_t>(i), static_cast<uint32_t>(j - i), kind, data.data() + i, rows.data() + i, steps.data() + i));
            i = j;
        }
    }

    template <typename Mat>
    std::vector<PairMatches> MatchPairsAny(const std::vector<Mat>& descriptors,
                                           const std::vector<std::pair<unsigned, unsigned>>& pairs)
    {
        std::vector<eacham_pair_t> p(pairs.size());
        for (size_t k = 0; k < pairs.size(); ++k) { p[k].first = pairs[k].first; p[k].second = pairs[k].second; }
        std::vector<eacham_pair_result_t> res(pairs.size());
        size_t used = 0;
        const size_t guess = pairs.size() * 192 + 1024;
        if (multi != nullptr)
        {
            Check(eacham_gpu_multi_clear(multi));
            Stage(descriptors, [this](uint32_t first, uint32_t n, int kind, const void* const* d, const uint32_t* r, const size_t* s)
                  { return eacham_gpu_multi_set_descriptors_batch(multi, first, n, kind, d, r, s); });
            Check(eacham_gpu_multi_commit(multi));
            PinnedMatches buf(guess);       // page-locked: every device copies its shard at full PCIe rate
            int rc = eacham_gpu_multi_match_pairs(multi, p.data(), p.size(), &opts, res.data(), buf.p, buf.n, &used);
            if (rc == EACHAM_ERR_BUFFER_TOO_SMALL)
            {
                buf = PinnedMatches(used);
                rc = eacham_gpu_multi_match_pairs(multi, p.data(), p.size(), &opts, res.data(), buf.p, buf.n, &used);
            }
            Check(rc);
            return Unpack(pairs, res, buf.p);
        }
        Check(eacham_gpu_clear(handle));
        Stage(descriptors, [this](uint32_t first, uint32_t n, int kind, const void* const* d, const uint32_t* r, const size_t* s)
              { return eacham_gpu_set_descriptors_batch(handle, first, n, kind, d, r, s); });
        Check(eacham_gpu_commit(handle));
        std::vector<eacham_match_t> buf(guess);
        int rc = eacham_gpu_match_pairs(handle, p.data(), p.size(), &opts, res.data(), buf.data(), buf.size(), &used);
        if (rc == EACHAM_ERR_BUFFER_TOO_SMALL)
        {
            buf.resize(used);
            rc = eacham_gpu_fetch_results(handle, res.data(), res.size(), buf.data(), buf.size(), &used);
        }
        Check(rc);
        return Unpack(pairs, res, buf.data());
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
    unsigned DeviceCount() const { return multi != nullptr ? eacham_gpu_multi_device_count(multi) : 1u; }

private:
    struct PinnedMatches
    {
        eacham_match_t* p = nullptr;
        size_t n = 0;
        explicit PinnedMatches(const size_t count) : p(static_cast<eacham_match_t*>(eacham_gpu_host_alloc((count ? count : 1) * sizeof(eacham_match_t)))), n(count)
        {
            if (p == nullptr) throw std::runtime_error(std::string("eacham_gpu: ") + eacham_gpu_last_error());
        }
        ~PinnedMatches() { eacham_gpu_host_free(p); }
        PinnedMatches(const PinnedMatches&) = delete;
        PinnedMatches& operator=(const PinnedMatches&) = delete;
        PinnedMatches& operator=(PinnedMatches&& o) noexcept
        {
            if (this != &o) { eacham_gpu_host_free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
            return *this;
        }
    };

    static std::vector<PairMatches> Unpack(const std::vector<std::pair<unsigned, unsigned>>& pairs,
                                           const std::vector<eacham_pair_result_t>& res, const eacham_match_t* buf)
    {
        std::vector<PairMatches> out(pairs.size());
        for (size_t k = 0; k < pairs.size(); ++k)
        {
            PairMatches& m = out[k];
            m.first = pairs[k].first; m.second = pairs[k].second;
            m.n12 = res[k].n12; m.n21 = res[k].n21;
            m.gated = (res[k].flags & EACHAM_PAIR_GATED) != 0;
            m.connected = (res[k].flags & EACHAM_PAIR_CONNECTED) != 0;
            m.bestMatches12.reserve(res[k].count);
            m.bestMatches21.reserve(res[k].count);
            for (uint64_t e = 0; e < res[k].count; ++e)
            {
                const eacham_match_t& mm = buf[res[k].offset + e];
                m.bestMatches12[mm.query] = mm.train;
                m.bestMatches21[mm.train] = mm.query;
            }
        }
        return out;
    }

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
    eacham_gpu_multi* multi = nullptr;
    eacham_match_opts opts{};
};

}
