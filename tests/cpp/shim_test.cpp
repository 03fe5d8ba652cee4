// Exercises include/eacham/FeatureMatcherGpu.h the way /root/reference/apps/sfm/main.cpp:70,98-147 uses its matcher:
// concurrent Match() calls on ONE matcher object from several threads, then the batched MatchPairs(), both
// checked against a brute-force restatement (test code). Mat is a cv::Mat-shaped stub (no OpenCV C++ in the image).
// Build: g++ -std=c++17 -O2 -Iinclude tests/cpp/shim_test.cpp -Leacham_b200 -leacham_gpu -lpthread
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>

#include "eacham/FeatureMatcherGpu.h"

struct Mat   // the cv::Mat members the shim touches
{
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    std::vector<unsigned char> store;
    int type() const { return 0; }   // CV_8U
    Mat(int r, int c, size_t st) : rows(r), cols(c), step(st), store(static_cast<size_t>(r) * st + 1) { data = store.data(); }
    unsigned char* row(int r) { return data + static_cast<size_t>(r) * step; }
    const unsigned char* row(int r) const { return data + static_cast<size_t>(r) * step; }
};

static int hamming(const unsigned char* a, const unsigned char* b)
{
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]);
    return d;
}

// FeatureMatcherFlann::Match with the exact matcher (strict <, ascending index) + ratio (float/float < 0.8)
static std::unordered_map<unsigned, unsigned> cpu_match(const Mat& q, const Mat& t)
{
    std::unordered_map<unsigned, unsigned> out;
    for (int i = 0; i < q.rows; ++i)
    {
        float d0 = 1e30f, d1 = 1e30f; int i0 = -1, i1 = -1;
        for (int j = 0; j < t.rows; ++j)
        {
            const float d = static_cast<float>(hamming(q.row(i), t.row(j)));
            if (d < d1) { if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; } else { d1 = d; i1 = j; } }
        }
        if (i1 >= 0 && d0 / d1 < 0.8) out[i] = i0;
    }
    return out;
}

int main()
{
    std::mt19937 rng(42);
    const int n_images = 6;
    std::vector<Mat> imgs;
    std::vector<std::vector<unsigned char>> pool(400, std::vector<unsigned char>(32));
    for (auto& p : pool) for (auto& b : p) b = static_cast<unsigned char>(rng());
    for (int k = 0; k < n_images; ++k)
    {
        const int rows = 500 + 37 * k;
        Mat m(rows, 32, k % 2 ? 48 : 32);          // odd images have a padded step (ROI-like)
        for (int r = 0; r < rows; ++r)
        {
            unsigned char* dst = m.row(r);
            if (r < 200) { std::memcpy(dst, pool[(r * 7 + k * 13) % 400].data(), 32); for (int f = 0; f < 12; ++f) { int bit = rng() % 256; dst[bit / 8] ^= (1u << (bit % 8)); } }
            else for (int b = 0; b < 32; ++b) dst[b] = static_cast<unsigned char>(rng());
        }
        imgs.push_back(std::move(m));
        imgs.back().data = imgs.back().store.data();
    }
    eacham::FeatureMatcherGpu matcher{0.8f};

    // (1) concurrent Match() on one object, as the reference's TBB workers do
    const auto pairs = eacham::FeatureMatcherGpu::ExhaustivePairs(n_images);
    std::vector<eacham::FeatureMatcherGpu::MatchType> got12(pairs.size()), got21(pairs.size());
    std::vector<std::thread> th;
    for (int w = 0; w < 4; ++w)
        th.emplace_back([&, w] {
            for (size_t k = w; k < pairs.size(); k += 4)
            {
                got12[k] = matcher.Match(imgs[pairs[k].first], imgs[pairs[k].second]);
                got21[k] = matcher.Match(imgs[pairs[k].second], imgs[pairs[k].first]);
            }
        });
    for (auto& t : th) t.join();
    int bad = 0;
    std::vector<std::unordered_map<unsigned, unsigned>> want12(pairs.size()), want21(pairs.size());
    for (size_t k = 0; k < pairs.size(); ++k)
    {
        want12[k] = cpu_match(imgs[pairs[k].first], imgs[pairs[k].second]);
        want21[k] = cpu_match(imgs[pairs[k].second], imgs[pairs[k].first]);
        bad += (want12[k] != got12[k]) + (want21[k] != got21[k]);
    }
    // (2) the batched path == main.cpp:111-146 applied to those maps
    const auto res = matcher.MatchPairs(imgs, pairs);
    int connected = 0;
    for (size_t k = 0; k < pairs.size(); ++k)
    {
        std::unordered_map<unsigned, unsigned> b12, b21;
        const bool gated = want12[k].size() < 30 || want21[k].size() < 30;
        if (!gated)
            for (const auto& [m1, m2] : want12[k])
                if (want21[k].count(m2) > 0 && want21[k].at(m2) == m1) { b12[m1] = m2; b21[m2] = m1; }
        const bool conn = b12.size() > 30;
        if (!conn) { b12.clear(); b21.clear(); }
        bad += (res[k].gated != gated) + (res[k].connected != conn) + (res[k].bestMatches12 != b12) + (res[k].bestMatches21 != b21) +
               (res[k].n12 != want12[k].size()) + (res[k].n21 != want21[k].size());
        connected += conn;
    }
    // (3) errors surface as exceptions, not crashes
    bool threw = false;
    try { std::vector<std::pair<unsigned, unsigned>> bogus{{0, 99}}; matcher.MatchPairs(imgs, bogus); } catch (const std::runtime_error&) { threw = true; }
    std::printf("shim_test: %zu pairs, %d connected, %d mismatches, error path %s\n", pairs.size(), connected, bad, threw ? "ok" : "MISSING");
    return (bad == 0 && threw && connected > 0) ? 0 : 1;
}
