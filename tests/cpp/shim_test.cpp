// Exercises the C++ side of the drop-in (include/eacham/*.h) the way /root/reference/apps/sfm/main.cpp uses its matcher,
// with no Python in the process:
//   (1) the reference's own call shape, unchanged:  std::async(std::launch::async, &Matcher::Match, &matcher, d1, d2)
//       (main.cpp:107-108) from several threads on ONE matcher object (main.cpp:98-109);
//   (2) the batched MatchPairs() == main.cpp:111-146 applied to those maps;
//   (3) MatchPairs() on every visible GPU at once (eacham_gpu_multi_*: NCCL broadcast, sharded pairs) == one GPU;
//   (4) MatchPhase() (include/eacham/MatchPhaseGpu.h) on a Graph / Node pair shaped like the reference's, with a frame that
//       has no node (main.cpp:75) and the Factor.quality fix (Graph.h:39-40);
//   (5) the match-graph dump (include/eacham/MatchGraphIO.h) round trip;
//   (6) errors surface as exceptions.
// All checked against a brute-force restatement (test code). cv::Mat is a stub with the members the shim touches (no OpenCV C++
// in this image). Build: g++ -std=c++17 -O2 -Iinclude tests/cpp/shim_test.cpp -Leacham_b200 -leacham_gpu -lpthread
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <future>
#include <map>
#include <random>
#include <thread>
#include <vector>

namespace cv
{
struct Mat   // the cv::Mat members the shim touches
{
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    std::vector<unsigned char> store;
    int type() const { return 0; }   // CV_8U
    Mat() = default;
    Mat(int r, int c, size_t st) : rows(r), cols(c), step(st), store(static_cast<size_t>(r) * st + 1) { data = store.data(); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), store(o.store) { data = store.data(); }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; step = o.step; store = o.store; data = store.data(); return *this; }
    unsigned char* row(int r) { return data + static_cast<size_t>(r) * step; }
    const unsigned char* row(int r) const { return data + static_cast<size_t>(r) * step; }
};
}
#define EACHAM_HAVE_CV_MAT 1

#include "eacham/FeatureMatcherGpu.h"
#include "eacham/MatchGraphIO.h"
#include "eacham/MatchPhaseGpu.h"

using match_t = std::unordered_map<unsigned, unsigned>;

// ---- the slice of the reference's data model MatchPhase touches (shapes of modules/sfm/data/Node.h, Graph.h) ----
struct Factor { unsigned id; float quality; match_t matches; };
struct Node
{
    unsigned id;
    cv::Mat descriptors;
    std::map<unsigned, Factor> factors;
    const cv::Mat& GetDescriptors() const { return descriptors; }
    Factor& AddFactor(const Node* n) { if (!factors.count(n->id)) factors.insert({n->id, {n->id, -1.f, {}}}); return factors[n->id]; }
    Factor& GetFactor(unsigned i) { return factors.at(i); }
};
struct Graph
{
    std::map<unsigned, Node*> nodes;
    const std::map<unsigned, Node*>& GetNodes() { return nodes; }
    void Connect(Node* n1, Node* n2, match_t&& matches)
    {
        auto& factor = n1->AddFactor(n2);
        factor.matches = std::move(matches);
        factor.quality = matches.size();       // the reference's bug, reproduced: reads the moved-from map
    }
};
struct Frame { unsigned id; };

static int hamming(const unsigned char* a, const unsigned char* b)
{
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]);
    return d;
}

// FeatureMatcherFlann::Match with the exact matcher (strict <, ascending index) + ratio (float/float < 0.8)
static match_t cpu_match(const cv::Mat& q, const cv::Mat& t)
{
    match_t out;
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
    std::vector<cv::Mat> imgs;
    std::vector<std::vector<unsigned char>> pool(400, std::vector<unsigned char>(32));
    for (auto& p : pool) for (auto& b : p) b = static_cast<unsigned char>(rng());
    for (int k = 0; k < n_images; ++k)
    {
        const int rows = 500 + 37 * k;
        cv::Mat m(rows, 32, k % 2 ? 48 : 32);          // odd images have a padded step (ROI-like)
        for (int r = 0; r < rows; ++r)
        {
            unsigned char* dst = m.row(r);
            if (r < 200) { std::memcpy(dst, pool[(r * 7 + k * 13) % 400].data(), 32); for (int f = 0; f < 12; ++f) { int bit = rng() % 256; dst[bit / 8] ^= (1u << (bit % 8)); } }
            else for (int b = 0; b < 32; ++b) dst[b] = static_cast<unsigned char>(rng());
        }
        imgs.push_back(m);
    }
    eacham::FeatureMatcherGpu matcher{0.8f};

    // (1) the reference's call, verbatim but for the type name, from 4 threads on one matcher object
    const auto pairs = eacham::FeatureMatcherGpu::ExhaustivePairs(n_images);
    std::vector<match_t> got12(pairs.size()), got21(pairs.size());
    std::vector<std::thread> th;
    for (int w = 0; w < 4; ++w)
        th.emplace_back([&, w] {
            for (size_t k = w; k < pairs.size(); k += 4)
            {
                auto t12 = std::async(std::launch::async, &eacham::FeatureMatcherGpu::Match, &matcher, imgs[pairs[k].first], imgs[pairs[k].second]);
                got12[k] = t12.get();
                auto t21 = std::async(std::launch::async, &eacham::FeatureMatcherGpu::Match, &matcher, imgs[pairs[k].second], imgs[pairs[k].first]);
                got21[k] = t21.get();
            }
        });
    for (auto& t : th) t.join();
    int bad = 0;
    std::vector<match_t> want12(pairs.size()), want21(pairs.size());
    for (size_t k = 0; k < pairs.size(); ++k)
    {
        want12[k] = cpu_match(imgs[pairs[k].first], imgs[pairs[k].second]);
        want21[k] = cpu_match(imgs[pairs[k].second], imgs[pairs[k].first]);
        bad += (want12[k] != got12[k]) + (want21[k] != got21[k]);
    }
    const int bad_match = bad;

    // (2) the batched path == main.cpp:111-146 applied to those maps
    std::vector<match_t> best12(pairs.size()), best21(pairs.size());
    std::vector<bool> conn(pairs.size());
    int connected = 0;
    for (size_t k = 0; k < pairs.size(); ++k)
    {
        const bool gated = want12[k].size() < 30 || want21[k].size() < 30;
        if (!gated)
            for (const auto& [m1, m2] : want12[k])
                if (want21[k].count(m2) > 0 && want21[k].at(m2) == m1) { best12[k][m1] = m2; best21[k][m2] = m1; }
        conn[k] = best12[k].size() > 30;
        if (!conn[k]) { best12[k].clear(); best21[k].clear(); }
        connected += conn[k];
    }
    auto check_batch = [&](const std::vector<eacham::PairMatches>& res) {
        int b = 0;
        for (size_t k = 0; k < pairs.size(); ++k)
        {
            const bool gated = want12[k].size() < 30 || want21[k].size() < 30;
            b += (res[k].gated != gated) + (res[k].connected != conn[k]) + (res[k].bestMatches12 != best12[k]) + (res[k].bestMatches21 != best21[k]) +
                 (res[k].n12 != want12[k].size()) + (res[k].n21 != want21[k].size());
        }
        return b;
    };
    const auto res = matcher.MatchPairs(imgs, pairs);
    const int bad_batch = check_batch(res);

    // (3) every visible GPU in this process (NCCL broadcast of the arena, sharded pair list); with one GPU the multi-device
    //     code path still runs, on a single shard
    const int n_dev = eacham_gpu_device_count();
    std::vector<int> devices;
    for (int d = 0; d < std::max(n_dev, 1); ++d) devices.push_back(d);
    int bad_multi = 0;
    unsigned multi_devices = 0;
    {
        eacham_gpu_multi* mh = nullptr;
        std::vector<int32_t> d32(devices.begin(), devices.end());
        eacham_gpu_config cfg{};
        if (eacham_gpu_create_multi(d32.data(), static_cast<uint32_t>(d32.size()), &cfg, &mh) != EACHAM_OK) { std::printf("create_multi: %s\n", eacham_gpu_last_error()); return 1; }
        multi_devices = eacham_gpu_multi_device_count(mh);
        eacham_gpu_destroy_multi(mh);
        if (devices.size() > 1)
        {
            eacham::FeatureMatcherGpu many{0.8f, devices};
            bad_multi = check_batch(many.MatchPairs(imgs, pairs)) + (many.DeviceCount() != devices.size());
        }
    }

    // (4) the whole match phase on a reference-shaped graph; frame 3 never got a node
    std::vector<Node> storage(n_images);
    Graph graph;
    std::vector<Frame> frames;
    for (int k = 0; k < n_images; ++k)
    {
        frames.push_back({100u + k});
        storage[k].id = 100u + k; storage[k].descriptors = imgs[k];
        if (k != 3) graph.nodes[100u + k] = &storage[k];
    }
    const auto stats = eacham::MatchPhase(graph, frames, matcher);
    int bad_phase = (stats.frames != (size_t)n_images) + (stats.nodes != (size_t)n_images - 1) + (stats.pairs != (size_t)(n_images - 1) * (n_images - 2) / 2);
    size_t phase_connected = 0;
    for (size_t k = 0; k < pairs.size(); ++k)
    {
        const unsigned i = pairs[k].first, j = pairs[k].second;
        const bool expect = conn[k] && i != 3 && j != 3;
        const bool has = storage[i].factors.count(100u + j) > 0;
        bad_phase += has != expect;
        if (has && expect)
        {
            ++phase_connected;
            const Factor &f12 = storage[i].GetFactor(100u + j), &f21 = storage[j].GetFactor(100u + i);
            bad_phase += (f12.matches != best12[k]) + (f21.matches != best21[k]) + (f12.quality != static_cast<float>(best12[k].size())) +
                         (f21.quality != static_cast<float>(best21[k].size()));
        }
    }
    bad_phase += phase_connected != stats.connected;

    // (5) dump -> load -> same edges
    int bad_io = 0;
    {
        const auto g = eacham::ToMatchGraph(res, n_images);
        const std::string path = "/tmp/eacham_shim_test.graph";
        eacham::SaveMatchGraph(path, g);
        const auto g2 = eacham::LoadMatchGraph(path);
        bad_io += (g2.n_images != (uint64_t)n_images) + (g2.pairs.size() != pairs.size()) + (g2.matches.size() != g.matches.size());
        for (size_t k = 0; k < pairs.size() && !bad_io; ++k)
            bad_io += (g2.Connected(k) != conn[k]) + (g2.Best12(k) != best12[k]) + (g2.Best21(k) != best21[k]) + (g2.pairs[k].first != pairs[k].first);
        bool threw = false;
        try { eacham::LoadMatchGraph("/tmp/eacham_shim_test.missing"); } catch (const std::runtime_error&) { threw = true; }
        bad_io += !threw;
        std::remove(path.c_str());
    }

    // (6) errors surface as exceptions, not crashes
    bool threw = false;
    try { std::vector<std::pair<unsigned, unsigned>> bogus{{0, 99}}; matcher.MatchPairs(imgs, bogus); } catch (const std::runtime_error&) { threw = true; }
    bad = bad_match + bad_batch + bad_multi + bad_phase + bad_io;
    std::printf("shim_test: %zu pairs, %d connected; mismatches: Match via std::async %d, MatchPairs %d, multi-device (%u device%s) %d, MatchPhase %d, "
                "graph io %d; error path %s\n", pairs.size(), connected, bad_match, bad_batch, multi_devices, multi_devices == 1 ? "" : "s", bad_multi,
                bad_phase, bad_io, threw ? "ok" : "MISSING");
    return (bad == 0 && threw && connected > 0) ? 0 : 1;
}
