// MatchGraphIO.h -- reader / writer of the on-disk match graph, C++ side of
// eacham_b200/graph_io.py. The reference keeps the result of its O(n^2) matching phase only in memory (Graph::Connect,
// /root/reference/modules/sfm/data/Graph.h:30-41; its only output file is transform.json, /root/reference/modules/sfm/utils/Saver.h:13-73);
// this dump makes the phase restartable: a later run loads the edges and calls Graph::Connect without matching again.
//
// Layout (little endian, identical to graph_io.py):
//     magic  "EACHAMG1"                       8 bytes
//     n_pairs u64, n_matches u64, n_images u64
//     pairs    [n_pairs]   eacham_pair_t        {first u32, second u32}
//     results  [n_pairs]   eacham_pair_result_t {n12 u32, n21 u32, n_mutual u32, flags u32, offset u64, count u64}
//     matches  [n_matches] eacham_match_t       {query u32, train u32}
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../eacham_gpu.h"

namespace eacham
{

struct MatchGraph
{
    uint64_t n_images = 0;
    std::vector<eacham_pair_t> pairs;
    std::vector<eacham_pair_result_t> results;     // one per pair; offset / count index into `matches`
    std::vector<eacham_match_t> matches;

    // What the reference hands to Graph::Connect for pair k (main.cpp:142-146): {a -> b} and {b -> a}. Empty unless connected.
    bool Connected(const size_t k) const { return (results[k].flags & EACHAM_PAIR_CONNECTED) != 0; }
    std::unordered_map<unsigned, unsigned> Best12(const size_t k) const
    {
        std::unordered_map<unsigned, unsigned> m;
        for (uint64_t e = 0; e < results[k].count; ++e) m[matches[results[k].offset + e].query] = matches[results[k].offset + e].train;
        return m;
    }
    std::unordered_map<unsigned, unsigned> Best21(const size_t k) const
    {
        std::unordered_map<unsigned, unsigned> m;
        for (uint64_t e = 0; e < results[k].count; ++e) m[matches[results[k].offset + e].train] = matches[results[k].offset + e].query;
        return m;
    }
};

namespace detail
{
struct File
{
    std::FILE* f = nullptr;
    File(const std::string& path, const char* mode) : f(std::fopen(path.c_str(), mode))
    {
        if (f == nullptr) throw std::runtime_error("MatchGraphIO: cannot open " + path);
    }
    ~File() { if (f != nullptr) std::fclose(f); }
    File(const File&) = delete;
    File& operator=(const File&) = delete;
};
}

inline void SaveMatchGraph(const std::string& path, const MatchGraph& g)
{
    static_assert(sizeof(eacham_pair_t) == 8 && sizeof(eacham_pair_result_t) == 32 && sizeof(eacham_match_t) == 8, "record layout");
    if (g.results.size() != g.pairs.size()) throw std::runtime_error("MatchGraphIO: one result per pair expected");
    for (const auto& r : g.results)
        if (r.offset + r.count > g.matches.size()) throw std::runtime_error("MatchGraphIO: a pair's match range lies outside the match buffer");
    detail::File out(path, "wb");
    const uint64_t head[3] = {g.pairs.size(), g.matches.size(), g.n_images};
    bool ok = std::fwrite("EACHAMG1", 1, 8, out.f) == 8 && std::fwrite(head, 8, 3, out.f) == 3;
    ok = ok && std::fwrite(g.pairs.data(), sizeof(eacham_pair_t), g.pairs.size(), out.f) == g.pairs.size();
    ok = ok && std::fwrite(g.results.data(), sizeof(eacham_pair_result_t), g.results.size(), out.f) == g.results.size();
    ok = ok && std::fwrite(g.matches.data(), sizeof(eacham_match_t), g.matches.size(), out.f) == g.matches.size();
    if (!ok) throw std::runtime_error("MatchGraphIO: short write to " + path);
}

inline MatchGraph LoadMatchGraph(const std::string& path)
{
    detail::File in(path, "rb");
    char magic[8];
    uint64_t head[3];
    if (std::fread(magic, 1, 8, in.f) != 8 || std::memcmp(magic, "EACHAMG1", 8) != 0) throw std::runtime_error("MatchGraphIO: " + path + " is not an eacham match graph");
    if (std::fread(head, 8, 3, in.f) != 3) throw std::runtime_error("MatchGraphIO: " + path + " is truncated");
    MatchGraph g;
    g.n_images = head[2];
    g.pairs.resize(head[0]); g.results.resize(head[0]); g.matches.resize(head[1]);
    bool ok = std::fread(g.pairs.data(), sizeof(eacham_pair_t), g.pairs.size(), in.f) == g.pairs.size();
    ok = ok && std::fread(g.results.data(), sizeof(eacham_pair_result_t), g.results.size(), in.f) == g.results.size();
    ok = ok && std::fread(g.matches.data(), sizeof(eacham_match_t), g.matches.size(), in.f) == g.matches.size();
    if (!ok) throw std::runtime_error("MatchGraphIO: " + path + " is truncated");
    for (const auto& r : g.results)
        if (r.offset + r.count > g.matches.size()) throw std::runtime_error("MatchGraphIO: " + path + " holds a match range outside its match buffer");
    return g;
}

// The batched path's output (std::vector<eacham::PairMatches> from FeatureMatcherGpu::MatchPairs) as a MatchGraph; matches of a
// pair are written sorted by query index, as the C ABI returns them.
template <typename PairMatchesVec>
MatchGraph ToMatchGraph(const PairMatchesVec& res, const uint64_t n_images)
{
    MatchGraph g;
    g.n_images = n_images;
    for (const auto& m : res)
    {
        eacham_pair_t p; p.first = m.first; p.second = m.second;
        eacham_pair_result_t r{};
        r.n12 = m.n12; r.n21 = m.n21;
        r.n_mutual = m.gated ? 0u : static_cast<uint32_t>(m.bestMatches12.size());
        r.flags = (m.gated ? EACHAM_PAIR_GATED : 0u) | (m.connected ? EACHAM_PAIR_CONNECTED : 0u);
        r.offset = g.matches.size();
        r.count = m.bestMatches12.size();
        std::vector<std::pair<unsigned, unsigned>> sorted(m.bestMatches12.begin(), m.bestMatches12.end());
        std::sort(sorted.begin(), sorted.end());
        for (const auto& ab : sorted) { eacham_match_t e; e.query = ab.first; e.train = ab.second; g.matches.push_back(e); }
        g.pairs.push_back(p);
        g.results.push_back(r);
    }
    return g;
}

}
