// MatchPhaseGpu.h -- the reference's whole match phase (/root/reference/apps/sfm/main.cpp:81-152) as ONE call.
//
//   eacham::MatchPhase(*graph, frames, matcher);
//
// replaces the pair list, the par_unseq loop with its per-pair std::async, the mutex-guarded checkedBuffer and the two
// Graph::Connect calls. It is a template on the reference's own types, used only through the members the reference itself uses:
//   Graph:  GetNodes() -> map<unsigned, Node*>   /root/reference/modules/sfm/data/Graph.h:54-57
//           Connect(Node*, Node*, match_t&&)      /root/reference/modules/sfm/data/Graph.h:35-41
//   Node:   GetDescriptors(), GetFactor(id)       /root/reference/modules/sfm/data/Node.h:136-139, 171-180
//   frame:  .id                                   /root/reference/apps/sfm/main.cpp:101-102
// Differences from the loop it replaces, all of them fixes:
//   * frames that never got a node (fewer than min_features_count features, main.cpp:75) are skipped; the reference
//     dereferences graph->Get(id) == nullptr for them (main.cpp:104-108);
//   * every unordered pair is matched once, both directions from one distance matrix (the reference evaluates it twice);
//   * no shared checkedBuffer, so neither the unlocked read (main.cpp:128) nor the colliding pair hash (main.cpp:116) exist;
//   * Factor.quality gets the match count: Graph::Connect reads matches.size() AFTER moving from it (Graph.h:39-40), which
//     leaves 0 in the reference.
// The match sets handed to Graph::Connect are exactly the reference's bestMatches12 / bestMatches21 (main.cpp:133-146).
#pragma once

#include <cstddef>
#include <utility>
#include <vector>

#include "FeatureMatcherGpu.h"

namespace eacham
{

struct MatchPhaseStats
{
    std::size_t frames = 0;        // frames given
    std::size_t nodes = 0;         // frames that have a node (the others are skipped)
    std::size_t pairs = 0;         // unordered pairs matched
    std::size_t connected = 0;     // pairs for which Graph::Connect was called (both ways)
};

template <typename GraphT, typename Frames>
MatchPhaseStats MatchPhase(GraphT& graph, const Frames& frames, FeatureMatcherGpu& matcher)
{
    MatchPhaseStats stats;
    const auto& nodes = graph.GetNodes();
    using NodePtr = typename std::decay<decltype(nodes.begin()->second)>::type;
    using Mat = typename std::decay<decltype(nodes.begin()->second->GetDescriptors())>::type;
    std::vector<NodePtr> present;
    std::vector<Mat> descriptors;
    for (const auto& frame : frames)
    {
        ++stats.frames;
        const auto it = nodes.find(frame.id);
        if (it == nodes.end() || it->second == nullptr) continue;      // main.cpp:75: no node for this frame
        present.push_back(it->second);
        descriptors.push_back(it->second->GetDescriptors());           // cv::Mat header copy: shares the pixel buffer
    }
    stats.nodes = present.size();
    const auto pairs = FeatureMatcherGpu::ExhaustivePairs(static_cast<unsigned>(present.size()));
    stats.pairs = pairs.size();
    for (auto& m : matcher.MatchPairsAny(descriptors, pairs))
    {
        if (!m.connected) continue;                                    // main.cpp:142
        NodePtr node1 = present[m.first];
        NodePtr node2 = present[m.second];
        const float quality = static_cast<float>(m.bestMatches12.size());
        graph.Connect(node1, node2, std::move(m.bestMatches12));       // main.cpp:144
        graph.Connect(node2, node1, std::move(m.bestMatches21));       // main.cpp:145
        node1->GetFactor(node2->id).quality = quality;                 // Graph.h:39-40 stores 0 here
        node2->GetFactor(node1->id).quality = quality;
        ++stats.connected;
    }
    return stats;
}

}
