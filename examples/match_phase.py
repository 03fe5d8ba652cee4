#!/usr/bin/env python
"""The reference's match phase (/root/reference/apps/sfm/main.cpp:62-152) with the GPU matcher swapped in:

    read images -> extract descriptors (cv2, as the reference does with cv::SIFT) -> ONE MatchPairs call
    -> connected edges (what main.cpp:142-146 hands to Graph::Connect) -> optional binary match-graph dump.

    python examples/match_phase.py 'images/*.jpg' [--orb] [--dump graph.bin]

Needs a B200 (there is no CPU fallback) and cv2 for the extraction step only.
"""
import argparse
import glob
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--orb", action="store_true", help="ORB (256-bit) instead of the reference's SIFT")
    ap.add_argument("--max-features", type=int, default=15000)       # config feature.max_features_count
    ap.add_argument("--min-features", type=int, default=100)         # config feature.min_features_count (main.cpp:75)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    import cv2
    import eacham_b200
    from eacham_b200 import synth, graph_io
    files = sorted(glob.glob(args.pattern))
    ext = cv2.ORB_create(min(args.max_features, 8000)) if args.orb else cv2.SIFT_create(args.max_features, 3, 0.009, 10, 1.3)
    descs, kept = [], []
    for f in files:
        img = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        if img is None:
            continue
        _, d = ext.detectAndCompute(img, None)
        if d is not None and d.shape[0] >= args.min_features:       # frames with too few features get no node (main.cpp:75)
            descs.append(np.ascontiguousarray(d)); kept.append(f)
    print(f"{len(kept)} of {len(files)} images have >= {args.min_features} features")
    pairs = synth.exhaustive_pairs(len(descs))
    with eacham_b200.FeatureMatcherGpu(0.8) as m:
        t0 = time.perf_counter()
        m.Upload(descs)
        res, buf = m.MatchPairsRaw(pairs)
        dt = time.perf_counter() - t0
    edges = list(graph_io.connected_edges(pairs, res, buf))
    print(f"[Match] time1: {dt * 1e3:.0f}ms   ({len(pairs)} pairs, {len(edges)} connected, {int(res['count'].sum())} mutual matches)")
    if args.dump:
        graph_io.save_match_graph(args.dump, pairs, res, buf, n_images=len(descs))
        print(f"match graph written to {args.dump}")


if __name__ == "__main__":
    main()
