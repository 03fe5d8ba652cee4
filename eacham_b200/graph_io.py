"""On-disk match graph.

The reference keeps the result of its O(n^2) matching phase only in memory (``Graph::Connect`` stores a ``match_t`` per
ordered direction, /root/reference/modules/sfm/data/Graph.h:30-41; the only file it ever writes is the final
``transform.json``, /root/reference/modules/sfm/utils/Saver.h:13-73). This module dumps exactly what the batched path
returns -- per-pair counts/flags plus the compacted mutual matches -- so the phase is restartable and parity runs have a
stable artefact.

Layout (little endian):
    magic  "EACHAMG1"                       8 bytes
    n_pairs u64, n_matches u64, n_images u64
    pairs    [n_pairs]   {first u32, second u32}
    results  [n_pairs]   eacham_pair_result_t (n12 u32, n21 u32, n_mutual u32, flags u32, offset u64, count u64)
    matches  [n_matches] eacham_match_t {query u32, train u32}
"""
from __future__ import annotations

import numpy as np

from . import _lib as L

MAGIC = b"EACHAMG1"


def save_match_graph(path: str, pairs: np.ndarray, results: np.ndarray, matches: np.ndarray, n_images: int = 0) -> None:
    pairs = np.ascontiguousarray(np.asarray(pairs, np.uint32).reshape(-1, 2))
    results = np.ascontiguousarray(results.astype(L.RESULT_DTYPE, copy=False))
    matches = np.ascontiguousarray(matches.astype(L.MATCH_DTYPE, copy=False))
    if results.shape[0] != pairs.shape[0]:
        raise ValueError("one result per pair expected")
    if results.shape[0] and int((results["offset"] + results["count"]).max()) > matches.shape[0]:
        raise ValueError("a pair's match range lies outside the match buffer")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([pairs.shape[0], matches.shape[0], n_images or (int(pairs.max()) + 1 if pairs.size else 0)], "<u8").tobytes())
        f.write(pairs.tobytes()); f.write(results.tobytes()); f.write(matches.tobytes())


def load_match_graph(path: str):
    """Returns (pairs[n,2] uint32, results record array, matches record array, n_images)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not an eacham match graph")
        n_pairs, n_matches, n_images = np.frombuffer(f.read(24), "<u8").tolist()
        pairs = np.frombuffer(f.read(8 * n_pairs), "<u4").reshape(-1, 2).copy()
        results = np.frombuffer(f.read(32 * n_pairs), L.RESULT_DTYPE).copy()
        matches = np.frombuffer(f.read(8 * n_matches), L.MATCH_DTYPE).copy()
    if results.shape[0] != n_pairs or matches.shape[0] != n_matches:
        raise ValueError(f"{path}: truncated")
    return pairs, results, matches, int(n_images)


def connected_edges(pairs: np.ndarray, results: np.ndarray, matches: np.ndarray):
    """Iterates what the reference hands to Graph::Connect (apps/sfm/main.cpp:142-146): for every connected pair
    (first, second, best12 {a -> b}, best21 {b -> a})."""
    for (i, j), r in zip(pairs.tolist(), results):
        if r["flags"] & L.PAIR_CONNECTED:
            m = matches[int(r["offset"]): int(r["offset"]) + int(r["count"])]
            yield i, j, dict(zip(m["query"].tolist(), m["train"].tolist())), dict(zip(m["train"].tolist(), m["query"].tolist()))
