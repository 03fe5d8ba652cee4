"""Hand-constructed matching cases that pin every edge listed in SURVEY.md section 8(a).

Shared by tests/golden/make_golden.py (which records OpenCV's answers) and the parity tests.
Rows are 256-bit ORB descriptors (uint8[32]) unless a case says otherwise.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def _rand_rows(rng, n):
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def flip_bits(row: np.ndarray, k: int, rng=None, start: int = 0) -> np.ndarray:
    """Copy of ``row`` with exactly k bits flipped (bits start..start+k-1, or random positions with rng)."""
    bits = np.unpackbits(row.copy())
    pos = np.arange(start, start + k) % 256 if rng is None else rng.choice(256, size=k, replace=False)
    bits[pos] ^= 1
    return np.packbits(bits)


def planted_pair(rng, n1: int, n2: int, n_good: int, d_good: int = 10, n_oneway: int = 0,
                 dup_filler: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Two images with exactly ``n_good`` rows that are each other's clear nearest neighbour (distance d_good,
    everything else ~128 away), and ``n_oneway`` extra rows of image 1 that ratio-pass towards a row of image 2
    which itself prefers someone else (so they count in |m12| only).  Filler rows of image 2 come in exact
    duplicate pairs and filler rows of image 1 are duplicated too, so that no filler row can pass the ratio test
    by accident (its best and second-best distances are equal)."""
    base = _rand_rows(rng, n_good)
    a_rows = [base[i] for i in range(n_good)]
    b_rows = [flip_bits(base[i], d_good, rng) for i in range(n_good)]
    # one-way: a' at distance d_good+6 from b_i (which still prefers a_i at d_good)
    for k in range(n_oneway):
        a_rows.append(flip_bits(b_rows[k % max(n_good, 1)], d_good + 6 + (k // max(n_good, 1)), rng))
    def filler(n):
        if n <= 0:
            return []
        if not dup_filler:
            return list(_rand_rows(rng, n))
        half = _rand_rows(rng, (n + 1) // 2)
        rows = list(half) + list(half)
        return rows[:n] if n % 2 == 0 else rows[:n - 1] + [rows[0]]
    a_rows += filler(n1 - len(a_rows))
    b_rows += filler(n2 - len(b_rows))
    a = np.stack(a_rows).astype(np.uint8); b = np.stack(b_rows).astype(np.uint8)
    return np.ascontiguousarray(a[rng.permutation(len(a))]), np.ascontiguousarray(b[rng.permutation(len(b))])


def orb_cases() -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    rng = np.random.default_rng(20261018)
    cases: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}

    # best-distance ties: train rows 2, 5, 7 identical and equal to query 0 -> (2, d 0), (5, d 0)   [Probe A1]
    q = _rand_rows(rng, 6); t = _rand_rows(rng, 12)
    t[2] = q[0]; t[5] = q[0]; t[7] = q[0]
    cases["ties_best"] = (q, t)

    # second-best ties: unique best at distance 3, rows 1, 4, 9 all at distance 40 -> second = row 1
    q = _rand_rows(rng, 4); t = _rand_rows(rng, 16)
    t[6] = flip_bits(q[1], 3)
    for r, s in ((1, 0), (4, 50), (9, 100)):
        t[r] = flip_bits(q[1], 40, start=s)
    cases["ties_second"] = (q, t)

    # duplicate train rows (d0 == d1 > 0 -> ratio 1 -> reject) and exact double duplicate (0/0 = NaN -> reject)
    q = _rand_rows(rng, 8); t = _rand_rows(rng, 20)
    t[3] = flip_bits(q[0], 7); t[11] = t[3]            # d0 = d1 = 7
    t[4] = q[1]; t[15] = q[1]                          # d0 = d1 = 0 -> NaN
    t[8] = q[2]                                        # d0 = 0, d1 ~ 100+ -> accept
    cases["duplicates_nan_zero"] = (q, t)

    # ratio exactly 4/5 is rejected (strict <): (4,5), (100,125), (8,10); inside: (3,4) = 0.75, (99,125); outside: (101,125)
    for d0, d1 in ((4, 5), (100, 125), (8, 10), (3, 4), (99, 125), (101, 125), (0, 5), (0, 0), (256, 256), (200, 256)):
        q = _rand_rows(rng, 1)
        t = np.stack([flip_bits(q[0], d1, start=128 if d1 <= 128 else 0), flip_bits(q[0], d0)])   # best is row 1
        cases[f"ratio_{d0}_{d1}"] = (q, t)

    # sizes: N != M, nothing a multiple of any tile; one-row and two-row trains; empty images
    cases["ragged_517x1031"] = planted_pair(rng, 517, 1031, 64)
    cases["ragged_1031x517"] = planted_pair(rng, 1031, 517, 40)
    cases["tiny_train_1"] = (_rand_rows(rng, 5), _rand_rows(rng, 1))
    cases["tiny_train_2"] = (_rand_rows(rng, 5), _rand_rows(rng, 2))
    cases["empty_train"] = (_rand_rows(rng, 5), np.zeros((0, 32), np.uint8))
    cases["empty_query"] = (np.zeros((0, 32), np.uint8), _rand_rows(rng, 5))

    # gates: per-direction count exactly 29 / 30, mutual count exactly 30 / 31
    cases["gate_dir_29"] = planted_pair(rng, 200, 220, 29)
    cases["gate_dir_30_mutual_30"] = planted_pair(rng, 200, 220, 30)       # passes :111, fails :142 (30 > 30 false)
    cases["gate_mutual_31"] = planted_pair(rng, 200, 220, 31)              # connected
    # only one direction passes the gate: 40 ratio-passing rows 1->2 but they share 20 partners
    cases["gate_one_direction"] = planted_pair(rng, 300, 260, 20, n_oneway=20)
    # a train row that is the NN of several queries: only the mutual one survives
    cases["shared_partner"] = planted_pair(rng, 400, 380, 35, n_oneway=35)
    # all-zero rows and all-ones rows (max distance 256)
    q = _rand_rows(rng, 40); t = _rand_rows(rng, 50)
    q[0] = 0; q[1] = 255; t[0] = 255; t[1] = 0; t[2] = 0
    cases["zeros_ones"] = (q, t)
    return cases


def sift_cases() -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """float32 [N,128] cases: OpenCV-shaped integer-valued rows, with exact ties / duplicates / zero distance."""
    from eacham_b200 import synth
    rng = np.random.default_rng(77)
    cases: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
    a, b = synth.sift_image_set(2, 300, seed=5, pool=900, share=0.4)
    cases["sift_int_300"] = (a, b)
    a, b = synth.sift_image_set(2, 300, seed=15, pool=260, share=0.4)
    cases["sift_int_connected"] = (a, b)
    a, b = synth.sift_image_set(2, 200, seed=16, pool=170, share=0.4, integer_valued=False)
    cases["sift_float_connected"] = (a, b)
    a, b = synth.sift_image_set(2, 257, seed=6, pool=700, share=0.4, integer_valued=False)
    cases["sift_float_257"] = (a, b[:131])
    q, t = synth.sift_image_set(2, 64, seed=7, pool=200)
    t[3] = q[0]; t[9] = q[0]               # exact duplicates: d0 = d1 = 0 -> NaN -> reject; lowest index first
    t[5] = q[1]                            # d0 = 0 accept
    t[20] = t[21] = q[2] + rng.integers(0, 2, 128).astype(np.float32)   # d0 == d1 > 0 -> reject
    cases["sift_ties"] = (q, t)
    cases["sift_tiny_train"] = (q[:5].copy(), t[:1].copy())
    return cases
