"""Adversarial cases for the tensor-core SIFT path (MatchPairs on F32X128): near-ties BELOW the resolution of its bf16 scorer
placed so that the scorer's second-best candidate is the wrong row, exactly at the ratio threshold. The batched path must still
return the exact matcher's match sets (0 differences): its certainty check (tc_match_kernels.cuh, rerank_ratio_checked) sends such
queries to an exact FP32 scan. Also: the distances the batched path's ratio test sees (eacham_gpu_debug_pair_knn2) against OpenCV.

All descriptors here are multiples of u = 2^-10 with at most 16 non-zero dimensions, so every squared distance is exact in fp32
whatever the summation order: OpenCV, the C oracle and the GPU agree bit for bit, and a difference is a real difference. 10-bit
values are NOT bf16-exact, so the scorer really is approximate on them."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
U = np.float32(2.0 ** -10)


def _ref_pair(a, b, **kw):
    return (O.cv2_match_pair_fast if O.have_cv2() else O.c_match_pair)(a, b, **kw)


def _scene(n_scen, n_train, placements, rng):
    """Scenario s lives around its own far-away offset (dimension 64 + s). Its query is the offset itself. Train rows:
         best      offset + 412u e0                                 d0 = 412u
         true 2nd  offset + 515u e1                                 d1 = 515u: ratio = 0.8 exactly -> (double)0.8f >= 0.8 -> NO match
         decoys    offset + 515u e1 + 1u e2 / + 1u e3               d  = sqrt(515^2 + 1)u: ratio = 0.7999985 -> a match if taken as 2nd
       515u needs 10 significant bits, so bf16 rounds the true second and the decoys to the same operand: the scorer sees a tie
       (the +-1u^2/2 difference is 4e-6 of D, below the 2^-15 key resolution) and keeps the rows with the LOWER indices.
       placements[s] = (best, decoy1, decoy2, true2nd) row indices in the train image."""
    q = np.zeros((n_scen, 128), np.float32)
    t = np.zeros((n_train, 128), np.float32)
    used = set()
    assert len({i for pl in placements for i in pl}) == 4 * n_scen
    for s in range(n_scen):
        off = np.zeros(128, np.float32); off[64 + s] = 900 * U
        q[s] = off
        ib, i1, i2, it = placements[s]
        t[ib] = off; t[ib, 0] = 412 * U
        t[i1] = off; t[i1, 1] = 515 * U; t[i1, 2] = 1 * U
        t[i2] = off; t[i2, 1] = 515 * U; t[i2, 3] = 1 * U
        t[it] = off; t[it, 1] = 515 * U
        used |= {ib, i1, i2, it}
    # filler rows: far from every scenario (their own offset dimension), sparse 10-bit values
    for j in range(n_train):
        if j in used:
            continue
        t[j, 40 + (j % 20)] = 700 * U
        dims = rng.choice(16, size=4, replace=False)
        t[j, dims] = rng.integers(1, 1024, 4).astype(np.float32) * U
    return q, t


def test_wrong_second_candidate_at_the_ratio_threshold():
    import eacham_b200
    rng = np.random.default_rng(5)
    placements = [
        (0, 3, 4, 730),            # decoys early, true second in a LATER 128-column tile
        (10, 133, 261, 645),       # decoys and true second at the same in-tile position (5) of different tiles
        (20, 21, 22, 23),          # all in one 32-column part of one tile, true second last
        (700, 300, 301, 302),      # best in a later tile than the tied group
        (50, 400, 401, 399),       # true second has the LOWEST index of the tied group: the scorer is right by luck
    ]
    n_scen, n_train = len(placements), 768
    q, t = _scene(n_scen, n_train, placements, rng)
    # the query image needs other rows too (they become train rows in the other direction): copies of scenario structure, transposed roles
    q_extra = np.zeros((200, 128), np.float32)
    for j in range(200):
        q_extra[j, 100 + (j % 20)] = 650 * U
        q_extra[j, rng.choice(16, 3, replace=False)] = rng.integers(1, 1024, 3).astype(np.float32) * U
    a = np.concatenate([q, q_extra]).astype(np.float32)
    with eacham_b200.FeatureMatcherGpu(0.8, min_dir=0, min_mutual=0) as m:
        for x, y in ((a, t), (t, a)):                       # both roles: the tied group among the columns and among the rows
            m.Upload([x, y])
            pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
            want = _ref_pair(x, y, min_dir=0, min_mutual=0)
            assert (pm.n12, pm.n21, pm.n_mutual) == (want["n12"], want["n21"], want["n_mutual"])
            assert np.array_equal(pm.matches.reshape(-1, 2), np.asarray(want["matches"]).reshape(-1, 2))
            assert m.timing()["exact_fallbacks"] >= n_scen - 1          # the certainty check caught them
        # the scenario queries must NOT match (ratio is exactly 0.8f, and (double)0.8f >= 0.8)
        m.Upload([a, t])
        i12, d12, i21, d21 = m.DebugPairKnn2(0, 1)
        for s, (ib, i1, i2, it) in enumerate(placements):
            assert i12[s, 0] == ib and i12[s, 1] == it, (s, i12[s])          # after the exact scan: the TRUE second neighbour
            assert d12[s, 0] == np.float32(412) * U and d12[s, 1] == np.float32(515) * U
        got = m.Match(a, t)                                 # the all-FP32 single-direction route agrees
        assert all(s not in got for s in range(n_scen))


@pytest.mark.parametrize("integer", [True, False])
def test_batched_path_distances_against_opencv(integer):
    """Distances of the tcgen05 path (not just of the FP32 knnMatch kernels): best within 1e-4 relative of OpenCV's everywhere;
    second within 1e-4 for integer-valued rows; for general floats it is the exact distance of the scorer's second candidate: never
    BELOW OpenCV's second-best and above it by no more than the bf16 scoring bound (the certainty check guarantees that this can
    not change a ratio-test outcome; where it could, the exact scan replaces it)."""
    import eacham_b200
    from eacham_b200 import synth
    imgs = synth.sift_image_set(2, 1500, seed=77, pool=2200, share=0.45, integer_valued=integer)
    a, b = imgs[0], np.ascontiguousarray(imgs[1][:1300])
    with eacham_b200.FeatureMatcherGpu(0.8) as m:
        m.Upload([a, b])
        i12, d12, i21, d21 = m.DebugPairKnn2(0, 1)
        for (q, t, idx, dist) in ((a, b, i12, d12), (b, a, i21, d21)):
            ri, rd = m.knnMatch(q, t)                       # exact FP32 kernels == OpenCV (tests/test_gpu_sift.py)
            if O.have_cv2():
                ci, cd = O.cv2_knn2(q, t)
                np.testing.assert_allclose(rd, cd, rtol=1e-6)
            np.testing.assert_allclose(dist[:, 0], rd[:, 0], rtol=1e-4)
            assert (idx[:, 0] == ri[:, 0]).mean() >= 0.999
            if integer:
                np.testing.assert_allclose(dist[:, 1], rd[:, 1], rtol=1e-4)
            else:
                assert np.all(dist[:, 1] >= rd[:, 1] * (1 - 1e-6))
                assert np.all(dist[:, 1] <= rd[:, 1] * 1.05)
        pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
        want = _ref_pair(a, b)
        if integer:
            assert np.array_equal(pm.matches.reshape(-1, 2), np.asarray(want["matches"]).reshape(-1, 2))
            assert m.timing()["exact_fallbacks"] <= 4


def test_passed_over_scores_decide_correctly():
    """The pruned sweep (tc_sift_kernels.cuh: low_bar / skip) does not insert scores >= ~0.68 x the running best; only their minimum is
    kept and folded into the decision. Scenarios with EXACT integer distances (query = 256 e_s; train rows = 256 e_s + d e_(64+s), so the
    distance is d; every other row is >= 320 away) placed so that the interesting rows arrive AFTER a bar exists (later tiles):
      a passed-over row is the true best (no match), the true second (match and no match), exactly at the ratio (no match: 0.8f >= 0.8
      as a double), and a chain of passed-over / inserted bests. Both engines that share the rule must equal OpenCV."""
    import eacham_b200
    scen = [
        # (distances in arrival order: (row index, d)), expected: index of the matched row or None
        ([(3, 100), (500, 85)], None),                      # 85 passed over, is the true best: 85 / 100 -> no match
        ([(4, 100), (501, 79)], 501),                       # 79 is inserted: 79 / 100 -> match
        ([(5, 50), (502, 62), (630, 63)], None),            # second passed over: 50 / 62 = 0.806 -> no match
        ([(6, 50), (503, 63), (631, 64)], 6),               # second passed over: 50 / 63 = 0.794 -> match
        ([(7, 40), (504, 50)], None),                       # exactly 0.8 -> no match
        ([(8, 100), (505, 90), (632, 81), (700, 73)], None),  # 90 passed, 81 inserted, 73 passed: best 73, second 81 -> no match
        ([(9, 100), (506, 90), (633, 60)], 633),            # 90 passed, 60 inserted: 60 / 90 -> match
        ([(10, 120), (300, 200), (507, 110), (634, 109)], None),   # two passed-over bests in a row: 109 / 110 -> no match
    ]
    n_train = 768
    q = np.zeros((len(scen), 128), np.float32)
    t = np.zeros((n_train, 128), np.float32)
    used = set()
    for s, (rows, _) in enumerate(scen):
        q[s, s] = 256
        for j, d in rows:
            assert j not in used
            used.add(j)
            t[j, s] = 256; t[j, 64 + s] = d
    rng = np.random.default_rng(9)
    for j in range(n_train):
        if j not in used:                                   # fillers: >= 320 from every query, integer-valued
            t[j, 32 + (j % 16)] = 200
            t[j, 100 + rng.integers(0, 20)] = rng.integers(1, 60)
    want = _ref_pair(q, t, min_dir=0, min_mutual=0)
    for engine in ("tensor", "tensor_v1"):
        with eacham_b200.FeatureMatcherGpu(0.8, sift_engine=engine, min_dir=0, min_mutual=0, cross_check=False) as m:
            m.Upload([q, t])
            got = m.Match(q, t)                             # per-call route: same kernel, single-direction mode
            for s, (_, exp) in enumerate(scen):
                assert got.get(s) == exp, (engine, s, got.get(s), exp)
            pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
            assert (pm.n12, pm.n21) == (want["n12"], want["n21"]), engine
        assert sum(e is not None for _, e in scen) == want["n12"]
