"""CPU: the oracle (C restatement, NumPy brute force, cv2) against the committed golden vectors and
hand-computed known answers. No GPU, no product code."""
import numpy as np
import pytest

from oracle import oracle as O
import cases

NONE = 0xFFFFFFFF


def _map_to_arr(m, n):
    a = np.full(n, NONE, np.uint32)
    for k, v in m.items():
        a[k] = v
    return a


def _check_case(g, knn2, match_pair, exact_dist=True):
    d1, d2 = g["d1"], g["d2"]
    for (q, t, idx, dist, m) in ((d1, d2, g["idx12"], g["dist12"], g["m12"]), (d2, d1, g["idx21"], g["dist21"], g["m21"])):
        i, s = knn2(q, t)
        assert np.array_equal(i, idx)
        if exact_dist:
            assert np.array_equal(s, dist)
        else:
            fin = np.isfinite(dist)
            assert np.array_equal(np.isfinite(s), fin)
            np.testing.assert_allclose(s[fin], dist[fin], rtol=1e-6, atol=1e-6)
        got = _map_to_arr(O.py_ratio_filter(i, s, 0.8), q.shape[0])
        assert np.array_equal(got, m)
    pr = match_pair(d1, d2)
    n12, n21, nm, gated, conn = g["pair"].tolist()
    assert (pr["n12"], pr["n21"], pr["n_mutual"], int(pr["gated"]), int(pr["connected"])) == (n12, n21, nm, gated, conn)
    assert np.array_equal(pr["matches"].reshape(-1, 2), g["matches"].reshape(-1, 2))


def test_c_oracle_matches_golden_orb(orb_golden):
    for name, g in orb_golden.items():
        _check_case(g, O.c_knn2, O.c_match_pair)


def test_numpy_oracle_matches_golden_orb(orb_golden):
    for name, g in orb_golden.items():
        _check_case(g, O.np_knn2, O.np_match_pair)


def test_c_oracle_matches_golden_orb_set(orb_set_golden):
    for name, g in orb_set_golden.items():
        _check_case(g, O.c_knn2, O.c_match_pair)


def test_c_oracle_matches_golden_sift(sift_golden):
    # integer-valued SIFT rows: every partial sum is an exact integer < 2^24 -> bit-equal to OpenCV's SIMD order
    for name, g in sift_golden.items():
        exact = "float" not in name
        if exact:
            _check_case(g, O.c_knn2, O.c_match_pair, exact_dist=True)
        else:
            i, s = O.c_knn2(g["d1"], g["d2"])
            assert (i == g["idx12"]).mean() >= 0.999
            np.testing.assert_allclose(s, g["dist12"], rtol=1e-5)


@pytest.mark.skipif(not O.have_cv2(), reason="cv2 not importable")
def test_cv2_still_matches_golden(orb_golden, sift_golden):
    """The fixtures were made with cv2 4.13.0; whatever cv2 is installed now must still agree (pins the pin)."""
    for g in list(orb_golden.values()) + [v for k, v in sift_golden.items() if "float" not in k]:
        _check_case(g, O.cv2_knn2, O.cv2_match_pair)
        pr = O.cv2_match_pair_fast(g["d1"], g["d2"])
        assert pr["n_mutual"] == g["pair"][2] and np.array_equal(pr["matches"].reshape(-1, 2), g["matches"].reshape(-1, 2))


def test_golden_cases_are_current():
    """tests/cases.py must still generate the inputs stored in the fixtures (else regenerate with make_golden.py)."""
    from conftest import load_golden
    g = load_golden("orb_cases")
    for name, (d1, d2) in cases.orb_cases().items():
        assert np.array_equal(g[name]["d1"], d1) and np.array_equal(g[name]["d2"], d2), name


def test_known_answers_ties():
    # Probe A1 of SURVEY.md: identical train rows 2, 5, 7 equal to query 0 -> (2, 0.0), (5, 0.0)
    q, t = cases.orb_cases()["ties_best"]
    i, s = O.c_knn2(q, t)
    assert i[0].tolist() == [2, 5] and s[0].tolist() == [0.0, 0.0]
    q, t = cases.orb_cases()["ties_second"]
    i, s = O.c_knn2(q, t)
    assert i[1].tolist() == [6, 1] and s[1].tolist() == [3.0, 40.0]


def test_swar_popcount_equals_builtin():
    # /root/reference/modules/base/tools/Tools3d.h:46-63 (SWAR) == __builtin_popcount form used by the oracle
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    rows[0] = 0; rows[1] = 255
    for k in range(0, 200, 2):
        a, b = rows[k], rows[k + 1]
        ref = int(np.unpackbits(a ^ b).sum())
        assert O.c_hamming256(a, b) == ref == O.c_hamming256(a, b, swar=True)


def test_ratio_integer_equivalence_exhaustive():
    """float32(d0)/float32(d1) < 0.8 (double)  <=>  5*d0 < 4*d1 for all Hamming distances (Probe A2); 0/0 rejects."""
    d0, d1 = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (d0.astype(np.float32) / d1.astype(np.float32)).astype(np.float64) < 0.8
    assert np.array_equal(r, 5 * d0 < 4 * d1)
    idx = np.array([[0, 1]], np.int32)
    for a, b, want in ((0, 0, False), (0, 3, True), (4, 5, False), (100, 125, False), (99, 125, True)):
        m = O.c_ratio_filter(idx, np.array([[a, b]], np.float32), 0.8)
        assert (m[0] != NONE) == want


def test_pair_logic_gates():
    m = {i: i for i in range(31)}
    r = O.py_pair_logic(m, dict(m), 30, 30)
    assert r["connected"] and r["n_mutual"] == 31
    m30 = {i: i for i in range(30)}
    r = O.py_pair_logic(m30, dict(m30), 30, 30)
    assert not r["gated"] and r["n_mutual"] == 30 and not r["connected"]      # main.cpp:142 is strict
    m29 = {i: i for i in range(29)}
    assert O.py_pair_logic(m29, dict(m), 30, 30)["gated"] and O.py_pair_logic(m, dict(m29), 30, 30)["gated"]
    # a train row that is NN of several queries: only the mutual one survives
    m12 = {i: i for i in range(40)}; m12[100] = 0; m12[101] = 1
    m21 = {i: i for i in range(40)}
    r = O.py_pair_logic(m12, m21, 30, 30)
    assert r["n12"] == 42 and r["n_mutual"] == 40 and [100, 0] not in r["matches"].tolist()


def test_l2_oracles_agree_on_random_floats():
    from eacham_b200 import synth
    a, b = synth.sift_image_set(2, 150, seed=9, pool=300, integer_valued=False)
    i1, s1 = O.c_knn2(a, b); i2, s2 = O.np_knn2(a, b)
    assert (i1 == i2).mean() > 0.999
    np.testing.assert_allclose(s1, s2, rtol=1e-5)
