"""Parity on REAL extractor output (tests/golden/real_descriptors.npz: cv2.ORB_create(2000) and the reference's
cv2.SIFT_create(2000, 3, 0.009, 10, 1.3) on two views of a rendered scene): the CPU oracle and, on a GPU, every engine
must reproduce what OpenCV's exact matcher + the reference's ratio / pair logic returned."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O


def _cases():
    g = load_golden("real_descriptors")
    g["real_sift"]["d1"] = g["real_sift"]["d1"].astype(np.float32)
    g["real_sift"]["d2"] = g["real_sift"]["d2"].astype(np.float32)
    return g


def test_c_oracle_on_real_descriptors():
    for name, g in _cases().items():
        i, s = O.c_knn2(g["d1"], g["d2"])
        assert np.array_equal(i, g["idx12"]) and np.array_equal(s, g["dist12"]), name
        pr = O.c_match_pair(g["d1"], g["d2"])
        n12, n21, nm, gated, conn = g["pair"].tolist()
        assert (pr["n12"], pr["n21"], pr["n_mutual"], int(pr["gated"]), int(pr["connected"])) == (n12, n21, nm, gated, conn), name
        assert np.array_equal(pr["matches"], g["matches"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tensor", "popc", "tensor_alu", "tensor_v1"])
def test_gpu_on_real_descriptors(engine):
    import eacham_b200
    g = _cases()
    with eacham_b200.FeatureMatcherGpu(0.8, orb_engine=engine) as m:
        for name in ("real_orb", "real_sift"):
            c = g[name]
            m.Upload([c["d1"], c["d2"]])
            pm = m.MatchPairs([(0, 1)])[0]
            n12, n21, nm, gated, conn = c["pair"].tolist()
            assert (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected) == (n12, n21, nm, bool(gated), bool(conn)), (engine, name)
            assert np.array_equal(pm.matches, c["matches"]), (engine, name)
            i, s = m.knnMatch(c["d1"], c["d2"])
            assert np.array_equal(i, c["idx12"]) and np.array_equal(s, c["dist12"]), (engine, name)
            got = np.full(c["d1"].shape[0], 0xFFFFFFFF, np.uint32)
            for k, v in m.Match(c["d1"], c["d2"]).items():
                got[k] = v
            assert np.array_equal(got, c["m12"]), (engine, name)
