"""GPU parity of eacham_gpu_verify_pairs against the NumPy restatement of OpenCV's hypothesis scoring
(oracle/verify_oracle.py, itself pinned against cv2 in tests/test_verify_oracle.py): residual medians, best hypothesis, sigma,
inlier counts and inlier masks -- bit for bit -- for essential-matrix and homography hypotheses, per-pair and shared, read straight
from the device-resident match lists of the preceding MatchPairs batch."""
import numpy as np
import pytest

from oracle import verify_oracle as V
from test_verify_oracle import two_views

pytestmark = pytest.mark.gpu


def _scene_images(rng, n, planar):
    """Two images whose ORB descriptors match one-to-one (a noisy, permuted copy) and whose keypoints are two views of a scene."""
    p1, p2, cam = two_views(rng, n, planar=planar)
    a = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    flips = np.packbits(rng.random((n, 256)) < 0.03, axis=1)
    perm = rng.permutation(n)
    b = np.ascontiguousarray((a ^ flips)[perm])
    return a, b, p1, np.ascontiguousarray(p2[perm]), cam


def _perturbed(M, rng, k, scale):
    out = [M]
    for _ in range(k - 1):
        out.append(M + rng.normal(0, scale * np.abs(M).max(), (3, 3)))
    return np.stack(out)


@pytest.mark.parametrize("model", ["essential", "homography"])
def test_verify_pairs_equals_oracle(model):
    import cv2
    import eacham_b200
    from eacham_b200 import _lib as L
    rng = np.random.default_rng(11 if model == "essential" else 12)
    imgs, kps, truth = [], [], []
    for s, n in enumerate((400, 233, 1000)):
        a, b, p1, p2, cam = _scene_images(rng, n, planar=(model == "homography"))
        imgs += [a, b]; kps += [p1, p2]
    f, cx, cy = cam
    pairs = [(0, 1), (2, 3), (4, 5), (0, 3), (5, 4)]                     # (0, 3): unrelated images -> not connected -> zeros
    n_hyp = 12
    with eacham_b200.FeatureMatcherGpu(0.8) as m:
        m.Upload(imgs)
        for i, k in enumerate(kps):
            m.SetKeypoints(i, k)
        res, buf = m.MatchPairsRaw(pairs)
        assert res["count"][3] == 0 and res["count"][0] > 300
        hyps = np.zeros((len(pairs), n_hyp, 3, 3))
        pts = []
        for k, (i, j) in enumerate(pairs):
            mm = buf[int(res[k]["offset"]): int(res[k]["offset"] + res[k]["count"])]
            q1, q2 = kps[i][mm["query"]], kps[j][mm["train"]]
            pts.append((q1, q2))
            if len(mm) > 8:
                if model == "essential":
                    M, _ = cv2.findEssentialMat(q1, q2, f, (cx, cy), cv2.LMEDS, 0.99, 4.0)
                else:
                    M, _ = cv2.findHomography(q1, q2, cv2.LMEDS, 4.0)
                hyps[k] = _perturbed(M, rng, n_hyp, 0.02)[rng.permutation(n_hyp)]
            else:
                hyps[k] = rng.normal(0, 1, (n_hyp, 3, 3))
        code = L.MODEL_ESSENTIAL if model == "essential" else L.MODEL_HOMOGRAPHY
        vres, med, mask = m.VerifyPairs(code, hyps, f, cx, cy, want_medians=True, want_mask=True)
        assert m.timing()["verify_ms"] > 0
        for k in range(len(pairs)):
            q1, q2 = pts[k]
            want = V.verify_pair(model, hyps[k], q1, q2, f, cx, cy)
            got_mask = mask[int(res[k]["offset"]): int(res[k]["offset"] + res[k]["count"])]
            assert int(vres[k]["best"]) == want["best"], k
            assert int(vres[k]["n_inliers"]) == want["n_inliers"], k
            assert vres[k]["median"] == want["median"] and vres[k]["sigma"] == want["sigma"], k
            assert np.array_equal(med[k], want["medians"]), k
            assert np.array_equal(got_mask, want["mask"]), k
        assert vres[0]["n_inliers"] > 150                               # the true model explains most matches of its scene
        # one shared hypothesis set for all pairs
        sres, smed, _ = m.VerifyPairs(code, hyps[0], f, cx, cy, want_medians=True)
        for k in range(len(pairs)):
            want = V.verify_pair(model, hyps[0], pts[k][0], pts[k][1], f, cx, cy)
            assert (int(sres[k]["best"]), int(sres[k]["n_inliers"])) == (want["best"], want["n_inliers"]) and np.array_equal(smed[k], want["medians"])


def test_verify_pairs_errors():
    import eacham_b200
    from eacham_b200 import _lib as L, synth
    imgs = synth.orb_image_set(2, 300, seed=3, pool=400)
    with eacham_b200.FeatureMatcherGpu(0.8) as m:
        m.Upload(imgs)
        with pytest.raises(L.EachamGpuError):                           # no batch yet
            m._last_n = 1
            m.VerifyPairs(L.MODEL_HOMOGRAPHY, np.eye(3)[None])
        m.MatchPairsRaw([(0, 1)])
        with pytest.raises(L.EachamGpuError) as e:                      # keypoints missing
            m.VerifyPairs(L.MODEL_HOMOGRAPHY, np.eye(3)[None])
        assert e.value.code == L.ERR_INVALID_ARG
        m.SetKeypoints(0, np.zeros((300, 2), np.float32)); m.SetKeypoints(1, np.zeros((299, 2), np.float32))
        with pytest.raises(L.EachamGpuError):                           # wrong keypoint count
            m.VerifyPairs(L.MODEL_HOMOGRAPHY, np.eye(3)[None])
        m.SetKeypoints(1, np.zeros((300, 2), np.float32))
        r, _, _ = m.VerifyPairs(L.MODEL_HOMOGRAPHY, np.eye(3)[None])
        assert r.shape == (1,)
