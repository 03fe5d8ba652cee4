"""CPU: the verification oracle (oracle/verify_oracle.py) pinned against OpenCV itself -- the reference's own dependency for
ReconstructionManager.cpp:58-74. No GPU."""
import numpy as np
import pytest

from oracle import verify_oracle as V

cv2 = pytest.importorskip("cv2")


def two_views(rng, n, outliers=0.25, noise=0.4, planar=False):
    """Random scene seen by two cameras (f = 600, pp = (400, 300)); returns float32 pixel points."""
    f, cx, cy = 600.0, 400.0, 300.0
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(4, 9, n) if not planar else np.full(n, 6.0)], 1)
    ang = 0.12
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.6, 0.05, 0.1])
    def proj(P):
        return np.stack([f * P[:, 0] / P[:, 2] + cx, f * P[:, 1] / P[:, 2] + cy], 1)
    p1 = proj(X) + rng.normal(0, noise, (n, 2))
    p2 = proj(X @ R.T + t) + rng.normal(0, noise, (n, 2))
    bad = rng.random(n) < outliers
    p2[bad] = np.stack([rng.uniform(0, 800, bad.sum()), rng.uniform(0, 600, bad.sum())], 1)
    return p1.astype(np.float32), p2.astype(np.float32), (f, cx, cy)


@pytest.mark.parametrize("seed,n", [(1, 200), (2, 131), (3, 64), (4, 500)])
def test_essential_scoring_reproduces_opencv_lmeds_mask(seed, n):
    """cv2.findEssentialMat(LMEDS) returns its best E and the LMedS inlier mask of that E: scoring E with the oracle must give the
    same mask (this pins the error function, the median rule, sigma and the inlier test)."""
    rng = np.random.default_rng(seed)
    p1, p2, (f, cx, cy) = two_views(rng, n)
    E, mask = cv2.findEssentialMat(p1, p2, f, (cx, cy), cv2.LMEDS, 0.99, 4.0)
    assert E is not None and E.shape == (3, 3)
    got = V.verify_pair("essential", E[None], p1, p2, f, cx, cy)
    assert np.array_equal(got["mask"], mask.reshape(-1).astype(np.uint8))
    assert got["n_inliers"] == int(mask.sum())


def test_homography_error_is_the_forward_transfer_error():
    rng = np.random.default_rng(7)
    p1, p2, _ = two_views(rng, 300, planar=True)
    H, _ = cv2.findHomography(p1, p2, cv2.LMEDS, 4.0)
    want = ((cv2.perspectiveTransform(p1.reshape(-1, 1, 2).astype(np.float64), H).reshape(-1, 2) - p2) ** 2).sum(1)
    got = V.homography_errors(H, p1, p2)
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-4)         # float32 arithmetic against a float64 evaluation
    # LMedS on a clean planar scene: most points are inliers of the returned (refined) H
    r = V.verify_pair("homography", H[None], p1, p2)
    assert r["n_inliers"] > 150


def test_lmeds_rules():
    errs = np.array([[4, 1, 3, 2], [0.5, 9, 9, 0.5], [1, 1, 1, 1]], np.float32)
    best, med, sigma, mask, meds = V.lmeds_select(np.tile(errs, (1, 3)), 4)     # n = 12 (even): mean of the two middle errors
    assert meds.tolist() == [2.5, 4.75, 1.0] and best == 2
    assert abs(float(sigma) - 2.5 * 1.4826 * (1 + 5 / 8) * 1.0) < 1e-6
    assert mask.sum() == 12
    e = V.essential_errors(np.zeros((3, 3)), np.ones((6, 2), np.float32), np.ones((6, 2), np.float32), 1.0, 0.0, 0.0)
    assert np.all(e == V.F32_MAX)                                             # 0/0 -> NaN -> as bad as it gets
