"""CPU restatement (NumPy) of the hypothesis scoring the reference obtains from OpenCV after matching -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product (eacham_b200/) never does.

Restates, for the geometric-verification step of the reference (ReconstructionManager::RecoverPoseTwoView,
/root/reference/modules/sfm/reconstruction/ReconstructionManager.cpp:47-86: cv::findEssentialMat(..., cv::LMEDS, 0.99, 4.0, 1000, mask)
and cv::findHomography(pts1, pts2, cv::LMEDS, 4.0, mask2, 100, 0.999)), the parts of OpenCV (un-vendored dependency, pinned
opencv/4.5.5 in /root/reference/conanfile.txt:3) that score a hypothesis:
    essential_errors     modules/calib3d/src/five-point.cpp   EMEstimatorCallback::computeError   (double, normalised coordinates, stored float)
    homography_errors    modules/calib3d/src/fundam.cpp       HomographyEstimatorCallback::computeError   (float)
    lmeds_select         modules/calib3d/src/ptsetreg.cpp     LMeDSPointSetRegistrator::run: median of sorted errors, least median wins,
                                                              sigma = 2.5 * 1.4826 * (1 + 5 / (n - modelPoints)) * sqrt(median) >= 0.001,
                                                              findInliers: err <= (float)(sigma * sigma)
Pinned (tests/test_verify_oracle.py) against cv2 itself: the mask returned by cv2.findEssentialMat(..., LMEDS) must equal
lmeds_select's mask for the E it returns; homography_errors against cv2.perspectiveTransform.
Every operation is elementwise IEEE arithmetic in a fixed order, the same order as eacham_b200/csrc/verify_kernels.cuh.
"""
from __future__ import annotations

import numpy as np

F32_MAX = np.float32(np.finfo(np.float32).max)


def essential_errors(E: np.ndarray, p1: np.ndarray, p2: np.ndarray, focal: float, cx: float, cy: float) -> np.ndarray:
    """p1, p2: [n, 2] float32 pixel coordinates. Returns float32 [n]."""
    E = np.asarray(E, np.float64).reshape(9)
    f, cx, cy = np.float64(focal), np.float64(cx), np.float64(cy)
    x1 = (p1[:, 0].astype(np.float64) - cx) / f; y1 = (p1[:, 1].astype(np.float64) - cy) / f
    x2 = (p2[:, 0].astype(np.float64) - cx) / f; y2 = (p2[:, 1].astype(np.float64) - cy) / f
    a0 = (E[0] * x1 + E[1] * y1) + E[2]
    a1 = (E[3] * x1 + E[4] * y1) + E[5]
    a2 = (E[6] * x1 + E[7] * y1) + E[8]
    b0 = (E[0] * x2 + E[3] * y2) + E[6]
    b1 = (E[1] * x2 + E[4] * y2) + E[7]
    x2tEx1 = (x2 * a0 + y2 * a1) + a2
    den = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1
    with np.errstate(divide="ignore", invalid="ignore"):
        e = ((x2tEx1 * x2tEx1) / den).astype(np.float32)
    return np.where(np.isnan(e), F32_MAX, e)


def homography_errors(H: np.ndarray, p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """Forward transfer error in float32, as OpenCV computes it from (float)H."""
    Hf = np.asarray(H, np.float64).reshape(9).astype(np.float32)
    u1, v1, u2, v2 = (p1[:, 0].astype(np.float32), p1[:, 1].astype(np.float32), p2[:, 0].astype(np.float32), p2[:, 1].astype(np.float32))
    one = np.float32(1.0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ww = one / ((Hf[6] * u1 + Hf[7] * v1) + Hf[8])
        dx = ((Hf[0] * u1 + Hf[1] * v1) + Hf[2]) * ww - u2
        dy = ((Hf[3] * u1 + Hf[4] * v1) + Hf[5]) * ww - v2
        e = (dx * dx + dy * dy).astype(np.float32)
    return np.where(np.isnan(e), F32_MAX, e)


def median_of(err: np.ndarray) -> np.float32:
    s = np.sort(err.astype(np.float32))
    n = s.shape[0]
    return s[n // 2] if n % 2 else np.float32((s[n // 2 - 1] + s[n // 2]) * np.float32(0.5))


def lmeds_select(errs: np.ndarray, model_points: int):
    """errs: [n_hyp, n] float32. Returns (best index, median, sigma, mask uint8 [n]) as OpenCV's LMedS registrator would."""
    n = errs.shape[1]
    meds = np.array([median_of(e) for e in errs], np.float32)
    best = int(np.argmin(meds))                       # first of the least medians
    med = meds[best]
    sigma = 2.5 * 1.4826 * (1.0 + 5.0 / float(n - model_points)) * float(np.sqrt(np.float64(med)))
    sigma = max(sigma, 0.001)
    thr = np.float32(sigma * sigma)
    mask = (errs[best] <= thr).astype(np.uint8)
    return best, med, np.float32(sigma), mask, meds


def verify_pair(model: str, hyps: np.ndarray, p1: np.ndarray, p2: np.ndarray, focal: float = 1.0, cx: float = 0.0, cy: float = 0.0):
    """What eacham_gpu_verify_pairs returns for one pair: dict(best, n_inliers, median, sigma, mask, medians)."""
    n = p1.shape[0]
    mp = 5 if model == "essential" else 4
    if n <= mp:
        return dict(best=0, n_inliers=0, median=np.float32(0), sigma=np.float32(0), mask=np.zeros(n, np.uint8), medians=np.zeros(len(hyps), np.float32))
    if model == "essential":
        errs = np.stack([essential_errors(h, p1, p2, focal, cx, cy) for h in hyps])
    else:
        errs = np.stack([homography_errors(h, p1, p2) for h in hyps])
    best, med, sigma, mask, meds = lmeds_select(errs, mp)
    return dict(best=best, n_inliers=int(mask.sum()), median=med, sigma=sigma, mask=mask, medians=meds)
