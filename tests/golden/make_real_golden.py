"""Golden fixtures from REAL extractor output: two views of a rendered textured scene, descriptors from
cv2.ORB_create(2000) and from cv2.SIFT_create(2000, 3, 0.009, 10, 1.3) -- the reference's own SIFT parameters
(/root/reference/modules/base/features/FeatureExtractorSift.cpp:8) -- matched by OpenCV's exact matcher.

    python tests/golden/make_real_golden.py        (build container, cv2 4.13.0)

SIFT descriptors are integer-valued 0..255 (checked) and stored as uint8 to keep the fixture small.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cv2  # noqa: E402
from make_golden import record  # noqa: E402


def scene(seed: int, size: int = 900) -> np.ndarray:
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (size, size), dtype=np.uint8), (0, 0), 3)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)
    for _ in range(260):                                  # corners and blobs for the detectors
        c = (int(rng.integers(0, size)), int(rng.integers(0, size)))
        col = int(rng.integers(0, 256))
        if rng.random() < 0.5:
            cv2.circle(img, c, int(rng.integers(4, 26)), col, -1)
        else:
            cv2.rectangle(img, c, (c[0] + int(rng.integers(8, 60)), c[1] + int(rng.integers(8, 60))), col, -1)
    return cv2.GaussianBlur(img, (0, 0), 1.0)


def second_view(img: np.ndarray, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    h, w = img.shape
    M = cv2.getRotationMatrix2D((w / 2, h / 2), 7.0, 1.06)
    M[:, 2] += (11.0, -6.0)
    out = cv2.warpAffine(img, M, (w, h), borderMode=cv2.BORDER_REFLECT)
    noise = rng.normal(0, 4.0, out.shape)
    return np.clip(out.astype(np.float32) * 0.93 + 9 + noise, 0, 255).astype(np.uint8)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    a = scene(1); b = second_view(a, 2)
    orb = cv2.ORB_create(2000)
    _, o1 = orb.detectAndCompute(a, None); _, o2 = orb.detectAndCompute(b, None)
    sift = cv2.SIFT_create(2000, 3, 0.009, 10, 1.3)
    _, s1 = sift.detectAndCompute(a, None); _, s2 = sift.detectAndCompute(b, None)
    assert o1.dtype == np.uint8 and o1.shape[1] == 32
    assert np.array_equal(s1, np.rint(s1)) and s1.min() >= 0 and s1.max() <= 255, "SIFT output is not integer-valued 0..255"
    out = {}
    p = record("real_orb", o1, o2, out)
    assert p["connected"], "the two views should connect"
    q = record("real_sift", s1.astype(np.float32), s2.astype(np.float32), out)
    assert q["connected"]
    for k in ("real_sift/d1", "real_sift/d2"):
        out[k] = out[k].astype(np.uint8)                  # lossless here; tests convert back to float32
    np.savez_compressed(os.path.join(here, "real_descriptors.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("/d1") or k.endswith("/d2")})


if __name__ == "__main__":
    main()
