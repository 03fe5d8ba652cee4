"""Synthetic descriptor sets shaped like the BASELINE.json configs.

The reference has no dataset in-tree (its images/ are README screenshots), so both the parity tests and
bench.py draw descriptors from these seeded generators.  Shapes follow what the reference's extractors
emit: ORB = N x 32 uint8 rows (8 x int32 per /root/reference/modules/base/tools/Tools3d.h:46-63);
SIFT = N x 128 float32, integer-valued 0..255 with row norm ~512
(/root/reference/modules/base/features/FeatureExtractorSift.cpp:8-26).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def orb_image_set(n_images: int, n_desc: int, seed: int, pool: int = 20000, share: float = 0.4,
                  flip: float = 0.06, dup_frac: float = 0.01, zero_rows: int = 2,
                  window: int = 0) -> List[np.ndarray]:
    """ORB-like images that really overlap: a global pool of ``pool`` random 256-bit landmarks; each image
    takes ``share * n_desc`` distinct landmarks, flips every bit with probability ``flip``, fills the rest
    with fresh random rows, plants ``dup_frac`` exact duplicate rows and ``zero_rows`` all-zero rows (forces
    best/second-best ties and the 0/0 ratio case), then permutes its rows.

    ``window`` > 0 draws each image's landmarks from a sliding window of the pool (KITTI-shaped sequences:
    neighbouring frames overlap, distant frames do not).
    """
    rng = np.random.default_rng(seed)
    landmarks = rng.integers(0, 256, (pool, 32), dtype=np.uint8)
    k = int(share * n_desc)
    out = []
    for im in range(n_images):
        if window > 0:
            lo = int((pool - window) * im / max(n_images - 1, 1))
            ids = lo + rng.choice(window, size=min(k, window), replace=False)
        else:
            ids = rng.choice(pool, size=min(k, pool), replace=False)
        seen = landmarks[ids]
        # bit flips: bernoulli(flip) mask packed into bytes
        noise = np.packbits(rng.random((seen.shape[0], 256)) < flip, axis=1)
        seen = seen ^ noise
        rest = rng.integers(0, 256, (n_desc - seen.shape[0], 32), dtype=np.uint8)
        d = np.concatenate([seen, rest], axis=0)
        n_dup = int(dup_frac * n_desc)
        if n_dup > 0:
            src = rng.integers(0, n_desc, n_dup)
            dst = rng.integers(0, n_desc, n_dup)
            d[dst] = d[src]
        if zero_rows > 0:
            d[rng.integers(0, n_desc, zero_rows)] = 0
        out.append(np.ascontiguousarray(d[rng.permutation(n_desc)]))
    return out


def _sift_like(rng: np.random.Generator, n: int, dim: int) -> np.ndarray:
    mag = rng.gamma(0.6, 1.0, (n, dim))
    mag /= np.linalg.norm(mag, axis=1, keepdims=True) + 1e-12
    mag = np.minimum(mag, 0.2)
    mag /= np.linalg.norm(mag, axis=1, keepdims=True) + 1e-12
    return mag


def sift_image_set(n_images: int, n_desc: int, seed: int, pool: int = 40000, share: float = 0.3,
                   noise: float = 6.0, integer_valued: bool = True, dim: int = 128) -> List[np.ndarray]:
    """SIFT-like images.  ``integer_valued`` = OpenCV-shaped (x512, rounded, clipped to [0,255]); otherwise
    unit-norm float rows (RootSIFT / learned-descriptor style) that exercise the FP32 re-rank."""
    rng = np.random.default_rng(seed)
    land = _sift_like(rng, pool, dim) * 512.0
    k = int(share * n_desc)
    out = []
    for _ in range(n_images):
        ids = rng.choice(pool, size=min(k, pool), replace=False)
        seen = land[ids] + rng.normal(0.0, noise, (len(ids), dim))
        rest = _sift_like(rng, n_desc - len(ids), dim) * 512.0
        d = np.concatenate([seen, rest], axis=0)
        if integer_valued:
            d = np.clip(np.rint(d), 0, 255)
        else:
            d = np.maximum(d, 0.0)
            d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
        out.append(np.ascontiguousarray(d[rng.permutation(n_desc)].astype(np.float32)))
    return out


def sift_image_set_pooled(n_images: int, n_desc: int, seed: int, pool: int = 400000, noise: float = 6.0,
                          integer_valued: bool = True, dim: int = 128) -> List[np.ndarray]:
    """Same statistics as ``sift_image_set`` at a fraction of the generation cost (bench.py, config 3: 500 x 8192 rows): every
    row of every image is a noisy view of one of ``pool`` SIFT-like landmarks, so two images share about n_desc^2 / pool of
    them (168 at 8192 / 400k; ``sift_image_set`` plants 151) and the rest behave as unrelated rows. Only the landmark pool
    needs the slow gamma draws; per image it is a gather plus float32 normal noise."""
    root = np.random.SeedSequence(seed)
    pool_seed, *image_seeds = root.spawn(n_images + 1)
    rng = np.random.default_rng(pool_seed)
    mag = rng.standard_gamma(0.6, (pool, dim), dtype=np.float32)
    mag /= np.linalg.norm(mag, axis=1, keepdims=True) + 1e-12
    np.minimum(mag, 0.2, out=mag)
    mag /= np.linalg.norm(mag, axis=1, keepdims=True) + 1e-12
    land = mag * np.float32(512.0)

    def one(ss):                                   # one generator per image: the result does not depend on the thread count
        r = np.random.default_rng(ss)
        ids = r.choice(pool, size=n_desc, replace=False)
        d = land[ids] + np.float32(noise) * r.standard_normal((n_desc, dim), dtype=np.float32)
        if integer_valued:
            d = np.clip(np.rint(d), 0, 255)
        else:
            d = np.maximum(d, 0.0)
            d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
        return np.ascontiguousarray(d.astype(np.float32))

    import os
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        return list(ex.map(one, image_seeds))


def exhaustive_pairs(n_images: int) -> np.ndarray:
    """The unordered pairs {(i, j): i < j} -- one per two ordered pairs of
    /root/reference/apps/sfm/main.cpp:84-92."""
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.uint32)


def window_pairs(n_images: int, width: int) -> np.ndarray:
    """Sliding-window pair list (BASELINE config 4): each frame against the next ``width`` frames."""
    out = []
    for f in range(n_images):
        hi = min(n_images, f + width + 1)
        if hi > f + 1:
            js = np.arange(f + 1, hi)
            out.append(np.stack([np.full_like(js, f), js], axis=1))
    return np.concatenate(out, axis=0).astype(np.uint32) if out else np.zeros((0, 2), np.uint32)
