"""GPU fuzz: many small ragged pairs of LOW-ENTROPY descriptors (massive distance ties, duplicates, zero distances) in one
batched call per engine, every pair compared with the C oracle. Ties are where index handling breaks first."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _orb_low_entropy(rng, n):
    """rows drawn from a small alphabet with a few flipped bits: distances take few distinct values"""
    if n == 0:
        return np.zeros((0, 32), np.uint8)
    alphabet = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    rows = alphabet[rng.integers(0, 6, n)].copy()
    flips = rng.integers(0, 4, n)
    for i in range(n):
        for _ in range(flips[i]):
            b = rng.integers(0, 256)
            rows[i, b // 8] ^= np.uint8(1 << (b % 8))
    return rows


def _sift_low_entropy(rng, n):
    if n == 0:
        return np.zeros((0, 128), np.float32)
    alphabet = rng.integers(0, 4, (5, 128)).astype(np.float32) * 40
    rows = alphabet[rng.integers(0, 5, n)].copy()
    rows += (rng.random((n, 128)) < 0.02) * rng.integers(1, 3, (n, 128))
    return rows.astype(np.float32)


def _check(m, descs, pairs, min_dir, min_mutual):
    m.Upload(descs)
    out = m.MatchPairs(pairs, emit_all=True)
    for (i, j), pm in zip(pairs, out):
        want = O.c_match_pair(descs[i], descs[j], 0.8, min_dir, min_mutual)
        got = (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected)
        assert got == (want["n12"], want["n21"], want["n_mutual"], want["gated"], want["connected"]), (i, j, descs[i].shape, descs[j].shape)
        assert np.array_equal(pm.matches.reshape(-1, 2), want["matches"].reshape(-1, 2)), (i, j)


@pytest.mark.parametrize("engine", ["tensor", "popc", "tensor_alu", "tensor_v1"])
def test_fuzz_orb_ties(engine):
    import eacham_b200
    rng = np.random.default_rng(2026)
    sizes = [0, 1, 2, 3, 31, 32, 33, 127, 128, 129, 255, 256, 257, 300, 511, 513, 640]
    descs = [_orb_low_entropy(rng, int(n)) for n in sizes]
    # add a few with planted unique partners so that some matches survive the ratio test
    for k in (9, 12, 15):
        n = min(descs[k].shape[0], descs[k + 1].shape[0], 60)
        uniq = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        descs[k][:n] = uniq; descs[k + 1][:n] = uniq
        descs[k + 1][:n, 0] ^= 1
    pairs = [(i, j) for i in range(len(descs)) for j in range(len(descs)) if i != j and (i + j) % 3 != 0]
    with eacham_b200.FeatureMatcherGpu(0.8, min_dir=3, min_mutual=2, orb_engine=engine) as m:
        _check(m, descs, pairs, 3, 2)


def test_fuzz_sift_ties():
    import eacham_b200
    rng = np.random.default_rng(77)
    sizes = [0, 1, 2, 5, 127, 128, 129, 200, 256, 257, 300, 384]
    descs = [_sift_low_entropy(rng, int(n)) for n in sizes]
    for k in (6, 8, 10):
        n = min(descs[k].shape[0], descs[k + 1].shape[0], 50)
        uniq = rng.integers(0, 256, (n, 128)).astype(np.float32)
        descs[k][:n] = uniq; descs[k + 1][:n] = uniq
        descs[k + 1][:n, 3] += 2
    pairs = [(i, j) for i in range(len(descs)) for j in range(len(descs)) if i != j and (i * 7 + j) % 4 != 0]
    with eacham_b200.FeatureMatcherGpu(0.8, min_dir=3, min_mutual=2) as m:
        _check(m, descs, pairs, 3, 2)
