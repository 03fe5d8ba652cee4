"""GPU parity for the SIFT (128-d float, L2) path through the C ABI.
Exact FP32 kernels (Match / knnMatch) and the tensor-core scorer + FP32 re-rank (MatchPairs).
Bar (BASELINE.json north_star): >= 99.9 % pair-set agreement, distances within 1e-4 relative; integer-valued
(OpenCV-shaped) descriptors are exact in bf16 x bf16 -> fp32, so those cases must agree exactly."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
NONE = 0xFFFFFFFF
RTOL = 1e-4


def _ref_pair(a, b, **kw):
    return (O.cv2_match_pair_fast if O.have_cv2() else O.c_match_pair)(a, b, **kw)


def _agreement(pm, want):
    """fraction of the union of both match sets on which they agree"""
    got = set(map(tuple, pm.matches.tolist())); ref = set(map(tuple, np.asarray(want["matches"]).reshape(-1, 2).tolist()))
    union = got | ref
    return 1.0 if not union else len(got & ref) / len(union)


def test_knn2_and_match_golden_exact_fp32(matcher, sift_golden):
    for name, g in sift_golden.items():
        for q, t, idx, dist, m in ((g["d1"], g["d2"], g["idx12"], g["dist12"], g["m12"]),
                                   (g["d2"], g["d1"], g["idx21"], g["dist21"], g["m21"])):
            i, s = matcher.knnMatch(q, t)
            fin = np.isfinite(dist)
            assert np.array_equal(np.isfinite(s), fin), name
            if "float" in name:
                assert (i == idx).mean() >= 0.999, name
                np.testing.assert_allclose(s[fin], dist[fin], rtol=RTOL)
            else:                                   # integer-valued rows: every partial sum is exact
                assert np.array_equal(i, idx), name
                assert np.array_equal(s[fin], dist[fin]), name
                got = np.full(q.shape[0], NONE, np.uint32)
                for k, v in matcher.Match(q, t).items():
                    got[k] = v
                assert np.array_equal(got, m), name


def test_match_pairs_golden_tensor_core(matcher, sift_golden):
    names, descs, pairs = [], [], []
    for name, g in sift_golden.items():
        names.append(name); pairs.append((len(descs), len(descs) + 1)); descs += [g["d1"], g["d2"]]
    matcher.Upload(descs)
    out = matcher.MatchPairs(pairs, emit_all=True)
    for name, pm in zip(names, out):
        g = sift_golden[name]
        n12, n21, nm, gated, conn = g["pair"].tolist()
        if "float" in name:
            assert abs(pm.n12 - n12) <= 1 and abs(pm.n21 - n21) <= 1, name
        else:
            assert (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected) == (n12, n21, nm, bool(gated), bool(conn)), name
            assert np.array_equal(pm.matches.reshape(-1, 2), g["matches"].reshape(-1, 2)), name


@pytest.mark.parametrize("n1,n2,integer", [(1000, 1500, True), (300, 129, True), (257, 513, True), (128, 128, True), (1, 300, True),
                                           (640, 2, True), (2048, 2048, True), (1200, 900, False), (2048, 2048, False),
                                           (8192, 8192, False)])
def test_match_pairs_sizes_vs_opencv(matcher, n1, n2, integer):
    from eacham_b200 import synth
    n = max(n1, n2)
    imgs = synth.sift_image_set(2, n, seed=n1 * 31 + n2, pool=int(n * 1.2) + 50, share=0.45, integer_valued=integer)
    a, b = np.ascontiguousarray(imgs[0][:n1]), np.ascontiguousarray(imgs[1][:n2])
    if integer and n1 > 20 and n2 > 20:
        b[3] = b[11]; a[5] = a[2]; b[7] = a[0]          # duplicates: ties, 0/0 and d0 = 0 cases
    matcher.Upload([a, b])
    pm = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    want = _ref_pair(a, b)
    if integer:
        assert (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected) == \
            (want["n12"], want["n21"], want["n_mutual"], want["gated"], want["connected"])
        assert np.array_equal(pm.matches.reshape(-1, 2), want["matches"].reshape(-1, 2))
    else:
        assert _agreement(pm, want) >= 0.999
        assert abs(pm.n12 - want["n12"]) <= max(1, want["n12"] // 500)
    pt = matcher.MatchPairs([(1, 0)], emit_all=True)[0]
    assert sorted(map(tuple, pt.matches[:, ::-1].tolist())) == sorted(map(tuple, pm.matches.tolist()))


def test_tensor_core_path_agrees_with_exact_fp32_path(matcher):
    """Same pairs through the tcgen05 scorer and through the all-FP32 kernels (cfg flag): identical on integer-valued
    data, >= 99.9 % on general floats."""
    import ctypes
    import eacham_b200
    from eacham_b200 import synth, _lib as L

    class Exact(eacham_b200.FeatureMatcherGpu):
        def __init__(self):
            self.inliersRatio = 0.8; self.ratio = 0.8; self.min_dir = 30; self.min_mutual = 30; self.cross_check = True
            self._lib = L.load(); self.device = 0
            cfg = L.Config(device=0, max_images=0, match_buffer_entries=0, flags=1)
            h = ctypes.c_void_p(); L.check(self._lib.eacham_gpu_create(ctypes.byref(cfg), ctypes.byref(h))); self._h = h

    for integer in (True, False):
        imgs = synth.sift_image_set(4, 1536, seed=99, pool=2500, share=0.4, integer_valued=integer)
        pairs = synth.exhaustive_pairs(4)
        matcher.Upload(imgs)
        tc = matcher.MatchPairs(pairs, emit_all=True)
        with Exact() as ex:
            ex.Upload(imgs)
            ref = ex.MatchPairs(pairs, emit_all=True)
        for a, b in zip(tc, ref):
            want = dict(matches=b.matches)
            if integer:
                assert (a.n12, a.n21, a.n_mutual) == (b.n12, b.n21, b.n_mutual) and np.array_equal(a.matches, b.matches)
            else:
                assert _agreement(a, want) >= 0.999


def test_sift_empty_and_tiny_images(matcher):
    from eacham_b200 import synth
    a = synth.sift_image_set(1, 200, seed=1, pool=300)[0]
    empty = np.zeros((0, 128), np.float32)
    matcher.Upload([a, empty, a[:1].copy(), a[:2].copy()])
    out = matcher.MatchPairs([(0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (1, 1)], emit_all=True)
    for pm, (x, y) in zip(out, [(a, empty), (empty, a), (a, a[:1]), (a[:1], a), (a, a[:2]), (empty, empty)]):
        want = _ref_pair(x, y)
        assert (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected) == (want["n12"], want["n21"], want["n_mutual"], want["gated"], want["connected"])
    assert matcher.Match(a, empty) == {} and matcher.Match(empty, a) == {}


def test_sift_full_size_properties(matcher):
    """BASELINE config-3 shape (8k SIFT per image): an image against a row-permuted copy of itself returns the
    permutation for every row (size-independent property; no CPU oracle needed at this size)."""
    from eacham_b200 import synth
    img = synth.sift_image_set(1, 8192, seed=3, pool=40000)[0]
    img = np.unique(img, axis=0)                    # drop exact duplicate rows so that every row has one zero-distance partner
    n = img.shape[0]
    perm = np.random.default_rng(3).permutation(n)
    matcher.Upload([img, np.ascontiguousarray(img[perm])])
    pm = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    assert pm.n12 == n and pm.n21 == n and pm.n_mutual == n
    assert np.array_equal(pm.matches[:, 0], np.arange(n)) and np.array_equal(pm.matches[:, 1], inv)


@pytest.mark.parametrize("sizes", [(1536, 1300, 257, 129), (8, 3000, 383, 1), (2049, 2, 640, 511)])
def test_sift_engines_agree_on_ragged_sizes(sizes):
    """The default engine (D and D^T on the tensor cores, pruned scans, tc_sift_kernels.cuh), the round-1 tensor kernel
    (EACHAM_CFG_SIFT_TC_V1) and the all-FP32 kernels return the same pairs on integer-valued rows of ragged sizes: partial last row
    blocks (one resident half only), partial last tiles, fewer rows than a tile, single rows; both roles of every image."""
    import eacham_b200
    from eacham_b200 import synth
    rng = np.random.default_rng(sum(sizes))
    pool = synth.sift_image_set(1, 6000, seed=17, pool=6000)[0]
    imgs = [np.ascontiguousarray(pool[rng.choice(pool.shape[0], n, replace=False)]) for n in sizes]
    pairs = [(i, j) for i in range(len(imgs)) for j in range(len(imgs)) if i != j]
    out = {}
    for engine in ("tensor", "tensor_v1", "fp32"):
        with eacham_b200.FeatureMatcherGpu(0.8, sift_engine=engine, min_dir=0, min_mutual=0) as m:
            m.Upload(imgs)
            out[engine] = m.MatchPairs(pairs, emit_all=True)
            if engine == "tensor":                  # the per-call route runs the same kernel in single-direction mode
                for i, j in pairs[:4]:
                    assert len(m.Match(imgs[i], imgs[j])) == _ref_pair(imgs[i], imgs[j], min_dir=0, min_mutual=0)["n12"], (i, j)
    for a, b, c, (i, j) in zip(out["tensor"], out["tensor_v1"], out["fp32"], pairs):
        assert (a.n12, a.n21, a.n_mutual) == (b.n12, b.n21, b.n_mutual) == (c.n12, c.n21, c.n_mutual), (i, j)
        assert np.array_equal(a.matches, b.matches) and np.array_equal(a.matches, c.matches), (i, j)
        ref = _ref_pair(imgs[i], imgs[j], min_dir=0, min_mutual=0)
        assert (a.n12, a.n21, a.n_mutual) == (ref["n12"], ref["n21"], ref["n_mutual"]), (i, j)


def test_sift_engines_agree_on_large_images():
    """20,000 x 9,000 rows: more than 128 tiles and more than 64 row blocks per pair (tile / row-block ids beyond one byte), partial
    last block and tile; the default engine, the round-1 tensor kernel and the all-FP32 kernels return the same pair."""
    import eacham_b200
    from eacham_b200 import synth
    a, b = synth.sift_image_set(2, 20000, seed=21, pool=60000, share=0.3)
    b = np.ascontiguousarray(b[:9000])
    out = {}
    for engine in ("tensor", "tensor_v1", "fp32"):
        with eacham_b200.FeatureMatcherGpu(0.8, sift_engine=engine) as m:
            m.Upload([a, b])
            out[engine] = m.MatchPairs([(0, 1), (1, 0)], emit_all=True)
    for x, y, z in zip(out["tensor"], out["tensor_v1"], out["fp32"]):
        assert x.n12 > 100 and x.n21 > 100
        assert (x.n12, x.n21, x.n_mutual) == (y.n12, y.n21, y.n_mutual) == (z.n12, z.n21, z.n_mutual)
        assert np.array_equal(x.matches, y.matches) and np.array_equal(x.matches, z.matches)


@pytest.mark.parametrize("ratio", [0.6, 0.95])
def test_sift_other_ratios(ratio):
    """The pruning bar of the default SIFT engine is derived from the ratio (ratio^2 x best): other thresholds than 0.8, against OpenCV,
    batched and per-call."""
    import eacham_b200
    from eacham_b200 import synth
    a, b = synth.sift_image_set(2, 1400, seed=31, pool=2200, share=0.5)
    b = np.ascontiguousarray(b[:1100])
    want = _ref_pair(a, b, ratio=ratio, min_dir=0, min_mutual=0)
    with eacham_b200.FeatureMatcherGpu(ratio, ratio=ratio, min_dir=0, min_mutual=0) as m:
        m.Upload([a, b])
        pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
        assert (pm.n12, pm.n21, pm.n_mutual) == (want["n12"], want["n21"], want["n_mutual"])
        assert np.array_equal(pm.matches, np.asarray(want["matches"]).reshape(-1, 2))
        one = m.Match(a, b)
        assert len(one) == want["n12"]
