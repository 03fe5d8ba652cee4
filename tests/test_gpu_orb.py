"""GPU parity for the ORB (256-bit Hamming) path, through the C ABI, against the golden vectors and the oracle.
Bit-exact: index sets, distances, counts, flags."""
import threading

import numpy as np
import pytest

from oracle import oracle as O
import cases

pytestmark = pytest.mark.gpu
NONE = 0xFFFFFFFF


@pytest.fixture(scope="module", params=["tensor", "popc", "tensor_alu", "tensor_v1"])
def matcher(request):
    """Every test below runs on both ORB engines: the tcgen05 FP8 engine (default) and the XOR+POPC kernel."""
    import eacham_b200
    m = eacham_b200.FeatureMatcherGpu(0.8, orb_engine=request.param)
    yield m
    m.close()


def _map_arr(m, n):
    a = np.full(n, NONE, np.uint32)
    for k, v in m.items():
        a[k] = v
    return a


def _assert_pair_equal(pm, want, name=""):
    assert (pm.n12, pm.n21, pm.n_mutual, pm.gated, pm.connected) == \
        (want["n12"], want["n21"], want["n_mutual"], bool(want["gated"]), bool(want["connected"])), name
    assert np.array_equal(pm.matches.reshape(-1, 2), np.asarray(want["matches"]).reshape(-1, 2)), name


def test_knn2_and_match_golden(matcher, orb_golden):
    for name, g in orb_golden.items():
        for q, t, idx, dist, m in ((g["d1"], g["d2"], g["idx12"], g["dist12"], g["m12"]),
                                   (g["d2"], g["d1"], g["idx21"], g["dist21"], g["m21"])):
            i, s = matcher.knnMatch(q, t)
            assert np.array_equal(i, idx), name
            assert np.array_equal(s, dist), name
            assert np.array_equal(_map_arr(matcher.Match(q, t), q.shape[0]), m), name


def test_match_pairs_golden_fused(matcher, orb_golden, orb_set_golden):
    """All golden cases in ONE batched call (ragged sizes, empty images, gates at 29/30/31, ties, NaN)."""
    names, descs, pairs = [], [], []
    for name, g in list(orb_golden.items()) + list(orb_set_golden.items()):
        names.append(name)
        pairs.append((len(descs), len(descs) + 1))
        descs += [g["d1"], g["d2"]]
    matcher.Upload(descs)
    allg = dict(orb_golden); allg.update(orb_set_golden)
    for emit_all in (True, False):
        out = matcher.MatchPairs(pairs, emit_all=emit_all)
        for name, pm in zip(names, out):
            n12, n21, nm, gated, conn = allg[name]["pair"].tolist()
            want = dict(n12=n12, n21=n21, n_mutual=nm, gated=gated, connected=conn,
                        matches=allg[name]["matches"] if (emit_all or conn) else np.zeros((0, 2), np.uint32))
            _assert_pair_equal(pm, want, name)


@pytest.mark.parametrize("n1,n2", [(4096, 4096), (4097, 511), (513, 4095), (1, 700), (700, 1), (2, 2), (5000, 4500), (9000, 300)])
def test_match_pairs_vs_c_oracle_sizes(matcher, n1, n2):
    """Row-block boundaries (512-row slots, 4096-row blocks), chunk boundaries (256 columns), tiny images."""
    rng = np.random.default_rng(n1 * 7919 + n2)
    a, b = cases.planted_pair(rng, n1, n2, min(n1, n2, 90) if min(n1, n2) > 40 else 0, dup_filler=False)
    # sprinkle duplicates and zero rows to force ties in both directions
    if n1 > 10 and n2 > 10:
        a[3] = a[7]; b[5] = b[9]; a[1] = 0; b[2] = 0; b[4] = 0
    matcher.Upload([a, b])
    pm = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    _assert_pair_equal(pm, O.c_match_pair(a, b), f"{n1}x{n2}")
    # the transposed pair must give the transposed answer
    pt = matcher.MatchPairs([(1, 0)], emit_all=True)[0]
    assert (pt.n12, pt.n21, pt.n_mutual) == (pm.n21, pm.n12, pm.n_mutual)
    assert sorted(map(tuple, pt.matches[:, ::-1].tolist())) == sorted(map(tuple, pm.matches.tolist()))
    # one-direction API agrees with the oracle too
    assert matcher.Match(a, b) == O.c_match(a, b)
    i, s = matcher.knnMatch(b, a)
    io, so = O.c_knn2(b, a)
    assert np.array_equal(i, io) and np.array_equal(s, so)


def test_exhaustive_set_vs_cv2(matcher):
    """A small exhaustive image set (BASELINE config-1 shape, scaled down): every pair equals OpenCV + reference logic."""
    from eacham_b200 import synth
    imgs = synth.orb_image_set(12, 2048, seed=1, pool=6000)
    pairs = synth.exhaustive_pairs(len(imgs))
    matcher.Upload(imgs)
    out = matcher.MatchPairs(pairs, emit_all=True)
    ref = O.cv2_match_pair_fast if O.have_cv2() else O.c_match_pair
    n_conn = 0
    for (i, j), pm in zip(pairs.tolist(), out):
        _assert_pair_equal(pm, ref(imgs[i], imgs[j]), f"pair {i},{j}")
        n_conn += pm.connected
    assert n_conn > 0


def test_options_ratio_and_gates(matcher):
    rng = np.random.default_rng(5)
    a, b = cases.planted_pair(rng, 600, 640, 50, d_good=60)
    import eacham_b200
    for ratio, min_dir, min_mutual in ((0.8, 30, 30), (0.7, 10, 5), (0.95, 0, 0), (0.5, 1, 0)):
        with eacham_b200.FeatureMatcherGpu(0.8, ratio=ratio, min_dir=min_dir, min_mutual=min_mutual, orb_engine=matcher.orb_engine) as m:
            m.Upload([a, b])
            pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
            _assert_pair_equal(pm, O.c_match_pair(a, b, ratio, min_dir, min_mutual), f"ratio {ratio}")
            assert m.Match(a, b) == O.c_match(a, b, ratio)
    with eacham_b200.FeatureMatcherGpu(0.8, cross_check=False, orb_engine=matcher.orb_engine) as m:      # first -> second direction only
        m.Upload([a, b])
        pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
        assert pm.best12() == O.c_match(a, b)


def test_ratio_above_one_keeps_opencv_tie_order(matcher):
    """With ratio > 1 a query whose two best neighbours TIE passes the test, so OpenCV's tie order (lowest train index first)
    becomes visible in the match set; the tensor engines hand such calls to the packed-key XOR+POPC kernels."""
    import eacham_b200
    rng = np.random.default_rng(17)
    a = rng.integers(0, 4, (300, 32), dtype=np.uint8)          # low entropy: ties everywhere
    b = rng.integers(0, 4, (280, 32), dtype=np.uint8)
    b[200] = b[3]; b[150] = b[3]; a[10] = b[3]
    with eacham_b200.FeatureMatcherGpu(0.8, ratio=1.25, min_dir=0, min_mutual=0, orb_engine=matcher.orb_engine) as m:
        m.Upload([a, b])
        _assert_pair_equal(m.MatchPairs([(0, 1)], emit_all=True)[0], O.c_match_pair(a, b, 1.25, 0, 0), "ratio 1.25")


def test_strided_rows(matcher):
    """cv::Mat with step > cols (a ROI): the ABI takes (ptr, rows, step)."""
    rng = np.random.default_rng(11)
    big = rng.integers(0, 256, (300, 48), dtype=np.uint8)
    q = big[:, :32]; t = big[::-1][:150, 8:40]
    assert not q.flags["C_CONTIGUOUS"]
    assert matcher.Match(q, t) == O.c_match(np.ascontiguousarray(q), np.ascontiguousarray(t))
    matcher.Upload([q, t])
    pm = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    _assert_pair_equal(pm, O.c_match_pair(np.ascontiguousarray(q), np.ascontiguousarray(t)))


def test_concurrent_match_calls_on_one_handle(matcher):
    """The reference calls Match concurrently on ONE matcher from TBB workers (main.cpp:98-109)."""
    rng = np.random.default_rng(3)
    sets = [cases.planted_pair(rng, 300 + 17 * k, 280 + 13 * k, 40) for k in range(8)]
    want = [O.c_match(a, b) for a, b in sets]
    got = [None] * len(sets)

    def work(k):
        for _ in range(3):
            got[k] = matcher.Match(*sets[k])

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(sets))]
    [t.start() for t in th]; [t.join() for t in th]
    assert got == want


def test_error_codes(matcher):
    from eacham_b200 import _lib as L
    import eacham_b200
    with eacham_b200.FeatureMatcherGpu(0.8, orb_engine=matcher.orb_engine) as m:
        with pytest.raises(L.EachamGpuError) as e:
            m.MatchPairs([(0, 1)])
        assert e.value.code == L.ERR_NOT_COMMITTED
        m.Upload([np.zeros((4, 32), np.uint8), np.zeros((4, 128), np.float32)])
        with pytest.raises(L.EachamGpuError) as e:
            m.MatchPairs([(0, 1)])
        assert e.value.code == L.ERR_KIND_MISMATCH
        with pytest.raises(L.EachamGpuError) as e:
            m.MatchPairs([(0, 7)])
        assert e.value.code == L.ERR_NOT_COMMITTED
        with pytest.raises(TypeError):
            m.Match(np.zeros((4, 31), np.uint8), np.zeros((4, 32), np.uint8))


def test_buffer_too_small_reports_need(matcher):
    rng = np.random.default_rng(8)
    a, b = cases.planted_pair(rng, 300, 300, 60)
    matcher.Upload([a, b])
    tiny = np.empty(10, dtype=[("query", "<u4"), ("train", "<u4")])
    res, buf = matcher.MatchPairsRaw([(0, 1)], buf=tiny)      # wrapper retries through fetch_results
    assert res["n_mutual"][0] == 60 and buf.shape[0] == 60


def test_full_size_properties(matcher):
    """BASELINE config-2 shape (4k ORB per image), size-independent properties instead of a CPU oracle:
    an image against a row-permuted copy of itself must return exactly the permutation for every unique row."""
    from eacham_b200 import synth
    img = synth.orb_image_set(1, 4096, seed=2, pool=20000, dup_frac=0.0, zero_rows=0)[0]
    rng = np.random.default_rng(2)
    perm = rng.permutation(4096)
    matcher.Upload([img, np.ascontiguousarray(img[perm])])
    pm = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    inv = np.empty(4096, np.int64); inv[perm] = np.arange(4096)
    assert pm.n12 == 4096 and pm.n21 == 4096 and pm.n_mutual == 4096 and pm.connected
    assert np.array_equal(pm.matches[:, 0], np.arange(4096)) and np.array_equal(pm.matches[:, 1], inv)
    # idempotence / determinism: same call, same bytes
    pm2 = matcher.MatchPairs([(0, 1)], emit_all=True)[0]
    assert np.array_equal(pm.matches, pm2.matches)


def test_cpp_shim_program():
    """include/eacham/FeatureMatcherGpu.h driven like apps/sfm/main.cpp drives its matcher (threads + batched)."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "shim_test.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "shim_test.cpp"),
                    "-L" + os.path.join(root, "eacham_b200"), "-leacham_gpu", "-lpthread", "-o", exe], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "eacham_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_images_beyond_fused_limit_use_gpu_fallback(matcher):
    """More than 16,384 descriptors per image exceeds the fused kernel's shared-memory budget: the library switches to
    its per-pair exact GPU kernels (never to the CPU) and the answer is unchanged."""
    rng = np.random.default_rng(17)
    a, b = cases.planted_pair(rng, 17000, 700, 80, dup_filler=False)
    matcher.Upload([a, b])
    for pair in ((0, 1), (1, 0)):
        pm = matcher.MatchPairs([pair], emit_all=True)[0]
        x, y = (a, b) if pair == (0, 1) else (b, a)
        _assert_pair_equal(pm, O.c_match_pair(x, y), str(pair))


def test_window_pairs_kitti_shape(matcher):
    """BASELINE config 4 shape, scaled down: a frame sequence whose neighbours overlap, sliding-window pair list."""
    from eacham_b200 import synth
    imgs = synth.orb_image_set(40, 2048, seed=4, pool=12000, window=4000)
    pairs = synth.window_pairs(len(imgs), 5)
    matcher.Upload(imgs)
    out = matcher.MatchPairs(pairs, emit_all=True)
    ref = O.cv2_match_pair_fast if O.have_cv2() else O.c_match_pair
    rng = np.random.default_rng(0)
    for k in rng.choice(len(pairs), size=40, replace=False):
        i, j = pairs[k].tolist()
        _assert_pair_equal(out[k], ref(imgs[i], imgs[j]), f"frames {i},{j}")
    near = [pm.n_mutual for pm in out if pm.second - pm.first == 1]
    far = [pm.n_mutual for pm in out if pm.second - pm.first == 5]
    assert np.mean(near) > np.mean(far)                 # neighbouring frames share more landmarks


def test_match_graph_dump_roundtrip_from_gpu(matcher, tmp_path):
    from eacham_b200 import synth, graph_io
    imgs = synth.orb_image_set(6, 1024, seed=8, pool=2500)
    pairs = synth.exhaustive_pairs(6)
    matcher.Upload(imgs)
    res, buf = matcher.MatchPairsRaw(pairs)
    p = str(tmp_path / "graph.bin")
    graph_io.save_match_graph(p, pairs, res, buf, n_images=6)
    p2, r2, m2, n = graph_io.load_match_graph(p)
    assert np.array_equal(r2, res) and np.array_equal(m2, buf) and n == 6
    edges = {(i, j): b12 for i, j, b12, _ in graph_io.connected_edges(p2, r2, m2)}
    for (i, j), b12 in edges.items():
        assert b12 == {int(a): int(b) for a, b in O.c_match_pair(imgs[i], imgs[j])["matches"]}
    assert len(edges) > 0


@pytest.mark.parametrize("kind", ["orb", "sift"])
def test_match_route_cache_and_concurrency(kind):
    """eacham_gpu_match on the tensor-core engines: descriptors cached on the device across calls (keyed by pointer + sampled-row
    hash), calls from several threads in flight at once. Same answers as the exact CPU matcher, also after a matrix is rewritten in
    place (a sampled row changes -> re-upload), with the cache switched off, and on the legacy kernels."""
    import eacham_b200
    from eacham_b200 import synth
    if kind == "orb":
        imgs = synth.orb_image_set(6, 1300, seed=41, pool=3000)
        imgs[2] = np.ascontiguousarray(imgs[2][:700])
    else:
        imgs = synth.sift_image_set(5, 700, seed=42, pool=900)
        imgs[1] = np.ascontiguousarray(imgs[1][:333])
    n = len(imgs)
    want = {(i, j): O.c_match(imgs[i], imgs[j]) for i in range(n) for j in range(n) if i != j}
    for kw in ({}, {"match_cache": False}, {"match_legacy": True}):
        with eacham_b200.FeatureMatcherGpu(0.8, **kw) as m:
            got = {}

            def work(w):
                for k, (i, j) in enumerate(sorted(want)):
                    if k % 4 == w:
                        got[(i, j)] = m.Match(imgs[i], imgs[j])

            th = [threading.Thread(target=work, args=(w,)) for w in range(4)]
            [t.start() for t in th]; [t.join() for t in th]
            assert got == want, kw
            if not kw:
                # rewrite image 0 in place: its first row is one of the sampled rows, so the cached copy must not be reused
                old = imgs[0].copy()
                imgs[0][:] = imgs[3][: imgs[0].shape[0]] if imgs[3].shape[0] >= imgs[0].shape[0] else np.roll(imgs[0], 7, axis=0)
                assert m.Match(imgs[0], imgs[1]) == O.c_match(imgs[0], imgs[1])
                assert m.Match(imgs[1], imgs[0]) == O.c_match(imgs[1], imgs[0])
                imgs[0][:] = old
                assert m.Match(imgs[0], imgs[1]) == want[(0, 1)]


def test_batched_staging_equals_per_image_staging():
    """eacham_gpu_set_descriptors_batch (what Upload uses; several host threads copy into the pinned staging buffer) lays out the same arena as
    one eacham_gpu_set_descriptors call per image: ragged, empty and strided images, > 4 MiB in total so that more than one worker runs;
    and its argument errors are reported before anything is staged."""
    import ctypes
    import eacham_b200
    from eacham_b200 import _lib as L
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 256, (30000, 40), dtype=np.uint8)
    imgs = []
    for k in range(160):
        n = int(rng.integers(0, 2049)) if k % 17 else 0
        rows = pool[rng.choice(pool.shape[0], n, replace=False)]
        imgs.append(rows[:, 4:36] if k % 3 == 0 else np.ascontiguousarray(rows[:, :32]))       # every third one strided (step 40)
    pairs = [(i, j) for i in range(len(imgs)) for j in range(i + 1, min(i + 4, len(imgs)))]
    with eacham_b200.FeatureMatcherGpu(0.8) as a, eacham_b200.FeatureMatcherGpu(0.8) as b:
        a.Upload(imgs)                                              # batched
        for i, d in enumerate(imgs):
            b.SetDescriptors(i, d)                                  # one call per image
        b.Commit()
        assert a.arena()[1] == b.arena()[1]
        ra, rb = a.MatchPairs(pairs, emit_all=True), b.MatchPairs(pairs, emit_all=True)
        for x, y in zip(ra, rb):
            assert (x.n12, x.n21, x.n_mutual, x.gated, x.connected) == (y.n12, y.n21, y.n_mutual, y.gated, y.connected)
            assert np.array_equal(x.matches, y.matches)
        # errors: nothing staged, the committed batch still answers
        ptrs = (ctypes.c_void_p * 2)(imgs[1].ctypes.data, None)
        rows = np.array([imgs[1].shape[0], 5], np.uint32)
        with pytest.raises(L.EachamGpuError) as e:
            L.check(a._lib.eacham_gpu_set_descriptors_batch(a._h, 0, 2, L.KIND_ORB256, ptrs, rows.ctypes.data_as(ctypes.c_void_p), None))
        assert e.value.code == L.ERR_INVALID_ARG
        rows[1] = 70000
        ptrs[1] = imgs[1].ctypes.data
        with pytest.raises(L.EachamGpuError) as e:
            L.check(a._lib.eacham_gpu_set_descriptors_batch(a._h, 0, 2, L.KIND_ORB256, ptrs, rows.ctypes.data_as(ctypes.c_void_p), None))
        assert e.value.code == L.ERR_TOO_LARGE
        again = a.MatchPairs(pairs[:8], emit_all=True)
        for x, y in zip(again, ra[:8]):
            assert np.array_equal(x.matches, y.matches)
