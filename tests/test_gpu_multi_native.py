"""The single-process multi-device route of the C ABI (eacham_gpu_multi_*): results must be byte-identical to one device's,
in input order, whatever the number of devices. On a 1-GPU box the sharding / re-interleaving / offset rebasing logic still runs
with the SAME device listed twice (parallel-H2D mode: NCCL refuses duplicate devices) and with one device through NCCL."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _single(imgs, pairs, emit_all, **kw):
    import eacham_b200
    with eacham_b200.FeatureMatcherGpu(0.8, **kw) as m:
        m.Upload(imgs)
        res, buf = m.MatchPairsRaw(pairs, emit_all=emit_all)
        return res.copy(), buf.copy()


def _same(res1, buf1, res2, buf2):
    assert len(res1) == len(res2)
    for k in range(len(res1)):
        a, b = res1[k], res2[k]
        assert (a["n12"], a["n21"], a["n_mutual"], a["flags"], a["count"]) == (b["n12"], b["n21"], b["n_mutual"], b["flags"], b["count"]), k
        assert np.array_equal(buf1[int(a["offset"]): int(a["offset"] + a["count"])], buf2[int(b["offset"]): int(b["offset"] + b["count"])]), k


@pytest.mark.parametrize("devices,parallel_h2d", [([0], False), ([0, 0], True), ([0, 0, 0], True)])
def test_multi_equals_single_on_one_gpu(devices, parallel_h2d):
    import eacham_b200
    from eacham_b200 import synth
    imgs = synth.orb_image_set(9, 1100, seed=21, pool=2500)
    imgs[4] = imgs[4][:700]                                   # ragged
    pairs = synth.exhaustive_pairs(len(imgs))
    want = _single(imgs, pairs, True)
    with eacham_b200.MultiGpuMatcher(devices, parallel_h2d=parallel_h2d) as m:
        assert m._lib.eacham_gpu_multi_device_count(m._h) == len(devices)
        m.Upload(imgs)
        for emit_all in (True, False):
            res, buf = m.MatchPairsRaw(pairs, emit_all=emit_all)
            _same(*_single(imgs, pairs, emit_all), res, buf)
        out = m.MatchPairs(pairs[:5], emit_all=True)
        for (i, j), pm in zip(pairs[:5].tolist(), out):
            w = O.c_match_pair(imgs[i], imgs[j])
            assert (pm.n12, pm.n21, pm.n_mutual) == (w["n12"], w["n21"], w["n_mutual"]) and np.array_equal(pm.matches, w["matches"])
        t = m.timing()
        assert t["kernel_ms_max"] > 0 and t["kernel_launches"] >= 1
        # a second upload with another set reuses the handles
        imgs2 = synth.orb_image_set(5, 600, seed=22, pool=900)
        m.Upload(imgs2)
        p2 = synth.exhaustive_pairs(5)
        _same(*_single(imgs2, p2, True), *m.MatchPairsRaw(p2, emit_all=True))
    assert int(want[0]["count"].sum()) > 0


def test_multi_sift_and_errors():
    import eacham_b200
    from eacham_b200 import synth, _lib as L
    s = synth.sift_image_set(4, 400, seed=5, pool=350)
    pairs = synth.exhaustive_pairs(4)
    with eacham_b200.MultiGpuMatcher([0, 0], parallel_h2d=True) as m:
        m.Upload(s)
        _same(*_single(s, pairs, True), *m.MatchPairsRaw(pairs, emit_all=True))
        with pytest.raises(L.EachamGpuError) as e:
            m.MatchPairsRaw([(0, 9)])
        assert e.value.code == L.ERR_NOT_COMMITTED
    with pytest.raises(L.EachamGpuError):
        eacham_b200.MultiGpuMatcher([99])


def test_multi_all_visible_devices_with_nccl():
    """Every visible GPU, arena replicated by ncclBroadcast (the configuration a C++ caller uses on an 8-GPU box)."""
    import eacham_b200
    from eacham_b200 import synth, _lib as L
    n = L.load().eacham_gpu_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    imgs = synth.orb_image_set(12, 2048, seed=31, pool=6000)
    pairs = synth.exhaustive_pairs(len(imgs))
    with eacham_b200.MultiGpuMatcher(list(range(n))) as m:
        m.Upload(imgs)
        _same(*_single(imgs, pairs, False), *m.MatchPairsRaw(pairs))
        assert m.timing()["broadcast_ms"] > 0
