"""The two ORB engines -- tcgen05 FP8 (default) and XOR+POPC (EACHAM_CFG_ORB_POPC) -- must return the same bytes, and
the tensor engine must handle extra shapes (sizes around its 128-row blocks). test_gpu_orb.py runs every parity test on both."""
import numpy as np
import pytest

from oracle import oracle as O
import cases
from test_gpu_orb import _assert_pair_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tensor", "tensor_alu", "tensor_v1"])
def tmatcher(request):
    import eacham_b200
    m = eacham_b200.FeatureMatcherGpu(0.8, orb_engine=request.param)
    yield m
    m.close()


def test_golden_cases_tensor_engine(tmatcher, orb_golden, orb_set_golden):
    names, descs, pairs = [], [], []
    allg = dict(orb_golden); allg.update(orb_set_golden)
    for name, g in allg.items():
        names.append(name); pairs.append((len(descs), len(descs) + 1)); descs += [g["d1"], g["d2"]]
    tmatcher.Upload(descs)
    for emit_all in (True, False):
        for name, pm in zip(names, tmatcher.MatchPairs(pairs, emit_all=emit_all)):
            n12, n21, nm, gated, conn = allg[name]["pair"].tolist()
            want = dict(n12=n12, n21=n21, n_mutual=nm, gated=gated, connected=conn,
                        matches=allg[name]["matches"] if (emit_all or conn) else np.zeros((0, 2), np.uint32))
            _assert_pair_equal(pm, want, name)


@pytest.mark.parametrize("n1,n2", [(4096, 4096), (4097, 511), (513, 4095), (1, 700), (700, 1), (2, 2), (5000, 4500), (129, 257)])
def test_sizes_tensor_engine(tmatcher, n1, n2):
    rng = np.random.default_rng(n1 * 7919 + n2)
    a, b = cases.planted_pair(rng, n1, n2, min(n1, n2, 90) if min(n1, n2) > 40 else 0, dup_filler=False)
    if n1 > 10 and n2 > 10:
        a[3] = a[7]; b[5] = b[9]; a[1] = 0; b[2] = 0; b[4] = 0; a[9] = 255; b[8] = 255
    tmatcher.Upload([a, b])
    _assert_pair_equal(tmatcher.MatchPairs([(0, 1)], emit_all=True)[0], O.c_match_pair(a, b), f"{n1}x{n2}")
    _assert_pair_equal(tmatcher.MatchPairs([(1, 0)], emit_all=True)[0], O.c_match_pair(b, a), f"{n2}x{n1}")


def test_engines_agree_on_an_exhaustive_set(tmatcher):
    import eacham_b200
    from eacham_b200 import synth
    matcher = eacham_b200.FeatureMatcherGpu(0.8, orb_engine="popc")
    imgs = synth.orb_image_set(10, 2048, seed=5, pool=6000)
    pairs = synth.exhaustive_pairs(len(imgs))
    matcher.Upload(imgs); tmatcher.Upload(imgs)
    r1, b1 = matcher.MatchPairsRaw(pairs, emit_all=True)
    r2, b2 = tmatcher.MatchPairsRaw(pairs, emit_all=True)
    for k in range(len(pairs)):
        assert (r1[k]["n12"], r1[k]["n21"], r1[k]["n_mutual"], r1[k]["flags"]) == (r2[k]["n12"], r2[k]["n21"], r2[k]["n_mutual"], r2[k]["flags"])
        assert np.array_equal(b1[int(r1[k]["offset"]): int(r1[k]["offset"] + r1[k]["count"])],
                              b2[int(r2[k]["offset"]): int(r2[k]["offset"] + r2[k]["count"])])
    assert int(r1["count"].sum()) > 0
    matcher.close()
