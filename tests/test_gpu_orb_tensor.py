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


def test_passed_over_scores_decide_correctly_orb():
    """The pruned rows path of the default ORB engine (tc_orb_kernels.cuh) does not merge scores >= 0.85 x the running best; only their
    minimum is kept and folded into the second best. Scenarios with chosen Hamming distances (query s = a 32-bit signature; its train rows
    = the signature + d more bits; everything else is >= 64 bits away), the interesting rows arriving in later tiles: a passed-over row is
    the true best, the true second (match and no match), exactly at the ratio, alternating chains. Batched and per-call routes, against the
    C oracle and the XOR+POPC engine."""
    import eacham_b200
    scen = [
        ([(3, 50), (500, 45)], None),
        ([(4, 50), (501, 39)], 501),
        ([(5, 30), (502, 37), (630, 38)], None),
        ([(6, 30), (503, 38), (631, 39)], 6),
        ([(7, 40), (504, 50)], None),
        ([(8, 60), (505, 54), (632, 50), (700, 44)], None),
        ([(9, 60), (506, 54), (633, 40)], 633),
    ]
    n_train = 768

    def bits_to_bytes(b):
        return np.packbits(b.astype(np.uint8), axis=-1, bitorder="little")

    qb = np.zeros((len(scen), 256), np.uint8)
    tb = np.zeros((n_train, 256), np.uint8)
    used = set()
    for s, (rows, _) in enumerate(scen):
        qb[s, s * 16:s * 16 + 32 if s * 16 + 32 <= 128 else 128] = 1
        for j, d in rows:
            assert j not in used
            used.add(j)
            tb[j] = qb[s]
            tb[j, 128:128 + d] = 1
    rng = np.random.default_rng(13)
    for j in range(n_train):
        if j not in used:
            tb[j] = rng.integers(0, 2, 256)
    q, t = bits_to_bytes(qb), bits_to_bytes(tb)
    want = O.c_match(q, t)
    assert all(want.get(s) == e for s, (_, e) in enumerate(scen)), want
    for engine in ("tensor", "popc"):
        with eacham_b200.FeatureMatcherGpu(0.8, orb_engine=engine, min_dir=0, min_mutual=0, cross_check=False) as m:
            assert m.Match(q, t) == want, engine
            m.Upload([q, t])
            pm = m.MatchPairs([(0, 1)], emit_all=True)[0]
            ref = O.c_match_pair(q, t, min_dir=0, min_mutual=0)
            assert (pm.n12, pm.n21, pm.n_mutual) == (ref["n12"], ref["n21"], ref["n_mutual"]), engine
