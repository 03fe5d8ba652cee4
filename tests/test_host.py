"""CPU: the C-ABI library loads and exports every symbol include/eacham_gpu.h declares; host-side helpers.
No compute calls (there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "eacham_gpu.h")).read()
    return sorted(set(re.findall(r"EACHAM_API\s+[\w\s\*]+?\b(eacham_gpu_\w+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from eacham_b200 import build, _lib
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/eacham_gpu.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms
    assert _lib.load().eacham_gpu_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    from eacham_b200 import _lib as L
    assert np.dtype(L.MATCH_DTYPE).itemsize == 8
    assert np.dtype(L.PAIR_DTYPE).itemsize == 8
    assert np.dtype(L.RESULT_DTYPE).itemsize == 32
    assert ctypes.sizeof(L.MatchOpts) == 24 and ctypes.sizeof(L.Config) == 24 and ctypes.sizeof(L.Timing) == 32
    assert ctypes.sizeof(L.MultiTiming) == 32


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from eacham_b200 import _lib as L
    import eacham_b200
    assert L.load().eacham_gpu_device_count() == 0
    with pytest.raises(L.EachamGpuError) as e:
        eacham_b200.FeatureMatcherGpu(0.8)
    assert e.value.code == L.ERR_NO_DEVICE
    with pytest.raises(L.EachamGpuError) as e:
        eacham_b200.MultiGpuMatcher([0, 1])
    assert e.value.code == L.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under eacham_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eacham_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_pair_enumeration():
    from eacham_b200 import synth
    p = synth.exhaustive_pairs(100)
    assert p.shape == (4950, 2) and (p[:, 0] < p[:, 1]).all()          # half of main.cpp:84-92's ordered pairs
    assert synth.exhaustive_pairs(500).shape[0] == 124750
    w = synth.window_pairs(4541, 20)
    assert w.shape[0] == 4541 * 20 - 210 and ((w[:, 1] - w[:, 0]) <= 20).all() and ((w[:, 1] - w[:, 0]) >= 1).all()


def test_synth_is_deterministic_and_shaped():
    from eacham_b200 import synth
    a = synth.orb_image_set(3, 256, seed=4, pool=1000)
    b = synth.orb_image_set(3, 256, seed=4, pool=1000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert a[0].shape == (256, 32) and a[0].dtype == np.uint8
    s = synth.sift_image_set(2, 64, seed=3, pool=100)
    assert s[0].shape == (64, 128) and s[0].dtype == np.float32 and np.array_equal(s[0], np.rint(s[0]))
    assert 400 < np.linalg.norm(s[0], axis=1).mean() < 620


def test_match_graph_roundtrip(tmp_path):
    from eacham_b200 import graph_io, _lib as L
    pairs = np.array([[0, 1], [0, 2], [1, 2]], np.uint32)
    res = np.zeros(3, L.RESULT_DTYPE)
    res[0] = (40, 41, 35, L.PAIR_CONNECTED, 0, 35); res[1] = (3, 50, 0, L.PAIR_GATED, 0, 0); res[2] = (31, 31, 31, L.PAIR_CONNECTED, 35, 31)
    m = np.zeros(66, L.MATCH_DTYPE); m["query"] = np.arange(66); m["train"] = np.arange(66)[::-1]
    p = str(tmp_path / "g.bin")
    graph_io.save_match_graph(p, pairs, res, m, n_images=3)
    p2, r2, m2, n = graph_io.load_match_graph(p)
    assert np.array_equal(p2, pairs) and np.array_equal(r2, res) and np.array_equal(m2, m) and n == 3
    edges = list(graph_io.connected_edges(p2, r2, m2))
    assert [(e[0], e[1], len(e[2])) for e in edges] == [(0, 1, 35), (1, 2, 31)]
    assert edges[0][3][65] == 0                      # best21 is the inverse map
    with pytest.raises(ValueError):
        res_bad = res.copy(); res_bad[2]["count"] = 99
        graph_io.save_match_graph(p, pairs, res_bad, m)
