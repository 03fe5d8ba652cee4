"""Host-side mirror of the reference's matcher interface over the C ABI.

``FeatureMatcherGpu`` keeps the shape of ``eacham::FeatureMatcherFlann``
(/root/reference/modules/base/features/FeatureMatcherFlann.h:11-24: ctor ``(float inliersRatio)``,
``MatchType Match(const cv::Mat&, const cv::Mat&)`` with ``MatchType = unordered_map<unsigned, unsigned>``) and of
``IFeatureMatcher<T>::Match`` (/root/reference/modules/base/features/IFeatureMatcher.h:18-19); ``MatchPairs``
subsumes the pair loop of /root/reference/apps/sfm/main.cpp:84-147. All arithmetic happens in libeacham_gpu.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


def _kind_of(d: np.ndarray) -> int:
    if d.dtype == np.uint8 and d.ndim == 2 and d.shape[1] == 32:
        return L.KIND_ORB256
    if d.dtype == np.float32 and d.ndim == 2 and d.shape[1] == 128:
        return L.KIND_F32X128
    raise TypeError(f"unsupported descriptor matrix {d.dtype} {d.shape}: expected uint8 [N,32] (ORB) or float32 [N,128] (SIFT)")


def _rows_ptr(d: np.ndarray):
    """(pointer, rows, row stride) of a 2-D array whose rows are contiguous (cv::Mat semantics: ptr, rows, step)."""
    if d.strides[1] != d.itemsize:
        d = np.ascontiguousarray(d)
    return d, d.ctypes.data_as(ctypes.c_void_p), d.shape[0], (d.strides[0] if d.shape[0] > 1 else d.shape[1] * d.itemsize)


@dataclass
class PairMatches:
    """Result of the batched path for one unordered pair (what main.cpp:142-146 hands to Graph::Connect)."""
    first: int
    second: int
    n12: int
    n21: int
    n_mutual: int
    gated: bool
    connected: bool
    matches: np.ndarray      # [count, 2] uint32: (row in first, row in second), sorted by row in first

    def best12(self) -> Dict[int, int]:
        return {int(a): int(b) for a, b in self.matches}

    def best21(self) -> Dict[int, int]:
        return {int(b): int(a) for a, b in self.matches}


def _upload_batched(fn, handle, descriptors: Sequence[np.ndarray]) -> None:
    """Runs of consecutive images of one kind go through one batch call each (pointer / row count / stride arrays)."""
    n = len(descriptors)
    keep = []                                   # the (possibly converted) arrays must outlive the call
    ptrs = (ctypes.c_void_p * max(n, 1))()
    rows = np.zeros(max(n, 1), np.uint32)
    strides = (ctypes.c_size_t * max(n, 1))()
    kinds = []
    for i, d in enumerate(descriptors):
        kinds.append(_kind_of(d))
        d, p, r, s = _rows_ptr(d)
        keep.append(d)
        ptrs[i] = p.value if isinstance(p, ctypes.c_void_p) else p
        rows[i] = r
        strides[i] = s
    i = 0
    while i < n:
        j = i
        while j < n and kinds[j] == kinds[i]:
            j += 1
        L.check(fn(handle, i, j - i, kinds[i], ctypes.byref(ptrs, i * ctypes.sizeof(ctypes.c_void_p)),
                   rows[i:].ctypes.data_as(ctypes.c_void_p), ctypes.byref(strides, i * ctypes.sizeof(ctypes.c_size_t))))
        i = j


class FeatureMatcherGpu:
    """B200 matcher with the reference's ``FeatureMatcherFlann`` shape.

    ``inliersRatio`` is accepted for signature compatibility; like the reference
    (FeatureMatcherFlann.cpp:23 hard-codes the double literal 0.8 and never reads the member) the ratio used is
    ``ratio`` (default 0.8).
    """

    MatchType = dict

    def __init__(self, inliersRatio: float = 0.8, *, ratio: float = 0.8, device: int = 0, min_dir: int = 30,
                 min_mutual: int = 30, cross_check: bool = True, match_buffer_entries: int = 0,
                 orb_engine: str = "tensor", sift_exact_fp32: bool = False, sift_engine: str = "tensor", match_cache: bool = True, match_legacy: bool = False):
        self.inliersRatio = float(inliersRatio)
        self.ratio = float(ratio)
        self.min_dir, self.min_mutual, self.cross_check = int(min_dir), int(min_mutual), bool(cross_check)
        self._lib = L.load()
        engines = {"tensor": 0, "popc": L.CFG_ORB_POPC, "tensor_v1": L.CFG_ORB_TC_V1, "tensor_alu": L.CFG_ORB_TC_ALU_SORT}
        if orb_engine not in engines:
            raise ValueError("orb_engine must be 'tensor' (FP8 tensor-core engine with F16 accumulators and packed epilogue, default), "
                             "'popc' (XOR+POPC kernel), 'tensor_v1' (round-1 tensor kernel) or 'tensor_alu' (default engine, sort-2 on the ALU pipe)")
        sift_engines = {"tensor": 0, "tensor_v1": L.CFG_SIFT_TC_V1, "fp32": L.CFG_SIFT_EXACT_FP32}
        if sift_engine not in sift_engines:
            raise ValueError("sift_engine must be 'tensor' (bf16 tensor-core scoring of both directions with pruned scans + exact re-rank, default), "
                             "'tensor_v1' (round-1 tensor kernel) or 'fp32' (all-FP32 kernels)")
        self.sift_engine = "fp32" if sift_exact_fp32 else sift_engine
        flags = (L.CFG_SIFT_EXACT_FP32 if sift_exact_fp32 else 0) | sift_engines[sift_engine] | engines[orb_engine] | (0 if match_cache else L.CFG_MATCH_NO_CACHE) | \
                (L.CFG_MATCH_LEGACY if match_legacy else 0)
        self.orb_engine = orb_engine
        cfg = L.Config(device=device, max_images=0, match_buffer_entries=match_buffer_entries, flags=flags)
        h = ctypes.c_void_p()
        L.check(self._lib.eacham_gpu_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.device = device

    # -- lifetime ---------------------------------------------------------------------------------------
    def close(self):
        for name in ("_pin_res", "_pin_buf"):
            b = getattr(self, name, None)
            if b is not None:
                b.close()
                setattr(self, name, None)
        if getattr(self, "_h", None):
            self._lib.eacham_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference's Match(): one direction, one pair -------------------------------------------------
    def Match(self, descriptor1: np.ndarray, descriptor2: np.ndarray) -> Dict[int, int]:
        """{queryIdx -> trainIdx} of the ratio-passing nearest neighbours of descriptor1's rows in descriptor2."""
        kind = _kind_of(descriptor1)
        if _kind_of(descriptor2) != kind:
            raise TypeError("descriptor kinds differ")
        q, qp, qn, qs = _rows_ptr(descriptor1)
        t, tp, tn, ts = _rows_ptr(descriptor2)
        out = np.empty(max(qn, 1), dtype=L.MATCH_DTYPE)
        n = ctypes.c_size_t()
        L.check(self._lib.eacham_gpu_match(self._h, kind, qp, qn, qs, tp, tn, ts, self.ratio,
                                           out.ctypes.data_as(ctypes.c_void_p), out.shape[0], ctypes.byref(n)))
        out = out[:n.value]
        return dict(zip(out["query"].tolist(), out["train"].tolist()))

    def knnMatch(self, descriptor1: np.ndarray, descriptor2: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """k=2 neighbours as (trainIdx[N,2] int32, distance[N,2] float32) -- the content of the
        vector<vector<DMatch>> that FeatureMatcherFlann.cpp:17 receives."""
        kind = _kind_of(descriptor1)
        if _kind_of(descriptor2) != kind:
            raise TypeError("descriptor kinds differ")
        q, qp, qn, qs = _rows_ptr(descriptor1)
        t, tp, tn, ts = _rows_ptr(descriptor2)
        idx = np.full((qn, 2), -1, np.int32)
        dist = np.full((qn, 2), np.inf, np.float32)
        L.check(self._lib.eacham_gpu_knn2(self._h, kind, qp, qn, qs, tp, tn, ts, idx.ctypes.data_as(ctypes.c_void_p),
                                          dist.ctypes.data_as(ctypes.c_void_p)))
        return idx, dist

    # -- descriptor arena ---------------------------------------------------------------------------------
    def SetDescriptors(self, image_id: int, descriptors: np.ndarray) -> None:
        kind = _kind_of(descriptors)
        d, p, n, s = _rows_ptr(descriptors)
        L.check(self._lib.eacham_gpu_set_descriptors(self._h, image_id, kind, p, n, s))

    def Reserve(self, image_id: int, kind: int, rows: int) -> None:
        L.check(self._lib.eacham_gpu_reserve(self._h, image_id, kind, rows))

    def Commit(self) -> None:
        L.check(self._lib.eacham_gpu_commit(self._h))

    def Clear(self) -> None:
        L.check(self._lib.eacham_gpu_clear(self._h))

    def Upload(self, descriptors: Sequence[np.ndarray]) -> None:
        """All images (ids 0..n-1) staged by one set_descriptors_batch call per descriptor kind, followed by one commit (one H2D copy)."""
        self.Clear()
        _upload_batched(self._lib.eacham_gpu_set_descriptors_batch, self._h, descriptors)
        self.Commit()

    def arena(self) -> Tuple[int, int]:
        ptr = ctypes.c_void_p(); nbytes = ctypes.c_size_t()
        L.check(self._lib.eacham_gpu_arena(self._h, ctypes.byref(ptr), ctypes.byref(nbytes)))
        return int(ptr.value or 0), int(nbytes.value)

    # -- the batched path -----------------------------------------------------------------------------------
    def _opts(self, emit_all: bool) -> L.MatchOpts:
        return L.MatchOpts(ratio=self.ratio, min_dir=self.min_dir, min_mutual=self.min_mutual,
                           cross_check=1 if self.cross_check else 0, emit_all=1 if emit_all else 0)

    @staticmethod
    def _pairs_array(pairs) -> np.ndarray:
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2))
        return arr

    def _pinned(self, which: str, n: int, dtype):
        """Reusable page-locked host buffer (eacham_gpu_host_alloc): D2H at full PCIe rate instead of through the driver's bounce buffers."""
        from .multi import PinnedBuffer
        cur = getattr(self, which, None)
        if cur is None or cur.n < n:
            if cur is not None:
                cur.close()
            cur = PinnedBuffer(self._lib, int(n * 1.25) + 1024, dtype)
            setattr(self, which, cur)
        return cur

    def MatchPairsRaw(self, pairs, emit_all: bool = False, buf: Optional[np.ndarray] = None):
        """C-ABI call with host buffers: returns (results record array [n_pairs], matches record array [used]). Unless `buf` is given
        both are views of pinned buffers owned by this object, valid until its next MatchPairsRaw / close."""
        arr = self._pairs_array(pairs)
        n = arr.shape[0]
        self._last_n = n
        opts = self._opts(emit_all)
        used = ctypes.c_size_t()
        res = self._pinned("_pin_res", max(n, 1), L.RESULT_DTYPE).array[:n]
        own = buf is None
        if own:
            buf = self._pinned("_pin_buf", max(n * 192, 1 << 16), L.MATCH_DTYPE).array
        rc = self._lib.eacham_gpu_match_pairs(self._h, arr.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(opts),
                                              res.ctypes.data_as(ctypes.c_void_p), buf.ctypes.data_as(ctypes.c_void_p),
                                              buf.shape[0], ctypes.byref(used))
        if rc == L.ERR_BUFFER_TOO_SMALL:
            buf = self._pinned("_pin_buf", used.value, L.MATCH_DTYPE).array if own else np.empty(used.value, dtype=L.MATCH_DTYPE)
            L.check(self._lib.eacham_gpu_fetch_results(self._h, res.ctypes.data_as(ctypes.c_void_p), n,
                                                       buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0], ctypes.byref(used)))
        else:
            L.check(rc)
        return res, buf[:used.value]

    def MatchPairs(self, pairs, emit_all: bool = False) -> List[PairMatches]:
        """For each unordered pair (first, second): both directions, ratio test, gates, mutual filter
        (main.cpp:84-147). Returns one PairMatches per input pair, in input order."""
        arr = self._pairs_array(pairs)
        res, buf = self.MatchPairsRaw(arr, emit_all=emit_all)
        out = []
        for k in range(arr.shape[0]):
            r = res[k]
            m = buf[int(r["offset"]): int(r["offset"]) + int(r["count"])]
            out.append(PairMatches(int(arr[k, 0]), int(arr[k, 1]), int(r["n12"]), int(r["n21"]), int(r["n_mutual"]),
                                   bool(r["flags"] & L.PAIR_GATED), bool(r["flags"] & L.PAIR_CONNECTED),
                                   np.stack([m["query"], m["train"]], axis=1).astype(np.uint32)))
        return out

    def MatchPairsDevice(self, pairs) -> int:
        """Device-resident variant: results stay in HBM (fetch with FetchResults). Returns the total match count."""
        arr = self._pairs_array(pairs)
        opts = self._opts(False)
        total = ctypes.c_size_t()
        L.check(self._lib.eacham_gpu_match_pairs_device(self._h, arr.ctypes.data_as(ctypes.c_void_p), arr.shape[0],
                                                        ctypes.byref(opts), ctypes.byref(total)))
        self._last_n = arr.shape[0]
        return int(total.value)

    def FetchResults(self, n_pairs: Optional[int] = None):
        n = self._last_n if n_pairs is None else n_pairs
        res = np.zeros(n, dtype=L.RESULT_DTYPE)
        used = ctypes.c_size_t()
        rc = self._lib.eacham_gpu_fetch_results(self._h, res.ctypes.data_as(ctypes.c_void_p), n, None, 0, ctypes.byref(used))
        L.check(rc, allow=(L.ERR_BUFFER_TOO_SMALL,))
        buf = np.empty(max(used.value, 1), dtype=L.MATCH_DTYPE)
        L.check(self._lib.eacham_gpu_fetch_results(self._h, res.ctypes.data_as(ctypes.c_void_p), n,
                                                   buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0], ctypes.byref(used)))
        return res, buf[:used.value]

    def device_results(self) -> Tuple[int, int, int, int]:
        """(results_ptr, matches_ptr, n_pairs, n_matches) of the last batch, device resident."""
        r = ctypes.c_void_p(); m = ctypes.c_void_p(); n = ctypes.c_size_t(); k = ctypes.c_size_t()
        L.check(self._lib.eacham_gpu_device_results(self._h, ctypes.byref(r), ctypes.byref(m), ctypes.byref(n), ctypes.byref(k)))
        return int(r.value or 0), int(m.value or 0), int(n.value), int(k.value)

    def timing(self) -> Dict[str, float]:
        t = L.Timing()
        L.check(self._lib.eacham_gpu_last_timing(self._h, ctypes.byref(t)))
        return dict(upload_ms=t.upload_ms, pairs_h2d_ms=t.pairs_h2d_ms, kernel_ms=t.kernel_ms, d2h_ms=t.d2h_ms,
                    kernel_launches=int(t.kernel_launches), prep_ms=t.prep_ms, exact_fallbacks=int(t.exact_fallbacks), verify_ms=t.verify_ms)

    # -- geometric verification of the last batch ------------------------------------------------
    def SetKeypoints(self, image_id: int, xy: np.ndarray) -> None:
        """xy[rows, 2] float32: keypoint r belongs to descriptor row r of the image."""
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        L.check(self._lib.eacham_gpu_set_keypoints(self._h, image_id, xy.ctypes.data_as(ctypes.c_void_p), xy.shape[0], 8))

    def VerifyPairs(self, model: int, hyps: np.ndarray, focal: float = 1.0, cx: float = 0.0, cy: float = 0.0, want_medians: bool = False,
                    want_mask: bool = False):
        """Scores hypotheses for every pair of the last MatchPairs* batch (eacham_gpu_verify_pairs). hyps: [n_hyp, 3, 3] shared by all
        pairs, or [n_pairs, n_hyp, 3, 3]. Returns (results record array, medians or None, mask or None)."""
        hyps = np.ascontiguousarray(hyps, dtype=np.float64)
        shared = hyps.ndim == 3
        n_hyp = hyps.shape[0] if shared else hyps.shape[1]
        n = self._last_n
        if not shared and hyps.shape[0] != n:
            raise ValueError("per-pair hypotheses must have one block per pair of the last batch")
        opts = L.VerifyOpts(focal=float(focal), cx=float(cx), cy=float(cy), n_hyp=int(n_hyp), shared=1 if shared else 0)
        res = np.zeros(n, dtype=L.VERIFY_DTYPE)
        med = np.zeros((n, n_hyp), np.float32) if want_medians else None
        _, _, _, n_matches = self.device_results()
        mask = np.zeros(max(n_matches, 1), np.uint8) if want_mask else None
        L.check(self._lib.eacham_gpu_verify_pairs(self._h, model, hyps.ctypes.data_as(ctypes.c_void_p), ctypes.byref(opts),
                                                  res.ctypes.data_as(ctypes.c_void_p),
                                                  med.ctypes.data_as(ctypes.c_void_p) if want_medians else None,
                                                  mask.ctypes.data_as(ctypes.c_void_p) if want_mask else None))
        return res, med, (mask[:n_matches] if want_mask else None)

    def DebugPairKnn2(self, first: int, second: int):
        """(idx12[n1,2], dist12[n1,2], idx21[n2,2], dist21[n2,2]): the kNN(k=2) the batched tensor-core SIFT path's ratio test saw
        for one uploaded pair (see eacham_gpu_debug_pair_knn2)."""
        k1 = ctypes.c_int(); n1 = ctypes.c_uint32(); k2 = ctypes.c_int(); n2 = ctypes.c_uint32()
        L.check(self._lib.eacham_gpu_image_info(self._h, first, ctypes.byref(k1), ctypes.byref(n1), None))
        L.check(self._lib.eacham_gpu_image_info(self._h, second, ctypes.byref(k2), ctypes.byref(n2), None))
        i12 = np.full((n1.value, 2), -1, np.int32); d12 = np.full((n1.value, 2), np.inf, np.float32)
        i21 = np.full((n2.value, 2), -1, np.int32); d21 = np.full((n2.value, 2), np.inf, np.float32)
        opts = self._opts(True)
        L.check(self._lib.eacham_gpu_debug_pair_knn2(self._h, first, second, ctypes.byref(opts), i12.ctypes.data_as(ctypes.c_void_p),
                                                     d12.ctypes.data_as(ctypes.c_void_p), i21.ctypes.data_as(ctypes.c_void_p),
                                                     d21.ctypes.data_as(ctypes.c_void_p)))
        return i12, d12, i21, d21

    def flush_l2(self, nbytes: int = 256 << 20) -> None:
        L.check(self._lib.eacham_gpu_flush_l2(self._h, nbytes))
