"""Several B200s of one box in ONE process, behind the C ABI (``eacham_gpu_multi_*`` in include/eacham_gpu.h).

This is the route a single-process C++ caller (the reference's ``apps/sfm/main.cpp``) takes to more than one GPU: the
library itself uploads the descriptors once, broadcasts the arena with NCCL over NVLink (``ncclCommInitAll``), shards the pair
list of /root/reference/apps/sfm/main.cpp:84-92 across the devices and lets every device copy its own shard of the results into
its slice of the caller's (pinned) buffers. Python only passes pointers.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .matcher import PairMatches, _kind_of, _rows_ptr, _upload_batched


class PinnedBuffer:
    """Page-locked host memory from ``eacham_gpu_host_alloc`` viewed as a numpy record array."""

    def __init__(self, lib, n: int, dtype):
        self._lib = lib
        self.dtype = np.dtype(dtype)
        self.n = max(int(n), 1)
        self.ptr = lib.eacham_gpu_host_alloc(self.n * self.dtype.itemsize)
        if not self.ptr:
            raise MemoryError(lib.eacham_gpu_last_error().decode("utf-8", "replace"))
        raw = (ctypes.c_uint8 * (self.n * self.dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=self.n)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._lib.eacham_gpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuMatcher:
    """``FeatureMatcherGpu``'s batched interface over ``devices`` (default: every visible GPU)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, *, ratio: float = 0.8, min_dir: int = 30, min_mutual: int = 30,
                 cross_check: bool = True, orb_engine: str = "tensor", sift_exact_fp32: bool = False, parallel_h2d: bool = False):
        self._lib = L.load()
        if devices is None:
            devices = list(range(self._lib.eacham_gpu_device_count()))
        self.devices = [int(d) for d in devices]
        engines = {"tensor": 0, "popc": L.CFG_ORB_POPC, "tensor_v1": L.CFG_ORB_TC_V1, "tensor_alu": L.CFG_ORB_TC_ALU_SORT}
        flags = engines[orb_engine] | (L.CFG_SIFT_EXACT_FP32 if sift_exact_fp32 else 0) | (L.CFG_MULTI_PARALLEL_H2D if parallel_h2d else 0)
        cfg = L.Config(device=0, max_images=0, match_buffer_entries=0, flags=flags)
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        h = ctypes.c_void_p()
        L.check(self._lib.eacham_gpu_create_multi(arr, len(self.devices), ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.ratio, self.min_dir, self.min_mutual, self.cross_check = float(ratio), int(min_dir), int(min_mutual), bool(cross_check)
        self._res: Optional[PinnedBuffer] = None
        self._buf: Optional[PinnedBuffer] = None

    def close(self):
        for b in (self._res, self._buf):
            if b is not None:
                b.close()
        self._res = self._buf = None
        if getattr(self, "_h", None):
            self._lib.eacham_gpu_destroy_multi(self._h)
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

    def Upload(self, descriptors: Sequence[np.ndarray]) -> None:
        """clear + set_descriptors for ids 0..n-1 + commit: one H2D copy to devices[0], one NCCL broadcast."""
        L.check(self._lib.eacham_gpu_multi_clear(self._h))
        _upload_batched(self._lib.eacham_gpu_multi_set_descriptors_batch, self._h, descriptors)
        L.check(self._lib.eacham_gpu_multi_commit(self._h))

    def _pinned(self, which: str, n: int, dtype) -> PinnedBuffer:
        cur = getattr(self, which)
        if cur is None or cur.n < n:
            if cur is not None:
                cur.close()
            cur = PinnedBuffer(self._lib, int(n * 1.25) + 1024, dtype)
            setattr(self, which, cur)
        return cur

    def MatchPairsRaw(self, pairs, emit_all: bool = False):
        """(results[n_pairs], matches[used]) in input order; both are views of pinned buffers owned by this object (valid until the
        next call)."""
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2))
        n = arr.shape[0]
        opts = L.MatchOpts(ratio=self.ratio, min_dir=self.min_dir, min_mutual=self.min_mutual, cross_check=1 if self.cross_check else 0,
                           emit_all=1 if emit_all else 0)
        res = self._pinned("_res", n, L.RESULT_DTYPE)
        buf = self._pinned("_buf", max(n * 192, 1 << 16), L.MATCH_DTYPE)
        used = ctypes.c_size_t()
        rc = self._lib.eacham_gpu_multi_match_pairs(self._h, arr.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(opts), res.ptr, buf.ptr,
                                                    buf.n, ctypes.byref(used))
        if rc == L.ERR_BUFFER_TOO_SMALL:                     # deterministic need: grow once and run again
            buf = self._pinned("_buf", used.value, L.MATCH_DTYPE)
            rc = self._lib.eacham_gpu_multi_match_pairs(self._h, arr.ctypes.data_as(ctypes.c_void_p), n, ctypes.byref(opts), res.ptr,
                                                        buf.ptr, buf.n, ctypes.byref(used))
        L.check(rc)
        return res.array[:n], buf.array[:used.value]

    def MatchPairs(self, pairs, emit_all: bool = False) -> List[PairMatches]:
        arr = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)
        res, buf = self.MatchPairsRaw(arr, emit_all=emit_all)
        out = []
        for k in range(arr.shape[0]):
            r = res[k]
            m = buf[int(r["offset"]): int(r["offset"]) + int(r["count"])]
            out.append(PairMatches(int(arr[k, 0]), int(arr[k, 1]), int(r["n12"]), int(r["n21"]), int(r["n_mutual"]),
                                   bool(r["flags"] & L.PAIR_GATED), bool(r["flags"] & L.PAIR_CONNECTED),
                                   np.stack([m["query"], m["train"]], axis=1).astype(np.uint32)))
        return out

    def timing(self) -> Dict[str, float]:
        t = L.MultiTiming()
        L.check(self._lib.eacham_gpu_multi_last_timing(self._h, ctypes.byref(t)))
        return dict(upload_ms=t.upload_ms, broadcast_ms=t.broadcast_ms, match_ms=t.match_ms, d2h_ms=t.d2h_ms,
                    kernel_ms_max=t.kernel_ms_max, prep_ms_max=t.prep_ms_max, kernel_launches=int(t.kernel_launches))
