"""ctypes binding of libeacham_gpu.so (include/eacham_gpu.h). Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EACHAM_GPU_LIB") or os.path.join(_HERE, "libeacham_gpu.so")      # EACHAM_GPU_LIB: another BUILD of this library (kernel experiments)

OK = 0
ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY = -1, -2, -3, -4
ERR_BUFFER_TOO_SMALL, ERR_NOT_COMMITTED, ERR_TOO_LARGE, ERR_KIND_MISMATCH, ERR_NCCL = -5, -6, -7, -8, -9
KIND_ORB256, KIND_F32X128 = 0, 1
NONE = 0xFFFFFFFF
PAIR_GATED, PAIR_CONNECTED = 1, 2
CFG_SIFT_EXACT_FP32, CFG_ORB_POPC, CFG_ORB_TC_V1, CFG_ORB_TC_ALU_SORT, CFG_MULTI_PARALLEL_H2D = 1, 2, 4, 8, 16
CFG_MATCH_NO_CACHE, CFG_MATCH_LEGACY, CFG_SIFT_TC_V1 = 32, 64, 128
ABI_VERSION = 2


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("max_images", ctypes.c_uint32),
                ("match_buffer_entries", ctypes.c_uint64), ("flags", ctypes.c_uint32)]


class MatchOpts(ctypes.Structure):
    _fields_ = [("ratio", ctypes.c_double), ("min_dir", ctypes.c_uint32), ("min_mutual", ctypes.c_uint32),
                ("cross_check", ctypes.c_uint32), ("emit_all", ctypes.c_uint32)]


class Timing(ctypes.Structure):
    _fields_ = [("upload_ms", ctypes.c_float), ("pairs_h2d_ms", ctypes.c_float), ("kernel_ms", ctypes.c_float),
                ("d2h_ms", ctypes.c_float), ("kernel_launches", ctypes.c_uint32), ("prep_ms", ctypes.c_float),
                ("exact_fallbacks", ctypes.c_uint32), ("verify_ms", ctypes.c_float)]


class VerifyOpts(ctypes.Structure):
    _fields_ = [("focal", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double), ("n_hyp", ctypes.c_uint32), ("shared", ctypes.c_uint32)]


MODEL_ESSENTIAL, MODEL_HOMOGRAPHY = 0, 1
VERIFY_DTYPE = [("best", "<u4"), ("n_inliers", "<u4"), ("median", "<f4"), ("sigma", "<f4")]


class MultiTiming(ctypes.Structure):
    _fields_ = [("upload_ms", ctypes.c_float), ("broadcast_ms", ctypes.c_float), ("match_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float),
                ("kernel_ms_max", ctypes.c_float), ("prep_ms_max", ctypes.c_float), ("kernel_launches", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


# numpy-compatible record layouts of eacham_match_t / eacham_pair_t / eacham_pair_result_t
MATCH_DTYPE = [("query", "<u4"), ("train", "<u4")]
PAIR_DTYPE = [("first", "<u4"), ("second", "<u4")]
RESULT_DTYPE = [("n12", "<u4"), ("n21", "<u4"), ("n_mutual", "<u4"), ("flags", "<u4"), ("offset", "<u8"), ("count", "<u8")]

# every symbol include/eacham_gpu.h declares (tests check the header against this list and the .so)
SYMBOLS = [
    "eacham_gpu_abi_version", "eacham_gpu_device_count", "eacham_gpu_last_error", "eacham_gpu_default_opts",
    "eacham_gpu_create", "eacham_gpu_destroy", "eacham_gpu_set_descriptors", "eacham_gpu_set_descriptors_batch", "eacham_gpu_reserve", "eacham_gpu_commit",
    "eacham_gpu_clear", "eacham_gpu_arena", "eacham_gpu_image_info", "eacham_gpu_match", "eacham_gpu_knn2",
    "eacham_gpu_match_pairs", "eacham_gpu_match_pairs_device", "eacham_gpu_fetch_results", "eacham_gpu_device_results", "eacham_gpu_last_timing",
    "eacham_gpu_flush_l2",
    "eacham_gpu_create_multi", "eacham_gpu_destroy_multi", "eacham_gpu_multi_device_count", "eacham_gpu_multi_set_descriptors", "eacham_gpu_multi_set_descriptors_batch",
    "eacham_gpu_multi_clear", "eacham_gpu_multi_commit", "eacham_gpu_multi_match_pairs", "eacham_gpu_multi_last_timing",
    "eacham_gpu_host_alloc", "eacham_gpu_host_free", "eacham_gpu_debug_pair_knn2", "eacham_gpu_set_keypoints", "eacham_gpu_verify_pairs",
]

_lib = None


class EachamGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libeacham_gpu error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Loads the CUDA library. Raises if it has not been built -- the product path never falls back to CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m eacham_b200.build` "
                           "(nvcc, sm_100a). eacham_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u32, i32, dbl = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_double
    P = ctypes.POINTER
    lib.eacham_gpu_abi_version.restype = i32
    lib.eacham_gpu_device_count.restype = i32
    lib.eacham_gpu_last_error.restype = ctypes.c_char_p
    lib.eacham_gpu_default_opts.argtypes = [P(MatchOpts)]
    lib.eacham_gpu_default_opts.restype = None
    lib.eacham_gpu_create.argtypes = [P(Config), P(vp)]
    lib.eacham_gpu_destroy.argtypes = [vp]
    lib.eacham_gpu_destroy.restype = None
    lib.eacham_gpu_set_descriptors.argtypes = [vp, u32, i32, vp, u32, sz]
    lib.eacham_gpu_set_descriptors_batch.argtypes = [vp, u32, u32, i32, vp, vp, vp]
    lib.eacham_gpu_reserve.argtypes = [vp, u32, i32, u32]
    lib.eacham_gpu_commit.argtypes = [vp]
    lib.eacham_gpu_clear.argtypes = [vp]
    lib.eacham_gpu_arena.argtypes = [vp, P(vp), P(sz)]
    lib.eacham_gpu_image_info.argtypes = [vp, u32, P(i32), P(u32), P(sz)]
    lib.eacham_gpu_match.argtypes = [vp, i32, vp, u32, sz, vp, u32, sz, dbl, vp, sz, P(sz)]
    lib.eacham_gpu_knn2.argtypes = [vp, i32, vp, u32, sz, vp, u32, sz, vp, vp]
    lib.eacham_gpu_match_pairs.argtypes = [vp, vp, sz, P(MatchOpts), vp, vp, sz, P(sz)]
    lib.eacham_gpu_match_pairs_device.argtypes = [vp, vp, sz, P(MatchOpts), P(sz)]
    lib.eacham_gpu_fetch_results.argtypes = [vp, vp, sz, vp, sz, P(sz)]
    lib.eacham_gpu_device_results.argtypes = [vp, P(vp), P(vp), P(sz), P(sz)]
    lib.eacham_gpu_last_timing.argtypes = [vp, P(Timing)]
    lib.eacham_gpu_flush_l2.argtypes = [vp, sz]
    lib.eacham_gpu_debug_pair_knn2.argtypes = [vp, u32, u32, P(MatchOpts), vp, vp, vp, vp]
    lib.eacham_gpu_set_keypoints.argtypes = [vp, u32, vp, u32, sz]
    lib.eacham_gpu_verify_pairs.argtypes = [vp, i32, vp, P(VerifyOpts), vp, vp, vp]
    lib.eacham_gpu_create_multi.argtypes = [P(ctypes.c_int32), u32, P(Config), P(vp)]
    lib.eacham_gpu_destroy_multi.argtypes = [vp]
    lib.eacham_gpu_destroy_multi.restype = None
    lib.eacham_gpu_multi_device_count.argtypes = [vp]
    lib.eacham_gpu_multi_device_count.restype = u32
    lib.eacham_gpu_multi_set_descriptors.argtypes = [vp, u32, i32, vp, u32, sz]
    lib.eacham_gpu_multi_set_descriptors_batch.argtypes = [vp, u32, u32, i32, vp, vp, vp]
    lib.eacham_gpu_multi_clear.argtypes = [vp]
    lib.eacham_gpu_multi_commit.argtypes = [vp]
    lib.eacham_gpu_multi_match_pairs.argtypes = [vp, vp, sz, P(MatchOpts), vp, vp, sz, P(sz)]
    lib.eacham_gpu_multi_last_timing.argtypes = [vp, P(MultiTiming)]
    lib.eacham_gpu_host_alloc.argtypes = [sz]
    lib.eacham_gpu_host_alloc.restype = vp
    lib.eacham_gpu_host_free.argtypes = [vp]
    lib.eacham_gpu_host_free.restype = None
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ("eacham_gpu_abi_version", "eacham_gpu_device_count", "eacham_gpu_multi_device_count"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc: int, allow=()) -> int:
    if rc != OK and rc not in allow:
        raise EachamGpuError(rc, load().eacham_gpu_last_error().decode("utf-8", "replace"))
    return rc
