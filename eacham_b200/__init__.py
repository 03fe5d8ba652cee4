"""eacham_b200 -- B200-native exhaustive descriptor matching for the eacham SfM pipeline.

Only the hot path: kNN(k=2) + Lowe ratio + mutual cross-check over image pairs
(/root/reference/apps/sfm/main.cpp:81-152 calling /root/reference/modules/base/features/FeatureMatcherFlann.cpp:14-30),
as hand-written sm_100a CUDA behind a C ABI (include/eacham_gpu.h). No CPU fallback.
"""
from .matcher import FeatureMatcherGpu, PairMatches  # noqa: F401
from .multi import MultiGpuMatcher  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["FeatureMatcherGpu", "MultiGpuMatcher", "PairMatches", "synth"]
