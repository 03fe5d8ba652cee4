"""Runs a few SIFT pairs on a trace build (-DEACHAM_EXP=16, EACHAM_GPU_LIB=...) so that the kernel prints its per-tile timeline for CTA 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eacham_b200
from eacham_b200 import synth
imgs = synth.sift_image_set_pooled(8, 8192, seed=3)
pairs = synth.exhaustive_pairs(8)
with eacham_b200.FeatureMatcherGpu(0.8) as m:
    m.Upload(imgs)
    m.MatchPairsDevice(pairs)
    print("kernel_ms", m.timing()["kernel_ms"], "pairs", len(pairs), flush=True)
