"""Runs ONE pair batch on a trace build (EACHAM_EXP & 8) so that the kernel prints its per-tile timeline for CTA 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eacham_b200
from eacham_b200 import synth
imgs = synth.orb_image_set(24, 4096, seed=2, pool=20000)
pairs = synth.exhaustive_pairs(24)
with eacham_b200.FeatureMatcherGpu(0.8) as m:
    m.Upload(imgs)
    m.MatchPairsDevice(pairs)
    print("kernel_ms", m.timing()["kernel_ms"], "pairs", len(pairs), flush=True)
