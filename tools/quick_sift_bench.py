"""Quick SIFT kernel timing on one GPU (kernel-only pairs/s on a small exhaustive set)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eacham_b200
from eacham_b200 import synth
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 40
imgs = synth.sift_image_set_pooled(n_img, 8192, seed=3)
pairs = synth.exhaustive_pairs(n_img)
engine = sys.argv[2] if len(sys.argv) > 2 else "tensor"
with eacham_b200.FeatureMatcherGpu(0.8, sift_engine=engine) as m:
    m.Upload(imgs)
    m.MatchPairsDevice(pairs)
    ms = []
    for _ in range(3):
        m.flush_l2(256 << 20)
        m.MatchPairsDevice(pairs)
        ms.append(m.timing()["kernel_ms"])
    print({"engine": engine, "pairs": len(pairs), "kernel_ms": [round(x, 2) for x in ms], "pairs_per_s": round(len(pairs) / (min(ms) * 1e-3)), "exact_fallbacks": m.timing()["exact_fallbacks"]})
