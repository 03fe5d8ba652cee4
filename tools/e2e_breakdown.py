"""Where one end-to-end step of the batched route spends its wall time (config 2 shape by default)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eacham_b200
from eacham_b200 import synth
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n_desc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
imgs = synth.orb_image_set(n_img, n_desc, seed=2, pool=20000)
pairs = np.asarray(synth.exhaustive_pairs(n_img), np.uint32)
with eacham_b200.FeatureMatcherGpu(0.8) as m:
    for it in range(4):
        t0 = time.perf_counter()
        m.Upload(imgs)
        t1 = time.perf_counter()
        res, buf = m.MatchPairsRaw(pairs)
        t2 = time.perf_counter()
        tm = m.timing()
        print({"it": it, "upload_wall_ms": round((t1 - t0) * 1e3, 2), "match_wall_ms": round((t2 - t1) * 1e3, 2),
               "lib": {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")}, "matches": len(buf)}, flush=True)
