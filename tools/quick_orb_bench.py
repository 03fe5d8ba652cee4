"""Quick A/B of the ORB engines on one GPU: kernel-only pairs/s on a small exhaustive set + agreement of the outputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eacham_b200
from eacham_b200 import synth

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 80
n_desc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
engines = sys.argv[3].split(",") if len(sys.argv) > 3 else ["tensor", "tensor_alu", "tensor_v1", "popc"]
imgs = synth.orb_image_set(n_img, n_desc, seed=2, pool=20000)
pairs = synth.exhaustive_pairs(n_img)
ref = None
for eng in engines:
    with eacham_b200.FeatureMatcherGpu(0.8, orb_engine=eng) as m:
        m.Upload(imgs)
        m.MatchPairsDevice(pairs)
        prep = m.timing()["prep_ms"]
        ms = []
        for _ in range(3):
            m.flush_l2(256 << 20)
            m.MatchPairsDevice(pairs)
            ms.append(m.timing()["kernel_ms"])
        res, buf = m.FetchResults()
        key = (res["n12"].tolist(), res["n21"].tolist(), res["n_mutual"].tolist(), res["flags"].tolist())
        lists = [buf[int(r["offset"]): int(r["offset"] + r["count"])].tobytes() for r in res]
        same = None if ref is None else (key == ref[0] and lists == ref[1])
        if ref is None:
            ref = (key, lists)
        print({"engine": eng, "pairs": len(pairs), "kernel_ms": [round(x, 3) for x in ms], "pairs_per_s": round(len(pairs) / (min(ms) * 1e-3)),
               "prep_ms": round(prep, 3), "matches": int(res["count"].sum()), "same_as_first": same}, flush=True)
