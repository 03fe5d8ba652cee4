"""Generates the golden fixtures under tests/golden/ with OpenCV's exact matcher.

Run in the build container (cv2 4.13.0; the reference pins opencv/4.5.5, /root/reference/conanfile.txt:3 --
cv::batchDistance's K-nearest insertion is unchanged between the two):

    python tests/golden/make_golden.py

For every case of tests/cases.py it stores the inputs and what
``cv2.BFMatcher(norm).knnMatch(q, t, k=2)`` returned in both directions, plus the ratio-filtered maps
(/root/reference/modules/base/features/FeatureMatcherFlann.cpp:21-27) and the pair result
(/root/reference/apps/sfm/main.cpp:111-146) computed from those OpenCV outputs by the transcribed Python logic.
The reference itself ships no golden vectors; these are the pins.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cv2  # noqa: E402
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402


def record(name, d1, d2, out):
    i12, s12 = O.cv2_knn2(d1, d2)
    i21, s21 = O.cv2_knn2(d2, d1)
    m12 = O.py_ratio_filter(i12, s12, 0.8)
    m21 = O.py_ratio_filter(i21, s21, 0.8)
    pr = O.py_pair_logic(m12, m21, 30, 30)
    f12 = np.full(d1.shape[0], 0xFFFFFFFF, np.uint32)
    f21 = np.full(d2.shape[0], 0xFFFFFFFF, np.uint32)
    for k, v in m12.items():
        f12[k] = v
    for k, v in m21.items():
        f21[k] = v
    out[name + "/d1"] = d1; out[name + "/d2"] = d2
    out[name + "/idx12"] = i12; out[name + "/dist12"] = s12
    out[name + "/idx21"] = i21; out[name + "/dist21"] = s21
    out[name + "/m12"] = f12; out[name + "/m21"] = f21
    out[name + "/pair"] = np.array([pr["n12"], pr["n21"], pr["n_mutual"], int(pr["gated"]), int(pr["connected"])], np.int64)
    out[name + "/matches"] = pr["matches"]
    print(f"{name:28s} {d1.shape} x {d2.shape}  n12={pr['n12']} n21={pr['n21']} mutual={pr['n_mutual']} "
          f"gated={pr['gated']} connected={pr['connected']}")
    return pr


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    orb = {}
    prs = {name: record(name, d1, d2, orb) for name, (d1, d2) in cases.orb_cases().items()}
    # the constructions must hit the edges they were built for
    assert prs["gate_dir_29"]["gated"] and prs["gate_dir_29"]["n12"] == 29 and prs["gate_dir_29"]["n21"] == 29
    p = prs["gate_dir_30_mutual_30"]; assert not p["gated"] and p["n_mutual"] == 30 and not p["connected"]
    p = prs["gate_mutual_31"]; assert p["n_mutual"] == 31 and p["connected"]
    p = prs["gate_one_direction"]; assert p["n12"] >= 30 and p["n21"] < 30 and p["gated"]
    p = prs["shared_partner"]; assert p["n12"] == 70 and p["n21"] == 35 and p["n_mutual"] == 35 and p["connected"]
    assert orb["ties_best/idx12"][0].tolist() == [2, 5] and orb["ties_best/dist12"][0].tolist() == [0.0, 0.0]
    assert orb["ties_second/idx12"][1].tolist() == [6, 1]
    NONE = 0xFFFFFFFF
    for key, want in (("ratio_4_5", NONE), ("ratio_100_125", NONE), ("ratio_8_10", NONE), ("ratio_3_4", 1), ("ratio_99_125", 1),
                      ("ratio_101_125", NONE), ("ratio_0_5", 1), ("ratio_0_0", NONE), ("ratio_256_256", NONE), ("ratio_200_256", 1)):
        assert orb[key + "/m12"][0] == want, (key, orb[key + "/m12"], orb[key + "/dist12"])
        d0, d1 = (int(x) for x in key.split("_")[1:])
        assert sorted(orb[key + "/dist12"][0].tolist()) == sorted([float(d0), float(d1)]), key
    np.savez_compressed(os.path.join(here, "orb_cases.npz"), **orb)
    sift = {}
    for name, (d1, d2) in cases.sift_cases().items():
        record(name, d1, d2, sift)
    np.savez_compressed(os.path.join(here, "sift_cases.npz"), **sift)
    # a mid-size random planted ORB image set (3 images x 1024) for pair-level parity
    from eacham_b200 import synth
    imgs = synth.orb_image_set(3, 1024, seed=1, pool=4000)
    mid = {}
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        record(f"set_{i}_{j}", imgs[i], imgs[j], mid)
    np.savez_compressed(os.path.join(here, "orb_set_3x1024.npz"), **mid)
    with open(os.path.join(here, "VERSIONS.txt"), "w") as f:
        f.write(f"cv2 {cv2.__version__}\nnumpy {np.__version__}\nreference pins opencv/4.5.5 (conanfile.txt:3)\n")


if __name__ == "__main__":
    main()
