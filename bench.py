#!/usr/bin/env python
"""bench.py -- image pairs matched / s for exhaustive 4k-ORB matching (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload orb4k|orb4k5|orb2k|sift8k|kitti2k]

A "step" is one full pass of the hot path over the workload's pair list: for every unordered image pair both kNN(k=2)
directions, ratio 0.8, gates and the mutual filter (/root/reference/apps/sfm/main.cpp:84-147).

N = 1: BASELINE.json configs[1] -- 500 synthetic images x 4096 ORB descriptors, all 124,750 pairs -- is the main line; the other
single-GPU configs (1: 100 x 2k ORB, 3: 500 x 8k SIFT, 4: KITTI-shaped window) run as short sub-records under `workloads`, each with
its own value / e2e / roofline / cpu_baseline / parity count.
N > 1 (torchrun, one rank per GPU): BASELINE.json configs[4] -- 2,000 images x 4096 ORB, 1,999,000 pairs -- STRONG scaling: the fixed
pair list is sharded rank::N after ONE NCCL broadcast of the descriptor arena; there is no data-path collective.

value    = pairs / s with descriptors resident in HBM: CUDA events around the matching kernel on the library's stream, summed over
           the K steps, max over ranks.
e2e      = the same through the C-ABI calls a user makes, host buffers in, host buffers out, wall clock. N = 1:
           eacham_gpu_set_descriptors + commit + match_pairs. N > 1: the single-process multi-device route a C++ caller has
           (eacham_gpu_multi_*: one H2D, one in-process ncclBroadcast, sharded kernels, every GPU copying its own shard into one
           pinned host buffer), driven by rank 0 over all N GPUs while the other ranks wait on a CPU (gloo) barrier.
roofline = the dominant kernel against the tensor pipe: algorithmic 2 * 256 * N * M FLOP per ORB pair (the distance matrix once, one
           e4m3 element per bit) over the kernel time, against the UTCQMMA rate measured on this pool's B200
           (tools/tc_peak_microbench.cu -> profiles/r02_tc_peak.jsonl). SURVEY.md 8(d)'s POPC figure is kept as `popc_yardstick`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, descriptors per image, images, BASELINE.json config index)
    "orb4k": ("orb", 4096, 500, 1),
    "orb4k5": ("orb", 4096, 2000, 4),
    "orb2k": ("orb", 2048, 100, 0),
    "sift8k": ("sift", 8192, 500, 2),
    # KITTI-shaped sequence, each frame against the next 20
    "kitti2k": ("orb", 2048, 4541, 3),
}
WINDOW = {"kitti2k": 20}
METRIC = "image pairs matched/sec (exhaustive, 4k ORB)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        p["_source"] = "measured"
    else:
        p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0, "_source": "fallback"}
    return p


def fp8_peak():
    """Highest UTCQMMA (e4m3) rate measured by tools/tc_peak_microbench.cu on this pool's B200."""
    best, src = None, None
    path = os.path.join(ROOT, "profiles", "r02_tc_peak.jsonl")
    if os.path.exists(path):
        for line in open(path):
            line = line.strip()
            if line.startswith("{") and '"mma"' in line:
                r = json.loads(line)
                if "e4m3" in r["mma"] and (best is None or r["tflops"] > best):
                    best, src = r["tflops"], f"{r['mma']} M=128 N={r['N']}: {r['cycles_per_mma']} cycles per MMA"
    if best is None:
        return 2 * peaks()["bf16_tflops"], "2 x bf16 burst peak (no UTCQMMA measurement on file)"
    return best, "measured, profiles/r02_tc_peak.jsonl (" + src + ")"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_images(name: str, n_images: int):
    from eacham_b200 import synth
    kind, n_desc, _, _ = WORKLOADS[name]
    if name == "kitti2k":
        # landmarks drawn from a sliding window of the pool: neighbouring frames overlap, distant ones do not
        return synth.orb_image_set(n_images, n_desc, seed=2, pool=max(20000, 40 * n_images), window=6000)
    if kind == "orb":
        return synth.orb_image_set(n_images, n_desc, seed=2, pool=20000)
    return synth.sift_image_set_pooled(n_images, n_desc, seed=3)


def make_pairs(name: str, n_images: int):
    from eacham_b200 import synth
    win = WINDOW.get(name, 0)
    return synth.window_pairs(n_images, win) if win else synth.exhaustive_pairs(n_images)


def cpu_match_pair(a, b):
    """The reference algorithm on the CPU: OpenCV exact matcher both directions + ratio + gates + mutual."""
    from oracle import oracle as O
    if O.have_cv2():
        return O.cv2_match_pair_fast(a, b), "reference-dependency"
    return O.c_match_pair(a, b), "port"


def cpu_threads():
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        return int(cv2.getNumThreads())
    except Exception:
        return 1


def time_cpu(images, pairs, budget_s: float, max_pairs: int, min_pairs: int = 4):
    """Times the CPU path on a bounded random sample of the workload's pairs. Returns (pairs/s, n_sampled, {pair index: result})."""
    rng = np.random.default_rng(123)
    order = rng.permutation(len(pairs))[:max_pairs]
    t0 = time.perf_counter()
    done, results = 0, {}
    for k in order:
        i, j = int(pairs[k][0]), int(pairs[k][1])
        results[int(k)] = cpu_match_pair(images[i], images[j])[0]
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= min_pairs:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, results


def parity_mismatches(cpu_res, res_all, buf_all):
    """GPU results (record arrays in input order) against the CPU results of the sampled pairs: bit for bit."""
    mism = 0
    for k, want in cpu_res.items():
        r = res_all[k]
        got = buf_all[int(r["offset"]): int(r["offset"]) + int(r["count"])]
        ok = (int(r["n12"]), int(r["n21"]), int(r["n_mutual"])) == (want["n12"], want["n21"], want["n_mutual"])
        if ok and want["connected"]:
            ok = np.array_equal(np.stack([got["query"], got["train"]], 1), want["matches"])
        mism += (not ok)
    return mism


def workload_config(name, n_images, n_pairs, gpus):
    kind, n_desc, _, cfg_idx = WORKLOADS[name]
    win = WINDOW.get(name, 0)
    return {"workload": f"BASELINE.json configs[{cfg_idx}]: synthetic {n_images} images x {n_desc} {'ORB-256bit' if kind == 'orb' else 'SIFT-128 f32'} descriptors, "
                        + (f"sliding window (each frame vs next {win}) = {n_pairs} unordered pairs" if win else f"exhaustive {n_pairs} unordered pairs")
                        + ", ratio 0.8 + cross-check, gates 30/30",
            "images": n_images, "descriptors_per_image": n_desc, "pairs": n_pairs, "pairs_per_gpu": n_pairs // max(gpus, 1),
            "parallelism": f"fixed pair list sharded over {gpus} GPUs in whole 16x16 image blocks (strong scaling), arena replicated by one NCCL broadcast" if gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB device memset outside the event-timed region); per-step CUDA events summed",
            "orb_engine": "tensor (tcgen05 kind::f8f6f4, F16 accumulators, packed epilogue)" if kind == "orb" else None}


def roofline_record(kind, n_desc, pairs_done, kernel_s, engine, traffic_key):
    pk = peaks()
    tj_path = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
    if kind == "orb":
        flop_per_pair = 2.0 * 256 * n_desc * n_desc
        peak_tf, peak_src = fp8_peak()
        achieved = flop_per_pair * pairs_done / kernel_s / 1e12
        popc_peak = 148 * 16.0 * pk["sm_max_mhz"] * 1e6
        popc_ach = 8.0 * n_desc * n_desc * pairs_done / kernel_s
        rec = {"bound": "tensor", "tensor_kind": "fp8 e4m3 -> f16 (UTCQMMA)", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
               "frac": achieved / peak_tf, "traffic": tj.get(traffic_key), "traffic_note": tj.get("_note"), "peak_source": peak_src,
               "work_per_pair": flop_per_pair, "kernel": "orb_tc_match_pairs_kernel" if engine == "tensor" else engine,
               "note": "algorithmic work = the N x M distance matrix ONCE per unordered pair, one e4m3 element per descriptor bit: "
                       "2 * 256 * N * M FLOP; the reference evaluates the matrix twice",
               "popc_yardstick": {"achieved_tpopc32": popc_ach / 1e12, "peak_tpopc32": popc_peak / 1e12, "ratio": popc_ach / popc_peak,
                                  "note": "SURVEY.md 8(d): 8 POPC32 per 256-bit distance against 148 SMs x 16 POPC/clk x sm_max_mhz; a yardstick "
                                          "only -- this engine issues no POPC in its hot loop"},
               "hbm": {"achieved_gbs": 2 * n_desc * 32 * pairs_done / kernel_s / 1e9, "peak_gbs": pk["hbm_gbs"],
                       "note": "algorithmic bytes = both images of every pair; far from the HBM bound"}}
        if engine == "popc":
            rec.update({"bound": "tensor", "note": "XOR+POPC engine measured against the same algorithmic-FLOP yardstick; its own binding pipe is the ALU "
                                                   "(DESIGN.md section 4)"})
        return rec
    flop_per_pair = 2.0 * 128 * n_desc * n_desc
    peak_tf = pk["bf16_tflops_sustained"]
    achieved = flop_per_pair * pairs_done / kernel_s / 1e12
    return {"bound": "tensor", "tensor_kind": "bf16 -> f32 (UTCHMMA)", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf, "traffic": tj.get(traffic_key), "peak_source": f"bf16_tflops_sustained ({pk['_source']}, MEASURED_PEAKS.json)",
            "work_per_pair": flop_per_pair, "kernel": "sift_tc_match_pairs_kernel",
            "note": "algorithmic FLOP: the kernel issues 2x this on the tensor cores (D and D^T, DESIGN.md section 7)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (OpenCV exact matcher, all host threads) on the same
    workload; each step is a bounded sample of the workload's pairs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or ("orb4k" if args.gpus == 1 else "orb4k5")
    kind, n_desc, n_images, _ = WORKLOADS[name]
    n_images = args.images or n_images
    n_pairs_total = len(make_pairs(name, n_images))
    # generating all images of the workload is not needed to time a sample: draw the sampled pairs' images only
    n_gen = min(n_images, 24)
    images = make_images(name, n_gen)
    from eacham_b200 import synth
    pairs = synth.exhaustive_pairs(n_gen)
    threads = cpu_threads()
    per_step = 48 if kind == "orb" else 4
    rng = np.random.default_rng(7)
    _, how = cpu_match_pair(images[0][:64], images[1][:64])

    def step():
        for k in rng.choice(len(pairs), size=per_step, replace=False):
            cpu_match_pair(images[int(pairs[k][0])], images[int(pairs[k][1])])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f"{per_step} random pairs per step drawn from the first {n_gen} images of the workload "
              f"({n_desc} descriptors each); cv2.BFMatcher-core batchDistance K=2 both directions + ratio + gates + mutual")
    line = {
        "impl": "reference", "metric": METRIC if name in ("orb4k", "orb4k5") else f"image pairs matched/sec ({name})",
        "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "u8" if kind == "orb" else "f32", "data": "synthetic",
        "config": workload_config(name, n_images, n_pairs_total, args.gpus),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference" if how == "reference-dependency" else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class Env:
    """Rank / world plumbing: torch.distributed only when launched under torchrun."""

    def __init__(self, gpus: int):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != gpus and self.world > 1:
            raise SystemExit(f"--gpus {gpus} but WORLD_SIZE={self.world}")
        self.dist = None
        self.cpu_group = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local_rank}"))
            self.cpu_group = dist.new_group(backend="gloo")       # CPU-side waits: no kernel spins on an idle rank's GPU
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def cpu_barrier(self):
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def device_timed(env, m, my_pairs, steps, warmup, sample_clocks=False):
    """W untimed warm-up passes, then K timed ones: L2 flushed, barrier, CUDA events around the kernel on the library's stream."""
    for _ in range(warmup):
        m.MatchPairsDevice(my_pairs)
    sampler = ClockSampler(env.local_rank) if sample_clocks else None
    env.barrier()
    if sampler:
        sampler.start()
    kernel_ms, launches = [], 0
    t0 = time.perf_counter()
    for _ in range(steps):
        m.flush_l2(256 << 20)
        env.barrier()
        m.MatchPairsDevice(my_pairs)
        t = m.timing()
        kernel_ms.append(t["kernel_ms"]); launches += t["kernel_launches"]
    env.barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    total_ms = env.max_over_ranks(float(sum(kernel_ms)))
    return total_ms, launches, wall_s, clocks


def verify_record(m, images, pairs, res, buf, n_hyp=32, cpu_budget_s=6.0):
    """SURVEY.md 8(f) N1 on the batch that is still resident in HBM: every pair's matches scored under n_hyp essential-matrix and n_hyp
    homography hypotheses (LMedS median, sigma, inlier count), eacham_gpu_verify_pairs; CPU = the NumPy restatement of OpenCV's scoring."""
    from eacham_b200 import _lib as L
    from oracle import verify_oracle as V
    rng = np.random.default_rng(99)
    kps = [np.stack([rng.uniform(0, 800, d.shape[0]), rng.uniform(0, 800, d.shape[0])], 1).astype(np.float32) for d in images]
    for i, k in enumerate(kps):
        m.SetKeypoints(i, k)
    f, cx, cy = 700.0, 400.0, 400.0
    hyp_e = rng.normal(0, 1, (n_hyp, 3, 3))
    hyp_h = np.eye(3)[None] + rng.normal(0, 0.01, (n_hyp, 3, 3)); hyp_h[:, 2, 2] = 1.0
    out = {}
    n_pairs = len(pairs)
    for model, code, hyps in (("essential", L.MODEL_ESSENTIAL, hyp_e), ("homography", L.MODEL_HOMOGRAPHY, hyp_h)):
        m.VerifyPairs(code, hyps, f, cx, cy)                                  # warm-up
        ms, wall = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            vres, _, _ = m.VerifyPairs(code, hyps, f, cx, cy)
            wall.append(time.perf_counter() - t0)
            ms.append(m.timing()["verify_ms"])
        order = np.random.default_rng(5).permutation(n_pairs)[:4096]
        t0 = time.perf_counter()
        done = mism = 0
        for k in order:
            r = res[k]
            mm = buf[int(r["offset"]): int(r["offset"] + r["count"])]
            i, j = int(pairs[k][0]), int(pairs[k][1])
            want = V.verify_pair(model, hyps, kps[i][mm["query"]], kps[j][mm["train"]], f, cx, cy)
            mism += (int(vres[k]["best"]), int(vres[k]["n_inliers"]), float(vres[k]["median"])) != (want["best"], want["n_inliers"], float(want["median"]))
            done += 1
            if time.perf_counter() - t0 > cpu_budget_s / 2 and done >= 16:
                break
        cpu_s = time.perf_counter() - t0
        matches = int(res["count"].sum())
        out[model] = {"value": n_pairs / (min(ms) * 1e-3), "unit": "pairs/s", "hypotheses_per_pair": n_hyp, "kernel_ms": min(ms),
                      "e2e": {"value": n_pairs / min(wall), "unit": "pairs/s", "includes": "hypotheses H2D + kernel + per-pair results D2H (match lists and keypoints already resident)"},
                      "residuals_per_s": matches * n_hyp / (min(ms) * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": (matches * 24 + n_pairs * 56) / (min(ms) * 1e-3) / 1e9, "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                                   "frac": (matches * 24 + n_pairs * 56) / (min(ms) * 1e-3) / 1e9 / peaks()["hbm_gbs"], "traffic": None,
                                   "note": "algorithmic bytes = match entry (8 B) + two gathered keypoints (16 B) per match + per-pair records; the kernel is bound by "
                                           "its shared-memory bitonic sorts (one per hypothesis), not by HBM -- see DESIGN.md"},
                      "cpu_baseline": {"value": done / cpu_s, "unit": "pairs/s", "cores": 1, "kind": "port",
                                       "sample": f"{done} random pairs of the batch, NumPy restatement of OpenCV's computeError + LMedS (oracle/verify_oracle.py), 1 thread",
                                       "parity_checked_pairs": done, "parity_mismatches": int(mism)}}
    return out


def run_single_gpu_workload(env, name, steps, warmup, e2e_steps, cpu_budget_s, orb_engine="tensor", sample_clocks=False, images_override=0,
                            with_engines=False, with_match_api=False, with_verify=False):
    """One workload on one GPU: value, e2e, roofline, cpu_baseline with in-run parity. Returns the record (and keeps nothing)."""
    import eacham_b200
    kind, n_desc, n_images, _ = WORKLOADS[name]
    n_images = images_override or n_images
    t_gen = time.perf_counter()
    images = make_images(name, n_images)
    gen_s = time.perf_counter() - t_gen
    pairs = make_pairs(name, n_images)
    n_pairs = pairs.shape[0]
    m = eacham_b200.FeatureMatcherGpu(0.8, device=env.local_rank, orb_engine=orb_engine)
    m.Upload(images)
    upload_ms = m.timing()["upload_ms"]
    arena_bytes = m.arena()[1]
    total_ms, launches, wall_s, clocks = device_timed(env, m, pairs, steps, max(warmup, 3), sample_clocks)
    prep_ms = m.timing()["prep_ms"]
    value = n_pairs * steps / (total_ms * 1e-3)

    # end to end through the C ABI: host descriptors in, host results out (first pass untimed: buffers reach their size)
    e2e_s, res, buf = 0.0, None, None
    for s in range(e2e_steps + 1):
        t0 = time.perf_counter()
        m.Upload(images)
        res, buf = m.MatchPairsRaw(pairs)
        dt = time.perf_counter() - t0
        if s > 0:
            e2e_s += dt
    e2e = {"value": n_pairs * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(arena_bytes + pairs.nbytes),
           "d2h_bytes_per_step": int(res.nbytes + buf.nbytes), "steps": e2e_steps,
           "includes": "set_descriptors + commit (pinned staging, one H2D) + tensor-core operand prep + pair list H2D + kernel + D2H of results and matches"}

    threads = cpu_threads()
    v, n_s, cpu_res = time_cpu(images, pairs, cpu_budget_s, 2048)
    _, how = cpu_match_pair(images[0][:64], images[1][:64])
    mism = parity_mismatches(cpu_res, res, buf)
    one_thread = None
    try:                                                  # also a 1-thread figure (2 pairs)
        import cv2
        cv2.setNumThreads(1)
        t1 = time.perf_counter()
        for k in list(cpu_res)[:2]:
            cpu_match_pair(images[int(pairs[k][0])], images[int(pairs[k][1])])
        one_thread = 2 / (time.perf_counter() - t1)
        cv2.setNumThreads(threads)
    except Exception:
        pass
    cpu = {"value": v, "unit": "pairs/s", "cores": threads, "one_thread_value": one_thread, "kind": "reference" if how == "reference-dependency" else "port",
           "sample": f"{n_s} random pairs of the workload (seed 123), OpenCV {('cv2 ' + __import__('cv2').__version__) if how == 'reference-dependency' else 'absent: C port'} "
                     f"batchDistance K=2 both directions + ratio + gates + mutual, {threads} threads",
           "parity_checked_pairs": n_s, "parity_mismatches": mism}

    rec = {"value": value, "unit": "pairs/s", "steps": steps, "warmup": max(warmup, 3), "ms_per_step": total_ms / steps,
           "config": workload_config(name, n_images, n_pairs, 1), "e2e": e2e, "gpu_launches": int(launches),
           "roofline": roofline_record(kind, n_desc, n_pairs * steps, total_ms * 1e-3, orb_engine if kind == "orb" else "sift", name),
           "cpu_baseline": cpu, "upload_ms": upload_ms, "prep_ms": prep_ms, "arena_bytes": int(arena_bytes), "generate_s": gen_s,
           "matches_per_step": int(res["count"].sum()), "wall_s_timed_region": wall_s}
    if clocks is not None:
        rec["clocks"] = clocks

    if with_verify:
        rec["verify"] = verify_record(m, images, pairs, res, buf)

    if with_engines and kind == "orb":
        engines = {orb_engine: {"value": value, "unit": "pairs/s", "steps": steps, "default": True}}
        for other in ("tensor_alu", "tensor_v1", "popc"):
            with eacham_b200.FeatureMatcherGpu(0.8, device=env.local_rank, orb_engine=other) as m2:
                m2.Upload(images)
                m2.MatchPairsDevice(pairs)
                ms = []
                for _ in range(2):
                    m2.flush_l2(256 << 20)
                    m2.MatchPairsDevice(pairs)
                    ms.append(m2.timing()["kernel_ms"])
                r2, b2 = m2.FetchResults()
                same = bool(np.array_equal(r2["count"], res["count"]) and np.array_equal(r2["n_mutual"], res["n_mutual"]) and
                            int(r2["count"].sum()) == len(b2))
                engines[other] = {"value": n_pairs * 2 / (sum(ms) * 1e-3), "unit": "pairs/s", "steps": 2, "default": False, "same_counts_as_default": same}
        engines["note"] = ("tensor = orb_tc_match_pairs_kernel (tcgen05 kind::f8f6f4, F16 accumulators, packed epilogue, sort-2 on the FMA pipe); "
                           "tensor_alu = same with sort-2 on the ALU pipe; tensor_v1 = round-1 kernel (F32 accumulators, 32-bit keys); "
                           "popc = orb_match_pairs_kernel (XOR + carry-save POPC). Bit-identical outputs (tests/test_gpu_orb*.py)")
        rec["engines"] = engines

    if kind == "sift":
        # A/B: the round-1 SIFT kernel (one MMA pass, REDUX column path, one query re-ranked at a time) on the same inputs
        with eacham_b200.FeatureMatcherGpu(0.8, device=env.local_rank, sift_engine="tensor_v1") as m2:
            m2.Upload(images)
            m2.MatchPairsDevice(pairs)
            m2.flush_l2(256 << 20)
            m2.MatchPairsDevice(pairs)
            ms1 = m2.timing()["kernel_ms"]
            r2, b2 = m2.FetchResults()
            same = bool(np.array_equal(r2["count"], res["count"]) and np.array_equal(r2["n_mutual"], res["n_mutual"]) and int(r2["count"].sum()) == len(b2))
        rec["engines"] = {"tensor": {"value": value, "unit": "pairs/s", "steps": steps, "default": True},
                          "tensor_v1": {"value": n_pairs / (ms1 * 1e-3), "unit": "pairs/s", "steps": 1, "default": False, "same_counts_as_default": same},
                          "note": "tensor = sift_tc_match_pairs_kernel (D and D^T on the tensor cores, pruned thread-local scans, 16-query re-rank); "
                                  "tensor_v1 = round-1 tc_match_pairs_kernel<sift>"}

    if with_match_api:
        # The reference-shaped per-call route driven the way apps/sfm/main.cpp:84-109 drives it: Match(d1, d2)
        # for every ORDERED pair of a node set, from 4 concurrent callers on one matcher object. The first pass over the set is not
        # timed (it uploads each image once: the device-side descriptor cache); the timed pass is the steady state of that loop.
        n_set = min(n_images, 24 if kind == "orb" else 12)
        calls = [(i, j) for i in range(n_set) for j in range(n_set) if i != j]

        # arguments are marshalled once; the timed loop is the C-ABI call itself (ctypes releases the GIL for its duration), which is
        # what a C++ caller of include/eacham/FeatureMatcherGpu.h sees
        import ctypes
        from eacham_b200 import _lib as L
        from eacham_b200.matcher import _kind_of, _rows_ptr
        lib = L.load()
        prepared = []
        for (i, j) in calls:
            q, qp, qn, qs = _rows_ptr(images[i]); t, tp, tn, ts = _rows_ptr(images[j])
            out = np.empty(max(qn, 1), dtype=L.MATCH_DTYPE)
            prepared.append((_kind_of(q), qp, qn, qs, tp, tn, ts, out, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t()))

        def run_calls(n_threads):
            def caller(w):
                for c in range(w, len(prepared), n_threads):
                    k, qp, qn, qs, tp, tn, ts, out, outp, n = prepared[c]
                    L.check(lib.eacham_gpu_match(m._h, k, qp, qn, qs, tp, tn, ts, 0.8, outp, out.shape[0], ctypes.byref(n)))
            t0 = time.perf_counter()
            th = [threading.Thread(target=caller, args=(w,)) for w in range(n_threads)]
            [t.start() for t in th]; [t.join() for t in th]
            return time.perf_counter() - t0

        cold_s = run_calls(4)
        dt4 = run_calls(4)
        dt1 = run_calls(1)
        rec["match_api"] = {"calls_per_s": len(calls) / dt4, "pairs_per_s": len(calls) / dt4 / 2, "calls": len(calls), "concurrent_callers": 4,
                            "images_in_set": n_set, "one_caller_calls_per_s": len(calls) / dt1, "first_pass_calls_per_s": len(calls) / cold_s,
                            "note": "eacham_gpu_match: host descriptors in, ratio-filtered map out, per call (one direction); the drop-in for "
                                    "FeatureMatcherFlann::Match with the reference's own loop over all ordered pairs (main.cpp:84-109). "
                                    "pairs_per_s = calls / 2. first_pass = the same loop with a cold device cache (each image uploaded once)"}
    m.close()
    return rec


def run_multi_gpu(env, args):
    """N > 1 under torchrun: BASELINE.json configs[4], strong scaling."""
    import eacham_b200
    from eacham_b200 import distributed as D
    name = args.workload or "orb4k5"
    kind, n_desc, n_images, _ = WORKLOADS[name]
    n_images = args.images or n_images
    all_pairs = make_pairs(name, n_images)
    n_pairs = all_pairs.shape[0]
    my_pairs = D.shard_pairs(all_pairs, env.rank, env.world)
    images = make_images(name, n_images) if env.rank == 0 else None
    m = eacham_b200.FeatureMatcherGpu(0.8, device=env.local_rank)
    D.upload_and_broadcast(m, images, src=0)
    upload_ms = m.timing()["upload_ms"]
    arena_bytes = m.arena()[1]
    total_ms, launches, wall_s, clocks = device_timed(env, m, my_pairs, args.steps, max(args.warmup, 3), sample_clocks=True)
    value = n_pairs * args.steps / (total_ms * 1e-3)

    # parity of the GATHERED result on rank 0: shards travel GPU -> GPU over NVLink, pair order restored, CPU-checked sample
    got = D.gather_results_device(m, all_pairs, dst=0)
    cpu = None
    if env.rank == 0:
        res_all, buf_all = got
        threads = cpu_threads()
        v, n_s, cpu_res = time_cpu(images, all_pairs, args.cpu_budget_s, 4096, min_pairs=256)
        _, how = cpu_match_pair(images[0][:64], images[1][:64])
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "reference" if how == "reference-dependency" else "port",
               "sample": f"{n_s} random pairs of the workload (seed 123), OpenCV batchDistance K=2 both directions + ratio + gates + mutual, {threads} threads",
               "parity_checked_pairs": n_s, "parity_mismatches": parity_mismatches(cpu_res, res_all, buf_all),
               "parity_of": f"the result gathered from all {env.world} ranks (NCCL gather of device-resident shards, original pair order)"}
        matches_per_step = int(res_all["count"].sum())
    m.close()                                                  # free this rank's arena before the single-process run below
    env.cpu_barrier()

    # e2e: the route a single-process C++ caller has (eacham_gpu_multi_*), driven by rank 0 over all N GPUs; host buffers in,
    # one pinned host buffer out. The other ranks wait on a CPU barrier with idle GPUs.
    e2e = None
    if env.rank == 0:
        e2e_steps = max(1, min(args.e2e_steps or args.steps, args.steps))
        with eacham_b200.MultiGpuMatcher(list(range(env.world))) as mm:
            e2e_s, phases = 0.0, []
            for s in range(e2e_steps + 1):                      # first pass untimed (NCCL warm-up, pinned buffers reach their size)
                t0 = time.perf_counter()
                mm.Upload(images)
                res, buf = mm.MatchPairsRaw(all_pairs)
                dt = time.perf_counter() - t0
                if s > 0:
                    e2e_s += dt
                    phases.append(mm.timing())
            same = bool(np.array_equal(res["count"], res_all["count"]) and np.array_equal(res["n_mutual"], res_all["n_mutual"]) and
                        parity_mismatches(cpu_res, res, buf) == 0)
            e2e = {"value": n_pairs * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(arena_bytes + all_pairs.nbytes),
                   "d2h_bytes_per_step": int(res.nbytes + buf.nbytes), "steps": e2e_steps, "route": "eacham_gpu_multi_* (one process, N devices)",
                   "includes": "set_descriptors + commit (one H2D to device 0 + one in-process ncclBroadcast over NVLink) + operand prep + pair sharding + "
                               "kernels on all GPUs + every GPU's own D2H into one pinned host buffer + re-interleaving into input order",
                   "phases_ms_last_step": phases[-1] if phases else None, "equals_gathered_result_and_cpu_sample": same}
    env.cpu_barrier()
    if env.rank != 0:
        return
    line = {
        "metric": METRIC if name in ("orb4k", "orb4k5") else f"image pairs matched/sec ({name})",
        "value": value, "unit": "pairs/s", "n_gpus": env.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 bits as fp8 e4m3 {0,1} -> f16 accumulate (exact multiples of 1/2)" if kind == "orb" else "bf16 scoring -> f32 accumulate, f32 re-rank",
        "data": "synthetic", "config": workload_config(name, n_images, n_pairs, env.world),
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline_record(kind, n_desc, len(my_pairs) * args.steps, total_ms * 1e-3, "tensor" if kind == "orb" else "sift", name),
        "cpu_baseline": cpu, "clocks": clocks, "upload_ms": upload_ms, "arena_bytes": int(arena_bytes), "wall_s_timed_region": wall_s,
        "matches_per_step": matches_per_step,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="default: orb4k at N = 1, orb4k5 at N > 1")
    ap.add_argument("--images", type=int, default=0, help="override the image count (debug)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: as many as --steps")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-sub-workloads", action="store_true", help="skip the configs 1 / 3 / 4 sub-records of the N = 1 run")
    ap.add_argument("--orb-engine", default="tensor", choices=["popc", "tensor", "tensor_alu", "tensor_v1"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args.gpus)
    if env.world > 1:
        run_multi_gpu(env, args)
        env.close()
        return

    name = args.workload or "orb4k"
    kind = WORKLOADS[name][0]
    e2e_steps = max(1, min(args.e2e_steps or args.steps, args.steps))
    main_rec = run_single_gpu_workload(env, name, args.steps, args.warmup, e2e_steps, args.cpu_budget_s, orb_engine=args.orb_engine,
                                       sample_clocks=True, images_override=args.images, with_engines=True, with_match_api=True, with_verify=True)
    line = {
        "metric": METRIC if name in ("orb4k", "orb4k5") else f"image pairs matched/sec ({name})",
        "value": main_rec.pop("value"), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": main_rec.pop("warmup"),
        "ms_per_step": main_rec.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 bits as fp8 e4m3 {0,1} -> f16 accumulate (exact multiples of 1/2)" if kind == "orb" else "bf16 scoring -> f32 accumulate, f32 re-rank",
        "data": "synthetic",
    }
    main_rec.pop("unit"); main_rec.pop("steps")
    line.update(main_rec)
    if not args.no_sub_workloads and args.workload is None:
        subs = {}
        for sub, (st, wu, es, budget) in {"orb2k": (5, 3, 3, 6.0), "sift8k": (2, 3, 1, 8.0), "kitti2k": (5, 3, 2, 6.0)}.items():
            t0 = time.perf_counter()
            r = run_single_gpu_workload(env, sub, st, wu, es, budget)
            r["wall_s_total"] = time.perf_counter() - t0
            subs[f"{sub}_cfg{WORKLOADS[sub][3] + 1}"] = r
        line["workloads"] = subs
    print(json.dumps(line))
    env.close()


if __name__ == "__main__":
    main()
