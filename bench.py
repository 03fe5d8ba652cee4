#!/usr/bin/env python
"""bench.py -- image pairs matched / s for exhaustive 4k-ORB matching (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload orb4k|orb2k|sift8k]

A "step" is one full pass of the hot path over the workload's pair list: for every unordered image pair both kNN(k=2)
directions, ratio 0.8, gates and the mutual filter (/root/reference/apps/sfm/main.cpp:84-147).  N=1 workload =
BASELINE.json configs[1]: 500 synthetic images x 4096 ORB descriptors, all 124,750 pairs.  For N>1 the image count
grows so that every rank keeps ~124,750 pairs (weak scaling); the pair list is sharded rank::N with no data-path
collective, after ONE NCCL broadcast of the descriptor arena.

value  = pairs / s with descriptors resident in HBM (CUDA events around the matching kernel on the library's stream).
e2e    = the same through the C-ABI calls a user makes, host buffers in, host buffers out (upload + broadcast +
         match + D2H of results inside the timed region).
roofline: the yardstick is SURVEY.md 8(d)'s POPC roofline (8 POPC32 per 256-bit distance, matrix once per pair, POPC pipe
         = 16 lane-ops/clk/SM measured: profiles/r01_pipe_microbench.jsonl): achieved = 8*N*M * pairs / kernel time.
         Two ORB engines produce identical bytes and are BOTH timed in every run (`engines`): the default tcgen05 FP8 engine
         (bits as e4m3 0/1: |a-b|^2 = hamming, exact) and the XOR+POPC kernel north_star describes.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 124750          # C(500, 2)
WORKLOADS = {
    # name: (kind, descriptors per image, images at N=1)
    "orb4k": ("orb", 4096, 500),
    "orb2k": ("orb", 2048, 100),
    "sift8k": ("sift", 8192, 500),
    # BASELINE config 4: KITTI-shaped sequence, each frame against the next 20 (fixed job: strong scaling for N > 1)
    "kitti2k": ("orb", 2048, 4541),
}
WINDOW = {"kitti2k": 20}


def images_for(n_gpus: int, base_images: int) -> int:
    if n_gpus == 1:
        return base_images
    target = n_gpus * base_images * (base_images - 1) // 2
    return int(math.ceil((1 + math.sqrt(1 + 8 * target)) / 2))


def peaks():
    p = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        p["_source"] = "measured"
    else:
        p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
             "_source": "fallback"}
    return p


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


def make_images(kind: str, n_images: int, n_desc: int, seed: int, window: int = 0):
    from eacham_b200 import synth
    if kind == "orb" and window:
        # landmarks drawn from a sliding window of the pool: neighbouring frames overlap, distant ones do not
        return synth.orb_image_set(n_images, n_desc, seed=seed, pool=max(20000, 40 * n_images), window=6000)
    if kind == "orb":
        return synth.orb_image_set(n_images, n_desc, seed=seed, pool=20000)
    return synth.sift_image_set(n_images, n_desc, seed=seed, pool=40000)


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


def time_cpu(images, pairs, budget_s: float, max_pairs: int):
    """Times the CPU path on a bounded sample of the workload's pairs. Returns (pairs/s, n_sampled, results)."""
    rng = np.random.default_rng(123)
    order = rng.permutation(len(pairs))[:max_pairs]
    t0 = time.perf_counter()
    done, results = 0, {}
    for k in order:
        i, j = int(pairs[k][0]), int(pairs[k][1])
        results[int(k)] = cpu_match_pair(images[i], images[j])[0]
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= 4:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, results


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (OpenCV exact matcher, all host threads) on the
    same workload; each step is a bounded sample of the workload's pairs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, n_desc, base_images = wl
    n_images = images_for(args.gpus, base_images)
    # generating all images of the workload is not needed to time a sample: draw the sampled pairs' images only
    n_gen = min(n_images, 24)
    images = make_images(kind, n_gen, n_desc, seed=2)
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
        "impl": "reference", "metric": "image pairs matched/sec (exhaustive, 4k ORB)" if args.workload == "orb4k" else f"image pairs matched/sec ({args.workload})",
        "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if kind == "orb" else "f32", "data": "synthetic",
        "config": workload_config(args, wl, n_images, n_images * (n_images - 1) // 2),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference" if how == "reference-dependency" else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, wl, n_images, n_pairs):
    kind, n_desc, _ = wl
    win = WINDOW.get(args.workload, 0)
    return {"workload": f"synthetic {n_images} images x {n_desc} {'ORB-256bit' if kind == 'orb' else 'SIFT-128 f32'} descriptors, "
                        + (f"sliding window (each frame vs next {win}) = {n_pairs} unordered pairs" if win else f"exhaustive {n_pairs} unordered pairs")
                        + ", ratio 0.8 + cross-check, gates 30/30",
            "images": n_images, "descriptors_per_image": n_desc, "pairs": n_pairs, "pairs_per_gpu": n_pairs // max(args.gpus, 1),
            "parallelism": f"pair list sharded rank::{args.gpus}, arena replicated by one NCCL broadcast" if args.gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB device memset outside the event-timed region); per-step CUDA events summed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="orb4k", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="override the image count (debug)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--orb-engine", default="tensor", choices=["popc", "tensor"],
                    help="tensor = FP8 tensor-core engine (bits as 0/1 e4m3 through tcgen05; default, fastest); "
                         "popc = XOR+POPC kernel (the engine BASELINE.json's north_star describes). Bit-identical results; "
                         "the other engine is timed too and reported under `engines`.")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    kind, n_desc, base_images = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import eacham_b200
    from eacham_b200 import synth, _lib as L
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    win = WINDOW.get(args.workload, 0)
    n_images = args.images or (base_images if win else images_for(world, base_images))
    all_pairs = synth.window_pairs(n_images, win) if win else synth.exhaustive_pairs(n_images)
    n_pairs = all_pairs.shape[0]
    from eacham_b200 import distributed as D
    my_pairs = D.shard_pairs(all_pairs, rank, world)
    images = make_images(kind, n_images, n_desc, seed=2, window=win) if rank == 0 else None

    m = eacham_b200.FeatureMatcherGpu(0.8, device=local_rank, orb_engine=args.orb_engine)
    # ---- descriptors resident in HBM --------------------------------------------------------------------
    if dist is None:
        m.Upload(images)
    else:
        D.upload_and_broadcast(m, images, src=0)
    upload_ms = m.timing()["upload_ms"]
    arena_bytes = m.arena()[1]

    for _ in range(max(args.warmup, 3)):
        m.MatchPairsDevice(my_pairs)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    kernel_ms, launches = [], 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        m.flush_l2(256 << 20)
        barrier()
        m.MatchPairsDevice(my_pairs)
        t = m.timing()
        kernel_ms.append(t["kernel_ms"]); launches += t["kernel_launches"]
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = max_over_ranks(float(sum(kernel_ms)))
    value = n_pairs * args.steps / (total_ms * 1e-3)

    # ---- the other ORB engine, same data, same pairs (so that both the tensor-core and the XOR+POPC numbers are in every run) ----
    engines = None
    if kind == "orb":
        other = "popc" if args.orb_engine == "tensor" else "tensor"
        m2 = eacham_b200.FeatureMatcherGpu(0.8, device=local_rank, orb_engine=other)
        if dist is None:
            m2.Upload(images)
        else:
            D.upload_and_broadcast(m2, images, src=0)
        alt_steps = max(1, min(2, args.steps))
        m2.MatchPairsDevice(my_pairs)
        alt_ms = []
        for _ in range(alt_steps):
            m2.flush_l2(256 << 20)
            barrier()
            m2.MatchPairsDevice(my_pairs)
            alt_ms.append(m2.timing()["kernel_ms"])
        barrier()
        alt_total = max_over_ranks(float(sum(alt_ms)))
        m2.close()
        popc_peak = 148 * 16.0 * peaks()["sm_max_mhz"] * 1e6
        def eng(v):
            return {"value": v, "unit": "pairs/s", "frac_of_popc_roofline": v * 8.0 * n_desc * n_desc / world / popc_peak}
        engines = {args.orb_engine: dict(eng(value), steps=args.steps, default=True),
                   other: dict(eng(n_pairs * alt_steps / (alt_total * 1e-3)), steps=alt_steps, default=False),
                   "note": "tensor = tc_match_pairs_kernel<orb> (tcgen05 kind::f8f6f4, bits as e4m3 0/1, exact); popc = orb_match_pairs_kernel "
                           "(XOR + carry-save POPC); identical outputs (tests/test_gpu_orb_tensor.py)"}

    # ---- end to end through the C ABI: host descriptors in, host results out ----------------------------
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_s, h2d, d2h = 0.0, 0, 0
    for s in range(e2e_steps + 1):          # first pass untimed (pinned staging / result buffers reach their size)
        barrier()
        t0 = time.perf_counter()
        if dist is None:
            m.Upload(images)
        else:
            D.upload_and_broadcast(m, images, src=0)
        if dist is None:
            res, buf = m.MatchPairsRaw(my_pairs)                      # host buffers in, host buffers out
        else:
            m.MatchPairsDevice(my_pairs)                              # shard stays in HBM ...
            got = D.gather_results_device(m, n_pairs, dst=0)          # ... NVLink gather, one D2H on rank 0
            if rank == 0:
                res, buf = got
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        if s > 0:
            e2e_s += dt
            h2d = arena_bytes + my_pairs.nbytes
            d2h = (res.nbytes + buf.nbytes) if rank == 0 else 0
    e2e_value = n_pairs * e2e_steps / e2e_s

    # ---- the reference-shaped per-call route: Match(d1, d2) one direction at a time ----
    match_api = None
    if rank == 0 and world == 1:
        n_calls = 128 if kind == "orb" else 16
        m.Match(images[0], images[1])
        t0 = time.perf_counter()
        for c in range(n_calls):
            i, j = c % (n_images - 1), (c % (n_images - 1)) + 1
            m.Match(images[i], images[j]) if c % 2 == 0 else m.Match(images[j], images[i])
        dt = time.perf_counter() - t0
        match_api = {"calls_per_s": n_calls / dt, "pairs_per_s": n_calls / dt / 2, "calls": n_calls,
                     "note": "eacham_gpu_match: host descriptors in, ratio-filtered map out, per call (upload + 2 kernels + D2H); "
                             "the drop-in for FeatureMatcherFlann::Match with the reference's own loop"}

    if rank == 0:
        pk = peaks()
        res_all, buf_all = res, buf
        # roofline of the dominant kernel (orb_match_pairs_kernel): POPC-pipe bound
        sm_count, popc_per_clk = 148, 16.0
        if kind == "orb":
            work_per_pair = 8.0 * n_desc * n_desc                       # POPC32 per unordered pair (distance matrix ONCE)
            peak = sm_count * popc_per_clk * pk["sm_max_mhz"] * 1e6     # lane-ops/s at max clock
            achieved = work_per_pair * (len(my_pairs) * args.steps) / (float(sum(kernel_ms)) * 1e-3)
            traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
            tkey = args.workload if args.orb_engine == "tensor" else args.workload + "_popc"
            tj = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
            traffic = tj.get(tkey)
            traffic_note = tj.get("_note")
            fp8_flops = 2.0 * 256 * n_desc * n_desc * (len(my_pairs) * args.steps) / (float(sum(kernel_ms)) * 1e-3)
            alu_ops_per_distance = 22.0        # 16 LOP3 + 6 VIMNMX issued on the ALU pipe per distance (SASS count)
            alu_peak = sm_count * 64.0 * pk["sm_max_mhz"] * 1e6
            alu_achieved = alu_ops_per_distance * n_desc * n_desc * (len(my_pairs) * args.steps) / (float(sum(kernel_ms)) * 1e-3)
            roof = {"bound": "int_popc", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TPOPC32/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                    "peak_source": f"16.00 POPC lane-ops/clk/SM measured (profiles/r01_pipe_microbench.jsonl) x 148 SMs x sm_max_mhz "
                                   f"{pk['sm_max_mhz']:.0f} ({pk['_source']})",
                    "work_per_pair": work_per_pair, "kernel": "orb_match_pairs_kernel" if args.orb_engine == "popc" else "tc_match_pairs_kernel<orb>",
                    "engine": args.orb_engine,
                    "note": ("algorithmic work = 8 POPC32 per 256-bit distance, distance matrix evaluated once per pair (SURVEY.md 8(d)). "
                             "frac > 1 is real: the XOR+POPC kernel compresses the 8 XOR words with carry-save adders and issues 4 POPC per "
                             "distance, so the ALU pipe (LOP3 + VIMNMX) is its limiter -- see alu_pipe.") if args.orb_engine == "popc" else
                            ("algorithmic work = 8 POPC32 per 256-bit distance, matrix once per pair (SURVEY.md 8(d)), kept as the yardstick; "
                             "this engine issues no POPC at all: distances come out of tcgen05 FP8 MMAs (bits as e4m3 0/1, |a-b|^2 = hamming, exact) "
                             "and the CUDA-core top-2 epilogue (ALU pipe) is the limiter. See `engines` for the XOR+POPC kernel on the same run."),
                    "tensor_pipe": None if args.orb_engine != "tensor" else {
                        "achieved_tflops_fp8": fp8_flops / 1e12, "peak_tflops_fp8": 2 * pk["bf16_tflops"],
                        "frac": fp8_flops / (2 * pk["bf16_tflops"] * 1e12),
                        "note": "algorithmic 2*256*N*M FLOP per pair through kind::f8f6f4 MMAs vs 2 x the measured bf16 burst peak (no measured FP8 "
                                "peak on file); ncu: tensor pipe 31%, ALU pipe 74%, issue 68% (profiles/r01d_ncu_orb_tensor_engine_summary.json) -- "
                                "the top-2 epilogue on the CUDA cores, not the tensor core, bounds this kernel"},
                    "alu_pipe": None if args.orb_engine != "popc" else {"achieved_tlaneops": alu_achieved / 1e12, "peak_tlaneops": alu_peak / 1e12, "frac": alu_achieved / alu_peak,
                                 "ops_per_distance": alu_ops_per_distance, "peak_source": "64 lane-ops/clk/SM (LOP3 63.2 measured) x 148 x sm_max_mhz"},
                    "hbm": {"achieved_gbs": 2 * n_desc * 32 * len(my_pairs) * args.steps / (float(sum(kernel_ms)) * 1e-3) / 1e9,
                            "peak_gbs": pk["hbm_gbs"], "note": "algorithmic bytes = both images of every pair; far from the HBM bound"}}
        else:
            work_per_pair = 2.0 * 128 * n_desc * n_desc
            peak = pk["bf16_tflops_sustained"] * 1e12
            achieved = work_per_pair * (len(my_pairs) * args.steps) / (float(sum(kernel_ms)) * 1e-3)
            roof = {"bound": "tensor", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": f"bf16_tflops_sustained ({pk['_source']})", "work_per_pair": work_per_pair}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = cpu_threads()
            v, n_s, cpu_res = time_cpu(images, all_pairs, args.cpu_budget_s, 2048)
            _, how = cpu_match_pair(images[0][:64], images[1][:64])
            # the CPU-timed pairs double as parity data: GPU result must equal OpenCV + reference logic bit for bit
            mism = 0
            for k, want in cpu_res.items():
                r = res_all[k]
                got = buf_all[int(r["offset"]): int(r["offset"]) + int(r["count"])]
                ok = (int(r["n12"]), int(r["n21"]), int(r["n_mutual"])) == (want["n12"], want["n21"], want["n_mutual"])
                if ok and want["connected"]:
                    ok = np.array_equal(np.stack([got["query"], got["train"]], 1), want["matches"])
                mism += (not ok)
            one_thread = None
            try:                                                  # also a 1-thread figure (3 pairs)
                import cv2
                cv2.setNumThreads(1)
                t1 = time.perf_counter()
                for k in list(cpu_res)[:3]:
                    cpu_match_pair(images[int(all_pairs[k][0])], images[int(all_pairs[k][1])])
                one_thread = 3 / (time.perf_counter() - t1)
                cv2.setNumThreads(threads)
            except Exception:
                pass
            cpu = {"value": v, "unit": "pairs/s", "cores": threads, "one_thread_value": one_thread, "kind": "reference" if how == "reference-dependency" else "port",
                   "sample": f"{n_s} random pairs of the workload (seed 123), OpenCV {('cv2 ' + __import__('cv2').__version__) if how == 'reference-dependency' else 'absent: C port'} "
                             f"batchDistance K=2 both directions + ratio + gates + mutual, {threads} threads",
                   "parity_checked_pairs": n_s, "parity_mismatches": mism}
        line = {
            "metric": "image pairs matched/sec (exhaustive, 4k ORB)" if args.workload == "orb4k" else f"image pairs matched/sec ({args.workload})",
            "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if win else "weak", "vs_baseline": None,
            "dtype": ("u8 (XOR+POPC)" if args.orb_engine == "popc" else "u8 bits as fp8 e4m3 {0,1} -> f32 accumulate (exact integers)") if kind == "orb"
                     else "bf16 scoring -> f32 accumulate, f32 re-rank",
            "data": "synthetic",
            "config": dict(workload_config(args, wl, n_images, n_pairs), orb_engine=args.orb_engine) if kind == "orb" else workload_config(args, wl, n_images, n_pairs),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "includes": "set_descriptors + commit (pinned staging, one H2D)" +
                    (" + NCCL arena broadcast" if world > 1 else "") + " + pair list H2D + kernel + D2H of results and matches" +
                    (" (shards gathered to rank 0 over NVLink, one D2H there)" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "upload_ms": upload_ms, "arena_bytes": int(arena_bytes), "wall_s_timed_region": wall_s,
            "matches_per_step": int(res_all["count"].sum()),
            "match_api": match_api,
            "engines": engines,
        }
        print(json.dumps(line))
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
