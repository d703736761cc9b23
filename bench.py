#!/usr/bin/env python
"""bench.py -- grouping images/s on synthetic 128x128x(18+30) maps, 30 persons/image (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                 # the reference's CPU algorithm (Python port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU, images sharded (weak scaling)

A "step" is one pass of the hot path (peaks -> candidate scoring -> greedy matching -> person assembly, the
window /root/reference/evaluate.py:507-513 times) over one batch of 256 images per GPU; at N > 1 it ends with an
NCCL gather of the person lists to rank 0.  One JSON line is printed by rank 0:

  value      whole-job images/s with the maps already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the C-ABI host entry point (spg_group_host): pinned HOST maps in, person lists
             back on the host, H2D/D2H inside the timed region
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json's HBM peak
  kernels    per-kernel event times of the same timed steps
  cpu_baseline  the Python/numpy port of the reference's algorithm, one process, same batch (N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "grouping images/sec @128x128 heatmaps, 30 persons/img"
UNIT = "images/s"
H = W = 128
BASE_SEED = 20260921
CAP_ROWS = 64  # person-row capacity of the handles (<= 40 persons/image in this workload; overflow would trip the status assert)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--persons", type=int, default=30)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"BASELINE configs[2] per-GPU shard: batch={args.batch}/GPU synthetic 128x128x(18+30) f32 maps, "
                        f"{args.persons} persons/img",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "persons": args.persons, "H": H, "W": W,
            "keypoint_channels": 18, "limb_channels": 30,
            "capacities": {"max_peaks_per_part": 64, "max_cands_per_limb": 1024, "max_person_rows": CAP_ROWS},
            "parallelism": f"image-sharded x{world}" + (" + NCCL gather of person lists to rank 0, overlapped with the next step (two workspaces used alternately)" if world > 1 else ""),
            "l2": f"inputs {args.batch * 48 * H * W * 4 / 1e6:.0f} MB/GPU > 126 MB L2: every step streams from HBM, no flush needed"}


def make_shard(args, rank):
    from improved_body_parts_b200 import synth

    return synth.make_batch(BASE_SEED + rank * args.batch, args.batch, H, W, args.persons)


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background while the timed regions run."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if any(a - 0.06 <= t <= b + 0.06 for a, b in self.windows)]
        used, scope = (inside, "timed regions") if len(inside) >= 2 else ([r for _, r in self.rows], "whole run (timed regions shorter than the sampling period)")
        sm, smax, reasons = [], [], set()
        for r in used:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 10:
                continue
            try:
                sm.append(float(f[2])); smax.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (Python/numpy port, oracle/grouping_port.py -- the reference
    itself is pure Python and cannot travel to this box) on all host cores, one image per task."""
    if rank != 0:
        return
    from improved_body_parts_b200 import skeleton
    from oracle import grouping_port as gp

    cores = os.cpu_count() or 1
    heat, paf = make_shard(args, 0)
    sample = min(args.batch, max(16, cores * 4))
    params = skeleton.default_params()
    pool, run = gp.make_pool(heat, paf, H, params, skeleton.LIMBS, workers=cores)
    try:
        for _ in range(max(args.warmup, 1)):
            run(range(sample))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(range(sample))
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = sample * args.steps / dt
    cfg = workload_config(args, world)
    desc = f"{sample} images of the batch per step, fork pool of {cores} processes, one image per task"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 maps; f64 coordinates", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from improved_body_parts_b200 import skeleton
    from improved_body_parts_b200.grouping import Grouper

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the grouping path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    heat_np, paf_np = make_shard(args, rank)
    B = args.batch
    params = skeleton.default_params()
    heat_pin = torch.from_numpy(heat_np).pin_memory()
    paf_pin = torch.from_numpy(paf_np).pin_memory()
    heat_d = heat_pin.to(dev, non_blocking=True)
    paf_d = paf_pin.to(dev, non_blocking=True)
    # Two handles used alternately: at N > 1 the person lists of step k are gathered (NCCL, side stream) straight out of
    # handle k%2's workspace while step k+1 runs on the other handle -- no staging copy, the transfer overlaps compute.
    # The timed region ends only after the last gather has completed.  At N = 1 only handle 0 is used.
    n_handles = 2 if world > 1 else 1
    groupers = [Grouper(max_batch=B, max_h=H, max_w=W, max_person_rows=CAP_ROWS, device=local_rank) for _ in range(n_handles)]
    g = groupers[0]
    all_views = [x.device_tensors() for x in groupers]
    views = all_views[0]
    from improved_body_parts_b200.sharding import gather_people
    locals_ = [{"n_persons": v["n_persons"][:B], "people_xy": v["people_xy"][:B], "people_score": v["people_score"][:B]}
               for v in all_views]
    gathered = [None]
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    ev_ready = [torch.cuda.Event() for _ in range(n_handles)]
    ev_done = [torch.cuda.Event() for _ in range(n_handles)]
    if world > 1:
        for e in ev_done:
            e.record(torch.cuda.current_stream())
    step_no = [0]

    def gather(hi):  # NCCL gather of handle hi's person lists to rank 0 (rank order == image order), on the side stream
        if world == 1:
            return
        main = torch.cuda.current_stream()
        ev_ready[hi].record(main)
        comm_stream.wait_event(ev_ready[hi])
        with torch.cuda.stream(comm_stream):
            gathered[0] = gather_people(locals_[hi], dst=0)
            ev_done[hi].record(comm_stream)

    def gather_join():  # make the main stream wait for the outstanding gathers
        if world > 1:
            for e in ev_done:
                torch.cuda.current_stream().wait_event(e)

    n_ev = 6
    stream = torch.cuda.current_stream()

    def step(evs=None):
        hi = step_no[0] % n_handles
        step_no[0] += 1
        gh = groupers[hi]
        if world > 1:
            stream.wait_event(ev_done[hi])  # this handle's previous person lists have left the GPU
        if evs: evs[0].record(stream)
        gh.nms_peaks(heat_d, params)
        if evs: evs[1].record(stream)
        gh.limb_score(paf_d, H, params)
        if evs: evs[2].record(stream)
        gh.limb_match(B, params)
        if evs: evs[3].record(stream)
        gh.assemble(B, params)
        if evs: evs[4].record(stream)
        gather(hi)
        if evs: evs[5].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # Per-kernel durations come from CUDA events recorded around every kernel of every `stage_every`-th step of the timed
    # region (an event between two kernels costs ~2.5 us of launch gap, 5 % of a step if every step carries six);
    # the region itself is bracketed by its own two events.
    stage_every = max(1, min(5, args.steps // 3))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(0, args.steps, stage_every)]
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sum(x.launch_count for x in groupers)
    w0 = time.time()
    ev_start.record(stream)
    for k in range(args.steps):
        step(evs[k // stage_every] if k % stage_every == 0 else None)
    gather_join()
    ev_end.record(stream)
    torch.cuda.synchronize()
    w1 = time.time()
    launches = sum(x.launch_count for x in groupers) - l0
    barrier()
    elapsed_ms = ev_start.elapsed_time(ev_end)
    stage_ms = [statistics.fmean(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(n_ev - 1)]
    if sampler:
        sampler.window(w0, w1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms / 1e3)

    # ---- correctness guard inside the bench: statuses clean, persons found
    r_status = views["status"][:B].cpu().numpy()
    r_np = views["n_persons"][:B].cpu().numpy()
    assert (r_status == 0).all(), f"status flags set: {np.unique(r_status)}"
    assert r_np.min() > 0, "no persons found -- the timed path did no work"
    if world > 1 and rank == 0:
        got = gathered[0]["n_persons"]
        assert got.shape[0] == world * B and bool((got[:B].cpu() == views["n_persons"][:B].cpu()).all()), "gathered lists are not in image order"

    # ---- e2e: HOST maps -> spg_group_host (H2D + kernels + D2H inside) -> person lists on the host
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    out = None
    for _ in range(2):
        out = g.group_host(heat_pin.numpy(), paf_pin.numpy(), H, params, out)
        gather(0)
    gather_join()
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = g.group_host(heat_pin.numpy(), paf_pin.numpy(), H, params, out)
        gather(0)
        gather_join()  # the next call overwrites handle 0's lists
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    w1 = time.time()
    if sampler:
        sampler.window(w0, w1)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * B * e2e_steps / e2e_s
    assert np.array_equal(out["n_persons"], r_np), "host entry point disagrees with the device entry point"
    h2d = world * (heat_np.nbytes + paf_np.nbytes)
    d2h = world * sum(out[k].nbytes for k in ("n_persons", "people_xy", "people_score", "status"))
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes per launch / event time)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    names = list(g.stage_kernels())  # the kernel variants that actually ran (ncu names)
    alg_bytes = [B * 18 * H * W * 4, B * 30 * H * W * 4, None, None]  # K1 reads heat once, K2a reads paf once
    kernels = {}
    for i, nme in enumerate(names):
        kernels[nme] = {"ms": stage_ms[i], "algorithmic_GBps": (alg_bytes[i] / (stage_ms[i] * 1e-3) / 1e9) if alg_bytes[i] else None}
    if world > 1:
        kernels["nccl_gather"] = {"ms": None, "algorithmic_GBps": None,
                                  "note": "runs on a side stream out of the other handle's workspace, overlapped with the next step"}
    dom = max(range(4), key=lambda i: stage_ms[i])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(names[dom])
    if alg_bytes[dom]:
        ach = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
        roofline = {"kernel": names[dom], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes[dom]}
    else:
        # a latency-bound kernel dominates: report the path's algorithmic bytes (3 145 728 B/image) over its time
        ach = B * 48 * H * W * 4 / (stage_ms[dom] * 1e-3) / 1e9
        roofline = {"kernel": names[dom], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": B * 48 * H * W * 4,
                    "note": "serial, latency-bound kernel; bytes are the whole path's per-image figure x batch"}
    # the north-star kernel always reported beside it
    ls = kernels[names[1]]
    roofline["limb_score_frac"] = ls["algorithmic_GBps"] / hbm_peak

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 maps; f64 coordinates/scores", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "spg_group_host (C ABI, pinned host maps in, person lists out)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "kernel_timing": f"CUDA events around every kernel of every {stage_every}th step of the timed region ({len(evs)} of {args.steps} steps)",
        "persons_found_per_image": float(r_np.mean()),
    }

    if world == 1 and not args.no_cpu_baseline:
        from oracle import grouping_port as gp
        from oracle import spg_oracle as so

        n_cpu = min(B, 256)
        gp.group_image(heat_np[0], paf_np[0], H, params, skeleton.LIMBS)
        t0 = time.perf_counter()
        found = [gp.group_image(heat_np[i], paf_np[i], H, params, skeleton.LIMBS)[3].shape[0] for i in range(n_cpu)]
        dt = time.perf_counter() - t0
        assert found == [int(v) for v in r_np[:n_cpu]], "CPU port and CUDA path disagree on person counts"
        result["cpu_baseline"] = {"value": n_cpu / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                  "sample": f"first {n_cpu} images of the same batch, Python/numpy port of the reference "
                                            f"(oracle/grouping_port.py), one process, {dt:.1f} s"}
        t0 = time.perf_counter()
        so.group_batch(heat_np, paf_np, skeleton.LIMBS, H, params, threads=1)
        dt1 = time.perf_counter() - t0
        nt = so.max_threads()
        t0 = time.perf_counter()
        so.group_batch(heat_np, paf_np, skeleton.LIMBS, H, params, threads=nt)
        dtn = time.perf_counter() - t0
        result["cpu_checker_c"] = {"single_thread": B / dt1, "all_threads": B / dtn, "threads": nt, "unit": UNIT,
                                   "note": "oracle/spg_oracle.c, a C rewrite used as test checker (not the reference's implementation)"}
    print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        raise SystemExit(f"--gpus {args.gpus} needs `python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py ...`")
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
