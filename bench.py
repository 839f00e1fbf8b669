#!/usr/bin/env python
"""Benchmark of the per-frame annotation hot path (BASELINE.json metric: annotated frames/s at
1080p / 100 instances; HBM GB/s of the mask-scan kernel vs the measured peak).

    python bench.py --gpus N --steps K --warmup W            # our arm (1 process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference-style numpy path on host cores

A *step* is one pass of the hot path over one batch of synthetic annotator tensors of
BASELINE config 2: 64 frames of 1920x1080, 100 instances — mask scan (K1) + per-object
projection / pose (K2) + occlusion ratios, compaction, record emission and class histogram
(K4).  ``value`` times K steps on the device with inputs resident in HBM (the 531 MB mask batch
is 4x the L2, so every step streams from HBM); ``e2e`` times the same batch through the public
``ConstructionLabelWriter.annotate_batch`` call with HOST (pinned) annotator arrays, i.e.
including the host table build, the H2D copies and the D2H read of the records.

One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "annotated frames/sec at 1080p/100 instances"
UNIT = "frames/s"
WORKLOAD = "c2: 64 x 1920x1080 frames, 100 instances, mask scan + 3D-box projection/pose + occlusion + emission"
BATCH = 64
CONFIG_KEY = "c2"


# --------------------------------------------------------------------------------------------
# synthetic batch
# --------------------------------------------------------------------------------------------
def make_frames(n: int, first: int = 0):
    from constructionsceneposeestimation_b200 import synthetic

    return synthetic.make_batch(synthetic.CONFIGS[CONFIG_KEY], n, first)


# --------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (= the reference's numpy path restated) on the host cores
# --------------------------------------------------------------------------------------------
def _cpu_frame(frame):
    from tests import helpers

    o = helpers.oracle_pipeline([frame])
    return int(o["n_out"][0])


def cpu_throughput(frames, procs: int, reps: int = 1):
    """frames/s of the oracle pipeline over `frames`, `procs` worker processes (best of reps)."""
    import multiprocessing as mp

    best = None
    if procs <= 1:
        for _ in range(reps):
            t0 = time.perf_counter()
            for fr in frames:
                _cpu_frame(fr)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return len(frames) / best
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_frame, frames[: procs])  # warm the workers (imports)
        for _ in range(reps):
            t0 = time.perf_counter()
            pool.map(_cpu_frame, frames, chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return len(frames) / best


def run_reference(args) -> int:
    """Reference arm: the reference-style numpy path (oracle port; the reference is pure Python and
    leaves most of this path unimplemented) on all host cores.  Each step is a bounded
    sample of the workload's frames."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = max(cores, min(BATCH, 2 * cores))
    frames = make_frames(sample)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_frame, frames, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_frame, frames, chunksize=1)
        dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_frames": BATCH, "resolution": "1920x1080", "instances": 100,
                   "frames_per_reference_step": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {BATCH} frames of one step per reference step, numpy oracle "
                                   f"(bincount+find_objects scan, per-object projection, emission), "
                                   f"{cores} worker processes"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    from constructionsceneposeestimation_b200 import _lib, ops
    from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for our arm (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()
    # keep stdout to the one JSON line: libraries (NCCL's version banner) write to fd 1 from C, so
    # point fd 1 at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: each rank owns its own frame range (weak scaling: 64 frames per GPU) -------
    uniq = int(os.environ.get("CSPE_BENCH_UNIQUE", str(BATCH)))
    frames = make_frames(uniq, first=rank * BATCH)
    if uniq < BATCH:
        frames = [frames[i % uniq] for i in range(BATCH)]
    # input tables come from the PRODUCT's host logic (classes.py / camera.py); the oracle is only
    # used below as the checker of one frame and as the CPU baseline
    from constructionsceneposeestimation_b200.sweep import build_host_tables

    lut, obj_record, slot_class, records, cam, _objs = build_host_tables(frames)
    H, W = frames[0]["instance_segmentation"]["data"].shape
    N = obj_record.shape[1]
    mask_host = torch.empty((BATCH, H, W), dtype=torch.int32, pin_memory=True)
    for i, fr in enumerate(frames):
        mask_host[i].copy_(torch.from_numpy(fr["instance_segmentation"]["data"].view(np.int32)))
    from constructionsceneposeestimation_b200.pipeline import LabelPipeline

    pipe = LabelPipeline(BATCH, H, W, N, records.shape[1], lut.shape[1], dev, per_frame_lut=True, min_pixels=1,
                         use_graph=args.graph)
    pipe.frame_base = rank * BATCH
    pipe.mask.copy_(mask_host, non_blocking=True)
    pipe.lut.copy_(torch.from_numpy(lut))
    pipe.obj_record.copy_(torch.from_numpy(obj_record))
    pipe.slot_class.copy_(torch.from_numpy(slot_class))
    pipe.records_in.copy_(torch.from_numpy(records))
    pipe.cam.copy_(torch.from_numpy(cam))
    d_mask, d_lut, scan, rec_out, n_out, class_hist = pipe.mask, pipe.lut, pipe.scan, pipe.records, pipe.n_out, pipe.class_hist
    torch.cuda.synchronize()

    def step():
        pipe.run()   # mask_scan (accumulate) || project_objects -> emit (+ scan-table reset)
    launches_per_step = pipe.launches_per_run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)   # samples clocks from the warm-up to the end of the e2e region
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:  # the histogram all-gather is part of the timed region: set its NCCL channels up here
        warm = torch.empty((world, _lib.NUM_CLASSES), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(warm, class_hist)
    barrier()

    # what the timed path produces for frame 0: compared with the oracle inside the CPU-baseline leg below
    gpu_frame0 = rec_out[0].cpu().numpy().view(_lib.RECORD_DTYPE).reshape(-1)[: int(n_out[0])].copy() if rank == 0 else None

    # ---- value: K steps, device-timed, inputs resident in HBM ---------------------------------
    class_hist.zero_()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    if world > 1:  # the path's one collective: all-gather of the per-class histogram at sweep end
        gathered = torch.empty((world, _lib.NUM_CLASSES), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, class_hist)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * BATCH * args.steps / (ms_total / 1000.0)

    # ---- roofline of the dominant kernel: the mask scan alone, K launches -----------------------
    barrier()
    reps = max(args.steps, 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.mask_scan(d_mask, d_lut, N, out=scan, accumulate=True)   # the scan kernel only (no init launch)
    e1.record()
    torch.cuda.synchronize()
    scan_ms = e0.elapsed_time(e1) / reps
    algo_bytes = 4.0 * H * W * BATCH + 20.0 * N * BATCH
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except (OSError, ValueError, KeyError, TypeError):
        pass
    traffic = None
    tfile = ROOT / "profiles" / "scan_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get("dram_bytes_per_launch")
        except (ValueError, OSError):
            traffic = None

    # ---- e2e: the public Writer call with HOST annotator arrays ----------------------------------
    # one dict of stacked annotators, the batch form of the writer's input: the mask batch is ONE pinned
    # [64,H,W] array (config 2 has no keypoint stage, so no skeleton / depth annotator)
    host_frames = {
        "instance_segmentation": {"data": mask_host, "info": [fr["instance_segmentation"]["info"] for fr in frames]},
        "bounding_box_3d": {"data": [fr["bounding_box_3d"]["data"] for fr in frames],
                            "info": [fr["bounding_box_3d"]["info"] for fr in frames]},
        "camera_pose": np.asarray([fr["camera_pose"] for fr in frames], dtype=np.float64),
        "camera_params": [fr["camera_params"] for fr in frames],
        "frame_id": rank * BATCH,
    }
    writer = ConstructionLabelWriter(None, device=dev, split_people=True)
    for _ in range(2):
        writer.annotate_batch(host_frames).synchronize()
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    d2h = 0
    # two batches in flight, like write_batch's queue (max_pending = 2): the host work of step k+1
    # (tables, enqueueing the copies) overlaps the PCIe transfer of step k; every step still copies its
    # inputs from pinned host memory and has its records read back on the host
    in_flight = None
    emitted = 0
    for _ in range(e2e_steps):
        labels = writer.annotate_batch(host_frames)
        if in_flight is not None:
            emitted += int(in_flight.n_out.sum())          # synchronises on that batch's event
        in_flight = labels
        d2h = labels._rec_host.numel() + labels._nout_host.numel() * 4
    emitted += int(in_flight.n_out.sum())
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(te.item())
    h2d = mask_host.numel() * 4 + lut.nbytes + obj_record.nbytes + slot_class.nbytes + records.nbytes + cam.nbytes
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, N = 1 only): numpy oracle on the host cores, bounded sample ---------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = max(cores, min(BATCH, 2 * cores))
        fps = cpu_throughput(frames[:sample], cores, reps=2)
        fps1 = cpu_throughput(frames[:4], 1, reps=1)
        # the CPU leg doubles as the checker of the GPU arm: frame 0 of the timed batch, record for record
        from tests import helpers

        want = helpers.oracle_pipeline(frames[:1], frame_base=0)
        helpers.assert_records_equal(gpu_frame0, want["recs"][0, : want["n_out"][0]])
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} of the step's {BATCH} frames through the numpy oracle pipeline "
                         f"(bincount+find_objects scan, per-object projection, emission), {cores} processes; "
                         f"1 core: {fps1:.2f} frames/s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_frames_per_gpu": BATCH, "resolution": f"{W}x{H}", "instances": N,
                       "unique_frames": uniq, "l2": "inputs (531 MB mask batch per step) larger than the 126 MB L2",
                       "parallelism": f"frames sharded, {world} rank(s), no data-path collective",
                       "step": "CUDA graph replay" if args.graph else "3 eager launches per step (PDL-chained)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "mask_scan_kernel", "ms_per_launch": scan_ms,
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "ConstructionLabelWriter.annotate_batch(stacked host annotator dict), 2 batches in flight",
                    "h2d_gbs_effective": h2d * e2e_steps / float(te.item()) / 1e9},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step as a CUDA graph (default: eager launches, which keep the programmatic "
                         "dependent-launch overlap between the kernels and measured faster)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
