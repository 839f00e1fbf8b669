"""Config 1 (one 1280x720 frame, ~20 instances): latency of one pass of the hot path, and of its kernels alone.

    python tools/c1_latency.py [reps]

Lines: the three-kernel chain as ONE CUDA-graph replay and as eager launches (back to back, CUDA events), then each
kernel on its own (scan = cspe_mask_scan_accumulate, project = cspe_project_objects, emit = cspe_emit_reset_scan).
Run it under `ncu --metrics gpu__time_duration.sum` for the per-launch device times.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from constructionsceneposeestimation_b200 import ops, synthetic
from constructionsceneposeestimation_b200.pipeline import LabelPipeline
from constructionsceneposeestimation_b200.sweep import build_host_tables

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda")


def timed(fn, n=reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        best = ms if best is None else min(best, ms)
    return best


def build(frames, B, use_graph):
    lut, obj_record, slot_class, records, cam, _ = build_host_tables(frames)
    lut = np.pad(lut, ((0, 0), (0, (-lut.shape[1]) % 4)), constant_values=-1)
    H, W = frames[0]["instance_segmentation"]["data"].shape
    N = obj_record.shape[1]
    pipe = LabelPipeline(B, H, W, N, records.shape[1], lut.shape[1], dev, use_graph=use_graph)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pipe.mask.copy_(t(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)))
    pipe.lut.copy_(t(lut))
    pipe.obj_record.copy_(t(obj_record))
    pipe.slot_class.copy_(t(slot_class))
    pipe.records_in.copy_(t(records.view(np.uint8).reshape(B, records.shape[1], -1)))
    pipe.cam.copy_(t(cam))
    return pipe, N


frames = synthetic.make_batch(synthetic.CONFIGS["c1"], 1)
for g in (True, False):
    pipe, N = build(frames, 1, g)
    ms = timed(pipe.run)
    print(json.dumps({"case": "c1 chain, " + ("one CUDA-graph replay" if g else "eager launches"), "us": round(ms * 1e3, 2),
                      "frames_per_s": round(1000 / ms)}), flush=True)
pipe, N = build(frames, 1, False)
H, W = pipe.H, pipe.W
scan_us = timed(lambda: ops.mask_scan(pipe.mask, pipe.lut, N, out=pipe.scan, accumulate=True)) * 1e3
proj_us = timed(lambda: ops.project_objects(pipe.records_in, pipe.obj_record, pipe.cam)) * 1e3
pipe.run()
emit_us = timed(lambda: ops.emit(pipe.scan, pipe.uv, pipe.z, pipe.pose, pipe.loose, pipe.flags, pipe.slot_class, H, W, 1, 0)) * 1e3
print(json.dumps({"case": "c1 kernels alone (eager, back to back, incl. their launch gaps and torch allocations)",
                  "scan_us": round(scan_us, 2), "project_us": round(proj_us, 2), "emit_us": round(emit_us, 2)}), flush=True)
