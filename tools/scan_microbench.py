"""Mask-scan kernel timing on different mask contents (B200).  Usage: python tools/scan_microbench.py [reps]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import ops, synthetic
from constructionsceneposeestimation_b200.sweep import build_host_tables

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["zeros", "c2", "noise16", "c4"]
dev = torch.device("cuda")
PEAK = 6454.3

def timeit(mask, lut, N, label):
    if lut.shape[-1] % 4:   # 16-byte LUT rows: the per-stage bulk copy of the LUT needs them
        lut = torch.nn.functional.pad(lut, (0, 4 - lut.shape[-1] % 4), value=-1).contiguous()
    out = ops.mask_scan(mask, lut, N)
    for _ in range(3):
        ops.mask_scan(mask, lut, N, out=out, accumulate=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.mask_scan(mask, lut, N, out=out, accumulate=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = mask.numel() * 4 / (ms * 1e-3) / 1e9
    print(json.dumps({"case": label, "shape": list(mask.shape), "ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / PEAK, 3)}), flush=True)

if "zeros" in which:
    m = torch.zeros((64, 1080, 1920), dtype=torch.int32, device=dev)
    timeit(m, torch.full((128,), -1, dtype=torch.int32, device=dev), 100, "zeros 64x1080p")
    del m
if "c2" in which:
    import os
    uniq = int(os.environ.get("CSPE_MICRO_UNIQUE", "16"))
    frames = synthetic.make_batch(synthetic.CONFIGS["c2"], uniq)
    lut, obj_record, *_ = build_host_tables(frames)
    m = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).to(dev).repeat(64 // uniq, 1, 1)
    l = torch.from_numpy(lut).to(dev).repeat(64 // uniq, 1)
    timeit(m, l, obj_record.shape[1], "synthetic c2 64x1080p")
    if "c2nolut" in which:   # same pixels, every id unmapped: the run decomposition without any table merge
        timeit(m, torch.full_like(l, -1), obj_record.shape[1], "synthetic c2 64x1080p, all ids unmapped")
    del m
if "noise16" in which:
    # 16x16-pixel blocks of random ids: many short runs
    g = torch.randint(2, 102, (64, 68, 120), device=dev, dtype=torch.int32)
    m = g.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :1080, :].contiguous()
    l = torch.arange(-2, 126, dtype=torch.int32, device=dev).clamp(min=-1)
    timeit(m, l, 100, "16x16 blocks 64x1080p")
    del m, g
if "c4" in which:
    frames = synthetic.make_batch(synthetic.CONFIGS["c4"], 4)
    lut, obj_record, *_ = build_host_tables(frames)
    m = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).to(dev).repeat(4, 1, 1)
    l = torch.from_numpy(lut).to(dev).repeat(4, 1)
    timeit(m, l, obj_record.shape[1], "synthetic c4 16x2160p")
