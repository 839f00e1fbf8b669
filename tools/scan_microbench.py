"""Mask-scan kernel timing on different mask contents (B200).

    python tools/scan_microbench.py [reps] [cases] [variants]

cases: comma list of zeros,c2,c2nolut,c2_dense,c2_textured,blocks16,checker2,noise2,noise100,c4
variants: comma list of FLUSH:SLOW pairs for the kernel's A/B switches (CSPE_SCAN_FLUSH / CSPE_SCAN_SLOW,
read per launch), e.g. "1:1,0:0"; default = the library default only.
One JSON line per (case, variant): CUDA-event time per launch over `reps` launches, algorithmic GB/s
(4*H*W bytes per frame) and the fraction of the measured HBM copy peak.
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from constructionsceneposeestimation_b200 import ops, synthetic
from constructionsceneposeestimation_b200.sweep import build_host_tables

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["zeros", "c2", "c2_dense", "c2_textured", "blocks16", "checker2",
                                                          "noise2", "noise100", "c4"]
variants = [tuple(v.split(":")) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [None]
dev = torch.device("cuda")
try:
    PEAK = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except (OSError, ValueError, KeyError):
    PEAK = 6454.3


def timeit(mask, lut, N, label):
    if lut.shape[-1] % 4:   # 16-byte LUT rows
        lut = torch.nn.functional.pad(lut, (0, 4 - lut.shape[-1] % 4), value=-1).contiguous()
    for var in variants:
        if var is not None:
            os.environ["CSPE_SCAN_FLUSH"], os.environ["CSPE_SCAN_SLOW"] = var
        out = ops.mask_scan(mask, lut, N)
        for _ in range(3):
            ops.mask_scan(mask, lut, N, out=out, accumulate=True)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):   # best of 3 windows of `reps` launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.mask_scan(mask, lut, N, out=out, accumulate=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            best = ms if best is None else min(best, ms)
        gbs = mask.numel() * 4 / (best * 1e-3) / 1e9
        print(json.dumps({"case": label, "variant": "default" if var is None else f"flush{var[0]}/slow{var[1]}",
                          "shape": list(mask.shape), "ms": round(best, 4), "GB/s": round(gbs, 1),
                          "frac": round(gbs / PEAK, 3)}), flush=True)


def synthetic_case(key, frames_total, uniq):
    frames = synthetic.make_batch(synthetic.CONFIGS[key], uniq)
    lut, obj_record, *_ = build_host_tables(frames)
    m = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).to(dev)
    m = m.repeat(frames_total // uniq, 1, 1)
    l = torch.from_numpy(lut).to(dev).repeat(frames_total // uniq, 1)
    return m, l, obj_record.shape[1]


B, H, W = 64, 1080, 1920
uniq = int(os.environ.get("CSPE_MICRO_UNIQUE", "16"))
flat_lut = torch.arange(-2, 126, dtype=torch.int32, device=dev).clamp(min=-1)   # ids 2..101 -> slots 0..99

if "zeros" in which:
    m = torch.zeros((B, H, W), dtype=torch.int32, device=dev)
    timeit(m, torch.full((128,), -1, dtype=torch.int32, device=dev), 100, "zeros 64x1080p")
    del m
for key in ("c2", "c2_dense", "c2_textured"):
    if key in which:
        m, l, N = synthetic_case(key, B, uniq)
        timeit(m, l, N, f"synthetic {key} 64x1080p")
        if key == "c2" and "c2nolut" in which:   # same pixels, every id unmapped: run decomposition without table merges
            timeit(m, torch.full_like(l, -1), N, "synthetic c2 64x1080p, all ids unmapped")
        del m, l
if "blocks16" in which or "noise16" in which:
    # 16x16-pixel blocks of random ids: two ids per 32-px strip row, a new pair every 16 rows
    g = torch.randint(2, 102, (B, 68, 120), device=dev, dtype=torch.int32)
    m = g.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :].contiguous()
    timeit(m, flat_lut, 100, "16x16 blocks 64x1080p")
    del m, g
if "checker2" in which:
    # two ids alternating at every pixel: the worst see-through texture (32 runs per strip row, 2 ids)
    ys = torch.arange(H, device=dev)[:, None]
    xs = torch.arange(W, device=dev)[None, :]
    m = torch.where(((xs + ys) & 1).bool(), 7, 42).to(torch.int32)[None].repeat(B, 1, 1).contiguous()
    timeit(m, flat_lut, 100, "checkerboard of 2 ids 64x1080p")
    del m
if "noise2" in which:
    m = torch.randint(4, 6, (B, H, W), device=dev, dtype=torch.int32)
    timeit(m, flat_lut, 100, "per-pixel noise, 2 ids 64x1080p")
    del m
if "noise100" in which:
    m = torch.randint(2, 102, (B, H, W), device=dev, dtype=torch.int32)
    timeit(m, flat_lut, 100, "per-pixel noise, 100 ids 64x1080p")
    del m
if "c4" in which:
    m, l, N = synthetic_case("c4", 16, 4)
    timeit(m, l, N, "synthetic c4 16x2160p")
