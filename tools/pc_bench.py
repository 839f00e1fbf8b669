"""f1 batched point cloud only: python tools/pc_bench.py [reps]   (64 x 1080p, RGBA; pass 1 alone via capacity = 0)"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import _lib, camera, synthetic
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda")
lib = _lib.load()
try:
    PEAK = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except (OSError, ValueError, KeyError):
    PEAK = 6454.3
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
depth = torch.from_numpy(np.stack([f["distance_to_image_plane"] for f in frames])).to(dev).repeat(8, 1, 1).contiguous()
B, H, W = depth.shape
rgb = torch.randint(0, 256, (B, H, W, 4), dtype=torch.uint8, device=dev)
cam = torch.from_numpy(np.stack([camera.pack_camera(frames[i % 8]["camera_pose"], frames[i % 8]["camera_params"]) for i in range(B)])).to(dev)
out = torch.empty((B * H * W, 6), dtype=torch.float64, device=dev)
off = torch.empty(B + 1, dtype=torch.int64, device=dev)
ws = torch.empty((lib.cspe_pointcloud_batch_workspace_bytes(B, H, W) + 7) // 8, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
def timed(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for label, cap in (("both passes", B * H * W), ("pass 1 only (capacity 0)", 0)):
    ms = timed(lambda: lib.cspe_depth_to_pointcloud_batch(depth.data_ptr(), rgb.data_ptr(), 4, B, H, W, cam.data_ptr(), out.data_ptr(),
                                                          cap, off.data_ptr(), ws.data_ptr(), s), reps)
    npts = int(off[-1].item())
    byts = B * H * W * 8 * (2 if cap else 1) + (npts * 48 if cap else 0)
    print(json.dumps({"case": f"f1 batch 64 x 1080p RGBA, {label}", "ms": round(ms, 4), "points": npts,
                      "algorithmic_GB/s": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)
# ---- what the HBM does on a pure WRITE stream and on a pure READ stream (the copy peak mixes both 50 / 50) --------
buf = torch.empty(1 << 32, dtype=torch.uint8, device=dev)            # 4 GiB
ms_w = timed(lambda: buf.fill_(7), 10)
src = buf.view(torch.int64)
ms_r = timed(lambda: src.sum(), 10)
half = buf[: 1 << 31]
ms_c = timed(lambda: buf[1 << 31:].copy_(half), 10)
print(json.dumps({"case": "HBM probes (torch kernels, 4 GiB)", "fill_GB/s": round((1 << 32) / ms_w / 1e6, 1),
                  "sum_read_GB/s": round((1 << 32) / ms_r / 1e6, 1), "copy_read+write_GB/s": round((1 << 32) / ms_c / 1e6, 1),
                  "measured_copy_peak": PEAK}), flush=True)
