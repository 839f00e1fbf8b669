"""Kernel-level timing of the "next" rows (f1 point cloud, f2 depth stats, f4 depth colormap) through the
C ABI with preallocated buffers (no Python allocation in the loop).  python tools/next_rows_bench.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import _lib, camera, synthetic
from oracle import labels as O
dev = torch.device("cuda"); PEAK = 6454.3
lib = _lib.load()
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
depth = torch.from_numpy(np.stack([f["distance_to_image_plane"] for f in frames])).to(dev).repeat(8, 1, 1).contiguous()
B, H, W = depth.shape
s = torch.cuda.current_stream().cuda_stream
def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
stats = torch.empty((B, 48), dtype=torch.uint8, device=dev)
ms = timed(lambda: lib.cspe_depth_stats(depth.data_ptr(), B, H, W, stats.data_ptr(), s))
print(json.dumps({"case": "f2 depth_stats 64x1080p (3 launches: init, reduce, finalize)", "ms": round(ms, 4),
                  "GB/s": round(B * H * W * 4 / ms / 1e6, 1), "frac": round(B * H * W * 4 / ms / 1e6 / PEAK, 3)}))
lut = torch.from_numpy(np.ascontiguousarray(O.jet_lut_bgr())).to(dev)
img = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
ms = timed(lambda: lib.cspe_depth_colormap(depth.data_ptr(), B, H, W, stats.data_ptr(), lut.data_ptr(), img.data_ptr(), s))
print(json.dumps({"case": "f4 depth_colormap 64x1080p", "ms": round(ms, 4), "GB/s": round(B * H * W * 7 / ms / 1e6, 1),
                  "frac": round(B * H * W * 7 / ms / 1e6 / PEAK, 3)}))
rgb = torch.randint(0, 256, (B, H, W, 4), dtype=torch.uint8, device=dev)
bgr = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
ms = timed(lambda: lib.cspe_rgb_to_bgr(rgb.data_ptr(), 4, B * H * W, bgr.data_ptr(), s))
print(json.dumps({"case": "f4 rgba_to_bgr 64x1080p", "ms": round(ms, 4), "GB/s": round(B * H * W * 7 / ms / 1e6, 1),
                  "frac": round(B * H * W * 7 / ms / 1e6 / PEAK, 3)}))
fr = frames[0]
cam = torch.from_numpy(camera.pack_camera(fr["camera_pose"], fr["camera_params"])).to(dev)
d0 = depth[0].contiguous(); rgb0 = rgb[0].contiguous()
out = torch.empty((H * W, 6), dtype=torch.float64, device=dev); n = torch.empty(1, dtype=torch.int64, device=dev)
ws = torch.empty((lib.cspe_pointcloud_workspace_bytes(H, W) + 7) // 8, dtype=torch.int64, device=dev)
ms = timed(lambda: lib.cspe_depth_to_pointcloud(d0.data_ptr(), rgb0.data_ptr(), 4, H, W, cam.data_ptr(), out.data_ptr(), H * W,
                                                n.data_ptr(), ws.data_ptr(), s))
npts = int(n.item()); byts = H * W * 4 + H * W * 4 + npts * 48   # depth + rgba read once, points written
print(json.dumps({"case": "f1 depth_to_pointcloud one 1080p frame", "ms": round(ms, 4), "points": npts,
                  "algorithmic_GB/s": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / PEAK, 3)}))
