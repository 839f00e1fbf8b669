"""Kernel-level timing of the "next" rows (f1 point cloud, f2 depth stats, f4 depth colormap) through the
C ABI with preallocated buffers (no Python allocation in the loop).  python tools/next_rows_bench.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import _lib, camera, synthetic
import cv2, io
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
lut = torch.from_numpy(np.ascontiguousarray(cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3))).to(dev)
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
# ---- f1 batched: all 64 frames in one pair of launches (5 GB of points: nothing stays in L2) ----------
try:
    PEAK = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass
camB = torch.from_numpy(np.stack([camera.pack_camera(frames[i % 8]["camera_pose"], frames[i % 8]["camera_params"])
                                  for i in range(B)])).to(dev)
outB = torch.empty((B * H * W, 6), dtype=torch.float64, device=dev)
offB = torch.empty(B + 1, dtype=torch.int64, device=dev)
wsB = torch.empty((lib.cspe_pointcloud_batch_workspace_bytes(B, H, W) + 7) // 8, dtype=torch.int64, device=dev)
for label, rgb_ptr, ch in (("RGBA", rgb.data_ptr(), 4), ("no colour", None, 0)):
    ms = timed(lambda: lib.cspe_depth_to_pointcloud_batch(depth.data_ptr(), rgb_ptr, ch, B, H, W, camB.data_ptr(), outB.data_ptr(),
                                                          B * H * W, offB.data_ptr(), wsB.data_ptr(), s), n=10)
    nptsB = int(offB[-1].item())
    bytsB = B * H * W * (4 + ch) * 2 + nptsB * 48      # depth (+ colour) read by both passes, 48 B per point written
    print(json.dumps({"case": f"f1 depth_to_pointcloud_batch 64 x 1080p, {label}", "ms": round(ms, 4), "ms_per_frame": round(ms / B, 5),
                      "points": nptsB, "algorithmic_GB/s": round(bytsB / ms / 1e6, 1), "frac": round(bytsB / ms / 1e6 / PEAK, 3)}))
del outB
# ---- f3: "%.6f" text of the point cloud and of the depth map (gcd.py:1752, 1688) --------------------
import time
tws = torch.empty((lib.cspe_text_workspace_bytes(H * W, 6) + 7) // 8, dtype=torch.int64, device=dev)
text = torch.empty((npts * 6 * 14 + 64,), dtype=torch.uint8, device=dev); nb = torch.empty(1, dtype=torch.int64, device=dev)
ms = timed(lambda: lib.cspe_format_fixed6(out.data_ptr(), 1, H * W, n.data_ptr(), 6, b"x y z r g b", text.data_ptr(), text.numel(),
                                          nb.data_ptr(), 0, None, tws.data_ptr(), s))
size = int(nb.item())
sample = out[:20000].cpu().numpy()
def savetxt(a, header=None):   # the reference's own calls, gcd.py:1688 / 1752
    b = io.BytesIO()
    np.savetxt(b, a, delimiter=" ", fmt="%.6f") if header is None else np.savetxt(b, a, fmt="%.6f", delimiter=" ", header=header, comments="")
    return b.getvalue()
t0 = time.perf_counter(); ref = savetxt(sample, "x y z r g b"); cpu_s = (time.perf_counter() - t0) * npts / len(sample)
assert text[: len(ref)].cpu().numpy().tobytes() == ref
print(json.dumps({"case": "f3 point-cloud text, one 1080p frame (np.savetxt bytes)", "ms": round(ms, 4), "points": npts,
                  "text_MB": round(size / 1e6, 1), "text_GB/s": round(size / ms / 1e6, 1),
                  "values_per_s": round(npts * 6 / ms * 1e3), "numpy_savetxt_s_per_frame_extrapolated": round(cpu_s, 2)}))
Bc = 8
dcsv = depth[:Bc].contiguous().view(Bc * H, W)
tws2 = torch.empty((lib.cspe_text_workspace_bytes(Bc * H, W) + 7) // 8, dtype=torch.int64, device=dev)
text2 = torch.empty((Bc * H * W * 14 + 64,), dtype=torch.uint8, device=dev)
split = torch.empty((Bc,), dtype=torch.int64, device=dev)
ms = timed(lambda: lib.cspe_format_fixed6(dcsv.data_ptr(), 0, Bc * H, None, W, None, text2.data_ptr(), text2.numel(), nb.data_ptr(),
                                          H, split.data_ptr(), tws2.data_ptr(), s))
size = int(nb.item())
t0 = time.perf_counter(); ref = savetxt(depth[0, :64].cpu().numpy()); cpu_s = (time.perf_counter() - t0) * H / 64
assert text2[: len(ref)].cpu().numpy().tobytes() == ref
print(json.dumps({"case": f"f3 depth CSV text, {Bc} x 1080p frames in one call", "ms": round(ms, 4), "ms_per_frame": round(ms / Bc, 4),
                  "text_MB_per_frame": round(size / Bc / 1e6, 1), "text_GB/s": round(size / ms / 1e6, 1),
                  "values_per_s": round(Bc * H * W / ms * 1e3), "numpy_savetxt_s_per_frame_extrapolated": round(cpu_s, 2)}))
