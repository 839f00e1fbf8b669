"""Small end-to-end pass for compute-sanitizer memcheck (tiny shapes, every kernel once)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import ops, synthetic, camera
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter
from tests import helpers

spec = synthetic.SceneSpec(320, 180, 14, 3, 17, config_id=8, with_rgb=True)
frames = synthetic.make_batch(spec, 3)
w = ConstructionLabelWriter(None, split_people=True)
labels = w.annotate_batch(frames).synchronize()
want = helpers.oracle_pipeline(frames)
assert np.array_equal(labels.n_out, want["n_out"])
# odd width (non-TMA producer), fused depth stats, point cloud
m = torch.randint(0, 9, (2, 37, 131), dtype=torch.int32, device="cuda")
lut = torch.tensor([-1, -1, 0, 1, 2, 3, 4, 5, 6], dtype=torch.int32, device="cuda")
ops.mask_scan(m, lut, 7)
d = torch.from_numpy(np.stack([f["distance_to_image_plane"] for f in frames])).cuda()
mm = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).cuda()
ops.mask_scan_depth_stats(mm, d, torch.from_numpy(want["lut"]).cuda(), want["obj_record"].shape[1])
ops.depth_stats(d)
cam = torch.from_numpy(camera.pack_camera(frames[0]["camera_pose"], frames[0]["camera_params"])).cuda()
ops.depth_to_pointcloud(d[0].contiguous(), torch.from_numpy(frames[0]["rgb"]).cuda(), cam)
torch.cuda.synchronize()
print("sanitize_small ok")
