import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
dbg = torch.zeros(296 * 2, dtype=torch.int64, device="cuda")
os.environ["CSPE_DBG_PTR"] = str(dbg.data_ptr())
from constructionsceneposeestimation_b200 import ops, synthetic
from tests import helpers
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 16)
lut, obj_record, *_ = helpers.host_tables(frames)
m = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).cuda().repeat(4, 1, 1)
l = torch.from_numpy(lut).cuda().repeat(4, 1)
N = obj_record.shape[1]
out = ops.mask_scan(m, l, N)
for _ in range(3):
    ops.mask_scan(m, l, N, out=out, accumulate=True)
torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(296, 2).astype(np.float64)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
dur = end - start
print(json.dumps({"start_us": [round(float(np.percentile(start, q)), 1) for q in (0, 50, 100)],
                  "end_us": [round(float(np.percentile(end, q)), 1) for q in (0, 5, 25, 50, 75, 95, 100)],
                  "dur_us": [round(float(np.percentile(dur, q)), 1) for q in (0, 50, 100)], "mean_end": round(float(end.mean()), 1)}))
