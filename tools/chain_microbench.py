"""Which part of the K1 -> K2 -> K4 chain is exposed?  python tools/chain_microbench.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.pipeline import LabelPipeline
from constructionsceneposeestimation_b200.sweep import build_host_tables
dev = torch.device("cuda")
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
lut, obj_record, slot_class, records, cam, _ = build_host_tables(frames)
lut = np.pad(lut, ((0, 0), (0, (-lut.shape[1]) % 4)), constant_values=-1)
B = 64; H, W = frames[0]["instance_segmentation"]["data"].shape; N = obj_record.shape[1]
pipe = LabelPipeline(B, H, W, N, records.shape[1], lut.shape[1], dev, use_graph=False)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
tile = lambda a: t(a).repeat((8,) + (1,) * (a.ndim - 1))
pipe.mask.copy_(tile(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)))
pipe.lut.copy_(tile(lut)); pipe.obj_record.copy_(tile(obj_record)); pipe.slot_class.copy_(tile(slot_class))
pipe.records_in.copy_(tile(records.view(np.uint8).reshape(8, records.shape[1], -1))); pipe.cam.copy_(tile(cam))
pipe.run(); torch.cuda.synchronize()
lib = pipe.lib; s = torch.cuda.current_stream().cuda_stream
def scan(): lib.cspe_mask_scan(pipe.mask.data_ptr(), B, H, W, pipe.lut.data_ptr(), pipe.L, pipe.lut_stride, N, pipe.scan.data_ptr(), s)
def scan_acc(): lib.cspe_mask_scan_accumulate(pipe.mask.data_ptr(), B, H, W, pipe.lut.data_ptr(), pipe.L, pipe.lut_stride, N, pipe.scan.data_ptr(), s)
def k2(fn): fn(pipe.records_in.data_ptr(), 96, pipe.R, pipe.obj_record.data_ptr(), pipe.cam.data_ptr(), B, N, pipe.uv.data_ptr(), pipe.z.data_ptr(), pipe.pose.data_ptr(), pipe.loose.data_ptr(), pipe.flags.data_ptr(), s)
def emit(): lib.cspe_emit(pipe.scan.data_ptr(), pipe.uv.data_ptr(), pipe.z.data_ptr(), pipe.pose.data_ptr(), pipe.loose.data_ptr(), pipe.flags.data_ptr(), pipe.slot_class.data_ptr(), B, N, H, W, 1, 0, pipe.records.data_ptr(), pipe.n_out.data_ptr(), pipe.class_hist.data_ptr(), s)
def timed(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1000, 2)
def graphed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g.replay
cases = {
  "scan_only(no init)": scan_acc,
  "init+scan": scan,
  "init+scan+k2_overlapped": lambda: (scan(), k2(lib.cspe_project_objects_overlapped)),
  "init+scan+k2_serial": lambda: (scan(), k2(lib.cspe_project_objects)),
  "init+scan+emit": lambda: (scan(), emit()),
  "init+scan+k2_overlapped+emit": lambda: (scan(), k2(lib.cspe_project_objects_overlapped), emit()),
}
out = {}
for k, fn in cases.items():
    out[k] = {"eager_us": timed(fn), "graph_us": timed(graphed(fn))}
print(json.dumps(out, indent=1))
