"""Time K2 / K4 alone on config-2 shaped tables (B200): python tools/emit_microbench.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import synthetic, _lib
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
def emit():
    lib.cspe_emit(pipe.scan.data_ptr(), pipe.uv.data_ptr(), pipe.z.data_ptr(), pipe.pose.data_ptr(), pipe.loose.data_ptr(),
                  pipe.flags.data_ptr(), pipe.slot_class.data_ptr(), B, N, H, W, 1, 0, pipe.records.data_ptr(),
                  pipe.n_out.data_ptr(), pipe.class_hist.data_ptr(), s)
def project():
    lib.cspe_project_objects(pipe.records_in.data_ptr(), 96, pipe.R, pipe.obj_record.data_ptr(), pipe.cam.data_ptr(), B, N,
                             pipe.uv.data_ptr(), pipe.z.data_ptr(), pipe.pose.data_ptr(), pipe.loose.data_ptr(),
                             pipe.flags.data_ptr(), s)
def init_scan():
    lib.cspe_mask_scan(pipe.mask.data_ptr(), B, H, W, pipe.lut.data_ptr(), pipe.L, pipe.lut_stride, N, pipe.scan.data_ptr(), s)
def timed(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print(json.dumps({"emit_us": round(timed(emit), 2), "project_us": round(timed(project), 2), "init+scan_us": round(timed(init_scan, 50), 2),
                  "pipeline_eager_us": round(timed(pipe.run, 50), 2)}))
pg = LabelPipeline.__new__(LabelPipeline); pg.__dict__.update(pipe.__dict__); pg.use_graph = True; pg.graph = None
print(json.dumps({"pipeline_graph_us": round(timed(pg.run, 100), 2)}))

# when does K2 finish relative to K1 when both are in flight?
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
for order in ("scan_first", "project_first"):
    res = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0, t_scan, t_proj = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record(main)
        side.wait_event(t0)
        def k1():
            lib.cspe_mask_scan(pipe.mask.data_ptr(), B, H, W, pipe.lut.data_ptr(), pipe.L, pipe.lut_stride, N, pipe.scan.data_ptr(), main.cuda_stream)
        def k2():
            lib.cspe_project_objects(pipe.records_in.data_ptr(), 96, pipe.R, pipe.obj_record.data_ptr(), pipe.cam.data_ptr(), B, N,
                                     pipe.uv.data_ptr(), pipe.z.data_ptr(), pipe.pose.data_ptr(), pipe.loose.data_ptr(),
                                     pipe.flags.data_ptr(), side.cuda_stream)
        if order == "scan_first":
            k1(); k2()
        else:
            k2(); k1()
        t_scan.record(main); t_proj.record(side)
        torch.cuda.synchronize()
        res.append((t0.elapsed_time(t_scan) * 1000, t0.elapsed_time(t_proj) * 1000))
    a = np.median(np.array(res), axis=0)
    print(json.dumps({"order": order, "scan_done_us": round(float(a[0]), 1), "project_done_us": round(float(a[1]), 1)}))
