"""Static-shape batch pipeline: K1 (mask scan) || K2 (projection/pose) [|| K3] -> K4 (emission).

All buffers are allocated once; one call to :meth:`LabelPipeline.run` enqueues the whole batch on
one stream.  K2 does not depend on K1: it is launched with programmatic stream serialisation and
K1 releases its dependents as soon as its persistent CTAs are resident, so K2 runs in the SM
resources K1 leaves free; the next batch's K1 starts streaming while this batch's K4 still copies
records.  :meth:`run_steps` captures a chain of k such steps into ONE CUDA graph — stream capture
keeps the programmatic edges between consecutive kernels (``cspe_graph_edge_kinds`` counts them), so
the replay has the overlap of the eager chain without its host launches.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import CAM_STRIDE, NUM_CLASSES, POSE_STRIDE, RECORD_DTYPE, SCAN_FIELDS


def graph_edge_kinds(graph: torch.cuda.CUDAGraph) -> Dict[str, int]:
    """{nodes, edges, programmatic} of a captured graph (needs ``CUDAGraph(keep_graph=True)``)."""
    n, e, p = C.c_int(0), C.c_int(0), C.c_int(0)
    _lib.check("cspe_graph_edge_kinds", _lib.load().cspe_graph_edge_kinds(
        graph.raw_cuda_graph(), C.byref(n), C.byref(e), C.byref(p)))
    return {"nodes": n.value, "edges": e.value, "programmatic": p.value}


class LabelPipeline:
    def __init__(self, batch: int, height: int, width: int, num_slots: int, recs_per_frame: int, lut_len: int,
                 device: torch.device, per_frame_lut: bool = True,  # lut_len: use a multiple of 4 (16-byte rows)
                 min_pixels: int = 1, use_graph: bool = True,
                 mask: Optional[torch.Tensor] = None, num_people: int = 0, num_joints: int = 0,
                 keypoint_tolerance: float = 0.15, overlapped: Optional[bool] = None):
        self.lib = _lib.load()
        self.B, self.H, self.W, self.N, self.R, self.L = batch, height, width, num_slots, recs_per_frame, lut_len
        self.device = torch.device(device)
        self.min_pixels = int(min_pixels)
        self.frame_base = 0
        dev = self.device
        i32, f64, u8 = torch.int32, torch.float64, torch.uint8
        B, N = batch, num_slots
        # The cross-batch overlap (`*_overlapped` entry points + double-buffered K2 / K3 outputs) is only
        # ordered correctly when the scan launches its full persistent grid: the scan of batch i+2 then
        # cannot become resident — and release K2(i+2), which rewrites the buffers K4(i) reads — before
        # every CTA of scan i+1 has exited, i.e. has waited for K4(i).  Small batches use the plain,
        # stream-ordered entry points (the few microseconds they serialise do not matter there).
        fills = bool(self.lib.cspe_mask_scan_fills_device(batch, height, width))
        self.overlapped = fills if overlapped is None else bool(overlapped and fills)
        self.mask = mask if mask is not None else torch.empty((B, height, width), dtype=i32, device=dev)
        if tuple(self.mask.shape) != (B, height, width) or not self.mask.is_contiguous():
            raise ValueError("mask must be a contiguous [B,H,W] tensor")
        self.lut = torch.full((B if per_frame_lut else 1, lut_len), -1, dtype=i32, device=dev)
        self.lut_stride = lut_len if per_frame_lut else 0
        self.obj_record = torch.full((B, N), -1, dtype=i32, device=dev)
        self.slot_class = torch.full((B, N), -1, dtype=i32, device=dev)
        self.records_in = torch.zeros((B, recs_per_frame, _lib.BBOX3D_RECORD_BYTES), dtype=u8, device=dev)
        self.cam = torch.zeros((B, CAM_STRIDE), dtype=f64, device=dev)
        # scan table starts as "no instance seen" (what cspe_mask_scan's init writes); K4 restores it per batch
        self.scan = torch.tensor([0, width, height, -1, -1], dtype=i32, device=dev).repeat(B, N, 1).contiguous()
        # K2's outputs are double-buffered: batch i+1's K2 starts beside batch i+1's scan, which itself
        # overlaps batch i's K4 — K4(i) must still be able to read batch i's projections
        self._k2 = [dict(uv=torch.empty((B, N, 8, 2), dtype=f64, device=dev),
                         z=torch.empty((B, N, 8), dtype=f64, device=dev),
                         pose=torch.empty((B, N, POSE_STRIDE), dtype=f64, device=dev),
                         loose=torch.empty((B, N, 4), dtype=f64, device=dev),
                         flags=torch.empty((B, N), dtype=u8, device=dev)) for _ in range(2)]
        self._parity = 0
        # optional K3 stage (config 3: skeleton keypoints + depth-buffer visibility)
        self.P, self.J, self.keypoint_tolerance = int(num_people), int(num_joints), float(keypoint_tolerance)
        self.joints = self.depth = None
        self._k3 = None
        if self.P > 0 and self.J > 0:
            self.joints = torch.zeros((B, self.P, self.J, 3), dtype=torch.float32, device=dev)
            self.depth = torch.full((B, height, width), float("inf"), dtype=torch.float32, device=dev)
            self._k3 = [dict(kp=torch.empty((B, self.P, self.J, 2), dtype=f64, device=dev),
                             kz=torch.empty((B, self.P, self.J), dtype=f64, device=dev),
                             vis=torch.empty((B, self.P, self.J), dtype=u8, device=dev)) for _ in range(2)]
        self.records = torch.empty((B, N, RECORD_DTYPE.itemsize), dtype=u8, device=dev)
        self.n_out = torch.empty((B,), dtype=i32, device=dev)
        self.class_hist = torch.zeros((NUM_CLASSES,), dtype=torch.int64, device=dev)
        # mask_scan (accumulate) + project_objects [+ keypoints] + emit (which re-initialises the scan table)
        self.launches_per_run = 3 + (1 if num_people > 0 and num_joints > 0 else 0)
        self.graphs = [None, None]  # one captured graph per K2 buffer parity
        self._step_graphs: Dict[Tuple[int, int], torch.cuda.CUDAGraph] = {}
        self.use_graph = use_graph

    # K2 outputs of the most recent run()
    @property
    def uv(self) -> torch.Tensor:
        return self._k2[self._parity ^ 1]["uv"]

    @property
    def z(self) -> torch.Tensor:
        return self._k2[self._parity ^ 1]["z"]

    @property
    def pose(self) -> torch.Tensor:
        return self._k2[self._parity ^ 1]["pose"]

    @property
    def loose(self) -> torch.Tensor:
        return self._k2[self._parity ^ 1]["loose"]

    @property
    def flags(self) -> torch.Tensor:
        return self._k2[self._parity ^ 1]["flags"]

    @property
    def keypoints(self):
        """(kp [B,P,J,2], kz [B,P,J], vis [B,P,J]) of the most recent run(), or None without a K3 stage."""
        if self._k3 is None:
            return None
        k3 = self._k3[self._parity ^ 1]
        return k3["kp"], k3["kz"], k3["vis"]

    # ---------------------------------------------------------------- enqueue
    def enqueue(self, parity: int, records: Optional[torch.Tensor] = None, n_out: Optional[torch.Tensor] = None,
                frame_base: Optional[int] = None, frame_base_dev: Optional[torch.Tensor] = None,
                frames: Optional[int] = None, first: int = 0) -> None:
        """scan -> project [-> keypoints] -> emit on the current stream, all chained by programmatic
        dependent launch.  The scan table is initialised once (constructor) and re-initialised by K4 as it
        reads it, so K1 runs in accumulate mode with no init launch.  In overlapped mode
          * K1 may start while the PREVIOUS batch's K4 is still running — K4 releases its dependents at
            entry — and waits for it only before its first merge into the table;
          * K2 (and K3) start as soon as K1's CTAs are resident and run in the SM resources K1 leaves
            free; they complete only after K1 does;
          * K4 waits for K2 (hence K1) and reads this batch's K2 buffers (double-buffered by ``parity``).
        ``records`` / ``n_out`` redirect K4's output (a sweep keeps one set per batch in flight);
        ``frame_base_dev`` (int32[1] on the device) adds a replay-time frame offset; ``frames`` < B runs a
        partial batch over the resident frames ``first .. first + frames`` (outputs start at row 0)."""
        lib, chk = self.lib, _lib.check
        k2 = self._k2[parity]
        B = self.B - first if frames is None else int(frames)
        if not (0 <= first and 0 < B and first + B <= self.B):
            raise ValueError(f"frames [{first}, {first + B}) outside the resident batch of {self.B}")
        ovl = self.overlapped and B == self.B
        records = self.records if records is None else records
        n_out = self.n_out if n_out is None else n_out
        frame_base = self.frame_base if frame_base is None else int(frame_base)
        main = torch.cuda.current_stream(self.device).cuda_stream
        j = int(first)   # input rows of the partial batch start at resident frame j

        def scan(fn):
            chk("cspe_mask_scan_accumulate", fn(
                self.mask.data_ptr() + j * self.H * self.W * 4, B, self.H, self.W,
                self.lut.data_ptr() + j * self.lut_stride * 4, self.L, self.lut_stride, self.N,
                self.scan.data_ptr(), main))

        def project(fn):
            chk("cspe_project_objects", fn(
                self.records_in.data_ptr() + j * self.R * _lib.BBOX3D_RECORD_BYTES, _lib.BBOX3D_RECORD_BYTES, self.R,
                self.obj_record.data_ptr() + j * self.N * 4, self.cam.data_ptr() + j * CAM_STRIDE * 8, B, self.N,
                k2["uv"].data_ptr(), k2["z"].data_ptr(), k2["pose"].data_ptr(), k2["loose"].data_ptr(),
                k2["flags"].data_ptr(), main))

        def keypoints(fn):
            k3 = self._k3[parity]
            chk("cspe_keypoints", fn(
                self.joints.data_ptr() + j * self.P * self.J * 12, B, self.P, self.J,
                self.depth.data_ptr() + j * self.H * self.W * 4, self.H, self.W,
                self.cam.data_ptr() + j * CAM_STRIDE * 8, self.keypoint_tolerance, k3["kp"].data_ptr(), k3["kz"].data_ptr(),
                k3["vis"].data_ptr(), main))

        if ovl:
            # full-size batch: K1 first (it may stream under the previous batch's K4), K2 / K3 beside it
            scan(lib.cspe_mask_scan_accumulate_overlapped)
            project(lib.cspe_project_objects_overlapped)
            if self._k3 is not None:  # K3 rides the same chain: beside the scan, done before K4
                keypoints(lib.cspe_keypoints_overlapped)
        else:
            # small batch (latency case, e.g. config 1 — one 720p frame): K2 [K3] FIRST as plain launches — they wait
            # at entry for whatever ran before (the previous batch's K4 still reading their output buffers) but
            # release their dependents before that wait — then K1 through its overlapped entry point: it streams
            # the mask beside K2's serial FP64 chain and waits for it (hence for everything before it) only
            # before its first merge into the scan table.  K4 waits for K1, which completes after K2 / K3.
            # Latency = max(K1, K2) + K4 instead of K1 + K2 + K4; nothing depends on the size of the grids.
            project(lib.cspe_project_objects)
            if self._k3 is not None:
                keypoints(lib.cspe_keypoints)
            scan(lib.cspe_mask_scan_accumulate_overlapped)
        common = (self.scan.data_ptr(), k2["uv"].data_ptr(), k2["z"].data_ptr(), k2["pose"].data_ptr(),
                  k2["loose"].data_ptr(), k2["flags"].data_ptr(), self.slot_class.data_ptr() + j * self.N * 4, B, self.N, self.H,
                  self.W, self.min_pixels, frame_base)
        if frame_base_dev is not None:
            chk("cspe_emit_reset_scan_indirect", lib.cspe_emit_reset_scan_indirect(
                *common, frame_base_dev.data_ptr(), records.data_ptr(), n_out.data_ptr(),
                self.class_hist.data_ptr(), main))
        else:
            chk("cspe_emit_reset_scan", lib.cspe_emit_reset_scan(
                *common, records.data_ptr(), n_out.data_ptr(), self.class_hist.data_ptr(), main))

    def _enqueue(self, parity: int) -> None:
        self.enqueue(parity)

    def run(self, frames: Optional[int] = None) -> None:
        """Enqueue one batch on the current stream (graph replay when enabled; ``frames`` < B = an eager
        partial batch).

        Inputs (mask, lut, obj_record, slot_class, records_in, cam) must not be rewritten by a kernel
        queued directly before run(): the overlapped launches read them without waiting for it.
        Rewrite them with copies / from another stream's event, or synchronise first."""
        parity = self._parity
        self._parity ^= 1
        with torch.cuda.device(self.device):
            if not self.use_graph or (frames is not None and frames != self.B):
                self.enqueue(parity, frames=frames)
                return
            if self.graphs[parity] is None:
                self.graphs[parity] = self._capture(lambda: self.enqueue(parity))
            self.graphs[parity].replay()

    def _capture(self, body) -> torch.cuda.CUDAGraph:
        """Capture ``body`` (which must leave the class histogram accumulated exactly once per replay)."""
        saved = self.class_hist.clone()
        body()  # warm-up outside capture (function attributes, lazy module load)
        torch.cuda.current_stream(self.device).synchronize()
        self.class_hist.copy_(saved)  # the capturing call does not execute; the warm-up must not count
        g = torch.cuda.CUDAGraph(keep_graph=True)
        with torch.cuda.graph(g):
            body()
        g.instantiate()
        return g

    def step_graph(self, steps: int) -> torch.cuda.CUDAGraph:
        """The CUDA graph of ``steps`` consecutive batches starting at the current buffer parity (captured once
        per (steps, parity)): 3-4 kernel nodes per step, every node but the first behind a PROGRAMMATIC edge."""
        key = (int(steps), self._parity)
        g = self._step_graphs.get(key)
        if g is None:
            p0 = self._parity

            def body():
                for i in range(steps):
                    self.enqueue(p0 ^ (i & 1))

            with torch.cuda.device(self.device):
                g = self._capture(body)
            self._step_graphs[key] = g
        return g

    def run_steps(self, steps: int) -> None:
        """``steps`` batches over the resident inputs as ONE graph launch (K1/K2/K4 of consecutive steps keep
        their programmatic-dependent-launch overlap inside the graph; graph launches serialise)."""
        g = self.step_graph(steps)
        self._parity ^= steps & 1
        with torch.cuda.device(self.device):
            g.replay()
