"""Static-shape batch pipeline: K1 (mask scan) || K2 (projection/pose) -> K4 (emission).

All buffers are allocated once; one call to :meth:`LabelPipeline.run` enqueues the whole batch on
one stream.  K2 does not depend on K1: it is launched with programmatic stream serialisation and
K1 releases its dependents as soon as its persistent CTAs are resident, so K2 runs in the SM
resources K1 leaves free.  With ``use_graph=True`` the four launches are captured once into a
CUDA graph and replayed (no per-launch host overhead, no tracing compiler — the kernels are the
hand-written ones behind the C ABI).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import CAM_STRIDE, NUM_CLASSES, POSE_STRIDE, RECORD_DTYPE, SCAN_FIELDS


class LabelPipeline:
    def __init__(self, batch: int, height: int, width: int, num_slots: int, recs_per_frame: int, lut_len: int,
                 device: torch.device, per_frame_lut: bool = True,  # lut_len: use a multiple of 4 (16-byte rows)
                 min_pixels: int = 1, use_graph: bool = True,
                 mask: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        self.B, self.H, self.W, self.N, self.R, self.L = batch, height, width, num_slots, recs_per_frame, lut_len
        self.device = torch.device(device)
        self.min_pixels = int(min_pixels)
        self.frame_base = 0
        dev = self.device
        i32, f64, u8 = torch.int32, torch.float64, torch.uint8
        B, N = batch, num_slots
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
        self.uv = torch.empty((B, N, 8, 2), dtype=f64, device=dev)
        self.z = torch.empty((B, N, 8), dtype=f64, device=dev)
        self.pose = torch.empty((B, N, POSE_STRIDE), dtype=f64, device=dev)
        self.loose = torch.empty((B, N, 4), dtype=f64, device=dev)
        self.flags = torch.empty((B, N), dtype=u8, device=dev)
        self.records = torch.empty((B, N, RECORD_DTYPE.itemsize), dtype=u8, device=dev)
        self.n_out = torch.empty((B,), dtype=i32, device=dev)
        self.class_hist = torch.zeros((NUM_CLASSES,), dtype=torch.int64, device=dev)
        self.launches_per_run = 3  # mask_scan (accumulate) + project_objects + emit (which re-initialises the scan table)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph

    # ---------------------------------------------------------------- enqueue
    def _enqueue(self) -> None:
        """scan -> project -> emit on ONE stream, 3 launches.  The scan table is initialised once
        (constructor) and re-initialised by K4 as it reads it, so K1 runs in accumulate mode with no
        init launch.  K2 is launched with programmatic stream serialisation and the scan releases its
        dependents as soon as its CTAs are resident, so K2 runs beside the scan (it does not read the
        scan's output) and K4 still sees both done."""
        lib, chk = self.lib, _lib.check
        main = torch.cuda.current_stream(self.device).cuda_stream
        chk("cspe_mask_scan_accumulate", lib.cspe_mask_scan_accumulate(
            self.mask.data_ptr(), self.B, self.H, self.W, self.lut.data_ptr(), self.L, self.lut_stride, self.N,
            self.scan.data_ptr(), main))
        chk("cspe_project_objects_overlapped", lib.cspe_project_objects_overlapped(
            self.records_in.data_ptr(), _lib.BBOX3D_RECORD_BYTES, self.R, self.obj_record.data_ptr(),
            self.cam.data_ptr(), self.B, self.N, self.uv.data_ptr(), self.z.data_ptr(), self.pose.data_ptr(),
            self.loose.data_ptr(), self.flags.data_ptr(), main))
        chk("cspe_emit_reset_scan", lib.cspe_emit_reset_scan(
            self.scan.data_ptr(), self.uv.data_ptr(), self.z.data_ptr(), self.pose.data_ptr(), self.loose.data_ptr(),
            self.flags.data_ptr(), self.slot_class.data_ptr(), self.B, self.N, self.H, self.W, self.min_pixels,
            self.frame_base, self.records.data_ptr(), self.n_out.data_ptr(), self.class_hist.data_ptr(), main))

    def run(self) -> None:
        """Enqueue one batch on the current stream (graph replay when enabled)."""
        with torch.cuda.device(self.device):
            if not self.use_graph:
                self._enqueue()
                return
            if self.graph is None:
                saved = self.class_hist.clone()
                self._enqueue()  # warm-up outside capture (function attributes, lazy module load)
                torch.cuda.current_stream(self.device).synchronize()
                self.class_hist.copy_(saved)  # one run() = one accumulation, also on the capturing call
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self.graph = g
            self.graph.replay()
