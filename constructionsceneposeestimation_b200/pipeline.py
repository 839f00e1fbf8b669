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
                 mask: Optional[torch.Tensor] = None, num_people: int = 0, num_joints: int = 0,
                 keypoint_tolerance: float = 0.15):
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
    def _enqueue(self, parity: int) -> None:
        """scan -> project -> emit on ONE stream, 3 launches, all chained by programmatic dependent
        launch.  The scan table is initialised once (constructor) and re-initialised by K4 as it
        reads it, so K1 runs in accumulate mode with no init launch.
          * K1 (overlapped) may start while the PREVIOUS batch's K4 is still running — K4 releases
            its dependents at entry — and waits for it only before its first merge into the table;
          * K2 (overlapped) starts as soon as K1's CTAs are resident and runs in the SM resources
            K1 leaves free; it completes only after K1 does;
          * K4 waits for K2 (hence K1) and reads this batch's K2 buffers (double-buffered)."""
        lib, chk = self.lib, _lib.check
        k2 = self._k2[parity]
        main = torch.cuda.current_stream(self.device).cuda_stream
        chk("cspe_mask_scan_accumulate_overlapped", lib.cspe_mask_scan_accumulate_overlapped(
            self.mask.data_ptr(), self.B, self.H, self.W, self.lut.data_ptr(), self.L, self.lut_stride, self.N,
            self.scan.data_ptr(), main))
        chk("cspe_project_objects_overlapped", lib.cspe_project_objects_overlapped(
            self.records_in.data_ptr(), _lib.BBOX3D_RECORD_BYTES, self.R, self.obj_record.data_ptr(),
            self.cam.data_ptr(), self.B, self.N, k2["uv"].data_ptr(), k2["z"].data_ptr(), k2["pose"].data_ptr(),
            k2["loose"].data_ptr(), k2["flags"].data_ptr(), main))
        if self._k3 is not None:  # K3 rides the same chain: beside the scan, done before K4
            k3 = self._k3[parity]
            chk("cspe_keypoints_overlapped", lib.cspe_keypoints_overlapped(
                self.joints.data_ptr(), self.B, self.P, self.J, self.depth.data_ptr(), self.H, self.W,
                self.cam.data_ptr(), self.keypoint_tolerance, k3["kp"].data_ptr(), k3["kz"].data_ptr(),
                k3["vis"].data_ptr(), main))
        chk("cspe_emit_reset_scan", lib.cspe_emit_reset_scan(
            self.scan.data_ptr(), k2["uv"].data_ptr(), k2["z"].data_ptr(), k2["pose"].data_ptr(),
            k2["loose"].data_ptr(), k2["flags"].data_ptr(), self.slot_class.data_ptr(), self.B, self.N, self.H, self.W,
            self.min_pixels, self.frame_base, self.records.data_ptr(), self.n_out.data_ptr(),
            self.class_hist.data_ptr(), main))

    def run(self) -> None:
        """Enqueue one batch on the current stream (graph replay when enabled).

        Inputs (mask, lut, obj_record, slot_class, records_in, cam) must not be rewritten by a kernel
        queued directly before run(): the overlapped launches read them without waiting for it.
        Rewrite them with copies / from another stream's event, or synchronise first."""
        parity = self._parity
        self._parity ^= 1
        with torch.cuda.device(self.device):
            if not self.use_graph:
                self._enqueue(parity)
                return
            if self.graphs[parity] is None:
                saved = self.class_hist.clone()
                self._enqueue(parity)  # warm-up outside capture (function attributes, lazy module load)
                torch.cuda.current_stream(self.device).synchronize()
                self.class_hist.copy_(saved)  # one run() = one accumulation, also on the capturing call
                # the warm-up consumed the table reset of this parity's emit; nothing else to restore
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(parity)
                self.graphs[parity] = g
            self.graphs[parity].replay()
