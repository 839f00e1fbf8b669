"""Per-frame quality log and run summary.

Keeps the counters and the per-frame entries of the reference's ``DataQualityLogger``
(gcd.py:236-464) and writes ``generation_summary.json`` with the same keys, so whatever reads
the reference's summary keeps working.  The depth entry comes from the statistics the GPU
already reduced (``cspe_depth_stats``, f2) instead of five numpy passes over the depth map
(gcd.py:318-330); label / point-cloud / RGB entries take the counts the kernels return.

``report()`` is the readable text of the reference's ``_generate_report`` (gcd.py:420-457), pinned by the report the
reference's own logger returned for a scripted run (``tests/golden/quality_log.json``).

The issue strings are data in that file — the reference's own report groups frames by the text
before the first ``:`` (gcd.py:449-453) — so they are kept verbatim.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Mapping, Optional, Sequence

ISSUE_DEPTH_ALL_ZERO = "深度图全为零"        # gcd.py:346
ISSUE_DEPTH_ALL_INF = "深度图全为无穷"       # gcd.py:350
ISSUE_DEPTH_FAILED = "深度图失败"            # gcd.py:358
ISSUE_NO_OBJECTS = "未识别到物体"            # gcd.py:370
ISSUE_CLOUD_EMPTY = "点云为空"               # gcd.py:294
ISSUE_CLOUD_SHORT = "点云不足"               # gcd.py:298
ISSUE_RGB_FAILED = "RGB失败"                 # gcd.py:310


def depth_quality_from_stats(st) -> Dict[str, object]:
    """One ``cspe_depth_stats_t`` record (``_lib.DEPTH_STATS_DTYPE``) -> the dict the reference stores under
    ``current_frame['depth']`` (gcd.py:333-342).  The mean is rounded to f32 like ``np.mean`` of an f32 array."""
    import numpy as np

    valid, total = int(st["valid_pixels"]), int(st["total_pixels"])
    mean = float(np.float32(st["depth_sum"] / valid)) if valid else 0.0
    return {"status": "valid", "valid_pixels": valid, "total_pixels": total,
            "valid_ratio": float(valid / total) if total else 0.0, "zero_pixels": int(st["zero_pixels"]),
            "inf_pixels": int(st["inf_pixels"]), "depth_range": [float(st["depth_min"]), float(st["depth_max"])],
            "depth_mean": mean}


def _empty_statistics() -> Dict[str, object]:
    # gcd.py:244-254
    return {
        "total_frames_attempted": 0,
        "successful_frames": 0,
        "failed_frames": 0,
        "retry_count": 0,
        "pointcloud_stats": {"valid": 0, "empty": 0, "insufficient": 0},
        "rgb_stats": {"valid": 0, "failed": 0},
        "depth_stats": {"valid": 0, "failed": 0, "all_zero": 0, "all_inf": 0},
        "label_stats": {"valid": 0, "empty": 0},
        "object_count": {"total": 0, "per_frame_avg": 0},
    }


class FrameQualityLog:
    """Accumulates what the reference's logger accumulates; no file is touched until ``save_summary``."""

    def __init__(self, log_dir: Optional[str] = None):
        self.log_dir = log_dir
        self.statistics = _empty_statistics()
        self.frame_logs: List[Dict[str, object]] = []
        self.current_frame: Optional[Dict[str, object]] = None

    # ---- per-frame events, same order as the capture loop (gcd.py:1567-2078) -------------------
    def frame_start(self, frame_id: int, camera_position: Sequence[float] = (0.0, 0.0, 0.0)) -> None:
        pos = camera_position.tolist() if hasattr(camera_position, "tolist") else list(camera_position)
        self.current_frame = {"frame_id": frame_id, "camera_position": pos, "retry_count": 0,
                              "status": "processing", "issues": []}

    def retry(self, retry_count: int) -> None:
        self.current_frame["retry_count"] = retry_count
        self.statistics["retry_count"] += 1

    def pointcloud(self, valid: bool, point_count: int = 0, reason: str = "") -> None:
        stats = self.statistics["pointcloud_stats"]
        if valid:
            stats["valid"] += 1
            self.current_frame["pointcloud"] = {"status": "valid", "points": point_count}
        elif point_count == 0:
            stats["empty"] += 1
            self.current_frame["issues"].append(f"{ISSUE_CLOUD_EMPTY}: {reason}")
        else:
            stats["insufficient"] += 1
            self.current_frame["issues"].append(f"{ISSUE_CLOUD_SHORT}: {point_count} 点")

    def rgb(self, valid: bool, reason: str = "") -> None:
        if valid:
            self.statistics["rgb_stats"]["valid"] += 1
            self.current_frame["rgb"] = {"status": "valid"}
        else:
            self.statistics["rgb_stats"]["failed"] += 1
            self.current_frame["issues"].append(f"{ISSUE_RGB_FAILED}: {reason}")

    def depth(self, quality: Optional[Mapping], reason: str = "") -> None:
        """``quality`` is ``BatchLabels.depth_quality(f)`` (the dict of gcd.py:333-342) or None = no depth."""
        stats = self.statistics["depth_stats"]
        if quality is None:
            stats["failed"] += 1
            self.current_frame["issues"].append(f"{ISSUE_DEPTH_FAILED}: {reason}")
            return
        self.current_frame["depth"] = dict(quality)
        total = quality["total_pixels"]
        if quality["zero_pixels"] == total:
            stats["all_zero"] += 1
            self.current_frame["issues"].append(ISSUE_DEPTH_ALL_ZERO)
        elif quality["inf_pixels"] == total:
            stats["all_inf"] += 1
            self.current_frame["issues"].append(ISSUE_DEPTH_ALL_INF)
        else:
            stats["valid"] += 1

    def labels(self, object_count: int) -> None:
        if object_count > 0:
            self.statistics["label_stats"]["valid"] += 1
            self.statistics["object_count"]["total"] += object_count
            self.current_frame["labels"] = {"status": "valid", "object_count": object_count}
        else:
            self.statistics["label_stats"]["empty"] += 1
            self.current_frame["issues"].append(ISSUE_NO_OBJECTS)

    def frame_end(self, success: bool) -> None:
        self.statistics["total_frames_attempted"] += 1
        self.statistics["successful_frames" if success else "failed_frames"] += 1
        self.current_frame["status"] = "success" if success else "failed"
        self.frame_logs.append(dict(self.current_frame))

    # ---- run summary (gcd.py:389-418) ---------------------------------------------------------
    def summary(self) -> Dict[str, object]:
        st = self.statistics
        if st["successful_frames"] > 0:
            st["object_count"]["per_frame_avg"] = st["object_count"]["total"] / st["successful_frames"]
        st["success_rate"] = st["successful_frames"] / max(1, st["total_frames_attempted"])
        return {"statistics": st, "frame_logs": self.frame_logs}

    def issue_counts(self) -> Dict[str, int]:
        """Issues grouped by their text before the first ':' — most frequent first (gcd.py:447-455)."""
        counts: Dict[str, int] = {}
        for frame in self.frame_logs:
            for issue in frame.get("issues", []):
                key = issue.split(":")[0]
                counts[key] = counts.get(key, 0) + 1
        return dict(sorted(counts.items(), key=lambda kv: kv[1], reverse=True))

    def report(self) -> str:
        """The readable run report of the reference's logger (``_generate_report``, gcd.py:420-457): the counters of
        ``summary()`` as text, then the issues grouped by type, most frequent first."""
        st = self.summary()["statistics"]
        pc, rgb, dp, lb, oc = st["pointcloud_stats"], st["rgb_stats"], st["depth_stats"], st["label_stats"], st["object_count"]
        lines = ["=== 数据生成汇总报告 ===", "",
                 "总体统计:",
                 f"  尝试帧数: {st['total_frames_attempted']}",
                 f"  成功帧数: {st['successful_frames']}",
                 f"  失败帧数: {st['failed_frames']}",
                 f"  成功率: {st['success_rate'] * 100:.1f}%",
                 f"  总重试次数: {st['retry_count']}", "",
                 "点云质量:",
                 f"  有效: {pc['valid']}", f"  为空: {pc['empty']}", f"  不足: {pc['insufficient']}", "",
                 "RGB图像:",
                 f"  成功: {rgb['valid']}", f"  失败: {rgb['failed']}", "",
                 "深度图:",
                 f"  有效: {dp['valid']}", f"  失败: {dp['failed']}", f"  全零: {dp['all_zero']}", f"  全无穷: {dp['all_inf']}", "",
                 "标签识别:",
                 f"  有效: {lb['valid']}", f"  为空: {lb['empty']}", f"  总物体数: {oc['total']}",
                 f"  平均每帧: {oc['per_frame_avg']:.2f}", "",
                 "常见问题:"]
        lines += [f"  {kind}: {count} 次" for kind, count in self.issue_counts().items()]
        return "\n".join(lines) + "\n"

    def save_summary(self, path: Optional[str] = None) -> Dict[str, object]:
        """``generation_summary.json`` (gcd.py:406-407) and, next to it, the readable report (the reference appends
        it to its detail log, gcd.py:410-413; here it is a file of its own, ``generation_report.txt`` — the per-frame
        lines of that detail log are not reproduced)."""
        data = self.summary()
        if path is None and self.log_dir is not None:
            os.makedirs(self.log_dir, exist_ok=True)
            path = os.path.join(self.log_dir, "generation_summary.json")   # gcd.py:259
        if path is not None:
            with open(path, "w", encoding="utf-8") as f:
                json.dump(data, f, indent=2, ensure_ascii=False)
            base = os.path.basename(path)
            name = base.replace("summary", "report").rsplit(".", 1)[0] + ".txt" if "summary" in base else base + ".report.txt"
            with open(os.path.join(os.path.dirname(path), name), "w", encoding="utf-8") as f:
                f.write(self.report())
        return data
