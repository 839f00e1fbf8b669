"""Host-side prim-path -> (object root, class) resolution and mesh -> object aggregation.

SURVEY §8a rows R1 and R2.  Behaviour follows the reference's class table
(gcd.py:69-106), crane child map (gcd.py:110-121), ``get_object_root`` (gcd.py:144-233) and
the aggregation loop (gcd.py:1858-1891); it stays on the host (a few hundred strings per
frame) and its output becomes the device tables the kernels consume: ``id -> slot`` for K1,
``slot -> record`` for K2 and ``slot -> class`` for K4.

The implementation is a rule table, memoised per path, rather than the reference's
if-chain; ``tests/test_oracle_pinning.py`` pins it against golden vectors produced by the
reference function itself (``tests/golden/object_roots.json``, ``frame_label.json``) and fuzzes it
against the live function where ``/root/reference`` exists.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from collections.abc import Mapping
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

# substring (lower case) -> class id; insertion order matters for the generic fallback
# (first key that occurs in the path wins), so it mirrors gcd.py:69-106 key for key.
CLASS_TABLE: Dict[str, int] = {
    "trafficcone": 0,
    "cone": 0,
    "tree": 1,
    "fence": 2,
    "fencing": 2,
    "construction_site": 2,
    "crane": 3,
    "pk7": 3,
    "cranebase": 6,
    "cranecolumn": 7,
    "craneboom": 8,
    "cranetelescopic": 9,
    "dumper": 4,
    "09684481": 4,
    "human": 5,
    "dhgen": 5,
    "skelroot": 5,
}

NUM_CLASSES = 10

# canonical name per class id, used for COCO categories / YOLO names
CLASS_NAMES: Tuple[str, ...] = (
    "trafficcone", "tree", "fence", "crane", "dumper", "human",
    "cranebase", "cranecolumn", "craneboom", "cranetelescopic",
)

# first-level child of the crane root (lower case) -> (part name, class id); gcd.py:110-121
CRANE_CHILD_PARTS: Dict[str, Tuple[str, int]] = {
    "s104gg03a_sw": ("cranebase", 6),
    "s104s01kb_sw": ("cranebase", 6),
    "s104hz01ka_sw": ("cranecolumn", 7),
    "s104h01kb_sw": ("cranecolumn", 7),
    "s104hz02ka_sw": ("cranecolumn", 7),
    "s104kz01ka_sw": ("cranecolumn", 7),
    "tn__s104ekb_as_sw_jj7": ("craneboom", 8),
    "s104kz02ka_sw": ("cranetelescopic", 9),
    "tn__hhk320ka_sw_lg": ("cranetelescopic", 9),
    "tn__hhk319_sw_od": ("cranetelescopic", 9),
}

CRANE_ROOT = "/World/GroundPlane/tn__Pk7501SLD_PNR3879_fPM"
DUMPER_ROOT = "/World/GroundPlane/tn__09684481_"
HUMAN_ROOT = "/World/GroundPlane/DHGen"

# keyword fallback for crane sub-paths, tested in this order (gcd.py:200-212)
_CRANE_KEYWORDS: Tuple[Tuple[str, Tuple[str, ...]], ...] = (
    ("cranebase", ("base", "chassis", "footer", "support", "grund", "fahrwerk")),
    ("cranecolumn", ("column", "turret", "mast", "tower", "saeule", "drehwerk", "oberwagen")),
    ("craneboom", ("boom", "arm", "jib", "ausleger")),
    ("cranetelescopic", ("telescop", "extension", "teleskop", "auszug")),
)

Resolved = Tuple[Optional[str], Optional[str], Optional[int]]
_UNMATCHED: Resolved = (None, None, None)


class ObjectRootResolver:
    """``resolve(path) -> (object_root, class_name, class_id)`` with per-path memoisation.

    ``crane_part_map`` is the run-time table the reference fills from the live stage
    (``build_crane_part_map`` gcd.py:1234-1279): exact mesh path -> (part name, class id).
    ``split_people`` is an extension (off by default = reference behaviour, where every
    ``dhgen`` path folds into the single root ``/World/GroundPlane/DHGen``): when on, the path
    segment that starts with ``dhgen`` is the person's root, so several people stay apart.
    """

    def __init__(self, crane_part_map: Optional[Mapping[str, Tuple[str, int]]] = None, split_people: bool = False):
        self.crane_part_map = dict(crane_part_map or {})
        self.split_people = split_people
        self._memo: Dict[str, Resolved] = {}

    def resolve(self, prim_path: str) -> Resolved:
        hit = self._memo.get(prim_path)
        if hit is None:
            hit = self._resolve(prim_path)
            self._memo[prim_path] = hit
        return hit

    __call__ = resolve

    # -- rules, in the reference's order ---------------------------------------------------
    def _resolve(self, path: str) -> Resolved:
        low = path.lower()
        segs = path.split("/")

        if "fencing_height_" in low:  # gcd.py:154-160 (segment test is case sensitive)
            for i, seg in enumerate(segs):
                if "Fencing_height_" in seg:
                    return "/".join(segs[: i + 1]), "fence", CLASS_TABLE["fence"]

        if "/world/tree/tree" in low and len(segs) >= 4:  # gcd.py:163-168
            return "/".join(segs[:4]), "tree", CLASS_TABLE["tree"]

        if "/cone001" in low:  # gcd.py:171-176
            for i, seg in enumerate(segs):
                if seg.lower().startswith("cone001"):
                    return "/".join(segs[: i + 1]), "trafficcone", CLASS_TABLE["trafficcone"]

        if "pk7" in low:  # gcd.py:179-217 ("pk7501sld" contains "pk7")
            return self._resolve_crane(path, low)

        if "09684481" in low:  # gcd.py:220-221
            return DUMPER_ROOT, "dumper", CLASS_TABLE["dumper"]

        if "dhgen" in low:  # gcd.py:224-225
            if self.split_people:
                for i, seg in enumerate(segs):
                    if seg.lower().startswith("dhgen"):
                        return "/".join(segs[: i + 1]), "human", CLASS_TABLE["human"]
            return HUMAN_ROOT, "human", CLASS_TABLE["human"]

        for key, cid in CLASS_TABLE.items():  # gcd.py:228-231
            if key in low:
                return path, key, cid
        return _UNMATCHED

    def _resolve_crane(self, path: str, low: str) -> Resolved:
        part = self.crane_part_map.get(path)
        if part is not None:
            return f"{CRANE_ROOT}#{part[0]}", part[0], part[1]
        prefix = CRANE_ROOT + "/"
        if path.startswith(prefix) or low.startswith(prefix.lower()):
            first = path[len(prefix):].split("/")[0].lower()
            part = CRANE_CHILD_PARTS.get(first)
            if part is not None:
                return f"{CRANE_ROOT}#{part[0]}", part[0], part[1]
        sub = low[low.find("pk7"):]
        for name, words in _CRANE_KEYWORDS:
            if any(w in sub for w in words):
                return f"{CRANE_ROOT}#{name}", name, CLASS_TABLE[name]
        return CRANE_ROOT, "crane", CLASS_TABLE["crane"]


@dataclass
class SceneObject:
    """One aggregated object = one label slot (fields of gcd.py:1878-1885)."""
    inst_idx: int
    class_id: int
    class_name: str
    prim_path: str
    mesh_paths: List[str] = field(default_factory=list)

    @property
    def mesh_count(self) -> int:
        return len(self.mesh_paths)

    @property
    def actual_prim_path(self) -> str:  # gcd.py:1929
        return self.prim_path.split("#")[0] if "#" in self.prim_path else self.prim_path


def aggregate_objects(prim_paths: Iterable[str], resolver: ObjectRootResolver) -> List[SceneObject]:
    """Group visible mesh paths by object root; inst_idx = first-seen order (gcd.py:1858-1886)."""
    by_root: Dict[str, SceneObject] = {}
    for path in prim_paths:
        root, name, cid = resolver.resolve(path)
        if root is None:
            continue
        obj = by_root.get(root)
        if obj is None:
            obj = SceneObject(len(by_root), cid, name, root)
            by_root[root] = obj
        obj.mesh_paths.append(path)
    return list(by_root.values())


RECORD_APPROX_BIT = 1 << 30   # = CSPE_OBJ_RECORD_APPROX_BIT (include/cspe.h)


def record_index_for(objects: Sequence[SceneObject], prim_paths: Sequence[str], fallback: str = "first_mesh",
                     mark_approx: bool = True) -> List[int]:
    """bbox3d record index per object, or -1.

    ``primPaths.index(root)`` first (gcd.py:1934); crane parts (virtual ``root#part``) then try
    their mesh paths in order (gcd.py:1953-1975).  ``fallback="first_mesh"`` extends that mesh
    rule to every object because the reference's other fallback — reading the live USD stage
    (gcd.py:1977-2023) — does not exist outside Isaac Sim; ``fallback="reference"`` keeps the
    reference's rule and leaves such objects without a record.  A first-mesh stand-in for anything but a
    crane part describes ONE mesh of a multi-mesh object, not the object: its index carries
    ``RECORD_APPROX_BIT`` so K2 sets ``OBJ_APPROX_RECORD`` in the record's flags and the label says so.

    ``fallback="union"`` gives those multi-mesh objects an object-level record instead: the u-th entry of
    ``union_members(objects, prim_paths)`` gets index ``len(prim_paths) + u`` (still with ``RECORD_APPROX_BIT``: the
    box is built from the mesh records, not read from the object's own prim) — the place where
    ``cspe_union_records`` writes the world-axis-aligned range of all its mesh records, which is the bound the
    reference's USD fallback computes (gcd.py:2000-2009).  A caller whose record array is padded to R > len(prim_paths)
    records re-bases those entries to ``R + u``.
    """
    if fallback not in ("first_mesh", "reference", "union"):
        raise ValueError(f"unknown record fallback {fallback!r}")
    first_index: Dict[str, int] = {}
    for i, p in enumerate(prim_paths):
        first_index.setdefault(p, i)
    out: List[int] = []
    n_union = 0
    for obj in objects:
        idx = first_index.get(obj.actual_prim_path, -1)
        if idx < 0 and ("#" in obj.prim_path or fallback != "reference"):
            multi = "#" not in obj.prim_path and len(obj.mesh_paths) > 1
            if fallback == "union" and multi:
                if any(mp in first_index for mp in obj.mesh_paths):
                    idx = (len(prim_paths) + n_union) | RECORD_APPROX_BIT
                    n_union += 1
            else:
                for mp in obj.mesh_paths:
                    idx = first_index.get(mp, -1)
                    if idx >= 0:
                        break
                # the reference's own rule for crane parts (gcd.py:1953-1975) is not an approximation of ours; a
                # single-mesh object is described exactly by its only mesh
                if idx >= 0 and mark_approx and multi:
                    idx |= RECORD_APPROX_BIT
        out.append(idx)
    return out


def union_members(objects: Sequence[SceneObject], prim_paths: Sequence[str]) -> List[Tuple[int, List[int]]]:
    """``[(slot, [record index of every mesh path that has a record])]`` for the objects ``fallback="union"`` builds
    an object-level record for, in slot order: no record under the object's own prim path, not a crane part (those
    keep the reference's first-mesh rule, gcd.py:1953-1975), several mesh paths, at least one of them with a record."""
    first_index: Dict[str, int] = {}
    for i, p in enumerate(prim_paths):
        first_index.setdefault(p, i)
    out: List[Tuple[int, List[int]]] = []
    for slot, obj in enumerate(objects):
        if obj.actual_prim_path in first_index or "#" in obj.prim_path or len(obj.mesh_paths) <= 1:
            continue
        mem = [first_index[mp] for mp in obj.mesh_paths if mp in first_index]
        if mem:
            out.append((slot, mem))
    return out


def pack_union(per_frame: Sequence[Sequence[Tuple[int, Sequence[int]]]]):
    """CSR tables of a batch for ``cspe_union_records``: (offsets int32 [B, U+1], members int32 [B, M], U) with U / M
    the largest count of the batch (frames with fewer union objects get empty ranges; M >= 1)."""
    import numpy as np

    B = len(per_frame)
    U = max((len(p) for p in per_frame), default=0)
    M = max(1, max((sum(len(m) for _, m in p) for p in per_frame), default=0))
    offsets = np.zeros((B, U + 1), dtype=np.int32)
    members = np.full((B, M), -1, dtype=np.int32)
    for f, plan in enumerate(per_frame):
        pos = 0
        for u, (_, mem) in enumerate(plan):
            members[f, pos: pos + len(mem)] = mem
            pos += len(mem)
            offsets[f, u + 1] = pos
        offsets[f, len(plan) + 1:] = pos
    return offsets, members, U


def label_path(label) -> Optional[str]:
    """Prim path of one ``idToLabels`` entry (a string, or a dict as tolerated at gcd.py:1829-1834)."""
    if isinstance(label, str):
        return label
    if isinstance(label, Mapping):
        for key in ("primPath", "prim_path", "path"):
            v = label.get(key)
            if isinstance(v, str):
                return v
    return None


def id_to_slot(id_to_labels: Mapping, objects: Sequence[SceneObject], resolver: ObjectRootResolver) -> Dict[int, int]:
    """instance id -> slot for ids whose prim resolves to one of ``objects`` (others omitted).

    Ids 0 (BACKGROUND) and 1 (UNLABELLED) never resolve to a class and drop out naturally.
    """
    slot_of_root = {o.prim_path: o.inst_idx for o in objects}
    out: Dict[int, int] = {}
    for key, label in id_to_labels.items():
        path = label_path(label)
        if path is None:
            continue
        root, _, _ = resolver.resolve(path)
        slot = slot_of_root.get(root) if root is not None else None
        if slot is not None:
            try:
                out[int(key)] = slot
            except (TypeError, ValueError):
                continue   # a non-numeric id key: skip the entry, never fatal (gcd.py:2024-2027 convention)
    return out
