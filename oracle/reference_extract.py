"""Execute the reference's own pure functions without importing the script.

TEST INFRASTRUCTURE.  ``generate_construction_data.py`` cannot be imported (it needs
``omni``/``pxr``/``isaacsim`` and has module-level side effects, gcd.py:13-19, 35,
1351-1375, 2098), but its numpy functions run fine once lifted out: we parse the file,
keep only top-level ``def``/``class`` nodes and the two table assignments, and ``exec``
them with numpy/scipy in the namespace.  Nothing is copied into this repository; the
source is read where it lies, so this only works where ``/root/reference`` exists (the
build container — not the GPU box, which uses the frozen fixtures in ``tests/golden``).
"""
from __future__ import annotations

import ast
import json
import os
import re
from pathlib import Path
from types import SimpleNamespace
from typing import Optional

import numpy as np

REFERENCE_SCRIPT = Path(os.environ.get("CSPE_REFERENCE_DIR", "/root/reference")) / "generate_construction_data.py"

_WANTED_ASSIGNS = {"construction_class", "CRANE_PART_CHILD_MAP", "_crane_part_map", "OBJECT_ROOT_PATTERNS"}
_cached: Optional[SimpleNamespace] = None


def available() -> bool:
    return REFERENCE_SCRIPT.exists()


def load() -> SimpleNamespace:
    """Namespace holding the reference's functions/classes/tables (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"{REFERENCE_SCRIPT} not present on this machine")
    from scipy.spatial.transform import Rotation as R  # what the script imports as R (gcd.py:23)

    tree = ast.parse(REFERENCE_SCRIPT.read_text(encoding="utf-8"))
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            keep.append(node)
        elif isinstance(node, ast.Assign):
            names = {t.id for t in node.targets if isinstance(t, ast.Name)}
            if names & _WANTED_ASSIGNS:
                keep.append(node)
    module = ast.Module(body=keep, type_ignores=[])
    ns = {"np": np, "R": R, "json": json, "os": os, "re": re, "Path": Path, "__name__": "gcd_extracted"}
    exec(compile(module, str(REFERENCE_SCRIPT), "exec"), ns)
    _cached = SimpleNamespace(**{k: v for k, v in ns.items() if not k.startswith("__")})
    _cached._ns = ns
    return _cached


def set_crane_part_map(mapping) -> None:
    """Fill the run-time crane table the reference builds from the live stage (gcd.py:124)."""
    ref = load()
    ref._ns["_crane_part_map"].clear()
    ref._ns["_crane_part_map"].update(mapping)
