"""CPU oracle for the per-frame annotation hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import this package.  The product package (``constructionsceneposeestimation_b200``)
never does; it fails loudly when libcspe.so is missing.

Reference: ``/root/reference/generate_construction_data.py`` (``gcd.py`` in citations).

Pinning status:
  * [REF] functions — ``bbox_to_transform`` (gcd.py:553-584), ``depth_to_pointcloud``
    (gcd.py:616-711), ``depth_stats`` (gcd.py:314-359), the class table / ``get_object_root``
    (gcd.py:69-121, 144-233): PINNED.  ``oracle/reference_extract.py`` AST-extracts the
    reference's own functions (no copy into this repo) and ``tests/golden/make_golden.py``
    froze their outputs on seeded inputs into ``tests/golden/*.npz|json``; ``tests/`` check the
    oracle against those fixtures everywhere and against the live functions wherever
    ``/root/reference`` exists.
  * ``camera_pose_from_usd_matrix`` (gcd.py:587-605) needs ``pxr``: restated from documented
    Gf semantics, PARITY UNPINNED.
  * [SPEC] stages the reference does not implement — mask scan (S1), corner projection (S2),
    object-in-camera pose (S3), keypoints (S4), occlusion ratios (S5), emission (S6), class
    histogram (S7): the reference holds no code, tests or golden vectors for them, so PARITY IS
    UNPINNED by the reference; this oracle *defines* them, following the
    reference's conventions (matrix layout gcd.py:568, intrinsics gcd.py:646-649, camera pose
    gcd.py:587-605, class ids gcd.py:69-106).  S1 is cross-checked by two independent numpy
    formulations plus a C restatement (``oracle/c/scan_oracle.c``).
"""
