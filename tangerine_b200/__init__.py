"""tangerine_b200: B200-native (sm_100a) SDF meshing hot path of Aeva/tangerine.

The product is ``libtangerine_b200.so`` (C ABI in ``include/tangerine_b200.h``, C++ host mirror of the
reference's export interface in ``include/tangerine_b200.hpp``).  This Python package is a thin ctypes
binding used by the tests and by ``bench.py``; it contains no compute and no CPU fallback.
"""
from .api import (  # noqa: F401
    EVAL_COLOR, EVAL_GRADIENT, EVAL_INTERP, EVAL_LIVE, EVAL_OCTREE, EVAL_TREE,
    MESH_COLORS, MESH_DEVICE_ONLY, MESH_FACE_NORMALS, MESH_FAST, MESH_LIVE_FIELD, MESH_NO_CULL, MESH_NORMALS, MESH_REBALANCE,
    Context, Grid, Mesh, Model, TangerineError, Tree, export_grid, lib, library_path,
)
