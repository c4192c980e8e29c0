"""Helpers shared by the parity tests: digests identical to tests/golden/make_golden.py."""
import hashlib

import numpy as np


def sort_rows(a):
    a = np.ascontiguousarray(a)
    if len(a) == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def vertex_records(pos, normal, color):
    cols = [np.ascontiguousarray(pos, np.float32).view(np.uint32), np.ascontiguousarray(normal, np.float32).view(np.uint32)]
    if color is not None:
        cols.append(np.asarray(color).astype(np.uint32))
    return sort_rows(np.concatenate(cols, axis=1))


def mesh_summary(pos, normal, color, tris):
    pos = np.ascontiguousarray(pos, np.float32)
    rec = vertex_records(pos, normal, color)
    tri = sort_rows(pos[np.asarray(tris, np.int64)].reshape(-1, 9).view(np.uint32))
    return {"vertices": int(len(pos)), "faces": int(len(tris)), "has_color": color is not None,
            "vertex_records_sha256": digest(rec), "triangle_coords_sha256": digest(tri),
            "positions_in_order_sha256": digest(pos)}


def same_floats(a, b):
    """Bit-for-bit equality up to the sign of zero and NaN payloads."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def ulp_diff(a, b):
    """Distance in units in the last place between two float32 arrays."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)
