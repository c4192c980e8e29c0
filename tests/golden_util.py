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


# ---- order-independent digests of the exported files -------------------------------------------------
# The reference writes triangles in the iteration order of a std::unordered_map (surface_nets.cpp:1124-1146 walks
# `active cell -> vertex id`) and MagicaVoxel voxels in the order its Pool() threads won the mutex (magica.cpp:44-66);
# neither order is part of the format.  Everything else in the files is compared byte for byte: the header and the
# vertex block of a PLY as one digest, face / triangle / voxel records as sorted multisets.

def _sorted_records(raw, width):
    rec = np.frombuffer(raw, np.uint8).reshape(-1, width)
    return sort_rows(rec).tobytes()


def ply_file_digests(path):
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode()
    nv = nf = 0
    for line in header.splitlines():
        if line.startswith("element vertex"):
            nv = int(line.split()[-1])
        if line.startswith("element face"):
            nf = int(line.split()[-1])
    stride = 24 + (3 if "property uchar red" in header else 0)
    faces_at = end + nv * stride
    assert len(data) == faces_at + nf * 13, "unexpected PLY size"
    return {"ply_bytes": len(data), "ply_header_and_vertices_sha256": hashlib.sha256(data[:faces_at]).hexdigest(),
            "ply_faces_sorted_sha256": hashlib.sha256(_sorted_records(data[faces_at:], 13)).hexdigest()}


def stl_file_digests(path):
    with open(path, "rb") as f:
        data = f.read()
    count = int(np.frombuffer(data[80:84], "<u4")[0])
    assert len(data) == 84 + 50 * count, "unexpected STL size"
    return {"stl_bytes": len(data), "stl_header_sha256": hashlib.sha256(data[:84]).hexdigest(),
            "stl_triangles_sorted_sha256": hashlib.sha256(_sorted_records(data[84:], 50)).hexdigest()}


def vox_file_digests(path):
    """MagicaVoxel: 'VOX ' version, then chunks (id, content bytes, children bytes, content, children).  The voxel
    list of every XYZI chunk is sorted; all other bytes stay in place."""
    with open(path, "rb") as f:
        data = bytearray(f.read())
    assert data[:4] == b"VOX "
    voxels = 0

    def walk(pos, end):
        nonlocal voxels
        while pos < end:
            cid = bytes(data[pos:pos + 4])
            n, m = (int(x) for x in np.frombuffer(bytes(data[pos + 4:pos + 12]), "<u4"))
            body = pos + 12
            if cid == b"XYZI":
                k = int(np.frombuffer(bytes(data[body:body + 4]), "<u4")[0])
                data[body + 4:body + 4 + 4 * k] = _sorted_records(bytes(data[body + 4:body + 4 + 4 * k]), 4)
                voxels += k
            walk(body + n, body + n + m)
            pos = body + n + m

    walk(8, len(data))
    return {"vox_bytes": len(data), "vox_voxels": voxels, "vox_canonical_sha256": hashlib.sha256(bytes(data)).hexdigest()}


# ---- layer-by-layer comparison with a tests/golden/slices_<workload>.json fixture (tests/golden/make_slices.py) -----
# Used by tests/test_gpu_bench_parity.py and by bench.py's `parity` key (outside the timed region).

def canonical_bits(a):
    """float32 -> uint32 bit patterns with -0 -> +0 and every NaN -> 0x7fc00000 (oracle/ref_tool.cpp CanonicalBits)."""
    a = np.ascontiguousarray(a, np.float32)
    bits = a.view(np.uint32).copy()
    bits[a == 0] = 0
    bits[np.isnan(a)] = 0x7FC00000
    return bits


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def layer_report(positions, normals, colors, triangles, fixture):
    """Cuts a whole-grid mesh into cell layers by the reference's counts and compares every layer's digests.
    Returns (layers, vertices compared, triangles compared, [(k, what) that differ])."""
    pos = canonical_bits(positions)
    nrm = canonical_bits(normals) if normals is not None else None
    tri = np.ascontiguousarray(triangles, np.uint32)
    bad = []
    v = t = 0
    if len(pos) != fixture["vertices"] or len(tri) != fixture["triangles"]:
        return len(fixture["layers"]), 0, 0, [(-1, "counts: %d / %d vertices, %d / %d triangles" % (len(pos), fixture["vertices"], len(tri), fixture["triangles"]))]
    for k, nv, nt, hp, hn, hc, ht in fixture["layers"]:
        if sha16(pos[v:v + nv]) != hp:
            bad.append((k, "positions"))
        if nrm is not None and fixture["attributes"] and sha16(nrm[v:v + nv]) != hn:
            bad.append((k, "normals"))
        if colors is not None and fixture["has_color"] and sha16(colors[v:v + nv]) != hc:
            bad.append((k, "colours"))
        if sha16(tri[t:t + nt]) != ht:
            bad.append((k, "triangles"))
        v += nv
        t += nt
    return len(fixture["layers"]), v, t, bad
