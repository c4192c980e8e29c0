"""CPU-only tests of the product's host side: the C-ABI library (symbols, tree builders, host octree + flattener
against the reference's golden octree hashes, writers) and the N>1 slab logic over a world_size-2 gloo group.

No compute entry point is called here: without a CUDA device those fail loudly (also checked).
"""
import ctypes as C
import os
import re
import socket
import struct

import numpy as np
import pytest

import oracle_lib as O
import tangerine_b200 as T
from tangerine_b200 import slabs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = ["basic_thing", "gear", "color-cube", "seaside_town", "kitchen_sink", "stencil_test", "cones", "scale", "flower", "synthetic200"]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tangerine_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"TG_API\s+[^;(]*?\b(tg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 60, names
    raw = C.CDLL(T.library_path())
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, "declared in include/tangerine_b200.h but not exported: %s" % missing
    # and the ctypes binding covers the same set (so the GPU tests really go through the C ABI)
    bound = set(T.lib()._tg_signatures)
    assert set(names) <= bound | {"tg_tree_eval"}, sorted(set(names) - bound)


def test_no_device_fails_loudly():
    """No CPU fallback: creating a context without a CUDA device is an error, not a silent host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(T.TangerineError) as e:
        T.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


@pytest.mark.parametrize("name", MODELS)
def test_host_octree_and_flattener_match_reference(name, golden):
    """tg_tree.cpp / tg_octree.cpp (SDFOctree::Create + Populate + Clip, sdf_evaluator.cpp:1609-1783, 782-850):
    node count and the hash over every node's pruned ProgramBuffer words equal the reference's."""
    tree = T.Tree.load(O.model_path(name))
    info = golden[name]["info"]
    s = tree.octree_stats()
    assert s["octree_hash"] == info["octree_hash"]
    assert s["octree_nodes"] == info["octree_nodes"]
    lo, hi = tree.bounds()
    assert np.allclose(lo, info["bounds_min"]) and np.allclose(hi, info["bounds_max"])
    assert tree.has_paint() == info["has_paint"]
    assert tree.leaf_count() == info["leaf_count"]


def test_tree_eval_matches_oracle_on_host():
    """SDFNode::Eval of the host tree (used by the octree build's Clip) against the oracle, bit for bit."""
    rng = np.random.default_rng(7)
    for name in ["kitchen_sink", "cones", "scale"]:
        tree = T.Tree.load(O.model_path(name))
        om = O.Model(name)
        lo, hi = tree.bounds()
        pts = (lo + (hi - lo) * rng.random((512, 3))).astype(np.float32)
        want = om.eval_tree(pts) if hasattr(om, "eval_tree") else None
        got = np.array([tree.eval(*map(float, p)) for p in pts], np.float32)
        if want is not None:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.all(np.isfinite(got))


def test_export_grid_follows_mesh_export_thread():
    """export.cpp:324-337: ModelMin -= 2 * Step; Extent = ceil((Max - Min) / Step)."""
    g = T.export_grid([-2, -2, -2], [2, 2, 2], np.float32(1 / 16))
    og = O.export_grid([-2, -2, -2], [2, 2, 2], np.float32(1 / 16))
    assert g.shape == (66, 66, 66) == og.shape
    assert (g.x, g.y, g.z, g.dx) == (og.x, og.y, og.z, og.dx)


def test_synthetic_tree_is_deterministic():
    """Config C4: mt19937(1234)-driven random CSG; same seed, same tree, on every build."""
    a = T.Tree.synthetic(200, 1234)
    b = T.Tree.synthetic(200, 1234)
    c = T.Tree.synthetic(200, 99)
    assert a.leaf_count() == b.leaf_count() == 201  # 200 brushes + the clipping box
    pts = np.random.default_rng(3).uniform(-5, 5, (64, 3)).astype(np.float32)
    va = [a.eval(*map(float, p)) for p in pts]
    assert va == [b.eval(*map(float, p)) for p in pts]
    assert va != [c.eval(*map(float, p)) for p in pts]
    sa, sb = a.octree_stats(), b.octree_stats()
    assert sa["octree_hash"] == sb["octree_hash"] and sa["octree_nodes"] > 1


# ---- slab logic ----------------------------------------------------------------------------------

def test_balanced_slabs_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for sz in (64, 66, 1024, 1000):
            nb = (sz + 7) // 8
            profile = rng.integers(0, 1000, nb)
            cut = slabs.balanced_slabs(profile, world, sz)
            assert len(cut) == world and cut[0][0] == 0 and cut[-1][1] == sz
            for r, (a, b) in enumerate(cut):
                assert a < b and a % 8 == 0 and (b % 8 == 0 or b == sz)
                if r:
                    assert a == cut[r - 1][1]
    # equal work when the profile allows it
    cut = slabs.balanced_slabs(np.ones(128), 8, 1024)
    assert all(abs((b - a) - 128) <= 8 for a, b in cut), cut
    # bottom-heavy profile: the first slab is thinner
    prof = np.r_[np.full(16, 100.0), np.full(112, 1.0)]
    cut = slabs.balanced_slabs(prof, 2, 1024)
    assert cut[0][1] < 512
    with pytest.raises(ValueError):
        slabs.balanced_slabs(np.ones(2), 4, 16)
    # finer alignment: cuts on any layer, closer to equal work
    prof = rng.integers(1, 1000, 128).astype(np.float64)
    for world in (2, 4, 8):
        coarse = slabs.balanced_slabs(prof, world, 1024, align=8)
        fine = slabs.balanced_slabs(prof, world, 1024, align=1)
        assert fine[0][0] == 0 and fine[-1][1] == 1024 and all(a < b for a, b in fine)
        assert all(fine[r][0] == fine[r - 1][1] for r in range(1, world))
        per_layer = np.repeat(prof / 8.0, 8)

        def spread(cut):
            w = [per_layer[a:b].sum() for a, b in cut]
            return max(w) / (sum(w) / len(w))
        assert spread(fine) <= spread(coarse) + 1e-9
        assert spread(fine) < 1.02


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slab_worker(rank, world, port, cuts, verts, cells, tris, shape, out_dir):
    """One rank of the N>1 path on CPU: owns the vertices / triangles of its z-slab in LOCAL numbering
    (halo layer first, then subtracted -- exactly what tg_export_mesh(slab) returns), exchanges the counts over
    gloo and rebases."""
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sx, sy, sz = shape
    k = cells // (sx * sy)
    k0, k1 = cuts[rank]
    own = (k >= k0) & (k < k1)
    first = int(np.argmax(own)) if own.any() else 0          # vertices are (k, j, i)-sorted: a slab is a contiguous run
    halo = int(((k == k0 - 1)).sum()) if k0 > 0 else 0
    # triangles are owned by the cell of their first index
    tri_own = own[tris[:, 0]]
    local = (tris[tri_own].astype(np.int64) - first).astype(np.uint32)   # local - halo numbering, may wrap below zero
    assert halo == 0 or (tris[tri_own].min() >= first - halo)
    base, total_v, total_f, per_rank = slabs.exchange_counts(int(own.sum()), int(tri_own.sum()), rank, world)
    assert base == first
    np.savez(os.path.join(out_dir, "part%d.npz" % rank), positions=verts[own], triangles=local, base=base,
             totals=np.array([total_v, total_f]), per_rank=per_rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_over_gloo(world, tmp_path):
    """world_size-2 (and 3) gloo run of the only collective on the path: per-slab counts -> index bases -> the
    stitched mesh equals the whole mesh (oracle surface nets on basic_thing)."""
    import torch.multiprocessing as mp
    om = O.Model("basic_thing")
    oc = O.Octree(om)
    grid = O.export_grid([-2, -2, -2], [2, 2, 2], np.float32(1 / 8))
    verts, cells, tris = oc.surface_nets(grid)
    assert len(verts) > 1000
    sz = grid.shape[2]
    k = cells // (grid.shape[0] * grid.shape[1])
    profile = np.bincount(k // 8, minlength=(sz + 7) // 8)
    cuts = slabs.balanced_slabs(profile, world, sz)
    port = _free_port()
    mp.spawn(_slab_worker, args=(world, port, cuts, verts, cells, tris.astype(np.int64), grid.shape, str(tmp_path)), nprocs=world, join=True)
    parts = []
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "part%d.npz" % r))
        assert tuple(z["totals"]) == (len(verts), len(tris))
        parts.append((z["positions"], z["triangles"], int(z["base"])))
    pos, tri = slabs.stitch(parts)
    assert np.array_equal(pos, verts)
    assert np.array_equal(tri, tris.astype(np.uint32))


# ---- writers (byte layouts of export.cpp:60-108, 198-280 and VoxWriter) ------------------------------

def _raw_mesh(positions, normals, colors, triangles, face_normals=None):
    from tangerine_b200.api import _Mesh
    m = _Mesh()
    keep = [np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(normals, np.float32),
            None if colors is None else np.ascontiguousarray(colors, np.uint8), np.ascontiguousarray(triangles, np.uint32),
            None if face_normals is None else np.ascontiguousarray(face_normals, np.float32)]
    m.positions = keep[0].ctypes.data_as(C.POINTER(C.c_float))
    m.normals = keep[1].ctypes.data_as(C.POINTER(C.c_float))
    if keep[2] is not None:
        m.colors = keep[2].ctypes.data_as(C.POINTER(C.c_uint8))
    m.triangles = keep[3].ctypes.data_as(C.POINTER(C.c_uint32))
    if keep[4] is not None:
        m.face_normals = keep[4].ctypes.data_as(C.POINTER(C.c_float))
    m.vertex_count = len(keep[0])
    m.triangle_count = len(keep[3])
    return m, keep


@pytest.mark.parametrize("with_color", [False, True])
def test_write_ply_layout(with_color, tmp_path):
    rng = np.random.default_rng(1)
    pos = rng.random((5, 3), np.float32)
    nrm = rng.random((5, 3), np.float32)
    col = rng.integers(0, 255, (5, 3)).astype(np.uint8) if with_color else None
    tri = np.array([[0, 1, 2], [2, 3, 4]], np.uint32)
    m, keep = _raw_mesh(pos, nrm, col, tri)
    path = str(tmp_path / "a.ply")
    assert T.lib().tg_write_ply(os.fsencode(path), C.byref(m)) == 0
    back = O.read_ply(path)
    assert np.array_equal(back["pos"], pos) and np.array_equal(back["normal"], nrm)
    assert np.array_equal(back["tris"], tri)
    assert (back["color"] is None) == (not with_color)
    if with_color:
        assert np.array_equal(back["color"], col)
    data = open(path, "rb").read()
    header, _, body = data.partition(b"end_header\n")
    assert header.startswith(b"ply\nformat binary_little_endian 1.0\n")
    assert b"element vertex 5\n" in header and b"element face 2\n" in header
    assert b"property list uchar uint vertex_indices\n" in header and b"comment Created by Tangerine\n" in header  # export.cpp:198-238
    assert len(body) == 5 * (24 + (3 if with_color else 0)) + 2 * 13


def test_write_stl_layout(tmp_path):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    nrm = np.zeros((4, 3), np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    fn = np.array([[0, 0, 1], [1, 0, 0]], np.float32)
    m, keep = _raw_mesh(pos, nrm, None, tri, fn)
    path = str(tmp_path / "a.stl")
    assert T.lib().tg_write_stl(os.fsencode(path), C.byref(m)) == 0
    data = open(path, "rb").read()
    assert len(data) == 80 + 4 + 2 * 50
    assert data[:80].rstrip(b"\0 ").startswith(b"STL generated by Tangerine")
    assert struct.unpack("<I", data[80:84])[0] == 2
    rec = struct.unpack("<12fH", data[84:134])
    assert rec[:3] == (0.0, 0.0, 1.0) and rec[3:12] == (0, 0, 0, 1, 0, 0, 0, 1, 0) and rec[12] == 0


def test_cli_info_and_loud_failure_without_a_device(tmp_path, capsys):
    """python -m tangerine_b200: `info` is host-only; `export` has no CPU path to fall back to (exit code 2)."""
    import torch
    from tangerine_b200.__main__ import main
    model = O.model_path("basic_thing")
    assert main(["info", model]) == 0
    assert "primitives 5" in capsys.readouterr().out
    assert main(["export", model, str(tmp_path / "x.obj"), "--grid", "8"]) == 2      # the reference has no OBJ writer either
    if not torch.cuda.is_available():
        assert main(["export", model, str(tmp_path / "x.ply"), "--grid", "8"]) == 2
        assert "no CPU fallback" in capsys.readouterr().err
        assert not (tmp_path / "x.ply").exists()



@pytest.mark.parametrize("name", ["basic_thing", "seaside_town", "gear", "kitchen_sink", "stencil_test", "cones", "scale", "flower", "color-cube", "synthetic200"])
def test_device_tables_do_not_depend_on_threads_or_refactors(name):
    """The five tables tg_model_create uploads (octree nodes, both instruction streams, regions, node ranks), hashed on
    the host: the same on 1, 3 and 8 threads, for the export's octree and the live mesher's, and equal to
    tests/golden/tables.json -- the tables every GPU parity test of this round ran on, so a host-side change of the
    builder or the flattener that alters them shows up without a device.  (The tree stream carries material ids, which
    are handed out per process in order of first use: it is compared across thread counts only.)"""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tables.json")) as f:
        want = json.load(f)[name]
    tree = T.Tree.load(O.model_path(name))
    stable = [0, 1, 3, 4]
    for live in (False, True):
        first = tree.tables_hash(1, live=live)
        assert [first[i] for i in stable] == [want["live" if live else "export"][i] for i in stable]
        for threads in (3, 8):
            assert tree.tables_hash(threads, live=live) == first
