#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE implementation itself.

Runs oracle/_ref/tangerine_ref (the reference's own hot-path sources, compiled unmodified by
oracle/Makefile) on the .tgm test models and stores, per model:

  * ``info``      bounds, tree statistics, octree node count and an FNV-1a hash over every octree
                  node's pivot / terminus / child mask / pruned postfix program  (SDFOctree::Create)
  * ``points``    seeded query points, and for each of them the reference's
                  SDFOctree::Eval, SDFNode::Eval (root tree), SDFInterpreter::Eval (root program),
                  SDFOctree::Gradient and export colour bytes (export.cpp:297-312)
  * ``mesh``      for a small export grid: vertex / face counts of the reference's PLY export and
                  SHA-256 digests of the sorted vertex records (position + normal + colour) and of
                  the sorted triangle coordinates, plus a strided sample of the raw vertices
  * ``cloud``     point-cloud export with 5 refinement iterations (export.cpp:384-469): digest + sample

Usage:  python tests/golden/make_golden.py [model ...]   (needs oracle/_ref/tangerine_ref; a few minutes for all)
The .tgm models themselves come from ``tangerine_ref dump-tgm`` on the reference's models/*.lua and
on tests/golden/models_src/*.lua (see tests/golden/README.md).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from golden_util import ply_file_digests, stl_file_digests, vox_file_digests  # noqa: E402

# model -> (cells per unit for the mesh export, points, point-cloud step or None)
CASES = {
    "basic_thing": (16, 4000, 0.125),
    "gear": (16, 4000, None),
    "color-cube": (5, 3000, None),
    "seaside_town": (4, 3000, None),
    "kitchen_sink": (12, 4000, 0.125),
    "stencil_test": (10, 4000, None),
    "cones": (6, 2000, None),
    "scale": (8, 2000, None),
    "flower": (4, 2000, None),
    # config C4 at test size: tangerine_b200.Tree.synthetic(200, 1234).save(...), see tests/golden/README.md
    "synthetic200": (8, 3000, None),
}
# Larger exports whose counts SURVEY.md section 6 recorded from the reference (counts + digests only).
BIG = {"basic_thing": 16, "gear": 32, "color-cube": 10, "seaside_town": 12.8}


VOX_GRID, VOX_COLOR = 6.0, 37


def sort_rows(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(a.T[::-1])]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def vertex_records(pos, normal, color):
    cols = [pos.view(np.uint32), normal.view(np.uint32)]
    if color is not None:
        cols.append(color.astype(np.uint32))
    return sort_rows(np.concatenate(cols, axis=1))


def mesh_summary(ply):
    rec = vertex_records(ply["pos"], ply["normal"], ply["color"])
    tri = sort_rows(ply["pos"][ply["tris"]].reshape(-1, 9).view(np.uint32))
    return {"vertices": int(len(ply["pos"])), "faces": int(len(ply["tris"])),
            "has_color": ply["color"] is not None,
            "vertex_records_sha256": digest(rec), "triangle_coords_sha256": digest(tri),
            "positions_in_order_sha256": digest(ply["pos"])}


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/tangerine_ref not built (make -C oracle ref)")
    tmp = tempfile.mkdtemp()
    manifest = {}
    only = sys.argv[1:]  # optional: regenerate just these models and merge them into the existing manifest
    if only:
        with open(os.path.join(HERE, "manifest.json")) as f:
            manifest = json.load(f)
    for name, (cpu, npts, cloud_step) in CASES.items():
        if only and name not in only:
            continue
        path = O.model_path(name)
        info = json.loads(O.ref_run("info", path))
        info.pop("octree_build_s")
        lo = np.array(info["bounds_min"], np.float32)
        hi = np.array(info["bounds_max"], np.float32)
        rng = np.random.default_rng(sum(map(ord, name)))
        pts = (rng.random((npts, 3), dtype=np.float32) * (hi - lo + np.float32(0.5)) + lo - np.float32(0.25)).astype(np.float32)
        arrays = {"points": pts}
        for mode in ("octree", "tree", "interp", "gradient", "color"):
            arrays[mode] = O.ref_eval(path, mode, pts, tmp)
        ply_path = os.path.join(tmp, name + ".ply")
        O.ref_run("export", path, cpu, 0, ply_path)
        ply = O.read_ply(ply_path)
        entry = {"info": info, "cells_per_unit": cpu, "mesh": mesh_summary(ply)}
        # file-level digests of the reference's writers (export.cpp:60-317, magica.cpp:27-72 + VoxWriter): PLY as above,
        # STL at the same grid, MagicaVoxel at VOX_GRID cells per unit; records whose order the format does not fix
        # (faces, voxels) are hashed as sorted multisets (tests/golden_util.py)
        stl_path = os.path.join(tmp, name + ".stl")
        O.ref_run("export", path, cpu, 0, stl_path)
        vox_path = os.path.join(tmp, name + ".vox")
        O.ref_run("vox", path, VOX_GRID, VOX_COLOR, vox_path)
        entry["files"] = dict(ply_file_digests(ply_path), **stl_file_digests(stl_path), **vox_file_digests(vox_path),
                              vox_grid_size=VOX_GRID, vox_color_index=VOX_COLOR)
        stride = max(1, len(ply["pos"]) // 256)
        arrays["mesh_pos_sample"] = sort_rows(ply["pos"])[::stride]
        if cloud_step:
            pc_path = os.path.join(tmp, name + "_pc.ply")
            O.ref_run("export-grid", path, *lo, *hi, cloud_step, 5, 1, pc_path)
            pc = O.read_ply(pc_path)
            entry["cloud"] = {"step": cloud_step, "refine": 5, "points": int(len(pc["pos"])),
                              "vertex_records_sha256": digest(vertex_records(pc["pos"], pc["normal"], pc["color"]))}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        manifest[name] = entry
        print(name, entry["mesh"]["vertices"], entry["mesh"]["faces"], flush=True)
    for name, cpu in BIG.items():
        if only and name not in only:
            continue
        ply_path = os.path.join(tmp, name + "_big.ply")
        O.ref_run("export", O.model_path(name), cpu, 0, ply_path)
        manifest[name]["mesh_big"] = dict(mesh_summary(O.read_ply(ply_path)), cells_per_unit=cpu)
        print(name, "big", manifest[name]["mesh_big"]["vertices"], manifest[name]["mesh_big"]["faces"], flush=True)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
