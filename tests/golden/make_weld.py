#!/usr/bin/env python
"""Golden welds from the REFERENCE's MeshGenerator (tangerine/mesh_generators.cpp, `tangerine_ref weld`): seeded vertex
streams with many repeats, signed zeros and near-equal values, and the triangle soup of a real mesh.  Output: weld.npz."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def cases():
    rng = np.random.default_rng(20231)
    lattice = rng.integers(-3, 4, (300, 3)).astype(np.float32) * np.float32(0.5)
    stream = lattice[rng.integers(0, len(lattice), 20000)].copy()
    stream[::7] *= np.float32(-1.0)                       # -0.0 where a component is zero
    yield "lattice", stream
    near = rng.random((500, 3), dtype=np.float32)
    near = np.concatenate([near, np.nextafter(near, np.float32(2.0)), near])   # one-ulp neighbours must stay apart
    yield "near", near[rng.permutation(len(near))]
    # the triangle soup of a mesh: vertices of the reference's own export, three per triangle
    tmp = tempfile.mkdtemp()
    ply = os.path.join(tmp, "m.ply")
    O.ref_run("export", O.model_path("basic_thing"), 8, 0, ply)
    mesh = O.read_ply(ply)
    yield "soup", mesh["pos"][mesh["tris"].reshape(-1)].astype(np.float32)
    yield "empty", np.zeros((0, 3), np.float32)
    yield "single", np.array([[1.0, -0.0, 3.0]], np.float32)


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/tangerine_ref not built (make -C oracle ref)")
    tmp = tempfile.mkdtemp()
    arrays = {}
    for name, stream in cases():
        vertices, indices = O.ref_weld(stream, tmp)
        arrays[name + "/in"], arrays[name + "/vertices"], arrays[name + "/indices"] = stream, vertices, indices
        print(name, len(stream), "->", len(vertices), flush=True)
    np.savez_compressed(os.path.join(HERE, "weld.npz"), **arrays)


if __name__ == "__main__":
    main()
