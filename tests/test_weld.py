"""MeshGenerator (tangerine/mesh_generators.cpp:20-80): the vertex weld of the lattice mesher.

tests/golden/weld.npz holds the reference's own answers (`tangerine_ref weld`, tests/golden/make_weld.py).  CPU: the C
oracle against them.  GPU: tg_weld (hash table + prefix scan, tg_engine.cu) against them and, on a large stream and on
the triangle soup of an exported mesh, against the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["lattice", "near", "soup", "empty", "single"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "weld.npz"))


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("case", CASES)
def test_oracle_weld_matches_the_reference(case, golden):
    vertices, indices = O.weld(golden[case + "/in"])
    assert same_bits(vertices, golden[case + "/vertices"].reshape(-1, 4))   # bits of the FIRST occurrence (-0 stays -0)
    assert np.array_equal(indices, golden[case + "/indices"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_weld_matches_the_reference(case, golden):
    import tangerine_b200 as T
    ctx = T.Context(0)
    vertices, indices = ctx.weld(golden[case + "/in"])
    assert same_bits(vertices, golden[case + "/vertices"].reshape(-1, 4))
    assert np.array_equal(indices, golden[case + "/indices"])
    ctx.close()


@pytest.mark.gpu
def test_gpu_weld_of_a_large_stream_and_of_an_exported_mesh():
    import tangerine_b200 as T
    ctx = T.Context(0)
    rng = np.random.default_rng(5)
    distinct = rng.standard_normal((200000, 3)).astype(np.float32)
    distinct[::11, 1] = 0.0
    stream = distinct[rng.integers(0, len(distinct), 3000000)].copy()
    stream[::5] *= np.float32(-1.0)
    vertices, indices = ctx.weld(stream)
    want_vertices, want_indices = O.weld(stream)
    assert same_bits(vertices, want_vertices) and np.array_equal(indices, want_indices)
    # a triangle soup made from an indexed mesh welds back to that mesh's referenced vertices, in first-use order
    tree = T.Tree.load(O.model_path("kitchen_sink"))
    lo, hi = tree.bounds()
    model = T.Model(ctx, tree)
    mesh = model.export_mesh(T.export_grid(lo, hi, np.float32(1.0 / 24.0)), flags=0)
    soup = mesh.positions[mesh.triangles.reshape(-1)]
    vertices, indices = ctx.weld(soup)
    assert np.array_equal(vertices[indices, :3], soup)
    assert len(vertices) <= len(np.unique(mesh.triangles))   # (two cells may put their vertices on the same point)
    want_vertices, want_indices = O.weld(soup)
    assert same_bits(vertices, want_vertices) and np.array_equal(indices, want_indices)
    mesh.close()
    model.close()
    ctx.close()
