"""Parity at the grid sizes bench.py measures (BASELINE.json configs), layer by layer against the reference.

The fixtures tests/golden/slices_<workload>.json come from the reference's own thunks run over the WHOLE grid
(tests/golden/make_slices.py -> oracle/_ref/tangerine_ref slices): per cell layer, counts and digests of positions,
normals, colours and triangle indices in the serial (k, j, i) order.  Here the CUDA export of the same grid -- the same
call bench.py times -- is cut into layers by the reference's counts and every layer's digest must agree, i.e. the whole
mesh is bit-identical (up to the sign of zero and NaN payloads, the equality tests/golden_util.same_floats defines).
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import tangerine_b200 as T
from golden_util import canonical_bits, layer_report

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixture(workload):
    path = os.path.join(HERE, "golden", "slices_%s.json" % workload)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % os.path.basename(path))
    with open(path) as f:
        return json.load(f)


def workload_tree(name):
    if name.startswith("synthetic:"):
        return T.Tree.synthetic(int(name.split(":")[1]), 1234)
    return T.Tree.load(O.model_path(name))


@pytest.fixture(scope="module")
def context():
    ctx = T.Context(0)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("workload", ["basic66", "gear512", "colorcube512", "seaside512", "seaside1024", "synthetic256", "synthetic512", "synthetic1024"])
def test_benched_grid_matches_reference_layer_by_layer(workload, context):
    fx = load_fixture(workload)
    tree = workload_tree(fx["model"])
    model = T.Model(context, tree)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32(fx["step"]))
    assert list(grid.shape) == fx["grid"]
    mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS, refine=0)
    assert mesh.vertex_count == fx["vertices"]
    assert mesh.triangle_count == fx["triangles"]
    layers, v, t, bad = layer_report(mesh.positions, mesh.normals, mesh.colors, mesh.triangles, fx)
    assert (v, t) == (fx["vertices"], fx["triangles"])
    assert not bad, "layers that differ from the reference: %s" % bad[:10]
    mesh.close()
    model.close()


def test_gear512_refine5_against_oracle(context):
    """BASELINE.json configs[1]: gear.lua 512x512x34 WITH refinement.  The reference's mesh export never refines
    (export.cpp:320-381), so the refined positions / normals are checked against the C oracle's restatement of the
    point-cloud refinement loop (export.cpp:433-469) on a strided vertex sample, and the unrefined mesh of the same
    grid against the reference (test above)."""
    name, step = "gear", np.float32(8.0 / 510.0)
    tree = T.Tree.load(O.model_path(name))
    model = T.Model(context, tree)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, step)
    plain = model.export_mesh(grid, refine=0)
    refined = model.export_mesh(grid, refine=5)
    assert refined.vertex_count == plain.vertex_count and refined.triangle_count == plain.triangle_count
    assert np.array_equal(refined.triangles, plain.triangles)
    om = O.Model(name)
    oc = O.Octree(om)
    pick = np.arange(0, plain.vertex_count, 37)
    want = oc.refine(plain.positions[pick].copy(), [step / 2] * 3, 5)
    got = refined.positions[pick]
    assert np.abs(got - want).max() <= 1e-3 * step          # north-star bound: Hausdorff <= 1e-3 of the grid step
    assert np.array_equal(canonical_bits(got), canonical_bits(want))
    assert np.array_equal(canonical_bits(refined.normals[pick]), canonical_bits(oc.gradient(want)))
    plain.close()
    refined.close()
    model.close()
