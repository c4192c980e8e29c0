"""The live mesher's path on the GPU (Sodapop, sodapop.cpp:227-247, 562-760) against the reference.

tg_model_create_live builds the octree the live mesher uses (no coalescing), TG_EVAL_LIVE / TG_MESH_LIVE_FIELD sample its
implicit function -- clamp(SDFOctree::Eval(p, Exact = false), -100, 100), sodapop.cpp:583-587 -- and tg_live_grid is
NaiveSurfaceNetsScratch's grid (:153-179).  The fixtures hold the reference's own values and, per cell layer, digests of
the mesh its vertex and face loops produce over the point cache of the octree leaves (tests/golden/make_live.py).
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import tangerine_b200 as T
from golden_util import layer_report, same_floats

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = ["basic_thing", "seaside_town", "gear", "kitchen_sink", "stencil_test", "cones", "scale", "flower", "color-cube"]


@pytest.fixture(scope="module")
def context():
    ctx = T.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "live.npz"))


def live_fixture(case):
    with open(os.path.join(HERE, "golden", "slices_live_%s.json" % case)) as f:
        return json.load(f)


def bits(*values):
    return [int(np.float32(v).view(np.uint32)) for v in values]


@pytest.mark.parametrize("name", MODELS)
def test_live_field_and_gradient_match_the_reference(name, context, golden):
    pts = np.load(os.path.join(HERE, "golden", name + ".npz"))["points"]
    model = T.Model(context, T.Tree.load(O.model_path(name)), live=True)
    assert same_floats(model.eval_points(pts, T.EVAL_LIVE), golden[name + "/live"])
    assert same_floats(model.eval_points(pts, T.EVAL_GRADIENT), golden[name + "/live_gradient"])
    model.close()


@pytest.mark.parametrize("case", ["basic20", "gear20", "kitchen20", "stencil20", "seaside20", "seaside50"])
def test_live_mesh_matches_the_reference_layer_by_layer(case, context):
    fx = live_fixture(case)
    model = T.Model(context, T.Tree.load(O.model_path(fx["model"])), live=True)
    grid = model.live_grid(fx["density"])
    assert list(grid.shape) == fx["grid"]
    assert bits(grid.x, grid.y, grid.z) == fx["live_grid"]["origin_bits"] and bits(grid.dx, grid.dy, grid.dz) == fx["live_grid"]["step_bits"]
    for extra in (0, T.MESH_NO_CULL):
        mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD | extra, refine=0)
        assert (mesh.vertex_count, mesh.triangle_count) == (fx["vertices"], fx["triangles"])
        layers, v, t, bad = layer_report(mesh.positions, mesh.normals, None, mesh.triangles, fx)
        assert (v, t) == (fx["vertices"], fx["triangles"]) and not bad, "layers that differ from the reference: %s" % bad[:10]
        mesh.close()
    model.close()


def test_live_lattice_matches_the_oracle(context):
    name = "kitchen_sink"
    model = T.Model(context, T.Tree.load(O.model_path(name)), live=True)
    oc = O.Octree(O.Model(name), live=True)
    grid = model.live_grid(20.0)
    oracle_grid = oc.live_grid(20.0)
    assert grid.shape == oracle_grid.shape and bits(grid.x, grid.dx) == bits(oracle_grid.x, oracle_grid.dx)
    mine, _ = model.eval_lattice(grid, flags=T.MESH_LIVE_FIELD)
    assert same_floats(mine, oc.lattice(oracle_grid))
    # and it is not the export's field: empty octants hold +100 instead of the parent's distance
    exact, _ = model.eval_lattice(grid)
    assert (mine == 100.0).any() and not same_floats(mine, exact)
    model.close()


def test_live_grid_needs_a_live_model(context):
    model = T.Model(context, T.Tree.load(O.model_path("basic_thing")))
    with pytest.raises(T.TangerineError):
        model.live_grid(20.0)
    model.close()


def test_live_mesh_on_two_gpus():
    """The live field goes through the same multi-device export (z-slabs, one stitched host mesh)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    fx = live_fixture("seaside50")
    ctx = T.Context(devices=[0, 1])
    model = T.Model(ctx, T.Tree.load(O.model_path(fx["model"])), live=True)
    mesh = model.export_mesh(model.live_grid(fx["density"]), flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD, refine=0)
    layers, v, t, bad = layer_report(mesh.positions, mesh.normals, None, mesh.triangles, fx)
    assert (v, t) == (fx["vertices"], fx["triangles"]) and not bad, bad[:10]
    mesh.close()
    model.close()
    ctx.close()
