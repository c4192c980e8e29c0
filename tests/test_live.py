"""The live mesher's path (Sodapop, sodapop.cpp:227-247, 562-760), CPU side: the C oracle against the REFERENCE's outputs.

tests/golden/live.npz and slices_live_<case>.json come from oracle/_ref/tangerine_ref (`eval live`, `eval live-gradient`,
`slices-live`; tests/golden/make_live.py).  The oracle restates the two-step octree construction (Create with Coalesce =
false, MaxDepth = 3, then Populate of the incomplete nodes), the inexact descent with the +-100 clamp
(sodapop.cpp:583-587) and NaiveSurfaceNetsScratch's grid (:153-179); here it is pinned to the reference bit for bit.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from golden_util import layer_report, same_floats

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = ["basic_thing", "seaside_town", "gear", "kitchen_sink", "stencil_test", "cones", "scale", "flower", "color-cube"]


def live_fixture(case):
    with open(os.path.join(HERE, "golden", "slices_live_%s.json" % case)) as f:
        return json.load(f)


def grid_bits(grid):
    return [int(np.float32(v).view(np.uint32)) for v in (grid.x, grid.y, grid.z)], [int(np.float32(v).view(np.uint32)) for v in (grid.dx, grid.dy, grid.dz)]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "live.npz"))


@pytest.mark.parametrize("name", MODELS)
def test_oracle_live_field_and_gradient_match_the_reference(name, golden):
    pts = np.load(os.path.join(HERE, "golden", name + ".npz"))["points"]
    oc = O.Octree(O.Model(name), live=True)
    assert same_floats(oc.eval(pts), golden[name + "/live"])
    assert same_floats(oc.gradient(pts), golden[name + "/live_gradient"])
    # the clamp is visible: nothing outside +-100, and the empty octants are exactly +100
    live = golden[name + "/live"]
    assert np.nanmax(live) <= 100.0 and np.nanmin(live) >= -100.0 and (live == 100.0).any()


@pytest.mark.parametrize("case", ["basic20", "gear20", "kitchen20", "stencil20", "seaside20"])
def test_oracle_live_mesh_matches_the_reference_layer_by_layer(case):
    fx = live_fixture(case)
    oc = O.Octree(O.Model(fx["model"]), live=True)
    grid = oc.live_grid(fx["density"])
    assert list(grid.shape) == fx["grid"]
    origin, step = grid_bits(grid)
    assert origin == fx["live_grid"]["origin_bits"] and step == fx["live_grid"]["step_bits"]
    # the whole grid is walked, the reference only walks its point cache: outside of it the clamped field is +100
    verts, cells, tris = oc.surface_nets(grid)
    layers, v, t, bad = layer_report(verts, oc.gradient(verts), None, tris, fx)
    assert (v, t) == (fx["vertices"], fx["triangles"]) and not bad, bad[:10]


@pytest.fixture(scope="module")
def live_octrees():
    with open(os.path.join(HERE, "golden", "live_octree.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", MODELS + ["synthetic200"])
def test_live_octree_of_library_and_oracle_match_the_reference(name, live_octrees):
    """The host octree builder with coalesce = false (tg_model_create_live's octree, no device needed) and the oracle's
    two-step construction against `tangerine_ref info-live`: nodes, leaves, program words, hash and the octree's Bounds."""
    import tangerine_b200 as T
    info = live_octrees[name]
    for threads in (1, 4):
        s = T.Tree.load(O.model_path(name)).octree_stats(0.25, threads, live=True)
        assert s["octree_hash"] == info["octree_hash"]
        assert (s["octree_nodes"], s["octree_leaves"], s["reference_words"]) == (info["octree_nodes"], info["octree_leaves"], info["octree_words"])
        assert [np.float32(v) for v in s["bounds_min"]] == [np.float32(v) for v in info["bounds_min"]]
        assert [np.float32(v) for v in s["bounds_max"]] == [np.float32(v) for v in info["bounds_max"]]
    oc = O.Octree(O.Model(name), live=True)
    stats = oc.stats()
    assert stats["hash"] == info["octree_hash"] and stats["nodes"] == info["octree_nodes"]
    lo, hi = oc.bounds()
    assert [np.float32(v) for v in lo] == [np.float32(v) for v in info["bounds_min"]] and [np.float32(v) for v in hi] == [np.float32(v) for v in info["bounds_max"]]
