"""TG_MESH_FAST: the opt-in fast build of the brick kernel (tangerine_b200/csrc/tg_fast.cu) against the exact one.

The default path is bit-identical to the reference (every other GPU test).  The fast path trades that for speed -- FMA
contraction, approximate sqrt / division, float where the reference promotes to double -- and has to stay inside the
tolerances BASELINE.json's north_star states for the whole project:
  * raw SDF samples within 1e-5 relative / 4 ULP,
  * identical cell sign classification except where |d| < 1e-6,
  * identical face counts on the test models,
  * vertex Hausdorff distance <= 1e-3 of the grid step after refinement.
"Relative" is taken against the larger of |d| and the scale of the coordinates that enter the distance (d = |p - c| - r
cancels to zero at the surface, so an error relative to d itself is unbounded for any arithmetic, the reference's
included); the bound used is 1e-5 * max(|d|, 1) on models whose coordinates are of order 1 .. 10.
"""
import numpy as np
import pytest

import oracle_lib as O
import tangerine_b200 as T
from golden_util import ulp_diff

pytestmark = pytest.mark.gpu
MODELS = ["basic_thing", "kitchen_sink", "gear", "color-cube", "seaside_town", "cones", "synthetic200"]


@pytest.fixture(scope="module")
def context():
    ctx = T.Context(0)
    yield ctx
    ctx.close()


def _grid(tree, cells_per_unit):
    lo, hi = tree.bounds()
    return T.export_grid(lo, hi, np.float32(1.0 / cells_per_unit))


@pytest.mark.parametrize("name", MODELS)
def test_fast_samples_within_tolerance(name, golden, context):
    tree = T.Tree.load(O.model_path(name))
    model = T.Model(context, tree)
    grid = _grid(tree, golden[name]["cells_per_unit"] * 2)
    exact, _ = model.eval_lattice(grid)
    fast, _ = model.eval_lattice(grid, flags=T.MESH_FAST)
    ok = np.isfinite(exact)
    err = np.abs(fast[ok] - exact[ok])
    bound = 1e-5 * np.maximum(np.abs(exact[ok]), 1.0)
    within = (err <= bound) | (ulp_diff(fast[ok], exact[ok]) <= 4)
    assert within.all(), "max error %.3g at |d| = %.3g" % (err[~within].max(), np.abs(exact[ok][~within]).min())
    # cell sign classification: is_scalar_positive is `d >= 0` (surface_nets.cpp:733-735)
    flips = (fast[ok] >= 0) != (exact[ok] >= 0)
    assert np.abs(exact[ok][flips]).max(initial=0.0) < 1e-6
    model.close()


@pytest.mark.parametrize("name", MODELS)
def test_fast_mesh_counts_and_hausdorff(name, golden, context):
    from scipy.spatial import cKDTree
    tree = T.Tree.load(O.model_path(name))
    model = T.Model(context, tree)
    grid = _grid(tree, golden[name]["cells_per_unit"] * 2)
    step = float(grid.dx)
    exact = model.export_mesh(grid, refine=5)
    fast = model.export_mesh(grid, refine=5, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_FAST)
    assert fast.triangle_count == exact.triangle_count      # identical face counts on the test models
    assert fast.vertex_count == exact.vertex_count
    a, b = exact.positions, fast.positions
    ok = np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1)
    # same cells in the same order: the distance vertex by vertex bounds the Hausdorff distance from above
    d = np.linalg.norm(a[ok] - b[ok], axis=1)
    if d.max() > 1e-3 * step:
        # the one-to-one bound was not enough somewhere: the true (two-sided) Hausdorff distance
        d = np.maximum(cKDTree(a[ok]).query(b[ok])[0].max(), cKDTree(b[ok]).query(a[ok])[0].max())
        assert d <= 1e-3 * step
    exact.close()
    fast.close()
    model.close()
