"""TG_MESH_FAST against the exact build: sample errors, sign flips, mesh counts, vertex distances.   python tools/fast_probe.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tangerine_b200 as T, oracle_lib as O
from golden_util import ulp_diff
from scipy.spatial import cKDTree
golden = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "manifest.json")))
ctx = T.Context(0)
for name in ["basic_thing", "kitchen_sink", "gear", "color-cube", "seaside_town", "cones", "synthetic200"]:
    tree = T.Tree.load(O.model_path(name)); model = T.Model(ctx, tree)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32(1.0 / (golden[name]["cells_per_unit"] * 2)))
    step = float(grid.dx)
    ex, _ = model.eval_lattice(grid); fa, _ = model.eval_lattice(grid, flags=T.MESH_FAST)
    ok = np.isfinite(ex)
    err = np.abs(fa[ok] - ex[ok]); rel = err / np.maximum(np.abs(ex[ok]), 1.0)
    flips = (fa[ok] >= 0) != (ex[ok] >= 0)
    near = int((np.abs(ex[ok]) < 1e-6).sum())
    print("%-13s samples %9d  max abs err %.3g  max err/max(|d|,1) %.3g  max ulp %d  flips %d (max |d| at a flip %.3g)  |d|<1e-6: %d" % (
        name, ok.sum(), err.max(), rel.max(), int(ulp_diff(fa[ok], ex[ok]).max()), flips.sum(), np.abs(ex[ok][flips]).max(initial=0), near))
    for refine in (0, 5):
        a = model.export_mesh(grid, refine=refine); b = model.export_mesh(grid, refine=refine, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_FAST)
        pa, pb = a.positions[np.isfinite(a.positions).all(1)], b.positions[np.isfinite(b.positions).all(1)]
        d1 = cKDTree(pa).query(pb)[0]; d2 = cKDTree(pb).query(pa)[0]
        d = np.concatenate([d1, d2]) / step
        print("    refine %d: V %d / %d  F %d / %d   nearest-vertex distance in steps: max %.3g  99.9%% %.3g  99%% %.3g  median %.3g" % (
            refine, a.vertex_count, b.vertex_count, a.triangle_count, b.triangle_count, d.max(), np.quantile(d, 0.999), np.quantile(d, 0.99), np.median(d)))
        a.close(); b.close()
    model.close()
