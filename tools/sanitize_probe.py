"""Small exports of every kind for compute-sanitizer.   compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
ctx = T.Context(0)
for name, cpu in (("kitchen_sink", 24.0), ("seaside_town", 12.0), ("color-cube", 10.0)):
    tree = T.Tree.load("tests/golden/models/%s.tgm" % name)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32(1.0 / cpu))
    model = T.Model(ctx, tree)
    for chunks in ("1", "3"):
        os.environ["TG_PIPELINE_CHUNKS"] = chunks
        for refine in (0, 2):
            mesh = model.export_mesh(grid, refine=refine)
            print(name, "chunks", chunks, "refine", refine, mesh.vertex_count, mesh.triangle_count, flush=True)
            mesh.close()
    mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_FAST | T.MESH_DEVICE_ONLY)
    print(name, "fast", mesh.vertex_count, flush=True)
    mesh.close()
    pts = (np.random.default_rng(1).random((2000, 3), dtype=np.float32) * (hi - lo) + lo).astype(np.float32)
    for mode in (T.EVAL_OCTREE, T.EVAL_INTERP, T.EVAL_TREE, T.EVAL_GRADIENT, T.EVAL_COLOR):
        model.eval_points(pts, mode)
    rays = np.concatenate([pts, np.tile(np.array([[0, 0, -1]], np.float32), (len(pts), 1))], axis=1)
    model.ray_cast(rays)
    print(name, "long programs", model.check_long_programs(), flush=True)
    cloud = model.export_points(lo, hi, np.float32(1.0 / cpu), refine=2)
    print(name, "cloud", cloud.vertex_count, flush=True)
    cloud.close()
    model.export_voxels(8.0)
    model.close()
    live = T.Model(ctx, tree, live=True)
    mesh = live.export_mesh(live.live_grid(20.0), flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD)
    live.eval_points(pts, T.EVAL_LIVE)
    print(name, "live", mesh.vertex_count, mesh.triangle_count, flush=True)
    mesh.close()
    live.close()
ctx.close()
print("done")
