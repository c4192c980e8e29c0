import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import tangerine_b200 as T
name, step = sys.argv[1], float(sys.argv[2])
tree = T.Tree.load("tests/golden/models/%s.tgm" % name)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
ctx = T.Context(0)
model = T.Model(ctx, tree)
for flush in (0, 1):
    for refine in (0, 5):
        for it in range(5):
            if flush: ctx.flush_l2()
            t1 = time.perf_counter()
            mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY, refine=refine); t2 = time.perf_counter()
            tm = mesh.timings
            mesh.close(); t3 = time.perf_counter()
        print("%s flush %d refine %d: export %.3f ms device %.3f (cull %.3f eval %.3f scan %.3f faces %.3f attr %.3f)" % (name, flush, refine, (t2 - t1) * 1e3, tm["total_device_ms"], tm["cull_ms"], tm["evaluate_ms"], tm["compact_ms"], tm["faces_ms"], tm["attributes_ms"]), flush=True)
