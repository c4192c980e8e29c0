"""Where does the end-to-end time go?  (diagnostic)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
tree = T.Tree.load("tests/golden/models/seaside_town.tgm")
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(10 / 1022))
ctx = T.Context(0)
model = T.Model(ctx, tree)
for it in range(6):
    t0 = time.perf_counter(); model.upload(); ctx.synchronize(); t1 = time.perf_counter()
    mesh = model.export_mesh(grid); t2 = time.perf_counter()
    tm = mesh.timings
    mesh.close(); t3 = time.perf_counter()
    print("upload %.2f ms  export %.2f ms (device %.2f, download %.2f)  close %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, tm["total_device_ms"], tm["download_ms"], (t3 - t2) * 1e3))
for it in range(3):
    t1 = time.perf_counter()
    mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY); t2 = time.perf_counter()
    tm = mesh.timings
    mesh.close()
    print("device-only export %.2f ms (device %.2f)" % ((t2 - t1) * 1e3, tm["total_device_ms"]))
