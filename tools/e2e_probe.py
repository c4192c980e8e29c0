"""Where does the end-to-end time go?  (diagnostic)   python tools/e2e_probe.py [chunks ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
tree = T.Tree.load("tests/golden/models/seaside_town.tgm")
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(10 / 1022))
ctx = T.Context(0)
model = T.Model(ctx, tree)
for chunks in (sys.argv[1:] or ["1", "4", "8", "16"]):
    os.environ["TG_PIPELINE_CHUNKS"] = chunks
    for it in range(5):
        t0 = time.perf_counter(); model.upload(); ctx.synchronize(); t1 = time.perf_counter()
        mesh = model.export_mesh(grid); t2 = time.perf_counter()
        tm = mesh.timings
        mesh.close(); t3 = time.perf_counter()
        if it >= 2:
            print("chunks %s: upload %.2f ms  export %.2f ms (device sum %.2f: cull %.2f eval %.2f scan %.2f faces %.2f attr %.2f; host-side collect %.2f; launches %d)  close %.2f ms" % (
                chunks, (t1 - t0) * 1e3, (t2 - t1) * 1e3, tm["total_device_ms"], tm["cull_ms"], tm["evaluate_ms"], tm["compact_ms"], tm["faces_ms"], tm["attributes_ms"],
                tm["download_ms"], tm["kernel_launches"], (t3 - t2) * 1e3), flush=True)
os.environ["TG_PIPELINE_CHUNKS"] = "1"
for it in range(3):
    t1 = time.perf_counter()
    mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY); t2 = time.perf_counter()
    tm = mesh.timings
    mesh.close()
    print("device-only export %.2f ms (device %.2f)" % ((t2 - t1) * 1e3, tm["total_device_ms"]))
