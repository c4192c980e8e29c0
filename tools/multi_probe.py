"""In-library multi-GPU export: per-rank stage times and the balance of the planned slabs.   python tools/multi_probe.py [workload] [N]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else "seaside1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name, step, refine, desc = bench.WORKLOADS[workload]
tree, _ = bench.load_workload_tree(T, name)
ctx = T.Context(devices=list(range(n)))
model = T.Model(ctx, tree)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
flags = T.MESH_NORMALS | T.MESH_COLORS
for mode in (("device", "host") if not os.environ.get("TG_PROBE_DEVICE_ONLY") else ("device",)):
    iters = int(os.environ.get("TG_PROBE_ITERS", "4"))
    walls = []
    for i in range(iters):
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.timer_begin()
        if mode == "host":
            model.upload()
        m = model.export_mesh(grid, flags=flags | (T.MESH_DEVICE_ONLY if mode == "device" else 0), refine=refine)
        ms = ctx.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        ranks = m.rank_info()
        if i >= 3:
            walls.append(wall)
        if i == iters - 1:
            print("%s: wall ms over %d exports: min %.3f median %.3f max %.3f" % (mode, len(walls), min(walls), sorted(walls)[len(walls) // 2], max(walls)))
            print("%s: V=%d F=%d  device %.3f ms  wall %.3f ms" % (mode, m.vertex_count, m.triangle_count, ms, wall))
            for r, (b, e, t) in enumerate(ranks):
                print("  rank %d  slab [%4d, %4d)  cull %.3f eval %.3f scan %.3f faces %.3f attr %.3f  total %.3f  bricks %d  flops/sample %.0f  ns/brick %.1f" % (
                    r, b, e, t["cull_ms"], t["evaluate_ms"], t["compact_ms"], t["faces_ms"], t["attributes_ms"], t["total_device_ms"], t["bricks_evaluated"],
                    t["algorithmic_flops"] / max(t["samples_evaluated"], 1), t["evaluate_ms"] * 1e6 / max(t["bricks_evaluated"], 1)))
        m.close()
model.close()
ctx.close()
