"""Static slab plan against the constants of the host-side cost estimate.   python tools/plan_sweep.py [workload] N "constant:straddle" ...
(TG_PLAN_CONSTANT / TG_PLAN_STRADDLE are read at every export and force a new plan.)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
import bench
workload = sys.argv[1]
n = int(sys.argv[2])
specs = sys.argv[3:]
name, step, refine, desc = bench.WORKLOADS[workload]
tree, _ = bench.load_workload_tree(T, name)
ctx = T.Context(devices=list(range(n)))
model = T.Model(ctx, tree)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
flags = T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY
for spec in specs:
    constant, straddle = spec.split(":")
    os.environ["TG_PLAN_CONSTANT"], os.environ["TG_PLAN_STRADDLE"] = constant, straddle
    times, totals = [], None
    for i in range(12):
        ctx.flush_l2()
        ctx.timer_begin()
        m = model.export_mesh(grid, flags=flags, refine=refine)
        ms = ctx.timer_end()
        if i >= 4:
            times.append(ms)
            totals = [t["total_device_ms"] for _, _, t in m.rank_info()]
            cuts = [b for b, _, _ in m.rank_info()]
        m.close()
    times.sort()
    print("N=%d constant %-5s straddle %-4s step median %.3f min %.3f  rank totals max %.3f min %.3f  cuts %s" % (n, constant, straddle, times[len(times) // 2], times[0], max(totals), min(totals), cuts), flush=True)
model.close()
ctx.close()
