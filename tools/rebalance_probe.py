"""TG_MESH_REBALANCE: per-rank totals export after export.   python tools/rebalance_probe.py [workload] [N] [rounds]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else "seaside1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
name, step, refine, desc = bench.WORKLOADS[workload]
tree, _ = bench.load_workload_tree(T, name)
ctx = T.Context(devices=list(range(n)))
model = T.Model(ctx, tree)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
flags = T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY
for i in range(rounds):
    ctx.flush_l2(); ctx.timer_begin()
    m = model.export_mesh(grid, flags=flags | (T.MESH_REBALANCE if i >= 2 else 0), refine=refine)
    ms = ctx.timer_end()
    r = m.rank_info()
    work = [t["cull_ms"] + t["evaluate_ms"] + t["compact_ms"] + t["attributes_ms"] for _, _, t in r]
    print("export %d%s: %.3f ms  cuts %s  work %s" % (i, " (rebalance)" if i >= 2 else "", ms, [b for b, _, _ in r][1:], [round(w, 3) for w in work]))
    m.close()
