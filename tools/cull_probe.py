"""K0 diagnostics: item counts per level and cull time for a workload.   TG_TRACE_CULL=1 python tools/cull_probe.py [workload]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
import bench
name, step, refine, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "seaside1024"]
tree, _ = bench.load_workload_tree(T, name)
ctx = T.Context(0)
model = T.Model(ctx, tree)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
for i in range(3):
    if i < 2:
        os.environ.pop("TG_TRACE_CULL", None)
    else:
        os.environ["TG_TRACE_CULL"] = "1"
    m = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_COLORS | T.MESH_DEVICE_ONLY)
    print({k: round(v, 3) if isinstance(v, float) else v for k, v in m.timings.items()})
    m.close()
