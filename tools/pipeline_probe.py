"""End-to-end export time against the slab schedule of the pipelined export.   python tools/pipeline_probe.py [workload] "chunks:share,share,..." ...
(TG_PIPELINE_CHUNKS / TG_PIPELINE_SHARES are read at every export; an empty share list means equal costs.)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tangerine_b200 as T
import bench
workload = sys.argv[1] if len(sys.argv) > 1 and ":" not in sys.argv[1] else "seaside1024"
specs = [a for a in sys.argv[1:] if ":" in a] or ["3:"]
name, step, refine, desc = bench.WORKLOADS[workload]
tree, _ = bench.load_workload_tree(T, name)
ctx = T.Context(0)
model = T.Model(ctx, tree)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(step))
flags = T.MESH_NORMALS | T.MESH_COLORS
for spec in specs:
    chunks, shares = spec.split(":")
    os.environ["TG_PIPELINE_CHUNKS"] = chunks
    if shares:
        os.environ["TG_PIPELINE_SHARES"] = shares
    else:
        os.environ.pop("TG_PIPELINE_SHARES", None)
    times = []
    for i in range(9):
        ctx.synchronize()
        t0 = time.perf_counter()
        model.upload()
        m = model.export_mesh(grid, flags=flags, refine=refine)
        times.append((time.perf_counter() - t0) * 1e3)
        dev = m.timings["total_device_ms"]
        m.close()
    times = sorted(times[3:])
    print("%-40s e2e min %.3f median %.3f ms   (device span %.3f)" % (spec, times[0], times[len(times) // 2], dev), flush=True)
model.close()
ctx.close()
