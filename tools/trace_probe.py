import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import tangerine_b200 as T
tree = T.Tree.load("tests/golden/models/seaside_town.tgm")
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(10 / 1022))
ctx = T.Context(0)
model = T.Model(ctx, tree)
for lanes, chunks in (("1", "8"), ("2", "8"), ("1", "4")):
    os.environ["TG_PIPELINE_CHUNKS"] = chunks
    os.environ["TG_PIPELINE_LANES"] = lanes
    os.environ.pop("TG_TRACE", None)
    for it in range(4):
        if it == 3:
            os.environ["TG_TRACE"] = "1"
        t1 = time.perf_counter()
        mesh = model.export_mesh(grid); t2 = time.perf_counter()
        tm = mesh.timings
        mesh.close()
    print("lanes %s chunks %s: export %.2f ms device span %.2f" % (lanes, chunks, (t2 - t1) * 1e3, tm["total_device_ms"]), flush=True)
    sys.stderr.flush()
