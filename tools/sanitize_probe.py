import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import tangerine_b200 as T
name, cpu = sys.argv[1], float(sys.argv[2])
tree = T.Tree.load("tests/golden/models/%s.tgm" % name)
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(1.0 / cpu))
ctx = T.Context(0)
model = T.Model(ctx, tree)
for chunks in sys.argv[3:]:
    os.environ["TG_PIPELINE_CHUNKS"] = chunks
    mesh = model.export_mesh(grid)
    print(chunks, mesh.vertex_count, mesh.triangle_count, flush=True)
    mesh.close()
