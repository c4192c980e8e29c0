import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tangerine_b200 as T, oracle_lib as O
rays = np.load(os.path.join(ROOT, "tests", "golden", "rays.npz"))
name = sys.argv[1] if len(sys.argv) > 1 else "kitchen_sink"
tree = T.Tree.load(O.model_path(name)); ctx = T.Context(0); model = T.Model(ctx, tree)
r = rays[name + "/rays"]
hit, travel, pos = model.ray_cast(r)
want = rays[name + "/raycast"]
wt = want[:, 1].copy().view(np.float32); wp = want[:, 2:5].copy().view(np.float32)
bad = np.where((travel != wt) & ~(np.isnan(travel) & np.isnan(wt)))[0]
print(len(bad), "rays differ in travel;", (hit != (want[:, 0] != 0)).sum(), "in hit")
om = O.Model(name)
for b in bad[:8]:
    print(b, "gpu", hit[b], travel[b], pos[b], "ref", want[b, 0], wt[b], wp[b], "ray", r[b])
    # march on the CPU oracle and on the GPU tree evaluator step by step
    d = r[b, 3:6] / np.sqrt(np.float32((r[b, 3:6] ** 2).sum(dtype=np.float32)))
    p = r[b, 0:3].copy(); t = np.float32(0)
    for it in range(100):
        dc = om.eval_tree(p[None])[0]; dg = model.eval_points(p[None], T.EVAL_TREE)[0]
        if dc != dg and not (np.isnan(dc) and np.isnan(dg)):
            print("   iteration", it, "point", p, "cpu", dc, "gpu", dg)
            break
        if dc <= 0.001: break
        t = np.float32(t + dc); p = (d * t + r[b, 0:3]).astype(np.float32)
