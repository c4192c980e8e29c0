import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import tangerine_b200 as T, oracle_lib as O
for name in ("seaside_town", "synthetic200"):
    tree = T.Tree.load(O.model_path(name)); ctx = T.Context(0); model = T.Model(ctx, tree)
    for dbg in (0, 1, 2, 3):
        os.environ["TG_LONG_DEBUG"] = str(dbg)
        print(name, "debug", dbg, [model.check_long_programs(r) for r in (0.05, 0.7, 4.0)])
