"""Diagnostic: field-level diff of our exported files against the reference tool's, on the GPU box."""
import os, sys, subprocess, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import tangerine_b200 as T
import json
man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
tmp = tempfile.mkdtemp()
L = T.lib()
for name in sys.argv[1:] or ["basic_thing"]:
    cpu = float(man[name]["cells_per_unit"])
    tree = T.Tree.load(O.model_path(name))
    for ext in (".ply", ".stl"):
        ours, ref = os.path.join(tmp, name + "_o" + ext), os.path.join(tmp, name + "_r" + ext)
        fn = L.tg_export_ply if ext == ".ply" else L.tg_export_stl
        assert fn(tree.h, cpu, 0, os.fsencode(ours), 0) == 0
        O.ref_run("export", O.model_path(name), cpu, 0, ref)
        a, b = open(ours, "rb").read(), open(ref, "rb").read()
        print(name, ext, "sizes", len(a), len(b), "identical" if a == b else "DIFFER")
        if a != b and ext == ".ply":
            pa, pb = O.read_ply(ours), O.read_ply(ref)
            for k in ("pos", "normal", "color", "tris"):
                if pa[k] is None: continue
                x, y = np.ascontiguousarray(pa[k]), np.ascontiguousarray(pb[k])
                if x.dtype == np.float32:
                    xv, yv = x.view(np.uint32), y.view(np.uint32)
                else:
                    xv, yv = x, y
                bad = np.argwhere(xv != yv)
                print("  ", k, "mismatching elements", len(bad), "of", xv.size)
                for r, c in bad[:8]:
                    print("     row %d col %d ours %r (%08x) ref %r (%08x)" % (r, c, x[r, c], int(xv[r, c]), y[r, c], int(yv[r, c])))
        if a != b and ext == ".stl":
            na = np.frombuffer(a[84:], np.uint8).reshape(-1, 50); nb = np.frombuffer(b[84:], np.uint8).reshape(-1, 50)
            fa = na[:, :48].copy().view(np.float32); fb = nb[:, :48].copy().view(np.float32)
            bad = np.argwhere(fa.view(np.uint32) != fb.view(np.uint32))
            print("   stl mismatching floats", len(bad), "of", fa.size, "header same", a[:84] == b[:84])
            for r, c in bad[:8]:
                print("     tri %d float %d ours %r ref %r" % (r, c, fa[r, c], fb[r, c]))
    files = man[name]["files"]
    ours, ref = os.path.join(tmp, name + "_o.vox"), os.path.join(tmp, name + "_r.vox")
    assert L.tg_export_magica_voxel(tree.h, files["vox_grid_size"], files["vox_color_index"], os.fsencode(ours), 0) == 0
    O.ref_run("vox", O.model_path(name), files["vox_grid_size"], files["vox_color_index"], ref)
    a, b = open(ours, "rb").read(), open(ref, "rb").read()
    print(name, ".vox sizes", len(a), len(b), "identical" if a == b else "DIFFER")
    if a != b:
        n = min(len(a), len(b))
        d = [i for i in range(n) if a[i] != b[i]]
        print("   first differing offsets", d[:10], "count", len(d))
