#!/usr/bin/env python
"""Golden data for the LIVE mesher (Sodapop, tangerine/sodapop.cpp:227-247 and 562-760) from the reference itself.

  live.npz                   per model: the implicit function of sodapop.cpp:583-587 -- clamp(SDFOctree::Eval(p, Exact =
                             false), -100, 100) on the octree of SDFOctree::Create(Evaluator, .25, false, 3, 0.0) with its
                             incomplete nodes populated -- at the seeded points of <model>.npz (`tangerine_ref eval live`),
                             and SDFOctree::Gradient on that octree (`live-gradient`, the live mesher's normals, :816)
  live_octree.json           per model: the live octree's statistics, hash (pre-order, as in manifest.json) and Bounds
  slices_live_<case>.json    per cell layer digests of the mesh of NaiveSurfaceNets' vertex and face loops over the point
                             cache of the octree leaves (:609-760) at a meshing density (`tangerine_ref slices-live`), with
                             the grid of NaiveSurfaceNetsScratch (:153-179)

    python tests/golden/make_live.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

MODELS = ["basic_thing", "seaside_town", "gear", "kitchen_sink", "stencil_test", "cones", "scale", "flower", "color-cube"]
# case -> (model, meshing density); 20 is Sodapop's default (sodapop.cpp:43), meshing_density_push adds to it (:221)
CASES = {"basic20": ("basic_thing", 20.0), "seaside20": ("seaside_town", 20.0), "gear20": ("gear", 20.0), "kitchen20": ("kitchen_sink", 20.0),
         "stencil20": ("stencil_test", 20.0), "seaside50": ("seaside_town", 50.0)}


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/tangerine_ref not built (make -C oracle ref)")
    tmp = tempfile.mkdtemp()
    arrays = {}
    for name in MODELS:
        pts = np.load(os.path.join(HERE, name + ".npz"))["points"]
        arrays[name + "/live"] = O.ref_eval(O.model_path(name), "live", pts, tmp)
        arrays[name + "/live_gradient"] = O.ref_eval(O.model_path(name), "live-gradient", pts, tmp)
        print(name, len(pts), int((arrays[name + "/live"] == 100.0).sum()), "samples at +100", flush=True)
    np.savez_compressed(os.path.join(HERE, "live.npz"), **arrays)
    # the live octree itself (`tangerine_ref info-live`): node / leaf counts, program words, hash, and the octree's Bounds
    octrees = {}
    for name in MODELS + ["synthetic200"]:
        info = json.loads(O.ref_run("info-live", O.model_path(name)))
        info.pop("octree_build_s")
        octrees[name] = info
    with open(os.path.join(HERE, "live_octree.json"), "w") as f:
        json.dump(octrees, f, indent=1, sort_keys=True)
    threads = os.cpu_count() or 1
    for case, (name, density) in CASES.items():
        out = os.path.join(HERE, "slices_live_%s.json" % case)
        args = [O.REF_TOOL, "slices-live", O.model_path(name), "%g" % density, str(threads), "1", out]
        print(" ".join(args), flush=True)
        print(subprocess.run(args, check=True, capture_output=True, text=True).stdout, flush=True)
        with open(out) as f:
            d = json.load(f)
        d["case"], d["model"], d["density"] = case, name, density
        with open(out, "w") as f:
            json.dump(d, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
