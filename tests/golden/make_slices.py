#!/usr/bin/env python
"""Per-layer digests of the REFERENCE mesh at the grid sizes bench.py measures (BASELINE.json configs).

Runs `oracle/_ref/tangerine_ref slices` -- the reference's own FirstLoopInnerThunk / SecondLoopThunk / WritePLY attribute
loop, compiled unmodified from /root/reference, driven by std::threads -- over the WHOLE grid of every benched workload and
stores, per cell layer k: vertex count, triangle count and SHA-256 prefixes of the positions, normals, colours and
triangle indices in the serial (k, j, i) order.  tests/test_gpu_bench_parity.py compares the CUDA export of the same
grids with these layer by layer, bit for bit (up to the sign of zero and NaN payloads).

    python tests/golden/make_slices.py [workload ...]       seaside1024 takes ~30-40 CPU-minutes on 8 cores

The synthetic scene is not a reference model: its tree comes from tg_make_synthetic (host code, no GPU needed) and is
handed to the reference tool as a .tgm file.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402

# workload -> (model, step, attributes)   -- the same table as bench.py WORKLOADS
WORKLOADS = {
    "basic66": ("basic_thing", 1.0 / 16.0),
    "gear512": ("gear", 8.0 / 510.0),
    "colorcube512": ("color-cube", 9.6 / 510.0),
    "seaside512": ("seaside_town", 10.0 / 510.0),
    "seaside1024": ("seaside_town", 10.0 / 1022.0),
    "synthetic256": ("synthetic:10000", 10.0 / 254.0),
    "synthetic512": ("synthetic:10000", 10.0 / 510.0),
    "synthetic1024": ("synthetic:10000", 10.0 / 1022.0),
}


def model_file(name, tmp):
    if name.startswith("synthetic:"):
        import tangerine_b200 as T
        path = os.path.join(tmp, name.replace(":", "_") + ".tgm")
        T.Tree.synthetic(int(name.split(":")[1]), 1234).save(path)
        return path
    return O.model_path(name)


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/tangerine_ref not built (make -C oracle ref)")
    tmp = tempfile.mkdtemp()
    threads = os.cpu_count() or 1
    for workload in (sys.argv[1:] or WORKLOADS):
        name, step = WORKLOADS[workload]
        path = model_file(name, tmp)
        info = json.loads(O.ref_run("info", path))
        lo, hi = info["bounds_min"], info["bounds_max"]
        out = os.path.join(HERE, "slices_%s.json" % workload)
        args = [O.REF_TOOL, "slices", path] + ["%.9g" % v for v in lo + hi] + ["%.9g" % float(np.float32(step)), str(threads), "0", "0", "1", out]
        print(" ".join(args), flush=True)
        print(subprocess.run(args, check=True, capture_output=True, text=True).stdout, flush=True)
        with open(out) as f:
            d = json.load(f)
        d["workload"], d["model"], d["step"] = workload, name, float(np.float32(step))
        with open(out, "w") as f:
            json.dump(d, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
