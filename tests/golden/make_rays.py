#!/usr/bin/env python
"""Golden ray casts from the REFERENCE itself (oracle/_ref/tangerine_ref eval <model> raycast|magnet): SDFNode::RayMarch
with the Lua binding's defaults (lua_sdf.cpp:410-444), on seeded rays.  Output: tests/golden/rays.npz."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

MODELS = ["basic_thing", "seaside_town", "kitchen_sink", "gear", "stencil_test"]


def rays_for(name, n=1500):
    om = O.Model(name)
    lo, hi = om.bounds()
    rng = np.random.default_rng(sum(map(ord, name)) + 17)
    origin = (lo - 1.0 + (hi - lo + 2.0) * rng.random((n, 3))).astype(np.float32)
    target = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    rays = np.concatenate([origin, (target - origin).astype(np.float32)], axis=1).astype(np.float32)
    # the way seaside_town.lua probes its terrain (models/seaside_town.lua:83-85): straight down from above
    rays[:200, 0:2] = (lo[:2] + (hi[:2] - lo[:2]) * rng.random((200, 2))).astype(np.float32)
    rays[:200, 2] = np.float32(50.0)
    rays[:200, 3:6] = np.array([0, 0, -1], np.float32)
    return rays


def ref_rays(path, mode, rays, tmp):
    pin, pout = os.path.join(tmp, "rays.f32"), os.path.join(tmp, "hits.bin")
    rays.astype(np.float32).tofile(pin)
    O.ref_run("eval", path, mode, pin, pout)
    raw = np.fromfile(pout, np.uint32).reshape(-1, 5)
    return raw


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/tangerine_ref not built (make -C oracle ref)")
    tmp = tempfile.mkdtemp()
    arrays = {}
    for name in MODELS:
        rays = rays_for(name)
        arrays[name + "/rays"] = rays
        arrays[name + "/raycast"] = ref_rays(O.model_path(name), "raycast", rays, tmp)
        magnet = rays.copy()
        magnet[:, 3:6] = (rays[:, 0:3] + rays[:, 3:6]).astype(np.float32)   # origin + target form
        arrays[name + "/magnet_rays"] = magnet
        arrays[name + "/magnet"] = ref_rays(O.model_path(name), "magnet", magnet, tmp)
        print(name, int(arrays[name + "/raycast"][:, 0].sum()), "hits of", len(rays), flush=True)
    np.savez_compressed(os.path.join(HERE, "rays.npz"), **arrays)


if __name__ == "__main__":
    main()
