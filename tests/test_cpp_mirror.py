"""The real drop-in surface: a compiled C++ caller of include/tangerine_b200.hpp (tests/cpp/mirror_test.cpp), used the
way the reference's UI uses tangerine/export.h:36-39 (MeshExport detached, GetExportProgress polled, CancelExport)."""
import os
import subprocess

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "tangerine_b200")


@pytest.fixture(scope="module")
def mirror(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "mirror_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"),
                    "-L", LIBDIR, "-ltangerine_b200", "-Wl,-rpath," + LIBDIR, "-o", exe], check=True)
    return exe


def run(exe, *args):
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300)


def test_mirror_links_and_fails_loudly_without_a_model(mirror):
    r = run(mirror, "link")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK link" in r.stdout


@pytest.mark.gpu
def test_async_export_reports_progress_in_steps(mirror, tmp_path):
    """MeshExport on seaside_town at 512^3: Generation is monotonic and moves through more than three distinct values
    (per brick batch, written by the kernels into page-locked memory), while a second thread polls concurrently."""
    r = run(mirror, "export", O.model_path("seaside_town"), tmp_path / "town.ply", 510)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK export" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("delay_ms", [1, 150, 400])
def test_cancel_mid_flight(mirror, tmp_path, delay_ms):
    """CancelExport(true) during the octree build / during the export => TG_ERR_CANCELLED, no file, and the next export works."""
    r = run(mirror, "cancel", O.model_path("seaside_town"), tmp_path / "town.ply", 1022, delay_ms)
    if "status after cancel: 0" in r.stdout:
        pytest.skip("the export finished before the cancel arrived (%d ms)" % delay_ms)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK cancel" in r.stdout


@pytest.mark.gpu
def test_populate_drawable_is_the_live_mesh(mirror, tmp_path):
    """PopulateDrawable (the mirror of Sodapop::Populate) returns the mesh tests/test_gpu_live.py pins to the reference."""
    import json
    import re

    import numpy as np
    import tangerine_b200 as T
    r = run(mirror, "live", O.model_path("basic_thing"), tmp_path / "unused.ply")
    assert r.returncode == 0 and "OK live" in r.stdout, r.stdout + r.stderr
    vertices, triangles, fnv = re.search(r"live vertices (\d+) triangles (\d+) fnv ([0-9a-f]+)", r.stdout).groups()
    with open(os.path.join(ROOT, "tests", "golden", "slices_live_basic20.json")) as f:
        fx = json.load(f)
    assert (int(vertices), int(triangles)) == (fx["vertices"], fx["triangles"])
    ctx = T.Context(0)
    model = T.Model(ctx, T.Tree.load(O.model_path("basic_thing")), live=True)
    mesh = model.export_mesh(model.live_grid(20.0), flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD)
    h = 0xCBF29CE484222325
    for byte in np.ascontiguousarray(mesh.positions, np.float32).tobytes():
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    assert "%016x" % h == fnv
    mesh.close()
    model.close()
    ctx.close()
