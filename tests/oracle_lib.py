"""ctypes binding for oracle/liboracle.so (TEST INFRASTRUCTURE ONLY).

The oracle is the CPU restatement of the reference algorithm (oracle/tg_oracle.c).  Only tests,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this module; the product
package ``tangerine_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
REF_TOOL = os.path.join(ORACLE_DIR, "_ref", "tangerine_ref")
MODELS = os.path.join(ROOT, "tests", "golden", "models")


def model_path(name):
    return os.path.join(MODELS, name + ".tgm")


class Grid(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float),
                ("sx", C.c_uint64), ("sy", C.c_uint64), ("sz", C.c_uint64)]

    @property
    def shape(self):
        return (int(self.sx), int(self.sy), int(self.sz))


class _Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("cells", C.POINTER(C.c_int64)),
                ("triangles", C.POINTER(C.c_uint32)), ("vertex_count", C.c_uint64),
                ("triangle_count", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nodes", "leaves", "words", "leaf_words", "max_words", "max_stack", "hash")]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, fp = C.c_void_p, C.POINTER(C.c_float)
        L.tgo_model_load.restype = vp
        L.tgo_model_load.argtypes = [C.c_char_p]
        L.tgo_model_free.argtypes = [vp]
        L.tgo_model_bounds.argtypes = [vp, fp, fp]
        L.tgo_model_has_paint.argtypes = [vp]
        L.tgo_model_leaf_count.argtypes = [vp]
        L.tgo_model_root_program.restype = C.c_uint64
        L.tgo_model_root_program.argtypes = [vp, C.POINTER(C.c_uint32), C.c_uint64]
        L.tgo_octree_create.restype = vp
        L.tgo_octree_create.argtypes = [vp, C.c_float]
        L.tgo_octree_create_live.restype = vp
        L.tgo_octree_create_live.argtypes = [vp, C.c_float]
        L.tgo_live_grid.argtypes = [vp, C.c_float, C.POINTER(Grid)]
        L.tgo_live_grid.restype = None
        L.tgo_weld.restype = C.c_uint64
        L.tgo_weld.argtypes = [fp, C.c_uint64, fp, C.POINTER(C.c_uint32)]
        L.tgo_octree_free.argtypes = [vp]
        L.tgo_octree_stats.argtypes = [vp, C.POINTER(_Stats)]
        L.tgo_octree_bounds.argtypes = [vp, fp, fp]
        L.tgo_octree_bounds.restype = None
        L.tgo_eval_octree.argtypes = [vp, fp, C.c_uint64, fp, C.c_int]
        L.tgo_eval_tree.argtypes = [vp, fp, C.c_uint64, fp, C.c_int]
        L.tgo_ray_march.argtypes = [vp, fp, C.c_uint64, C.c_int, C.c_float, C.c_int, fp]
        L.tgo_ray_march.restype = None
        L.tgo_eval_interp.argtypes = [vp, fp, C.c_uint64, fp]
        L.tgo_gradient.argtypes = [vp, fp, C.c_uint64, fp]
        L.tgo_color.argtypes = [vp, fp, C.c_uint64, C.POINTER(C.c_uint8)]
        L.tgo_export_grid.argtypes = [fp, fp, fp, C.POINTER(Grid)]
        L.tgo_surface_nets.argtypes = [vp, C.POINTER(Grid), C.POINTER(_Mesh), C.c_int]
        L.tgo_mesh_free.argtypes = [C.POINTER(_Mesh)]
        L.tgo_lattice_samples.argtypes = [vp, C.POINTER(Grid), fp, C.c_int]
        L.tgo_refine.argtypes = [vp, fp, C.c_uint64, fp, C.c_int]
        L.tgo_point_cloud.restype = C.c_uint64
        L.tgo_point_cloud.argtypes = [vp, fp, fp, fp, C.POINTER(fp)]
        L.tgo_voxels.restype = C.c_uint64
        L.tgo_voxels.argtypes = [vp, C.c_float, C.POINTER(C.c_int32), fp, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
        L.tgo_free.argtypes = [vp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


THREADS = os.cpu_count() or 1


class Model:
    def __init__(self, path):
        if not os.path.exists(path):
            path = model_path(path)
        self.path = path
        self.h = lib().tgo_model_load(path.encode())
        if not self.h:
            raise IOError("cannot load " + path)

    def __del__(self):
        if getattr(self, "h", None):
            lib().tgo_model_free(self.h)
            self.h = None

    def bounds(self):
        lo = np.zeros(3, np.float32)
        hi = np.zeros(3, np.float32)
        lib().tgo_model_bounds(self.h, _fp(lo), _fp(hi))
        return lo, hi

    def has_paint(self):
        return bool(lib().tgo_model_has_paint(self.h))

    def leaf_count(self):
        return lib().tgo_model_leaf_count(self.h)

    def root_program(self):
        n = lib().tgo_model_root_program(self.h, None, 0)
        out = np.zeros(n, np.uint32)
        lib().tgo_model_root_program(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        return out

    def eval_tree(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros(len(pts), np.float32)
        lib().tgo_eval_tree(self.h, _fp(pts), len(pts), _fp(out), THREADS)
        return out

    def ray_march(self, rays, max_iterations=100, epsilon=0.001, magnet=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 5), np.float32)
        lib().tgo_ray_march(self.h, _fp(rays), len(rays), max_iterations, epsilon, 1 if magnet else 0, _fp(out))
        return out[:, 0] != 0, out[:, 1].copy(), out[:, 2:5].copy()

    def eval_interp(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros(len(pts), np.float32)
        lib().tgo_eval_interp(self.h, _fp(pts), len(pts), _fp(out))
        return out

    def voxels(self, grid_size):
        size = (C.c_int32 * 3)()
        radius = C.c_float()
        ptr = C.POINTER(C.c_int32)()
        n = lib().tgo_voxels(self.h, grid_size, size, C.byref(radius), C.byref(ptr), THREADS)
        xyz = np.ctypeslib.as_array(ptr, shape=(max(n, 1), 3))[:n].copy()
        lib().tgo_free(ptr)
        return tuple(size), radius.value, xyz


class Octree:
    def __init__(self, model, target_size=0.25, live=False):
        """live=True: the live mesher's octree (sodapop.cpp:240, 568-571); eval / lattice / surface_nets then sample its
        clamped inexact field (sodapop.cpp:583-587)."""
        self.model = model
        self.h = (lib().tgo_octree_create_live if live else lib().tgo_octree_create)(model.h, target_size)
        if not self.h:
            raise ValueError("octree could not be built")

    def bounds(self):
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib().tgo_octree_bounds(self.h, _fp(lo), _fp(hi))
        return lo, hi

    def live_grid(self, density=20.0):
        g = Grid()
        lib().tgo_live_grid(self.h, density, C.byref(g))
        return g

    def __del__(self):
        if getattr(self, "h", None):
            lib().tgo_octree_free(self.h)
            self.h = None

    def stats(self):
        s = _Stats()
        lib().tgo_octree_stats(self.h, C.byref(s))
        d = {n: int(getattr(s, n)) for n, _ in _Stats._fields_}
        d["hash"] = "%016x" % d["hash"]
        return d

    def eval(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros(len(pts), np.float32)
        lib().tgo_eval_octree(self.h, _fp(pts), len(pts), _fp(out), THREADS)
        return out

    def gradient(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros((len(pts), 3), np.float32)
        lib().tgo_gradient(self.h, _fp(pts), len(pts), _fp(out))
        return out

    def color(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros((len(pts), 3), np.uint8)
        lib().tgo_color(self.h, _fp(pts), len(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def lattice(self, grid):
        sx, sy, sz = grid.shape
        out = np.zeros((sz + 1, sy + 1, sx + 1), np.float32)
        lib().tgo_lattice_samples(self.h, C.byref(grid), _fp(out), THREADS)
        return out

    def surface_nets(self, grid):
        m = _Mesh()
        rc = lib().tgo_surface_nets(self.h, C.byref(grid), C.byref(m), THREADS)
        if rc != 0:
            raise MemoryError("oracle surface nets failed")
        nv, nt = int(m.vertex_count), int(m.triangle_count)
        verts = np.ctypeslib.as_array(m.vertices, shape=(max(nv, 1), 3))[:nv].copy()
        cells = np.ctypeslib.as_array(m.cells, shape=(max(nv, 1),))[:nv].copy()
        tris = np.ctypeslib.as_array(m.triangles, shape=(max(nt, 1), 3))[:nt].copy()
        lib().tgo_mesh_free(C.byref(m))
        return verts, cells, tris

    def refine(self, pts, half, iterations):
        pts = np.array(pts, np.float32).reshape(-1, 3).copy()
        half = _f3(half)
        lib().tgo_refine(self.h, _fp(pts), len(pts), _fp(half), iterations)
        return pts

    def point_cloud(self, lo, hi, step):
        lo, hi, step = _f3(lo), _f3(hi), _f3(step)
        ptr = C.POINTER(C.c_float)()
        n = lib().tgo_point_cloud(self.h, _fp(lo), _fp(hi), _fp(step), C.byref(ptr))
        pts = np.ctypeslib.as_array(ptr, shape=(max(n, 1), 3))[:n].copy()
        lib().tgo_free(ptr)
        return pts


def export_grid(lo, hi, step):
    g = Grid()
    lo, hi = _f3(lo), _f3(hi)
    step = _f3(np.broadcast_to(np.asarray(step, np.float32), (3,)))
    lib().tgo_export_grid(_fp(lo), _fp(hi), _fp(step), C.byref(g))
    return g


# ---------------------------------------------------------------------------------------------
# Reference tool (oracle/_ref/tangerine_ref) helpers; available only where it was built.
# ---------------------------------------------------------------------------------------------

def have_ref():
    return os.path.exists(REF_TOOL)


def ref_run(*args):
    return subprocess.run([REF_TOOL] + [str(a) for a in args], check=True, capture_output=True, text=True).stdout


def weld(vertices):
    """MeshGenerator::Accumulate over a vertex stream (oracle): (distinct vertices (n, 4), index per input vertex)."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    out = np.zeros((max(len(v), 1), 4), np.float32)
    idx = np.zeros(max(len(v), 1), np.uint32)
    n = lib().tgo_weld(_fp(v), len(v), _fp(out), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out[:n].copy(), idx[:len(v)].copy()


def ref_weld(vertices, tmpdir):
    """The reference's own MeshGenerator (oracle/_ref/tangerine_ref weld)."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    pin, pout = os.path.join(str(tmpdir), "weld_in.f32"), os.path.join(str(tmpdir), "weld_out.bin")
    v.tofile(pin)
    subprocess.run([REF_TOOL, "weld", pin, pout], check=True)
    raw = open(pout, "rb").read()
    n = int(np.frombuffer(raw[:4], np.uint32)[0])
    return np.frombuffer(raw[4:4 + 16 * n], np.float32).reshape(-1, 4).copy(), np.frombuffer(raw[4 + 16 * n:], np.uint32).copy()


def ref_eval(model, mode, pts, tmpdir):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    pin = os.path.join(str(tmpdir), "pts.f32")
    pout = os.path.join(str(tmpdir), "out.bin")
    pts.tofile(pin)
    ref_run("eval", model, mode, pin, pout)
    if mode in ("gradient", "live-gradient"):
        return np.fromfile(pout, np.float32).reshape(-1, 3)
    if mode == "color":
        return np.fromfile(pout, np.uint8).reshape(-1, 3)
    return np.fromfile(pout, np.float32)


def read_ply(path):
    """Minimal reader for the binary PLY layout the reference writes (export.cpp:198-280)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode()
    nv = nf = 0
    color = "property uchar red" in header
    for line in header.splitlines():
        if line.startswith("element vertex"):
            nv = int(line.split()[-1])
        if line.startswith("element face"):
            nf = int(line.split()[-1])
    vdt = [("p", "<f4", 3), ("n", "<f4", 3)] + ([("c", "u1", 3)] if color else [])
    v = np.frombuffer(data, dtype=np.dtype(vdt), count=nv, offset=end)
    off = end + nv * np.dtype(vdt).itemsize
    fdt = np.dtype([("n", "u1"), ("i", "<u4", 3)])
    f = np.frombuffer(data, dtype=fdt, count=nf, offset=off)
    return {"pos": v["p"].copy(), "normal": v["n"].copy(), "color": v["c"].copy() if color else None,
            "tris": f["i"].copy() if nf else np.zeros((0, 3), np.uint32), "header": header}
