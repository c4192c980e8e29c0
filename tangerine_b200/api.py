"""ctypes binding of libtangerine_b200.so (see include/tangerine_b200.h for the contract)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

EVAL_OCTREE, EVAL_INTERP, EVAL_TREE, EVAL_GRADIENT, EVAL_COLOR, EVAL_LIVE = range(6)
MESH_NORMALS, MESH_COLORS, MESH_NO_CULL, MESH_DEVICE_ONLY, MESH_FACE_NORMALS, MESH_KEEP_CANCEL, MESH_FAST, MESH_REBALANCE = 1, 2, 4, 8, 16, 32, 64, 128
MESH_LIVE_FIELD = 256


class TangerineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("tangerine_b200 error %d: %s" % (code, message))
        self.code = code


def library_path():
    # TANGERINE_B200_LIB: load an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("TANGERINE_B200_LIB") or os.path.join(_HERE, "libtangerine_b200.so")


def build_library():
    """Compile the CUDA extension in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc"), "-j8"])


class Grid(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float),
                ("sx", C.c_uint64), ("sy", C.c_uint64), ("sz", C.c_uint64)]

    @property
    def shape(self):
        return (int(self.sx), int(self.sy), int(self.sz))

    @property
    def cells(self):
        return int(self.sx) * int(self.sy) * int(self.sz)

    @property
    def samples(self):
        return (int(self.sx) + 1) * (int(self.sy) + 1) * (int(self.sz) + 1)


class MeshOptions(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("refine_iterations", C.c_int32), ("scale", C.c_float),
                ("slab_begin", C.c_uint64), ("slab_end", C.c_uint64)]


class MeshTimings(C.Structure):
    _fields_ = [("cull_ms", C.c_float), ("evaluate_ms", C.c_float), ("compact_ms", C.c_float), ("faces_ms", C.c_float),
                ("attributes_ms", C.c_float), ("total_device_ms", C.c_float), ("download_ms", C.c_float),
                ("bricks_total", C.c_uint64), ("bricks_evaluated", C.c_uint64), ("samples_evaluated", C.c_uint64),
                ("algorithmic_flops", C.c_uint64), ("kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class _Mesh(C.Structure):
    _fields_ = [("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("colors", C.POINTER(C.c_uint8)),
                ("triangles", C.POINTER(C.c_uint32)), ("face_normals", C.POINTER(C.c_float)),
                ("vertex_count", C.c_uint64), ("triangle_count", C.c_uint64), ("halo_vertices", C.c_uint64),
                ("layer_vertices", C.POINTER(C.c_uint32)), ("layer_vertex_cost", C.POINTER(C.c_double)), ("layer_count", C.c_uint64),
                ("timings", MeshTimings), ("opaque", C.c_void_p)]


class ModelStats(C.Structure):
    _fields_ = [("octree_nodes", C.c_uint64), ("octree_leaves", C.c_uint64), ("reference_words", C.c_uint64),
                ("reference_leaf_words", C.c_uint64), ("reference_max_words", C.c_uint64), ("max_stack", C.c_uint64),
                ("octree_hash", C.c_uint64), ("device_bytes", C.c_uint64), ("build_seconds", C.c_double),
                ("upload_seconds", C.c_double), ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3),
                ("has_paint", C.c_int32), ("leaf_count", C.c_int32)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["bounds_min"] = [float(v) for v in self.bounds_min]
        d["bounds_max"] = [float(v) for v in self.bounds_max]
        d["octree_hash"] = "%016x" % self.octree_hash
        return d


_lib = None


def lib():
    """Load the shared library; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("libtangerine_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C tangerine_b200/csrc`")
    L = C.CDLL(path)
    vp, fp, u32, u64, i32 = C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "tg_last_error": (C.c_char_p, []),
        "tg_version": (C.c_char_p, []),
        "tg_make_sphere": (vp, [C.c_float]),
        "tg_make_ellipsoid": (vp, [C.c_float] * 3),
        "tg_make_box": (vp, [C.c_float] * 3),
        "tg_make_torus": (vp, [C.c_float] * 2),
        "tg_make_cylinder": (vp, [C.c_float] * 2),
        "tg_make_plane": (vp, [C.c_float] * 3),
        "tg_make_cone": (vp, [C.c_float] * 2),
        "tg_make_coninder": (vp, [C.c_float] * 3),
        "tg_make_union": (vp, [vp, vp]),
        "tg_make_diff": (vp, [vp, vp]),
        "tg_make_inter": (vp, [vp, vp]),
        "tg_make_blend_union": (vp, [C.c_float, vp, vp]),
        "tg_make_blend_diff": (vp, [C.c_float, vp, vp]),
        "tg_make_blend_inter": (vp, [C.c_float, vp, vp]),
        "tg_make_flate": (vp, [vp, C.c_float]),
        "tg_make_stencil": (vp, [vp, vp, u32, i32]),
        "tg_tree_copy": (vp, [vp]),
        "tg_tree_free": (None, [vp]),
        "tg_tree_move": (i32, [vp] + [C.c_float] * 3),
        "tg_tree_rotate": (i32, [vp] + [C.c_float] * 4),
        "tg_tree_rotate_x": (i32, [vp, C.c_float]),
        "tg_tree_rotate_y": (i32, [vp, C.c_float]),
        "tg_tree_rotate_z": (i32, [vp, C.c_float]),
        "tg_tree_scale": (i32, [vp, C.c_float]),
        "tg_tree_align": (i32, [vp] + [C.c_float] * 3),
        "tg_tree_paint": (i32, [vp, u32, i32]),
        "tg_material_create": (u32, [C.c_float] * 3),
        "tg_tree_eval": (C.c_float, [vp] + [C.c_float] * 3),
        "tg_tree_bounds": (i32, [vp, fp, fp]),
        "tg_tree_has_paint": (i32, [vp]),
        "tg_tree_has_finite_bounds": (i32, [vp]),
        "tg_tree_leaf_count": (i32, [vp]),
        "tg_tree_load": (vp, [C.c_char_p]),
        "tg_tree_save": (i32, [vp, C.c_char_p]),
        "tg_make_synthetic": (vp, [u32, u32]),
        "tg_context_create": (vp, [i32]),
        "tg_context_destroy": (None, [vp]),
        "tg_context_device": (i32, [vp]),
        "tg_context_create_multi": (vp, [C.POINTER(C.c_int), i32]),
        "tg_context_device_count": (i32, [vp]),
        "tg_mesh_rank_info": (i32, [C.POINTER(_Mesh), i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(MeshTimings)]),
        "tg_model_create": (vp, [vp, vp, C.c_float, i32]),
        "tg_model_create_live": (vp, [vp, vp, C.c_float, i32]),
        "tg_live_grid": (i32, [vp, C.c_float, C.POINTER(Grid)]),
        "tg_rearm": (i32, [vp]),
        "tg_debug_tables_hash": (i32, [vp, C.c_float, i32, i32, C.POINTER(u64)]),
        "tg_weld": (i32, [vp, C.POINTER(C.c_float), u64, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(u64)]),
        "tg_tree_octree_stats": (i32, [vp, C.c_float, i32, C.POINTER(ModelStats)]),
        "tg_tree_octree_stats_live": (i32, [vp, C.c_float, i32, C.POINTER(ModelStats)]),
        "tg_model_destroy": (None, [vp]),
        "tg_tree_plan_slabs": (i32, [vp, C.c_float, C.POINTER(Grid), i32, C.POINTER(u64), C.POINTER(C.c_double)]),
        "tg_model_get_stats": (i32, [vp, C.POINTER(ModelStats)]),
        "tg_eval_points": (i32, [vp, i32, fp, u64, vp]),
        "tg_export_grid": (i32, [fp, fp, fp, C.POINTER(Grid)]),
        "tg_debug_check_long_programs": (i32, [vp, C.c_float, C.POINTER(u64)]),
        "tg_ray_cast": (i32, [vp, fp, u64, i32, C.c_float, i32, fp]),
        "tg_debug_node_program": (u64, [vp, u32, C.POINTER(u32), u64, C.POINTER(u32)]),
        "tg_export_mesh": (i32, [vp, C.POINTER(Grid), C.POINTER(MeshOptions), C.POINTER(_Mesh)]),
        "tg_mesh_free": (None, [C.POINTER(_Mesh)]),
        "tg_mesh_download": (i32, [C.POINTER(_Mesh), u32]),
        "tg_eval_lattice": (i32, [vp, C.POINTER(Grid), fp, fp]),
        "tg_eval_lattice_flags": (i32, [vp, C.POINTER(Grid), u32, fp, fp]),
        "tg_export_points": (i32, [vp, fp, fp, fp, i32, u32, C.c_float, C.POINTER(_Mesh)]),
        "tg_export_voxels": (i32, [vp, C.c_float, C.POINTER(C.c_int32), fp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(u64)]),
        "tg_free": (None, [vp]),
        "tg_progress": (i32, [vp, fp, C.POINTER(i32)]),
        "tg_cancel": (i32, [vp, i32]),
        "tg_export_ply": (i32, [vp, C.c_float, i32, C.c_char_p, i32]),
        "tg_export_stl": (i32, [vp, C.c_float, i32, C.c_char_p, i32]),
        "tg_export_magica_voxel": (i32, [vp, C.c_float, i32, C.c_char_p, i32]),
        "tg_write_ply": (i32, [C.c_char_p, C.POINTER(_Mesh)]),
        "tg_write_stl": (i32, [C.c_char_p, C.POINTER(_Mesh)]),
        "tg_timer_begin": (i32, [vp]),
        "tg_timer_end": (i32, [vp, fp]),
        "tg_measure_fp32_peak": (i32, [vp, C.POINTER(C.c_double)]),
        "tg_flush_l2": (i32, [vp]),
        "tg_context_synchronize": (i32, [vp]),
        "tg_model_upload": (i32, [vp]),
        "tg_brick_profile": (i32, [vp, C.POINTER(Grid), C.POINTER(C.c_uint32), u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._tg_signatures = sig
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise TangerineError(rc, lib().tg_last_error().decode(errors="replace"))


def _handle(h):
    if not h:
        raise TangerineError(-1, lib().tg_last_error().decode(errors="replace"))
    return h


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float32), (3,)))


class Tree:
    """A CSG tree handle (tg_tree).  Mirrors the reference's SDF:: constructors and Lua modifiers."""

    def __init__(self, handle):
        self.h = _handle(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().tg_tree_free(self.h)
            self.h = None

    # brushes ------------------------------------------------------------------------------------
    @staticmethod
    def sphere(radius):
        return Tree(lib().tg_make_sphere(radius))

    @staticmethod
    def ellipsoid(rx, ry, rz):
        return Tree(lib().tg_make_ellipsoid(rx, ry, rz))

    @staticmethod
    def box(ex, ey, ez):
        return Tree(lib().tg_make_box(ex, ey, ez))

    @staticmethod
    def torus(major, minor):
        return Tree(lib().tg_make_torus(major, minor))

    @staticmethod
    def cylinder(radius, extent):
        return Tree(lib().tg_make_cylinder(radius, extent))

    @staticmethod
    def plane(nx, ny, nz):
        return Tree(lib().tg_make_plane(nx, ny, nz))

    @staticmethod
    def cone(radius, height):
        return Tree(lib().tg_make_cone(radius, height))

    @staticmethod
    def coninder(radius_l, radius_h, height):
        return Tree(lib().tg_make_coninder(radius_l, radius_h, height))

    @staticmethod
    def synthetic(primitives, seed=1234):
        return Tree(lib().tg_make_synthetic(primitives, seed))

    @staticmethod
    def load(path):
        return Tree(lib().tg_tree_load(os.fsencode(path)))

    # operators ----------------------------------------------------------------------------------
    def union(self, other):
        return Tree(lib().tg_make_union(self.h, other.h))

    def diff(self, other):
        return Tree(lib().tg_make_diff(self.h, other.h))

    def inter(self, other):
        return Tree(lib().tg_make_inter(self.h, other.h))

    def blend_union(self, other, threshold):
        return Tree(lib().tg_make_blend_union(threshold, self.h, other.h))

    def blend_diff(self, other, threshold):
        return Tree(lib().tg_make_blend_diff(threshold, self.h, other.h))

    def blend_inter(self, other, threshold):
        return Tree(lib().tg_make_blend_inter(threshold, self.h, other.h))

    def flate(self, radius):
        return Tree(lib().tg_make_flate(self.h, radius))

    def stencil(self, mask, material, apply_to_negative=True):
        return Tree(lib().tg_make_stencil(self.h, mask.h, material, 1 if apply_to_negative else 0))

    def copy(self):
        return Tree(lib().tg_tree_copy(self.h))

    # modifiers return a modified copy, like the Lua layer (lua_sdf.cpp:56-62) -------------------------
    def move(self, x, y, z):
        t = self.copy()
        _check(lib().tg_tree_move(t.h, x, y, z))
        return t

    def rotate(self, qx, qy, qz, qw):
        t = self.copy()
        _check(lib().tg_tree_rotate(t.h, qx, qy, qz, qw))
        return t

    def rotate_x(self, degrees):
        t = self.copy()
        _check(lib().tg_tree_rotate_x(t.h, degrees))
        return t

    def rotate_y(self, degrees):
        t = self.copy()
        _check(lib().tg_tree_rotate_y(t.h, degrees))
        return t

    def rotate_z(self, degrees):
        t = self.copy()
        _check(lib().tg_tree_rotate_z(t.h, degrees))
        return t

    def scale(self, s):
        t = self.copy()
        _check(lib().tg_tree_scale(t.h, s))
        return t

    def align(self, x, y, z):
        t = self.copy()
        _check(lib().tg_tree_align(t.h, x, y, z))
        return t

    def paint(self, material, force=False):
        t = self.copy()
        _check(lib().tg_tree_paint(t.h, material, 1 if force else 0))
        return t

    # queries ------------------------------------------------------------------------------------
    def eval(self, x, y, z):
        return float(lib().tg_tree_eval(self.h, x, y, z))

    def bounds(self):
        lo = np.zeros(3, np.float32)
        hi = np.zeros(3, np.float32)
        _check(lib().tg_tree_bounds(self.h, _fp(lo), _fp(hi)))
        return lo, hi

    def has_paint(self):
        return bool(lib().tg_tree_has_paint(self.h))

    def leaf_count(self):
        return int(lib().tg_tree_leaf_count(self.h))

    def save(self, path):
        _check(lib().tg_tree_save(self.h, os.fsencode(path)))

    def tables_hash(self, threads=0, live=False, target_size=0.25):
        """FNV-1a of the five device tables (nodes, interpreter stream, tree stream, regions, node ranks) as hex strings."""
        out = (C.c_uint64 * 5)()
        _check(lib().tg_debug_tables_hash(self.h, target_size, threads, 1 if live else 0, out))
        return ["%016x" % v for v in out]

    def plan_slabs(self, grid, ranks, target_size=0.25):
        """(cuts, per-layer cost estimate) of a multi-GPU export of this tree over `ranks` devices (host only)."""
        cuts = (C.c_uint64 * (ranks + 1))()
        cost = np.zeros(grid.shape[2], np.float64)
        _check(lib().tg_tree_plan_slabs(self.h, target_size, C.byref(grid), ranks, cuts, cost.ctypes.data_as(C.POINTER(C.c_double))))
        return [int(c) for c in cuts], cost

    def octree_stats(self, target_size=0.25, threads=0, live=False):
        """live=True: the live mesher's octree; bounds_min / bounds_max are then the octree's own Bounds."""
        s = ModelStats()
        _check((lib().tg_tree_octree_stats_live if live else lib().tg_tree_octree_stats)(self.h, target_size, threads, C.byref(s)))
        return s.as_dict()


def material(r, g, b):
    return int(lib().tg_material_create(r, g, b))


def export_grid(lo, hi, step):
    g = Grid()
    lo, hi, step = _f3(lo), _f3(hi), _f3(step)
    _check(lib().tg_export_grid(_fp(lo), _fp(hi), _fp(step), C.byref(g)))
    return g


class Context:
    """One CUDA device, or -- Context(devices=[0, 1, ...]) -- the GPUs of one box driven together (tg_context_create_multi)."""

    def __init__(self, device=0, devices=None):
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.h = _handle(lib().tg_context_create_multi(arr, len(devices)))
        else:
            self.h = _handle(lib().tg_context_create(device))

    @property
    def device_count(self):
        return int(lib().tg_context_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().tg_context_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def timer_begin(self):
        _check(lib().tg_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float()
        _check(lib().tg_timer_end(self.h, C.byref(ms)))
        return ms.value

    def fp32_peak_tflops(self):
        v = C.c_double()
        _check(lib().tg_measure_fp32_peak(self.h, C.byref(v)))
        return v.value

    def flush_l2(self):
        _check(lib().tg_flush_l2(self.h))

    def synchronize(self):
        _check(lib().tg_context_synchronize(self.h))

    def progress(self):
        ratios = (C.c_float * 4)()
        stage = C.c_int()
        _check(lib().tg_progress(self.h, ratios, C.byref(stage)))
        return stage.value, [float(r) for r in ratios]

    def cancel(self, halt=True):
        _check(lib().tg_cancel(self.h, 1 if halt else 0))

    def rearm(self):
        _check(lib().tg_rearm(self.h))

    def weld(self, vertices):
        """MeshGenerator::Accumulate over a vertex stream: (distinct vertices as (n, 4) with w = 1, index per input vertex)."""
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        out = np.zeros((max(len(v), 1), 4), np.float32)
        idx = np.zeros(max(len(v), 1), np.uint32)
        unique = C.c_uint64(0)
        _check(lib().tg_weld(self.h, _fp(v), len(v), _fp(out), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(unique)))
        return out[:unique.value].copy(), idx[:len(v)].copy()


class Mesh:
    """Result of an export; owns the library-side buffers until closed."""

    def __init__(self, raw, model=None):
        self.raw = raw
        # the arrays are blocks of the context's pinned cache: keep model and context alive for as long as the mesh is
        self._model = model
        self.layer_vertices = (np.ctypeslib.as_array(raw.layer_vertices, shape=(int(raw.layer_count),)).copy()
                               if raw.layer_vertices and raw.layer_count else np.zeros(0, np.uint32))
        self.layer_vertex_cost = (np.ctypeslib.as_array(raw.layer_vertex_cost, shape=(int(raw.layer_count),)).copy()
                                  if raw.layer_vertex_cost and raw.layer_count else np.zeros(0, np.float64))
        self._bind()

    def download(self, index_base=0):
        """Bring a MESH_DEVICE_ONLY result to the host, adding index_base to the triangle indices on the device."""
        _check(lib().tg_mesh_download(C.byref(self.raw), int(index_base) & 0xFFFFFFFF))
        self._bind()

    def _bind(self):
        raw = self.raw
        nv, nt = int(raw.vertex_count), int(raw.triangle_count)
        self.vertex_count, self.triangle_count = nv, nt
        self.halo_vertices = int(raw.halo_vertices)
        self.timings = raw.timings.as_dict()

        def arr(ptr, n, width, dtype):
            if not ptr or n == 0:
                return None
            return np.ctypeslib.as_array(ptr, shape=(n, width))

        self.positions = arr(raw.positions, nv, 3, np.float32)
        self.normals = arr(raw.normals, nv, 3, np.float32)
        self.colors = arr(raw.colors, nv, 3, np.uint8)
        self.triangles = arr(raw.triangles, nt, 3, np.uint32)
        self.face_normals = arr(raw.face_normals, nt, 3, np.float32)
        if nv == 0:
            self.positions = np.zeros((0, 3), np.float32)
        if nt == 0:
            self.triangles = np.zeros((0, 3), np.uint32)

    def rank_info(self):
        """Multi-GPU exports: [(slab_begin, slab_end, timings dict)] per rank; [] for a single-device result."""
        out = []
        rank = 0
        while True:
            b, e, t = C.c_uint64(), C.c_uint64(), MeshTimings()
            n = lib().tg_mesh_rank_info(C.byref(self.raw), rank, C.byref(b), C.byref(e), C.byref(t))
            if n <= 0:
                break
            out.append((int(b.value), int(e.value), t.as_dict()))
            rank += 1
            if rank >= n:
                break
        return out

    def write_ply(self, path):
        _check(lib().tg_write_ply(os.fsencode(path), C.byref(self.raw)))

    def write_stl(self, path):
        _check(lib().tg_write_stl(os.fsencode(path), C.byref(self.raw)))

    def close(self):
        if self.raw is not None:
            self.positions = self.normals = self.colors = self.triangles = self.face_normals = None
            lib().tg_mesh_free(C.byref(self.raw))
            self.raw = None

    def __del__(self):
        self.close()


class Model:
    def __init__(self, context, tree, target_size=0.25, threads=0, live=False):
        """live=True: the live mesher's octree (tg_model_create_live) instead of the export's."""
        self.context = context
        create = lib().tg_model_create_live if live else lib().tg_model_create
        self.h = _handle(create(context.h, tree.h, target_size, threads))

    def live_grid(self, density=20.0):
        """The live mesher's grid for a meshing density (sodapop.cpp:153-179); needs live=True."""
        grid = Grid()
        _check(lib().tg_live_grid(self.h, density, C.byref(grid)))
        return grid

    def close(self):
        if getattr(self, "h", None):
            lib().tg_model_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def stats(self):
        s = ModelStats()
        _check(lib().tg_model_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def upload(self):
        _check(lib().tg_model_upload(self.h))

    def brick_profile(self, grid):
        n = (grid.shape[2] + 7) // 8
        out = np.zeros(n, np.uint32)
        _check(lib().tg_brick_profile(self.h, C.byref(grid), out.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return out

    def eval_points(self, points, mode=EVAL_OCTREE):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        n = len(pts)
        if mode == EVAL_GRADIENT:
            out = np.zeros((n, 3), np.float32)
        elif mode == EVAL_COLOR:
            out = np.zeros((n, 3), np.uint8)
        else:
            out = np.zeros(n, np.float32)
        _check(lib().tg_eval_points(self.h, mode, _fp(pts), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def ray_cast(self, rays, max_iterations=100, epsilon=0.001, magnet=False):
        """rays: (n, 6) origin + direction (or + target with magnet).  Returns (hit bool[n], travel[n], position[n, 3])."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 5), np.float32)
        _check(lib().tg_ray_cast(self.h, _fp(rays), len(rays), max_iterations, epsilon, 1 if magnet else 0, _fp(out)))
        return out[:, 0] != 0, out[:, 1].copy(), out[:, 2:5].copy()

    def check_long_programs(self, reach=0.5):
        out = (C.c_uint64 * 3)()
        _check(lib().tg_debug_check_long_programs(self.h, reach, out))
        return tuple(int(v) for v in out)

    def eval_lattice(self, grid, download=True, flags=0):
        sx, sy, sz = grid.shape
        ms = C.c_float()
        out = np.zeros((sz + 1, sy + 1, sx + 1), np.float32) if download else None
        _check(lib().tg_eval_lattice_flags(self.h, C.byref(grid), flags, _fp(out) if download else None, C.byref(ms)))
        return out, ms.value

    def export_mesh(self, grid, flags=MESH_NORMALS | MESH_COLORS, refine=0, scale=1.0, slab=None):
        opt = MeshOptions(flags, refine, scale, 0, 0)
        if slab is not None:
            opt.slab_begin, opt.slab_end = int(slab[0]), int(slab[1])
        raw = _Mesh()
        _check(lib().tg_export_mesh(self.h, C.byref(grid), C.byref(opt), C.byref(raw)))
        return Mesh(raw, self)

    def export_points(self, lo, hi, step, refine=0, flags=MESH_NORMALS | MESH_COLORS, scale=1.0):
        lo, hi, step = _f3(lo), _f3(hi), _f3(step)
        raw = _Mesh()
        _check(lib().tg_export_points(self.h, _fp(lo), _fp(hi), _fp(step), refine, flags, scale, C.byref(raw)))
        return Mesh(raw, self)

    def export_voxels(self, grid_size):
        size = (C.c_int32 * 3)()
        radius = C.c_float()
        ptr = C.POINTER(C.c_int32)()
        count = C.c_uint64()
        _check(lib().tg_export_voxels(self.h, grid_size, size, C.byref(radius), C.byref(ptr), C.byref(count)))
        n = int(count.value)
        xyz = np.ctypeslib.as_array(ptr, shape=(max(n, 1), 3))[:n].copy()
        lib().tg_free(ptr)
        return tuple(size), radius.value, xyz


def export_ply(tree, grid_size, refine, path, device=0):
    _check(lib().tg_export_ply(tree.h, grid_size, refine, os.fsencode(path), device))


def export_stl(tree, grid_size, refine, path, device=0):
    _check(lib().tg_export_stl(tree.h, grid_size, refine, os.fsencode(path), device))


def export_magica_voxel(tree, grid_size, color_index, path, device=0):
    _check(lib().tg_export_magica_voxel(tree.h, grid_size, color_index, os.fsencode(path), device))
