/*
 * tangerine_b200 -- C ABI of the B200 (sm_100a) SDF meshing path.
 *
 * This is the drop-in boundary for the reference's export hot path (Aeva/tangerine).  Host code stays
 * C++ and reaches CUDA only through these entry points: plain pointers and sizes, `int` status per call
 * (0 = TG_OK) plus a thread-local error string; nothing here ever aborts the host process (the
 * reference's own convention is print + abort(), tangerine/errors.cpp:23-32).
 *
 * Each group cites the reference interface it replaces.  The precedent for a C ABI over this path is
 * the reference's own legacy FFI: tangerine/c_sdf.cpp:27-180 (tree building, EvalTree),
 * tangerine/export.cpp:611-622 (ExportSTL / ExportPLY) and tangerine/magica.cpp:77-84
 * (ExportMagicaVoxel), bound from Racket in package/tangerine/export.rkt:26-28 and eval.rkt:37-62.
 *
 * Threading: a tg_context is used by one thread at a time (the reference runs an export on one
 * detached std::thread, export.cpp:576-578); tg_progress and tg_cancel may be called concurrently
 * from another thread, like GetExportProgress / CancelExport are from the UI thread.  Trees are plain
 * values with no hidden sharing.
 */
#ifndef TANGERINE_B200_H
#define TANGERINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TG_API __declspec(dllexport)
#else
#define TG_API __attribute__((visibility("default")))
#endif

enum
{
	TG_OK = 0,
	TG_ERR_INVALID = 1,   /* bad argument / malformed input */
	TG_ERR_NO_DEVICE = 2, /* CUDA device or kernel image unavailable: there is no CPU fallback */
	TG_ERR_CUDA = 3,      /* a CUDA call failed; see tg_last_error() */
	TG_ERR_MEMORY = 4,
	TG_ERR_IO = 5,
	TG_ERR_CANCELLED = 6,
	TG_ERR_UNSUPPORTED = 7 /* e.g. CSG nesting deeper than the device operand stack */
};

/* Thread-local description of the most recent failure on this thread. */
TG_API const char* tg_last_error(void);
TG_API const char* tg_version(void);

/* ------------------------------------------------------------------------------------------------
 * CSG trees.  Replaces: SDF:: constructors (tangerine/sdf_evaluator.h:223-264, .cpp:1206-1371) and the
 * legacy C ABI Make*Brush / Make*Op / MoveTree / RotateTree / ... (tangerine/c_sdf.cpp:27-180).
 * Arguments have the reference's meaning (radii / half extents, not the Lua layer's diameters).
 * Operators copy their operands (the Lua layer deep-copies on every modifier too, lua_sdf.cpp:56-62);
 * the caller keeps ownership of everything it was handed and frees each tree with tg_tree_free.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tg_tree tg_tree;

TG_API tg_tree* tg_make_sphere(float radius);
TG_API tg_tree* tg_make_ellipsoid(float radipode_x, float radipode_y, float radipode_z);
TG_API tg_tree* tg_make_box(float extent_x, float extent_y, float extent_z);
TG_API tg_tree* tg_make_torus(float major_radius, float minor_radius);
TG_API tg_tree* tg_make_cylinder(float radius, float extent);
TG_API tg_tree* tg_make_plane(float normal_x, float normal_y, float normal_z);
TG_API tg_tree* tg_make_cone(float radius, float height);
TG_API tg_tree* tg_make_coninder(float radius_l, float radius_h, float height);

TG_API tg_tree* tg_make_union(const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_diff(const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_inter(const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_blend_union(float threshold, const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_blend_diff(float threshold, const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_blend_inter(float threshold, const tg_tree* lhs, const tg_tree* rhs);
TG_API tg_tree* tg_make_flate(const tg_tree* child, float radius);
/* SDF::Stencil (sdf_evaluator.cpp:1361-1371): `material` overrides the child's paint where the mask is
 * negative (apply_to_negative != 0, Lua `stencil`) or non-negative (apply_to_negative == 0, Lua `mask`). */
TG_API tg_tree* tg_make_stencil(const tg_tree* child, const tg_tree* mask, uint32_t material, int apply_to_negative);

TG_API tg_tree* tg_tree_copy(const tg_tree* tree);
TG_API void tg_tree_free(tg_tree* tree);

/* In-place modifiers: SDFNode::Move / Rotate / Scale / ApplyMaterial, SDF::Align, SDF::RotateX/Y/Z. */
TG_API int tg_tree_move(tg_tree* tree, float x, float y, float z);
TG_API int tg_tree_rotate(tg_tree* tree, float quat_x, float quat_y, float quat_z, float quat_w);
TG_API int tg_tree_rotate_x(tg_tree* tree, float degrees);
TG_API int tg_tree_rotate_y(tg_tree* tree, float degrees);
TG_API int tg_tree_rotate_z(tg_tree* tree, float degrees);
TG_API int tg_tree_scale(tg_tree* tree, float scale);
TG_API int tg_tree_align(tg_tree* tree, float anchor_x, float anchor_y, float anchor_z);
TG_API int tg_tree_paint(tg_tree* tree, uint32_t material, int force);

/* A material is what the export path reads from one: its sampled sRGB base colour,
 * SampleColor(Material->GuessColor()) (tangerine/export.cpp:303-307).  Returns a process-wide id. */
TG_API uint32_t tg_material_create(float red, float green, float blue);

/* Queries: SDFNode::Eval (legacy EvalTree), Bounds, HasPaint, HasFiniteBounds, LeafCount. */
TG_API float tg_tree_eval(const tg_tree* tree, float x, float y, float z);
TG_API int tg_tree_bounds(const tg_tree* tree, float out_min[3], float out_max[3]);
TG_API int tg_tree_has_paint(const tg_tree* tree);
TG_API int tg_tree_has_finite_bounds(const tg_tree* tree);
TG_API int tg_tree_leaf_count(const tg_tree* tree);

/* Portable tree files (.tgm, layout in tangerine_b200/csrc/tg_tree.h). */
TG_API tg_tree* tg_tree_load(const char* path);
TG_API int tg_tree_save(const tg_tree* tree, const char* path);

/* The synthetic benchmark scene, config C4 (BASELINE.json configs[3]): `primitives` random brushes (mt19937(seed)) in
 * clusters of 8 folded left to right with smooth unions / differences, the clusters joined by a balanced tree of
 * unions, clipped to a 10-unit cube.  (Not one deep fold: SDFOctree::Create is exponential in blend-chain depth,
 * sdf_evaluator.cpp:793-816, so the reference could not build that tree.) */
TG_API tg_tree* tg_make_synthetic(uint32_t primitives, uint32_t seed);

/* ------------------------------------------------------------------------------------------------
 * Contexts and models.  A context owns one CUDA device, its streams and scratch memory.
 * tg_model_create replaces SDFOctree::Create(Evaluator, 0.25) (tangerine/sdf_evaluator.cpp:1609-1783,
 * called from export.cpp:322): it builds the pruning octree on host threads, flattens every node's
 * pruned subtree into device instruction streams and uploads them.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tg_context tg_context;
typedef struct tg_model tg_model;

TG_API tg_context* tg_context_create(int cuda_device);
/* Meshes and models may outlive their context: their arrays are blocks of the context's caches, so a context that still
 * has live meshes or models is torn down when the last of them is freed (it can start no new export meanwhile).  A
 * context is worth keeping for the life of the application: it owns the scratch arena, the result caches and the
 * page-locked staging memory, which cost more to make than an export. */
TG_API void tg_context_destroy(tg_context* context);
TG_API int tg_context_device(const tg_context* context);
/* Multi-GPU context: the devices of one box, driven from this one process (SURVEY.md 8b `tg_ctx_create(devices[], n)`).
 * Models created on it are replicated to every device; tg_export_mesh then cuts the grid into z-slabs, one per device,
 * combines the per-slab vertex counts with an ncclAllGather over NVLink and returns ONE host mesh, exactly as on one
 * device.  Needs libnccl.so.2 at run time when count > 1 (TG_ERR_UNSUPPORTED otherwise); every other
 * call on such a context runs on its first device. */
TG_API tg_context* tg_context_create_multi(const int* cuda_devices, int count);
TG_API int tg_context_device_count(const tg_context* context);

typedef struct tg_model_stats
{
	uint64_t octree_nodes;
	uint64_t octree_leaves;
	uint64_t reference_words;      /* total program size in the reference's word encoding */
	uint64_t reference_leaf_words;
	uint64_t reference_max_words;
	uint64_t max_stack;            /* largest SDFNode::StackSize */
	uint64_t octree_hash;          /* FNV-1a over (pivot, terminus, child mask, reference words), pre-order */
	uint64_t device_bytes;         /* size of the uploaded tables */
	double build_seconds;          /* host octree build + flatten */
	double upload_seconds;
	float bounds_min[3];
	float bounds_max[3];
	int32_t has_paint;
	int32_t leaf_count;
} tg_model_stats;

TG_API tg_model* tg_model_create(tg_context* context, const tg_tree* tree, float octree_target_size, int host_threads);
/* The live mesher's model (Sodapop, tangerine/sodapop.cpp:227-247 and 562-600): the octree of
 * SDFOctree::Create(Evaluator, .25, false, 3, 0.0) with every incomplete node populated -- nothing is coalesced.  Use it
 * with tg_live_grid, TG_MESH_LIVE_FIELD and TG_EVAL_LIVE; every other call works on it as on any model. */
TG_API tg_model* tg_model_create_live(tg_context* context, const tg_tree* tree, float octree_target_size, int host_threads);
/* Host half of tg_model_create only (octree build + flattening, no device needed): fills the octree_*,
 * reference_*, max_stack, bounds, has_paint, leaf_count and build_seconds fields. */
TG_API int tg_tree_octree_stats(const tg_tree* tree, float octree_target_size, int host_threads, tg_model_stats* out);
/* The same for the live mesher's octree (tg_model_create_live); bounds_min / bounds_max are then the octree's own Bounds,
 * from which tg_live_grid makes its grid. */
TG_API int tg_tree_octree_stats_live(const tg_tree* tree, float octree_target_size, int host_threads, tg_model_stats* out);
TG_API void tg_model_destroy(tg_model* model);
/* Copies the model's tables (octree nodes, both instruction streams, material colours) host -> device again.
 * tg_model_create already did this once; bench.py calls it inside its end-to-end timed region. */
TG_API int tg_model_upload(tg_model* model);
TG_API int tg_model_get_stats(const tg_model* model, tg_model_stats* out);

/* ------------------------------------------------------------------------------------------------
 * Point queries (parity dumps and the Lua-side eval / gradient calls).
 *   TG_EVAL_OCTREE    SDFOctree::Eval       (sdf_evaluator.cpp:1969-1989)  what the mesher samples
 *   TG_EVAL_INTERP    SDFInterpreter::Eval on the unpruned model (sdf_evaluator.cpp:1386-1605)
 *   TG_EVAL_TREE      SDFNode::Eval on the unpruned model        (magica.cpp:61 samples this)
 *   TG_EVAL_GRADIENT  SDFOctree::Gradient   (sdf_evaluator.h:327-331)      3 floats per point
 *   TG_EVAL_COLOR     export colour bytes   (export.cpp:297-312)           3 bytes per point
 *   TG_EVAL_LIVE      the live mesher's implicit function (sodapop.cpp:583-587): clamp(SDFOctree::Eval(p, Exact = false),
 *                     -100, 100) -- an empty octant yields +infinity, i.e. 100, instead of the parent's program
 * `points` and `out` are host pointers; count points of 3 floats.
 * ---------------------------------------------------------------------------------------------- */
enum { TG_EVAL_OCTREE = 0, TG_EVAL_INTERP = 1, TG_EVAL_TREE = 2, TG_EVAL_GRADIENT = 3, TG_EVAL_COLOR = 4, TG_EVAL_LIVE = 5 };
TG_API int tg_eval_points(tg_model* model, int mode, const float* points, uint64_t count, void* out);

/* Batched ray casts: SDFNode::RayMarch (tangerine/sdf_evaluator.cpp:336-354) on the unpruned model, the call behind the
 * Lua methods `ray_cast` and `magnet` (lua_sdf.cpp:410-444, 745-749; their defaults are max_iterations = 100,
 * epsilon = 0.001).  rays: 6 floats each, origin then direction (magnet != 0: origin then TARGET, the direction being
 * normalize(target - origin)).  out_hits: 5 floats per ray -- hit (1 or 0), travel (infinity on a miss), position. */
TG_API int tg_ray_cast(tg_model* model, const float* rays, uint64_t count, int max_iterations, float epsilon, int magnet, float* out_hits);

/* Vertex welding: MeshGenerator::Accumulate (tangerine/mesh_generators.cpp:20-80, used by the lattice mesher's "combined
 * vertices" mode, sodapop.cpp:1287, 1508) over a stream of `count` vertices (3 floats each; a triangle soup is three per
 * triangle).  out_indices[i] is the index of vertex i among the distinct vertices in order of first occurrence -- equal
 * means numerically equal per component, so -0 welds with +0 -- and out_vertices4 receives those as (x, y, z, 1) with
 * the bits of their first occurrence; both have room for `count` entries, *out_unique tells how many vertices there are. */
TG_API int tg_weld(tg_context* context, const float* vertices, uint64_t count, float* out_vertices4, uint32_t* out_indices, uint64_t* out_unique);

/* Diagnostics for the tests: FNV-1a hashes of the five device tables tg_model_create would upload for this tree (octree
 * nodes, interpreter stream, tree stream, regions, node cost ranks), built on `host_threads` threads (live != 0: the live
 * mesher's octree).  The tables must not depend on the thread count, and host-side refactors must not change them. */
TG_API int tg_debug_tables_hash(const tg_tree* tree, float octree_target_size, int host_threads, int live, uint64_t out_hashes[5]);

/* Self-check used by the tests: the culling pass evaluates long programs cooperatively (a warp or a block per point,
 * parallel fold); this runs every long program of the model at 9 points within `reach` of its octree node's pivot both
 * ways and returns { probes, disagreements of the block form, disagreements of the warp form } -- the last two must be 0. */
TG_API int tg_debug_check_long_programs(tg_model* model, float reach, uint64_t out_counts[3]);
/* Diagnostics: the device instruction stream (kStreamInterp words, tangerine_b200/csrc/tg_program.h) of one octree node,
 * built on the host.  Returns the word count (0 = no such node). */
TG_API uint64_t tg_debug_node_program(const tg_tree* tree, uint32_t node, uint32_t* out_words, uint64_t capacity, uint32_t* out_instruction_count);

/* ------------------------------------------------------------------------------------------------
 * Mesh export.  Replaces the span export.cpp:324-365 (grid set-up, isosurface::par_surface_nets with
 * the octree as implicit function, mesh conversion) plus the per-vertex attribute loops
 * export.cpp:297-314 (PLY) / 130-140 (STL), and the refinement loop export.cpp:433-469 applied to the
 * mesh vertices when refine_iterations > 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tg_grid
{
	float x, y, z;       /* origin:      isosurface::regular_grid_t (regular_grid.h:29-37) */
	float dx, dy, dz;    /* cell size */
	uint64_t sx, sy, sz; /* cells per axis */
} tg_grid;

/* The live mesher's grid (NaiveSurfaceNetsScratch, sodapop.cpp:153-179) for a meshing density (Sodapop's default is 20,
 * sodapop.cpp:43, 221): floor(density) samples per unit of the octree's bounds, at least 8 per axis, two cells of margin
 * below and one above.  `model` must come from tg_model_create_live. */
TG_API int tg_live_grid(const tg_model* model, float density, tg_grid* out);

/* MeshExportThread's grid: ModelMin -= 2 * Step; Extent = ceil((ModelMax - ModelMin) / Step) (export.cpp:324-337). */
TG_API int tg_export_grid(const float model_min[3], const float model_max[3], const float step[3], tg_grid* out);

enum
{
	TG_MESH_NORMALS = 1u << 0,      /* per-vertex SDFOctree::Gradient */
	TG_MESH_COLORS = 1u << 1,       /* per-vertex export colour; ignored (white) when the model has no paint */
	TG_MESH_NO_CULL = 1u << 2,      /* evaluate every brick, as the reference does (dense sweep) */
	TG_MESH_DEVICE_ONLY = 1u << 3,  /* leave results in HBM: out->positions etc. are NULL, counts are valid */
	TG_MESH_FACE_NORMALS = 1u << 4, /* per-triangle gradient at the centroid (WriteSTL, export.cpp:130-140) */
	TG_MESH_KEEP_CANCEL = 1u << 5,  /* do not re-arm the context: a tg_cancel issued before this call still cancels it */
	/* Opt-in fast arithmetic for the lattice evaluation: FMA contraction, approximate sqrt / division, float where the
	 * reference promotes to double.  Samples stay within BASELINE.json's tolerance (1e-5 relative / 4 ULP) of the
	 * reference but are no longer bit-identical, so a cell whose corner value is within that tolerance of zero may change
	 * its classification.  Without this flag every sample, vertex, normal and colour is bit-identical to the reference. */
	TG_MESH_FAST = 1u << 6,
	/* Multi-GPU contexts: after this export, move the slab cuts of the next export of the same model and grid by the
	 * per-device times just measured.  Without it every export runs on the cuts of the host-side estimate alone. */
	TG_MESH_REBALANCE = 1u << 7,
	/* Mesh (or sample, tg_eval_lattice_flags) the live mesher's field instead of the export's: TG_EVAL_LIVE at every
	 * lattice point.  On a tg_model_create_live model over tg_live_grid this is Sodapop's NaiveSurfaceNets mesh
	 * (sodapop.cpp:562-760): its first loop only visits the cells of the octree leaves' boxes ("point cache", :624-652), and
	 * outside of them the clamped field is +100 everywhere, so the meshes are the same (tests/test_gpu_live.py). */
	TG_MESH_LIVE_FIELD = 1u << 8
};

typedef struct tg_mesh_options
{
	uint32_t flags;
	int32_t refine_iterations; /* 0 = positions exactly as surface nets produced them */
	float scale;               /* positions are multiplied by this last (export.cpp:313); 0 means 1 */
	/* z-slab for multi-GPU runs: this call owns cell layers [slab_begin, slab_end) of the grid (any layer
	 * boundary; bricks that straddle a boundary are evaluated by both neighbours, each for its own layers).
	 * Both zero = the whole grid.  See tg_mesh.halo_vertices. */
	uint64_t slab_begin, slab_end;
} tg_mesh_options;

typedef struct tg_mesh_timings
{
	/* device milliseconds measured with CUDA events on the context's stream */
	float cull_ms, evaluate_ms, compact_ms, faces_ms, attributes_ms, total_device_ms;
	float download_ms; /* device -> host copies (0 with TG_MESH_DEVICE_ONLY) */
	uint64_t bricks_total, bricks_evaluated;
	uint64_t samples_evaluated;      /* lattice samples actually run through the interpreter */
	uint64_t algorithmic_flops;      /* sum over evaluated samples of their program's FLOP count */
	uint64_t kernel_launches;
} tg_mesh_timings;

typedef struct tg_mesh
{
	/* Library-owned host arrays (pinned); release with tg_mesh_free.  Vertices are ordered by cell,
	 * lexicographically in (k, j, i) -- the order the reference's serial loop produces
	 * (surface_nets.cpp:976-999).  Triangles are ordered by owning cell, then edge 0..2. */
	float* positions;      /* 3 per vertex */
	float* normals;        /* 3 per vertex, or NULL */
	uint8_t* colors;       /* 3 per vertex, or NULL */
	uint32_t* triangles;   /* 3 per triangle */
	float* face_normals;   /* 3 per triangle, or NULL */
	uint64_t vertex_count;
	uint64_t triangle_count;
	/* Slab runs: the first halo_vertices entries of the *local* numbering belong to the cell layer below
	 * the slab (owned by the previous rank) and are not part of positions[]; triangle indices are local
	 * numbers minus halo_vertices, so adding the previous ranks' vertex total makes them global. */
	uint64_t halo_vertices;
	/* Owned vertices per 8-layer brick layer of the grid (absolute layer index, 0 outside the slab): the measured
	 * profile a multi-GPU driver feeds back into its next slab cut (bench.py).  layer_count = ceil(sz / 8). */
	uint32_t* layer_vertices;
	double* layer_vertex_cost; /* same indexing: sum over the layer's vertices of their octree node's program FLOPs */
	uint64_t layer_count;
	tg_mesh_timings timings;
	void* opaque;
} tg_mesh;

TG_API int tg_export_mesh(tg_model* model, const tg_grid* grid, const tg_mesh_options* options, tg_mesh* out);
/* The z-slab cuts a multi-GPU export of `tree` on `grid` over `ranks` devices would use (host only, no device needed):
 * out_cuts receives ranks + 1 cell-layer indices, 0 ... grid->sz.  They come from a host-side estimate of the work per
 * cell layer made from the octree's terminus cells (optional out_layer_cost, grid->sz entries); no warm-up exports. */
TG_API int tg_tree_plan_slabs(const tg_tree* tree, float octree_target_size, const tg_grid* grid, int ranks, uint64_t* out_cuts, double* out_layer_cost);
/* Multi-GPU exports only: number of ranks (or -1 for a single-device result), and for `rank` its z-slab [begin, end) and
 * the stage timings of its device (the timings in tg_mesh are the maximum over the ranks). */
TG_API int tg_mesh_rank_info(const tg_mesh* mesh, int rank, uint64_t* out_slab_begin, uint64_t* out_slab_end, tg_mesh_timings* out_timings);
TG_API void tg_mesh_free(tg_mesh* mesh);
/* Second half of a TG_MESH_DEVICE_ONLY export: adds index_base to every triangle index on the device (multi-GPU:
 * the vertex total of the lower ranks, known once the per-slab counts were exchanged) and copies the arrays into
 * library-owned pinned host memory, filling the NULL pointers of *mesh. */
TG_API int tg_mesh_download(tg_mesh* mesh, uint32_t index_base);

/* Raw lattice samples of the grid through SDFOctree::Eval: (sx+1)*(sy+1)*(sz+1) floats, x fastest.
 * `out` may be NULL to time the evaluator alone; elapsed device milliseconds are returned in *out_ms. */
TG_API int tg_eval_lattice(tg_model* model, const tg_grid* grid, float* out, float* out_ms);
/* The same with mesh flags: TG_MESH_FAST evaluates with the fast arithmetic (tolerance tests). */
TG_API int tg_eval_lattice_flags(tg_model* model, const tg_grid* grid, uint32_t flags, float* out, float* out_ms);

/* ------------------------------------------------------------------------------------------------
 * Point-cloud export.  Replaces PointCloudExportThread's two Pool() passes (export.cpp:393-469).
 * ---------------------------------------------------------------------------------------------- */
/* `scale` multiplies the positions last, after normals and colours were sampled (WritePLY, export.cpp:313, 476); 0 means 1. */
TG_API int tg_export_points(tg_model* model, const float model_min[3], const float model_max[3], const float step[3],
	int refine_iterations, uint32_t flags, float scale, tg_mesh* out);

/* ------------------------------------------------------------------------------------------------
 * Voxel occupancy.  Replaces the Pool() loop of VoxExport (tangerine/magica.cpp:27-69).
 * out_xyz receives library-owned int32 triples in flat-index order; free with tg_free.
 * ---------------------------------------------------------------------------------------------- */
TG_API int tg_export_voxels(tg_model* model, float grid_size, int32_t out_size[3], float* out_radius,
	int32_t** out_xyz, uint64_t* out_count);
TG_API void tg_free(void* pointer);

/* ------------------------------------------------------------------------------------------------
 * Progress / cancel.  Replaces GetExportProgress / CancelExport (tangerine/export.h:27-39,
 * export.cpp:483-492, 582-592).  stage: 0 idle, 1 generate, 2 refine, 3 attributes + write.
 * ---------------------------------------------------------------------------------------------- */
TG_API int tg_progress(const tg_context* context, float out_ratios[4], int* out_stage);
TG_API int tg_cancel(tg_context* context, int halt);
/* Clears a tg_cancel so that a long-lived context can run its next export (MeshExport sets ExportActive again, export.cpp:568).
 * Exports started without TG_MESH_KEEP_CANCEL do this themselves. */
TG_API int tg_rearm(tg_context* context);

/* ------------------------------------------------------------------------------------------------
 * File-level entry points with the signatures of the reference's legacy FFI:
 *   ExportPLY / ExportSTL(tree, GridSize, RefineIterations, Path)   export.cpp:611-622
 *   ExportMagicaVoxel(tree, GridSize, ColorIndex, Path)             magica.cpp:77-84
 * They run ExportCommon's sequence (export.cpp:595-607) on CUDA device `cuda_device`.
 * ---------------------------------------------------------------------------------------------- */
TG_API int tg_export_ply(const tg_tree* tree, float grid_size, int refine_iterations, const char* path, int cuda_device);
TG_API int tg_export_stl(const tg_tree* tree, float grid_size, int refine_iterations, const char* path, int cuda_device);
TG_API int tg_export_magica_voxel(const tg_tree* tree, float grid_size, int color_index, const char* path, int cuda_device);

/* Writers on already-extracted data (WritePLY / WriteSTL byte layouts, export.cpp:60-108, 198-280). */
TG_API int tg_write_ply(const char* path, const tg_mesh* mesh);
TG_API int tg_write_stl(const char* path, const tg_mesh* mesh);

/* ------------------------------------------------------------------------------------------------
 * Measurement helpers for bench.py: device-side timing on the context's own stream, and an FP32
 * FMA-chain peak measurement (the roofline denominator for the evaluator; MEASURED_PEAKS.json has none).
 * ---------------------------------------------------------------------------------------------- */
TG_API int tg_timer_begin(tg_context* context);
TG_API int tg_timer_end(tg_context* context, float* out_ms);
TG_API int tg_measure_fp32_peak(tg_context* context, double* out_tflops);
TG_API int tg_flush_l2(tg_context* context);
TG_API int tg_context_synchronize(tg_context* context);

/* Multi-GPU partitioning: estimated work in each brick layer (layer b = cell layers [8b, 8b+8)):
 * every 8^3-cell brick that survives culling counts 64 + the FLOPs of the program at its centre; out_layers needs
 * ceil(sz / 8) entries.  Every rank computes the same profile and cuts the grid into z-slabs of equal work
 * without communicating. */
TG_API int tg_brick_profile(tg_model* model, const tg_grid* grid, uint32_t* out_layers, uint32_t layer_count);

#ifdef __cplusplus
}
#endif
#endif
