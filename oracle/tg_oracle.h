/*
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
 *
 * tg_oracle: a plain-C, CPU restatement of the reference's SDF meshing hot path
 * (Aeva/tangerine: tangerine/sdf_evaluator.cpp, tangerine/export.cpp, tangerine/magica.cpp,
 * third_party/naive-surface-nets/src/surface_nets.cpp).  Every function cites the reference
 * file:line it follows.  It is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own (SURVEY.md
 * section 4), so this restatement is pinned against outputs of the reference implementation itself,
 * compiled unmodified into oracle/_ref/tangerine_ref (tests/test_oracle_vs_ref.py, run wherever
 * that binary exists) and against the committed fixtures under tests/golden/ generated from it by
 * tests/golden/make_golden.py.
 */
#ifndef TG_ORACLE_H
#define TG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TgoModel TgoModel;   /* CSG tree + material colours (a .tgm file) */
typedef struct TgoOctree TgoOctree; /* SDFOctree with one pruned tree + postfix program per node */

typedef struct TgoGrid
{
	float x, y, z;       /* origin            (regular_grid.h:29-37) */
	float dx, dy, dz;    /* cell size                                */
	uint64_t sx, sy, sz; /* cells per axis                           */
} TgoGrid;

typedef struct TgoMesh
{
	float* vertices;     /* 3 floats per vertex, lexicographic (k, j, i) cell order */
	int64_t* cells;      /* linear cell index i + j*sx + k*sx*sy per vertex         */
	uint32_t* triangles; /* 3 indices per triangle, cell order then edge 0..2       */
	uint64_t vertex_count;
	uint64_t triangle_count;
} TgoMesh;

typedef struct TgoOctreeStats
{
	uint64_t nodes, leaves, words, leaf_words, max_words, max_stack, hash;
} TgoOctreeStats;

TgoModel* tgo_model_load(const char* tgm_path);
void tgo_model_free(TgoModel* model);
void tgo_model_bounds(const TgoModel* model, float out_min[3], float out_max[3]);
int tgo_model_has_paint(const TgoModel* model);
int tgo_model_leaf_count(const TgoModel* model);
uint64_t tgo_model_root_program(const TgoModel* model, uint32_t* out_words, uint64_t capacity);

TgoOctree* tgo_octree_create(const TgoModel* model, float target_size);
/* The live mesher's octree (sodapop.cpp:240, 568-571: no coalescing, populated in two steps).  tgo_eval_octree,
 * tgo_lattice_samples and tgo_surface_nets on it sample the live mesher's implicit function (sodapop.cpp:583-587). */
TgoOctree* tgo_octree_create_live(const TgoModel* model, float target_size);
void tgo_octree_free(TgoOctree* octree);
void tgo_octree_stats(const TgoOctree* octree, TgoOctreeStats* out);
void tgo_octree_bounds(const TgoOctree* octree, float out_min[3], float out_max[3]);

/* SDFOctree::Eval (interpreter of the node picked by Descend), SDFNode::Eval on the root tree,
 * SDFInterpreter::Eval on the root program, SDFOctree::Gradient, export colour bytes. */
void tgo_eval_octree(const TgoOctree* octree, const float* points, uint64_t count, float* out, int threads);
void tgo_eval_tree(const TgoModel* model, const float* points, uint64_t count, float* out, int threads);
void tgo_eval_interp(const TgoModel* model, const float* points, uint64_t count, float* out);
/* SDFNode::RayMarch (sdf_evaluator.cpp:336-354): rays = 6 floats (origin, direction), out5 = hit, travel, position */
void tgo_ray_march(const TgoModel* model, const float* rays, uint64_t count, int max_iterations, float epsilon, int magnet, float* out5);
void tgo_gradient(const TgoOctree* octree, const float* points, uint64_t count, float* out3);
void tgo_color(const TgoOctree* octree, const float* points, uint64_t count, uint8_t* out3);

/* MeshExportThread's grid (export.cpp:324-337) from model bounds and a step. */
void tgo_export_grid(const float model_min[3], const float model_max[3], const float step[3], TgoGrid* out);

/* MeshGenerator::Accumulate(vertex) for every vertex of a stream, in order (tangerine/mesh_generators.cpp:20-50): the
 * distinct vertices as (x, y, z, 1) in order of first occurrence and one index per input vertex.  Equal = LessVec3's
 * equivalence (numeric equality per component: -0 is +0).  Returns the number of distinct vertices. */
uint64_t tgo_weld(const float* vertices, uint64_t count, float* out_vertices4, uint32_t* out_indices);

/* NaiveSurfaceNetsScratch's grid (sodapop.cpp:153-179) from a live octree's bounds and the meshing density. */
void tgo_live_grid(const TgoOctree* octree, float meshing_density, TgoGrid* out);

/* par_surface_nets restated; lattice samples are computed once and shared between cells. */
int tgo_surface_nets(const TgoOctree* octree, const TgoGrid* grid, TgoMesh* out, int threads);
void tgo_mesh_free(TgoMesh* mesh);

/* Lattice samples only: (sx+1)*(sy+1)*(sz+1) floats, x fastest. */
void tgo_lattice_samples(const TgoOctree* octree, const TgoGrid* grid, float* out, int threads);

/* Refinement loop of export.cpp:433-469 applied in place to `count` points. */
void tgo_refine(const TgoOctree* octree, float* points, uint64_t count, const float half[3], int iterations);

/* PointCloudExportThread generation pass (export.cpp:393-428); returns count, points malloc'd. */
uint64_t tgo_point_cloud(const TgoOctree* octree, const float min[3], const float max[3], const float step[3], float** out_points);

/* VoxExport occupancy (magica.cpp:27-69); voxels as x,y,z int32 triples in flat-index order. */
uint64_t tgo_voxels(const TgoModel* model, float grid_size, int32_t out_size[3], float* out_radius, int32_t** out_xyz, int threads);

void tgo_free(void* pointer);

#ifdef __cplusplus
}
#endif
#endif
