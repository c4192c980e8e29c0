// Host octree build + flattening into the device tables.
//
// BuildFlatModel reproduces SDFOctree::Create / the SDFOctree constructor / Populate
// (tangerine/sdf_evaluator.cpp:1609-1783) decision for decision -- same bounding cube, pivots, clip
// radii, empty-octant pruning and coalescing rule -- so that every sample the GPU evaluates runs the
// pruned program the reference would have picked for it (SDFOctree::Descend, :1801-1835).  The
// octants of the first two levels are built on worker threads, each with a private copy of the node
// pool; results are merged by a pre-order walk that also emits both device instruction streams.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "tg_program.h"
#include "tg_tree.h"

namespace tg
{

constexpr uint32_t kMixedMaterial = 0xFFFFFFFEu;

struct FlatModelStats
{
	uint64_t nodes = 0;        // octree nodes
	uint64_t leaves = 0;       // terminal nodes
	uint64_t ref_words = 0;    // size of all programs in the reference's word encoding
	uint64_t ref_leaf_words = 0;
	uint64_t ref_max_words = 0;
	uint64_t max_stack = 0;    // largest SDFNode::StackSize of any node program
	uint64_t hash = 0;         // FNV-1a over (pivot, terminus, child mask, reference words) in pre-order
	bool reference_done = false; // ref_* and hash were computed (BuildFlatModel's reference_stats)
	uint64_t interp_words = 0; // device stream sizes
	uint64_t tree_words = 0;
	double build_seconds = 0.0;
};

struct FlatModel
{
	std::vector<FlatNode> nodes;      // node 0 is the octree root
	std::vector<uint32_t> interp;     // kStreamInterp programs, addressed by FlatNode::interp_offset
	std::vector<uint32_t> tree;       // kStreamTree programs, addressed by FlatNode::tree_offset
	std::vector<FlatRegion> regions;  // evaluation regions (tg_program.h), pre-order
	std::vector<uint32_t> node_rank;  // position of every node when the nodes are sorted by program cost, costliest first
	// per node: the material GetMaterial returns WHEREVER its program is evaluated (every brush that can win carries the same
	// one; kNoMaterial counts as one), or kMixedMaterial when it depends on the point and the walk has to run
	std::vector<uint32_t> node_material;
	// Terminus cells of the octree, for the multi-GPU slab planner: node index, cell span, and -- filled by
	// Model::Create on the device, one bit per 1/4-span sub-cell (x + 4 y + 16 z) -- whether the cell's program can have
	// a zero there (|d(sub-cell centre)| <= its half diagonal).  Empty mask vector = nothing known (all set is assumed).
	std::vector<uint32_t> leaf_nodes;
	std::vector<float> leaf_span;
	std::vector<uint64_t> leaf_mask;
	std::vector<float> material_rgb;  // 3 floats per material id; one extra trailing entry = default white
	uint32_t root_tree_offset = 0;    // kStreamTree program of the *unpruned* model (VoxExport samples it, magica.cpp:61)
	uint32_t root_interp_offset = 0;  // kStreamInterp program of the unpruned model
	uint32_t root_flops = 0;
	Box3 bounds;                      // Evaluator->Bounds()
	bool live_octree = false;         // built with coalesce = false
	Box3 live_bounds;                 // live_octree: SDFOctree::Bounds of the root (the live mesher's grid comes from it, sodapop.cpp:153-179)
	bool has_paint = false;           // Octree->Evaluator->HasPaint() (export.cpp:285)
	FlatModelStats stats;
};

// Returns false (and sets error) when the tree has no finite bounds, prunes away entirely, or nests
// deeper than the device stack supports.  threads <= 0 picks std::thread::hardware_concurrency().
// reference_stats: also compile every node program into the reference's word encoding for FlatModelStats::ref_* and
// ::hash (diagnostics that pin the octree against the reference's; the device tables do not depend on them).
// coalesce = false builds the live mesher's octree instead of the export's (SDFOctree::Create(Evaluator, .25, false, 3, 0.0)
// followed by Populate(false, 3, -1) of every node the depth limit left incomplete, sodapop.cpp:240, 568-571): no node is
// coalesced, and FlatModel::live_bounds is the root's Bounds as that two-step construction leaves them.
bool BuildFlatModel(const Tree& tree, float target_size, int threads, FlatModel& out, std::string& error, bool reference_stats = true, bool coalesce = true);

} // namespace tg
