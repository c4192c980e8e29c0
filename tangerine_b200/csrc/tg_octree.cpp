// See tg_octree.h.  Line references are to the reference's tangerine/sdf_evaluator.cpp.
#include "tg_octree.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <future>
#include <memory>
#include <thread>

namespace tg
{

namespace
{

struct Subtree;

struct BuildNode
{
	Box3 bounds;
	Vec3 pivot;
	bool terminus = false;
	uint32_t evaluator = kNoNode; // index into the owning Subtree's pool
	int32_t evaluator_leaves = 0;
	// Empty octants (children[i] == -1) whose subtree the clip removed outright: the parent's program, evaluated at
	// the octant's centre empty_center[i], was empty_value[i] > the octant's half diagonal.  0 = nothing known.
	float clip_value = 0.0f;  // value of the parent's program at this node's pivot, when the clip pruned everything
	Vec3 empty_center[8];
	float empty_value[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	// >= 0: index into Subtree::nodes; -1: empty octant; <= -2: root of Subtree::spawned[-2 - value]
	int32_t children[8] = { -1, -1, -1, -1, -1, -1, -1, -1 };
};

struct Subtree
{
	NodePool pool;
	std::vector<BuildNode> nodes;
	std::vector<std::unique_ptr<Subtree>> spawned;
	int32_t root = -1;
};

struct Builder
{
	float target_size;
	int parallel_depth; // nodes at depth <= parallel_depth build their octants on worker threads

	// SDFOctree::SDFOctree :1641-1700 with Coalesce = true, MaxDepth = -1 (what MeshExportThread asks for)
	int32_t Construct(Subtree& st, uint32_t in_evaluator, Box3 bounds, int depth)
	{
		const int32_t self = int32_t(st.nodes.size());
		st.nodes.emplace_back();
		Vec3 extent = bounds.max - bounds.min;
		float span = std::fmax(std::fmax(extent.x, extent.y), extent.z);
		Vec3 pivot = Vec3(float(span * 0.5)) + bounds.min;
		float radius = float(Length(Vec3(span)) * 0.5);
		float clip_value = 0.0f;
		uint32_t evaluator = st.pool.Clip(in_evaluator, pivot, radius, &clip_value);
		{
			BuildNode& n = st.nodes[self];
			n.clip_value = evaluator == kNoNode && clip_value > radius ? clip_value : 0.0f;
			n.bounds = bounds;
			n.pivot = pivot;
			n.evaluator = evaluator;
			n.evaluator_leaves = evaluator != kNoNode ? st.pool.nodes[evaluator].leaf_count : 0;
			n.terminus = span <= target_size || evaluator == kNoNode;
		}
		if (!st.nodes[self].terminus)
		{
			Populate(st, self, depth);
		}
		return self;
	}

	static Box3 Octant(const Box3& bounds, Vec3 pivot, int i) // :1714-1738
	{
		Box3 cb = bounds;
		if (i & 1) cb.min.x = pivot.x; else cb.max.x = pivot.x;
		if (i & 2) cb.min.y = pivot.y; else cb.max.y = pivot.y;
		if (i & 4) cb.min.z = pivot.z; else cb.max.z = pivot.z;
		return cb;
	}

	// SDFOctree::Populate :1703-1783
	void Populate(Subtree& st, int32_t self, int depth)
	{
		const Box3 bounds = st.nodes[self].bounds;
		const Vec3 pivot = st.nodes[self].pivot;
		const uint32_t evaluator = st.nodes[self].evaluator;
		bool uniform = true;
		bool penultimate = true;
		int live = 0;

		if (depth <= parallel_depth)
		{
			// Each octant prunes against a private copy of the pool, so the eight can run concurrently.
			std::future<std::unique_ptr<Subtree>> jobs[8];
			for (int i = 0; i < 8; ++i)
			{
				Box3 cb = Octant(bounds, pivot, i);
				jobs[i] = std::async(std::launch::async, [this, &st, evaluator, cb, depth]()
				{
					std::unique_ptr<Subtree> child(new Subtree);
					child->pool.nodes.reserve(st.pool.nodes.size() * 4 + 4096); // pruning appends set nodes: grow without re-copying
					child->pool = st.pool;
					child->root = Construct(*child, evaluator, cb, depth + 1);
					return child;
				});
			}
			for (int i = 0; i < 8; ++i)
			{
				std::unique_ptr<Subtree> child = jobs[i].get();
				const BuildNode& cn = child->nodes[child->root];
				if (cn.evaluator != kNoNode)
				{
					// The child's pool is a superset of this pool, so `evaluator` means the same node there.
					uniform &= child->pool.Equal(evaluator, cn.evaluator);
					penultimate &= cn.terminus;
					live++;
					st.nodes[self].children[i] = -2 - int32_t(st.spawned.size());
					st.spawned.push_back(std::move(child));
				}
				else
				{
					st.nodes[self].empty_center[i] = cn.pivot;
					st.nodes[self].empty_value[i] = cn.clip_value;
				}
			}
		}
		else
		{
			for (int i = 0; i < 8; ++i)
			{
				int32_t child = Construct(st, evaluator, Octant(bounds, pivot, i), depth + 1);
				const BuildNode& cn = st.nodes[child];
				if (cn.evaluator != kNoNode)
				{
					uniform &= st.pool.Equal(evaluator, cn.evaluator);
					penultimate &= cn.terminus;
					live++;
					st.nodes[self].children[i] = child;
				}
				else
				{
					const Vec3 child_pivot = cn.pivot;
					const float child_value = cn.clip_value;
					st.nodes[self].empty_center[i] = child_pivot;
					st.nodes[self].empty_value[i] = child_value;
				}
			}
		}

		BuildNode& n = st.nodes[self];
		if (live == 0)
		{
			n.evaluator = kNoNode; // :1753-1759
			n.terminus = true;
		}
		else if ((penultimate && uniform) || n.evaluator_leaves <= (depth > 3 ? depth : 3)) // :1769
		{
			for (int i = 0; i < 8; ++i)
			{
				n.children[i] = -1;
			}
			n.terminus = true;
		}
	}
};

// ------------------------------------------------------------------------------------------------
// Device stream generation
// ------------------------------------------------------------------------------------------------

struct StreamGen
{
	const NodePool& pool;
	std::vector<uint32_t>& out;
	const bool tree_stream;
	int depth = 0;      // values currently held (accumulator + spilled)
	int max_slots = 0;  // deepest spill slot used + 1
	uint32_t flops = 0;
	bool cullable = true;
	std::vector<uint32_t> starts; // kStreamInterp: quad offset of every instruction (Stop excluded), relative to the first
	const Mat4* inverse = nullptr; // CompiledInverseMatrix per brush node (optional), see NodePool::CompileReference
	uint32_t inverse_count = 0;

	StreamGen(const NodePool& p, std::vector<uint32_t>& o, bool tree) : pool(p), out(o), tree_stream(tree) {}

	void PushF(float f) { out.push_back(FloatBits(f)); }

	void EmitBrush(const Node& n, uint32_t node_index, uint32_t op, float threshold, uint32_t flags)
	{
		if (!tree_stream)
		{
			EmitBrushQuads(n, node_index, op, threshold, flags);
			return;
		}
		const size_t start = out.size();
		out.push_back(0);
		uint32_t xform = kXformNone;
		const bool identity = n.rotation.IsIdentity();
		const bool unit = n.scalation == 1.0f;
		const bool moved = !(n.translation == Vec3(0.0f, 0.0f, 0.0f));
		// Transform::ApplyInv: rotate(inverse(q), p - t) / s.  With q = identity and s = 1 that is p - t exactly.
		if (!identity || !unit)
		{
			xform = kXformQuat;
			Quat iq = Inverse(n.rotation);
			PushF(iq.w); PushF(iq.x); PushF(iq.y); PushF(iq.z);
			PushF(n.translation.x); PushF(n.translation.y); PushF(n.translation.z);
			PushF(n.scalation);
		}
		else if (moved)
		{
			xform = kXformOffset;
			PushF(-n.translation.x); PushF(-n.translation.y); PushF(-n.translation.z);
		}
		for (int i = 0; i < BrushParamCount(n.kind); ++i)
		{
			PushF(n.params[i]);
		}
		if (!unit)
		{
			flags |= kHdrScaleBit;
			PushF(n.scalation);
			flops += 1;
		}
		flags |= kHdrMaterialBit;
		out.push_back(n.material);
		uint32_t slot = kNoSlot;
		if (op == kOpPush)
		{
			if (depth >= 1)
			{
				slot = uint32_t(depth - 1);
				if (depth > max_slots) max_slots = depth;
			}
			depth++;
		}
		else if (op >= kOpBlendUnion && op <= kOpBlendDiff)
		{
			PushF(threshold);
		}
		flops += uint32_t(XformFlops(xform) + BrushFlops(n.kind) + OpFlops(op));
		if (n.kind == kKindEllipsoid)
		{
			cullable = false;
		}
		out[start] = MakeHeader(n.kind, xform, op, slot, flags, uint32_t(out.size() - start));
	}

	// kStreamInterp: 16-byte quads (tg_program.h).  [header p0 p1 p2] [3 quads matrix | 1 quad offset]? [scale threshold 0 0]?
	void EmitBrushQuads(const Node& n, uint32_t node_index, uint32_t op, float threshold, uint32_t flags)
	{
		const size_t start = out.size();
		starts.push_back(uint32_t(start / 4));
		out.push_back(0);
		for (int i = 0; i < 3; ++i)
		{
			PushF(i < BrushParamCount(n.kind) ? n.params[i] : 0.0f);
		}
		uint32_t xform = kXformNone;
		const bool identity = n.rotation.IsIdentity();
		const bool unit = n.scalation == 1.0f;
		const bool moved = !(n.translation == Vec3(0.0f, 0.0f, 0.0f));
		// EvaluatorTransform::Compile :409-429
		if (!identity || !unit)
		{
			xform = kXformMatrix;
			Mat4 inv = (inverse && node_index < inverse_count) ? inverse[node_index] : CompiledInverseMatrix(n);
			for (int c = 0; c < 4; ++c)
			{
				for (int r = 0; r < 3; ++r)
				{
					PushF(inv.m[c][r]);
				}
			}
		}
		else if (moved)
		{
			xform = kXformOffset;
			PushF(-n.translation.x); PushF(-n.translation.y); PushF(-n.translation.z); PushF(0.0f);
		}
		const bool blend = op >= kOpBlendUnion && op <= kOpBlendDiff;
		if (!unit)
		{
			flags |= kHdrScaleBit;
			flops += 1;
		}
		if (!unit || blend)
		{
			flags |= kHdrTailBit;
			PushF(n.scalation); PushF(blend ? threshold : 0.0f); PushF(0.0f); PushF(0.0f);
		}
		uint32_t slot = kNoSlot;
		if (op == kOpPush)
		{
			if (depth >= 1)
			{
				slot = uint32_t(depth - 1);
				if (depth > max_slots) max_slots = depth;
			}
			depth++;
		}
		flops += uint32_t(XformFlops(xform) + BrushFlops(n.kind) + OpFlops(op));
		if (n.kind == kKindEllipsoid)
		{
			cullable = false;
		}
		out[start] = MakeHeader(n.kind, xform, op, slot, flags, uint32_t((out.size() - start) / 4));
	}

	void EmitOp(uint32_t op, float param, uint32_t word, bool has_word, uint32_t flags)
	{
		const size_t start = out.size();
		if (!tree_stream) starts.push_back(uint32_t(start / 4));
		out.push_back(0);
		uint32_t slot = kNoSlot;
		if (op != kOpFlate && op != kOpStop)
		{
			slot = uint32_t(depth - 2); // the left operand / stencil child lives one below the accumulator
			depth--;
		}
		if ((op >= kOpBlendUnion && op <= kOpBlendDiff) || op == kOpFlate)
		{
			PushF(param);
		}
		if (has_word)
		{
			out.push_back(word);
		}
		flops += uint32_t(OpFlops(op));
		if (!tree_stream)
		{
			while ((out.size() - start) % 4 != 0) out.push_back(0); // [header param 0 0]
			out[start] = MakeHeader(kBrushNone, kXformNone, op, slot, flags, uint32_t((out.size() - start) / 4));
			return;
		}
		out[start] = MakeHeader(kBrushNone, kXformNone, op, slot, flags, uint32_t(out.size() - start));
	}

	void Gen(uint32_t index)
	{
		const Node& n = pool.nodes[index];
		if (IsBrush(n.kind))
		{
			EmitBrush(n, index, kOpPush, 0.0f, 0);
		}
		else if (IsSet(n.kind))
		{
			Gen(n.a);
			const Node& r = pool.nodes[n.b];
			uint32_t flags = (pool.nodes[n.a].has_paint ? kHdrLhsPaintBit : 0) | (r.has_paint ? kHdrRhsPaintBit : 0);
			const uint32_t op = n.kind - 8;
			if (IsBrush(r.kind))
			{
				EmitBrush(r, n.b, op, n.params[0], flags); // fused: accumulator = op(accumulator, brush)
			}
			else
			{
				Gen(n.b);
				EmitOp(op, n.params[0], 0, false, flags);
			}
		}
		else if (n.kind == kKindFlate)
		{
			Gen(n.a);
			EmitOp(kOpFlate, n.params[0], 0, false, 0);
		}
		else
		{
			Gen(n.a);
			if (tree_stream)
			{
				Gen(n.b);
				EmitOp(kOpStencil, 0.0f, n.material, true, n.kind == kKindStencilNeg ? kHdrStencilNegBit : 0);
			}
		}
	}

	void Finish()
	{
		out.push_back(MakeHeader(kBrushNone, kXformNone, kOpStop, kNoSlot, 0, 1));
		if (!tree_stream)
		{
			for (int i = 0; i < 3; ++i) out.push_back(0);
		}
	}
};

inline uint64_t Fnv(uint64_t hash, const void* data, size_t bytes)
{
	const uint8_t* c = static_cast<const uint8_t*>(data);
	for (size_t i = 0; i < bytes; ++i)
	{
		hash ^= c[i];
		hash *= 0x100000001B3ull;
	}
	return hash;
}

struct Flattener
{
	FlatModel& model;
	std::vector<uint32_t> ref_words;
	std::vector<uint32_t> program; // kStreamInterp program of the node being emitted
	std::vector<Mat4> inverse;     // CompiledInverseMatrix of every brush of the model, by node index
	int max_slots = 0;

	// Pre-order walk (same order as the octree hash in oracle/ref_tool.cpp) emitting one FlatNode per octree node.
	uint32_t Walk(const Subtree& st, int32_t index, const float (&lo)[3], const float (&hi)[3])
	{
		const BuildNode& bn = st.nodes[index];
		const uint32_t self = uint32_t(model.nodes.size());
		model.nodes.emplace_back();
		{
			FlatNode fn;
			fn.pivot[0] = bn.pivot.x;
			fn.pivot[1] = bn.pivot.y;
			fn.pivot[2] = bn.pivot.z;
			fn.terminus = bn.terminus ? 1u : 0u;
			for (int i = 0; i < 8; ++i) fn.children[i] = -1;
			fn.tree_offset = uint32_t(model.tree.size());
			program.clear();
			StreamGen interp(st.pool, program, false);
			interp.inverse = inverse.data();
			interp.inverse_count = uint32_t(inverse.size());
			interp.Gen(bn.evaluator);
			interp.Finish();
			fn.flags = interp.cullable ? kNodeCullable : 0u;
			const size_t count = interp.starts.size();
			fn.flags |= uint32_t(std::min<size_t>(count, (1u << 24) - 1u)) << kNodeCountShift;
			if (count >= kLongProgram)
			{
				// Long programs carry a table of their instructions' quad offsets right in front of them (padded to whole
				// quads): K0 evaluates such a program with one warp, every lane fetching its own instructions directly.
				fn.flags |= kNodeLong;
				for (size_t i = 0; i < count; ++i) model.interp.push_back(interp.starts[i]);
				while (model.interp.size() % 4 != 0) model.interp.push_back(0);
			}
			fn.interp_offset = uint32_t(model.interp.size());
			model.interp.insert(model.interp.end(), program.begin(), program.end());
			StreamGen tree(st.pool, model.tree, true);
			tree.Gen(bn.evaluator);
			tree.Finish();
			fn.flops = interp.flops;
			if (tree.max_slots > max_slots) max_slots = tree.max_slots;
			model.nodes[self] = fn;
		}
		// Reference-format words: statistics + hash only.
		ref_words.clear();
		st.pool.CompileReference(bn.evaluator, ref_words, inverse.empty() ? nullptr : inverse.data());
		ref_words.push_back(0); // OpcodeT::Stop (:1381)
		uint32_t child_mask = 0;
		for (int i = 0; i < 8; ++i)
		{
			if (bn.children[i] != -1) child_mask |= 1u << i;
		}
		FlatModelStats& s = model.stats;
		const uint32_t terminus = bn.terminus ? 1u : 0u;
		const uint64_t words = ref_words.size();
		s.nodes++;
		s.ref_words += words;
		if (terminus)
		{
			s.leaves++;
			s.ref_leaf_words += words;
		}
		if (words > s.ref_max_words) s.ref_max_words = words;
		const uint64_t stack = st.pool.nodes[bn.evaluator].stack_size;
		if (stack > s.max_stack) s.max_stack = stack;
		s.hash = Fnv(s.hash, &bn.pivot, 12);
		s.hash = Fnv(s.hash, &terminus, 4);
		s.hash = Fnv(s.hash, &child_mask, 4);
		s.hash = Fnv(s.hash, ref_words.data(), words * 4);

		if (bn.terminus)
		{
			FlatRegion region = { { lo[0], lo[1], lo[2] }, { hi[0], hi[1], hi[2] }, { 0.0f, 0.0f, 0.0f }, 0.0f, self, 0 };
			model.regions.push_back(region);
			const Vec3 extent = bn.bounds.max - bn.bounds.min;
			model.leaf_nodes.push_back(self);
			model.leaf_span.push_back(std::fmax(std::fmax(extent.x, extent.y), extent.z));
			return self;
		}
		const float pivot[3] = { bn.pivot.x, bn.pivot.y, bn.pivot.z };
		for (int i = 0; i < 8; ++i)
		{
			// octant i of SDFOctree::Descend: bit set <=> coordinate > pivot
			float clo[3], chi[3];
			for (int a = 0; a < 3; ++a)
			{
				const bool upper = (i >> a) & 1;
				clo[a] = upper ? std::fmax(lo[a], pivot[a]) : lo[a];
				chi[a] = upper ? hi[a] : std::fmin(hi[a], pivot[a]);
			}
			const int32_t c = bn.children[i];
			if (c == -1)
			{
				const Vec3 ec = bn.empty_center[i];
				FlatRegion region = { { clo[0], clo[1], clo[2] }, { chi[0], chi[1], chi[2] }, { ec.x, ec.y, ec.z }, bn.empty_value[i], self, 0 };
				model.regions.push_back(region);
				continue;
			}
			uint32_t child_index;
			if (c >= 0)
			{
				child_index = Walk(st, c, clo, chi);
			}
			else
			{
				const Subtree& sub = *st.spawned[size_t(-2 - c)];
				child_index = Walk(sub, sub.root, clo, chi);
			}
			model.nodes[self].children[i] = int32_t(child_index);
		}
		return self;
	}
};

} // namespace

bool BuildFlatModel(const Tree& tree, float target_size, int threads, FlatModel& out, std::string& error)
{
	const auto t0 = std::chrono::steady_clock::now();
	out = FlatModel();
	if (!tree.Valid())
	{
		error = "empty tree";
		return false;
	}
	// SDFOctree::Create :1609-1638
	if (!tree.HasFiniteBounds())
	{
		error = "Unable to construct SDF octree for infinite area evaluator.";
		return false;
	}
	auto degenerate = [](const Box3& b)
	{
		for (int i = 0; i < 3; ++i)
		{
			if (std::isinf(b.min[i]) || std::isinf(b.max[i]) || std::isnan(b.min[i]) || std::isnan(b.max[i]) || b.max[i] <= b.min[i]) return true;
		}
		return false;
	};
	out.bounds = tree.Bounds();
	if (degenerate(out.bounds))
	{
		error = "model bounds are degenerate";
		return false;
	}
	// AABB::BoundingCube (tangerine/aabb.cpp) then operator+(Margin = 0)
	Vec3 extent = out.bounds.max - out.bounds.min;
	float longest = std::fmax(std::fmax(extent.x, extent.y), extent.z);
	Vec3 padding = (Vec3(longest) - extent) * Vec3(0.5f);
	Box3 cube = { out.bounds.min - padding, out.bounds.max + padding };
	cube.min = cube.min - Vec3(0.0f);
	cube.max = cube.max + Vec3(0.0f);
	if (degenerate(cube))
	{
		error = "model bounding cube is degenerate";
		return false;
	}

	if (threads <= 0)
	{
		threads = int(std::thread::hardware_concurrency());
	}
	Builder builder;
	builder.target_size = target_size;
	builder.parallel_depth = threads >= 4 ? 2 : threads >= 2 ? 1 : 0; // 64 tasks: the octants are very unequal (most of a scene sits in a few of them)

	Subtree top;
	top.pool = tree.pool;
	top.root = builder.Construct(top, tree.root, cube, 1);
	if (top.nodes[top.root].evaluator == kNoNode)
	{
		error = "octree pruned the whole model away";
		return false;
	}

	const double construct_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (std::getenv("TG_TRACE_HOST")) std::fprintf(stderr, "octree build: construct %.1f ms (threads %d)\n", construct_seconds * 1e3, threads);
	out.stats.hash = 0xCBF29CE484222325ull;
	Flattener flattener{ out, {}, {}, {}, 0 };
	flattener.inverse.resize(tree.pool.nodes.size());
	for (size_t i = 0; i < tree.pool.nodes.size(); ++i)
	{
		const Node& n = tree.pool.nodes[i];
		if (IsBrush(n.kind) && (!n.rotation.IsIdentity() || n.scalation != 1.0f)) flattener.inverse[i] = CompiledInverseMatrix(n);
	}
	{
		const float lo[3] = { -INFINITY, -INFINITY, -INFINITY }, hi[3] = { INFINITY, INFINITY, INFINITY };
		flattener.Walk(top, top.root, lo, hi);
	}
	if (flattener.max_slots > kMaxStackSlots)
	{
		error = "CSG tree nests deeper than the device operand stack (" + std::to_string(kMaxStackSlots) + " slots)";
		return false;
	}
	out.has_paint = top.pool.nodes[top.nodes[top.root].evaluator].has_paint;
	{
		// The attribute pass works through the vertices grouped by octree node; with the costly programs first its
		// persistent warps finish on short batches (the tail of a slab's attribute kernels is a fixed cost per slab).
		std::vector<uint64_t> by_cost(out.nodes.size()); // (~flops, index): ascending = costliest first, ties in node order
		for (size_t i = 0; i < by_cost.size(); ++i) by_cost[i] = (uint64_t(~out.nodes[i].flops) << 32) | uint64_t(i);
		std::sort(by_cost.begin(), by_cost.end());
		out.node_rank.resize(by_cost.size());
		for (size_t r = 0; r < by_cost.size(); ++r) out.node_rank[uint32_t(by_cost[r])] = uint32_t(r);
	}

	// Unpruned model programs (VoxExport and whole-tree point queries).
	{
		out.root_interp_offset = uint32_t(out.interp.size());
		StreamGen interp(tree.pool, out.interp, false);
		interp.Gen(tree.root);
		interp.Finish();
		out.root_flops = interp.flops;
		out.root_tree_offset = uint32_t(out.tree.size());
		StreamGen tstream(tree.pool, out.tree, true);
		tstream.Gen(tree.root);
		tstream.Finish();
		if (tstream.max_slots > kMaxStackSlots)
		{
			error = "CSG tree nests deeper than the device operand stack";
			return false;
		}
	}
	for (int i = 0; i < 4 * 40; ++i) out.interp.push_back(0); // the interpreter fetches one quad and prefetches 512 B past an instruction
	SnapshotMaterials(out.material_rgb);
	out.material_rgb.push_back(1.0f); // default material (GetDefaultMaterial :34-38), addressed by kNoMaterial
	out.material_rgb.push_back(1.0f);
	out.material_rgb.push_back(1.0f);
	out.stats.interp_words = out.interp.size();
	out.stats.tree_words = out.tree.size();
	out.stats.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (std::getenv("TG_TRACE_HOST")) std::fprintf(stderr, "octree build: total %.1f ms\n", out.stats.build_seconds * 1e3);
	return true;
}

} // namespace tg
