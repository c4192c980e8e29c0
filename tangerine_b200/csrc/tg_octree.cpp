// See tg_octree.h.  Line references are to the reference's tangerine/sdf_evaluator.cpp.
#include "tg_octree.h"

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
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

// madvise(MADV_HUGEPAGE) on the 2 MB-aligned part of a freshly reserved buffer (no-op where the kernel offers none).
static void AdviseHugePages(const void* data, size_t bytes)
{
#if defined(__linux__) && defined(MADV_HUGEPAGE)
	const uintptr_t huge = uintptr_t(2) << 20;
	const uintptr_t begin = (reinterpret_cast<uintptr_t>(data) + huge - 1) & ~(huge - 1);
	const uintptr_t end = (reinterpret_cast<uintptr_t>(data) + bytes) & ~(huge - 1);
	if (end > begin) madvise(reinterpret_cast<void*>(begin), size_t(end - begin), MADV_HUGEPAGE);
#else
	(void)data;
	(void)bytes;
#endif
}

static size_t CountBuildNodes(const Subtree& st)
{
	size_t n = st.nodes.size();
	for (const auto& sub : st.spawned) n += CountBuildNodes(*sub);
	return n;
}

// A fixed set of worker threads and one queue of tasks.  A thread that waits for tasks it submitted runs queued tasks
// itself meanwhile (HelpUntil), so the recursion of the octree build can fan out at any depth without ever holding a
// thread idle or creating one per task.
class TaskPool
{
public:
	explicit TaskPool(int threads)
	{
		for (int t = 1; t < threads; ++t) workers.emplace_back([this] { Work(); }); // the caller is the first thread
	}
	~TaskPool()
	{
		{
			std::lock_guard<std::mutex> guard(lock);
			stop = true;
		}
		wake.notify_all();
		for (std::thread& t : workers) t.join();
	}
	void Submit(std::function<void()> task)
	{
		{
			std::lock_guard<std::mutex> guard(lock);
			queue.push_back(std::move(task));
		}
		wake.notify_one();
	}
	void HelpUntil(const std::atomic<int>& pending)
	{
		while (pending.load(std::memory_order_acquire) != 0)
		{
			std::function<void()> task;
			{
				std::lock_guard<std::mutex> guard(lock);
				if (!queue.empty())
				{
					task = std::move(queue.back()); // newest first: the deepest, smallest work stays with its creator
					queue.pop_back();
				}
			}
			if (task) task();
			else std::this_thread::yield();
		}
	}

private:
	void Work()
	{
		for (;;)
		{
			std::function<void()> task;
			{
				std::unique_lock<std::mutex> guard(lock);
				wake.wait(guard, [this] { return stop || !queue.empty(); });
				if (queue.empty()) return; // stop
				task = std::move(queue.front()); // oldest first: the largest subtrees spread over the workers
				queue.pop_front();
			}
			task();
		}
	}
	std::vector<std::thread> workers;
	std::mutex lock;
	std::condition_variable wake;
	std::deque<std::function<void()>> queue;
	bool stop = false;
};

struct Builder
{
	float target_size;
	TaskPool* tasks = nullptr; // null: build serially
	std::atomic<bool> failed{ false };
	// false: the live mesher's octree (no coalescing).  Its root Bounds are the union of the boxes of the nodes that were
	// live when SDFOctree::Create stopped at MaxDepth = 3 (:1669-1684, :1763-1768: a parent sums up its children's Bounds
	// when IT is populated; the nodes of depth 3 are populated later and no longer reach their parents).
	bool coalesce = true;
	std::mutex live_lock;
	Box3 live_box;
	bool live_any = false;
	// A node hands its eight octants to the task pool when its pruned tree still has this many brushes (below that a
	// task's bookkeeping costs more than the octant) and it is not too deep.
	static constexpr int kSpawnLeaves = 100;
	static constexpr int kSpawnDepth = 5;
	// ... and near the root whenever there is anything to split: models of few brushes have deep octrees too
	static constexpr int kSpawnAlwaysLeaves = 4;
	static constexpr int kSpawnAlwaysDepth = 3;

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
			if (!coalesce && evaluator != kNoNode && (depth == 3 || (depth < 3 && n.terminus)))
			{
				std::lock_guard<std::mutex> guard(live_lock);
				if (!live_any) live_box = bounds;
				else
				{
					live_box.min = Vec3(std::fmin(live_box.min.x, bounds.min.x), std::fmin(live_box.min.y, bounds.min.y), std::fmin(live_box.min.z, bounds.min.z));
					live_box.max = Vec3(std::fmax(live_box.max.x, bounds.max.x), std::fmax(live_box.max.y, bounds.max.y), std::fmax(live_box.max.z, bounds.max.z));
				}
				live_any = true;
			}
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

	// Whether the node of `bounds` -- already clipped to `evaluator` -- would keep its evaluator: a node lives when its
	// span is at most the target size, or when any of its octants clips to something that lives (:1649-1661, :1740-1757).
	bool AnyLive(NodePool& pool, uint32_t evaluator, const Box3& bounds) const
	{
		const Vec3 extent = bounds.max - bounds.min;
		const float span = std::fmax(std::fmax(extent.x, extent.y), extent.z);
		if (span <= target_size) return true;
		const Vec3 pivot = Vec3(float(span * 0.5)) + bounds.min;
		for (int i = 0; i < 8; ++i)
		{
			const Box3 cb = Octant(bounds, pivot, i);
			const Vec3 ce = cb.max - cb.min;
			const float cs = std::fmax(std::fmax(ce.x, ce.y), ce.z);
			const Vec3 cp = Vec3(float(cs * 0.5)) + cb.min;
			const float cr = float(Length(Vec3(cs)) * 0.5);
			const uint32_t child = pool.Clip(evaluator, cp, cr, nullptr);
			if (child != kNoNode && AnyLive(pool, child, cb)) return true;
		}
		return false;
	}

	// SDFOctree::Populate :1703-1783
	void Populate(Subtree& st, int32_t self, int depth)
	{
		const Box3 bounds = st.nodes[self].bounds;
		const Vec3 pivot = st.nodes[self].pivot;
		const uint32_t evaluator = st.nodes[self].evaluator;
		if (coalesce && st.nodes[self].evaluator_leaves <= (depth > 3 ? depth : 3))
		{
			// This node coalesces whatever its children turn out to be (:1769: EvaluatorLeaves <= max(Depth, 3)) -- unless
			// none of them lives, in which case it dies (:1751-1757).  Only that question is answered here: the reference
			// builds the whole subtree below such a node and then deletes it (six nodes in seven of seaside_town's).
			BuildNode& n = st.nodes[self];
			bool any = false;
			for (int i = 0; i < 8 && !any; ++i)
			{
				const Box3 cb = Octant(bounds, pivot, i);
				const Vec3 ce = cb.max - cb.min;
				const float cs = std::fmax(std::fmax(ce.x, ce.y), ce.z);
				const Vec3 cp = Vec3(float(cs * 0.5)) + cb.min;
				const float cr = float(Length(Vec3(cs)) * 0.5);
				const uint32_t child = st.pool.Clip(evaluator, cp, cr, nullptr);
				any = child != kNoNode && AnyLive(st.pool, child, cb);
			}
			if (!any) n.evaluator = kNoNode;
			n.terminus = true;
			return;
		}
		bool uniform = true;
		bool penultimate = true;
		int live = 0;

		const int leaves_here = st.nodes[self].evaluator_leaves;
		if (tasks && ((depth <= kSpawnDepth && leaves_here >= kSpawnLeaves) || (depth <= kSpawnAlwaysDepth && leaves_here >= kSpawnAlwaysLeaves)))
		{
			// Each octant prunes in a pool of its own laid over this one (NodePool::Overlay: nothing is copied).  This pool
			// does not grow while they run: its owner -- this thread -- only helps with other tasks until they are done.
			std::unique_ptr<Subtree> results[8];
			std::atomic<int> pending{ 8 };
			for (int i = 0; i < 8; ++i)
			{
				const Box3 cb = Octant(bounds, pivot, i);
				tasks->Submit([this, &st, &results, &pending, evaluator, cb, depth, i]()
				{
					try
					{
						std::unique_ptr<Subtree> child(new Subtree);
						child->pool.Overlay(st.pool);
						child->root = Construct(*child, evaluator, cb, depth + 1);
						results[i] = std::move(child);
					}
					catch (...)
					{
						failed.store(true);
					}
					pending.fetch_sub(1, std::memory_order_acq_rel);
				});
			}
			tasks->HelpUntil(pending);
			for (int i = 0; i < 8; ++i)
			{
				std::unique_ptr<Subtree> child = std::move(results[i]);
				if (!child) continue; // (failed: reported by the caller)
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
		else if (coalesce && ((penultimate && uniform) || n.evaluator_leaves <= (depth > 3 ? depth : 3))) // :1769
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
	std::vector<uint32_t> own_starts;
	std::vector<uint32_t>& starts; // kStreamInterp: quad offset of every instruction (Stop excluded), relative to the first
	const Mat4* inverse = nullptr; // CompiledInverseMatrix per brush node (optional), see NodePool::CompileReference
	uint32_t inverse_count = 0;

	StreamGen(const NodePool& p, std::vector<uint32_t>& o, bool tree) : pool(p), out(o), tree_stream(tree), starts(own_starts) {}
	// (with a caller's vector for the offsets: a generator per octree node must not allocate one each)
	StreamGen(const NodePool& p, std::vector<uint32_t>& o, bool tree, std::vector<uint32_t>& offsets) : pool(p), out(o), tree_stream(tree), starts(offsets) { starts.clear(); }

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

// Flattening in three passes.  (1) A serial pre-order walk over the octree -- the order of the reference's own walks,
// oracle/ref_tool.cpp -- lays out the node table, the evaluation regions and the list of programs to generate; it does
// no heavy work.  (2) Every node's two device programs and the statistics / hash of its program in the reference's word
// encoding are generated independently, on the task pool.  (3) A serial pass gives the programs their offsets and
// concatenates them.
// The material SDFNode::GetMaterial returns at every point, when it does not depend on the point (sdf_evaluator.cpp:
// 537-540 brush, 957-1012 set operators: a difference asks its left operand only, a union the operand that wins, an
// intersection the winner among the operands that have paint, :1130-1133 the wrappers their child).  A stencil
// (:666-679) decides by the sign of its mask: mixed.
static uint32_t UniformMaterial(const NodePool& pool, uint32_t index)
{
	const Node& n = pool.nodes[index];
	if (IsBrush(n.kind)) return n.material;
	if (n.kind == kKindFlate) return UniformMaterial(pool, n.a);
	if (!IsSet(n.kind)) return kMixedMaterial;
	const Family family = SetFamily(n.kind);
	const uint32_t l = UniformMaterial(pool, n.a);
	if (family == Family::Diff) return l;
	const uint32_t r = UniformMaterial(pool, n.b);
	if (family == Family::Union) return l == r ? l : kMixedMaterial;
	const bool l_valid = pool.nodes[n.a].has_paint, r_valid = pool.nodes[n.b].has_paint;
	if (l_valid && r_valid) return l == r ? l : kMixedMaterial;
	return l_valid ? l : r;
}

struct Flattener
{
	struct Job
	{
		const Subtree* st;
		int32_t index;
		// where the node's streams sit in its batch's buffers (kGenerateBatch consecutive nodes share a pair of buffers: one
		// allocation and one copy per batch instead of two per node)
		uint32_t interp_at = 0, interp_size = 0; // starts table (long programs) + program, whole quads
		uint32_t interp_program_at = 0;          // word offset of the program inside that
		uint32_t tree_at = 0, tree_size = 0;
		uint32_t flags = 0, flops = 0;
		int max_slots = 0;
		uint64_t ref_words = 0, stack = 0, node_hash = 0;
		uint32_t material = kMixedMaterial;
	};

	struct Batch
	{
		std::vector<uint32_t> interp, tree;
	};
	static constexpr size_t kGenerateBatch = 64;

	FlatModel& model;
	std::vector<Mat4> inverse;     // CompiledInverseMatrix of every brush of the model, by node index
	std::vector<Job> jobs;         // one per octree node, pre-order
	std::vector<Batch> batches;    // streams of jobs [b * kGenerateBatch, (b + 1) * kGenerateBatch), in job order
	int max_slots = 0;
	bool reference_stats = true;

	// Pass 1: pre-order walk emitting one FlatNode (pivot, terminus, children) per octree node and the regions.
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
			fn.interp_offset = fn.tree_offset = fn.flags = fn.flops = 0u;
			model.nodes[self] = fn;
		}
		jobs.emplace_back();
		jobs.back().st = &st;
		jobs.back().index = index;
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

	// Pass 2: one node's programs and reference-format statistics (thread-safe: touches only its job and -- in job order,
	// from the one thread that works through the batch -- its batch).
	void Generate(Job& job, Batch& batch) const
	{
		const Subtree& st = *job.st;
		const BuildNode& bn = st.nodes[job.index];
		// (scratch vectors live per thread: six allocations per node were a fifth of a large octree's flattening)
		static thread_local std::vector<uint32_t> program, offsets, ref_words;
		program.clear();
		StreamGen interp(st.pool, program, false, offsets);
		interp.inverse = inverse.data();
		interp.inverse_count = uint32_t(inverse.size());
		interp.Gen(bn.evaluator);
		interp.Finish();
		job.flags = interp.cullable ? kNodeCullable : 0u;
		const size_t count = interp.starts.size();
		job.flags |= uint32_t(std::min<size_t>(count, (1u << 24) - 1u)) << kNodeCountShift;
		job.interp_at = uint32_t(batch.interp.size());
		if (count >= kLongProgram)
		{
			// Long programs carry a table of their instructions' quad offsets right in front of them (padded to whole
			// quads): K0 evaluates such a program with a group of threads, each fetching its own instructions directly.
			job.flags |= kNodeLong;
			for (size_t i = 0; i < count; ++i) batch.interp.push_back(interp.starts[i]);
			while ((batch.interp.size() - job.interp_at) % 4 != 0) batch.interp.push_back(0);
		}
		job.interp_program_at = uint32_t(batch.interp.size()) - job.interp_at;
		batch.interp.insert(batch.interp.end(), program.begin(), program.end());
		job.interp_size = uint32_t(batch.interp.size()) - job.interp_at;
		job.tree_at = uint32_t(batch.tree.size());
		StreamGen tree(st.pool, batch.tree, true);
		tree.Gen(bn.evaluator);
		tree.Finish();
		job.tree_size = uint32_t(batch.tree.size()) - job.tree_at;
		job.flops = interp.flops;
		job.max_slots = tree.max_slots;
		job.stack = st.pool.nodes[bn.evaluator].stack_size;
		job.material = UniformMaterial(st.pool, bn.evaluator);
		if (!reference_stats) return;
		// Reference-format words: statistics + hash only.
		ref_words.clear();
		st.pool.CompileReference(bn.evaluator, ref_words, inverse.empty() ? nullptr : inverse.data());
		ref_words.push_back(0); // OpcodeT::Stop (:1381)
		uint32_t child_mask = 0;
		for (int i = 0; i < 8; ++i)
		{
			if (bn.children[i] != -1) child_mask |= 1u << i;
		}
		const uint32_t terminus = bn.terminus ? 1u : 0u;
		job.ref_words = ref_words.size();
		uint64_t h = 0xCBF29CE484222325ull;
		h = Fnv(h, &bn.pivot, 12);
		h = Fnv(h, &terminus, 4);
		h = Fnv(h, &child_mask, 4);
		h = Fnv(h, ref_words.data(), ref_words.size() * 4);
		job.node_hash = h;
	}

	// Pass 3: offsets, statistics, concatenation (the copies in batches on the task pool when there is one).  The octree
	// hash is FNV-1a over the nodes' own hashes in pre-order.
	template <class Pool> void Assemble(Pool* tasks)
	{
		model.node_material.assign(jobs.size(), kMixedMaterial);
		std::vector<size_t> interp_base(batches.size()), tree_base(batches.size());
		size_t interp_words = model.interp.size(), tree_words = model.tree.size();
		for (size_t b = 0; b < batches.size(); ++b)
		{
			interp_base[b] = interp_words;
			tree_base[b] = tree_words;
			interp_words += batches[b].interp.size();
			tree_words += batches[b].tree.size();
		}
		FlatModelStats& s = model.stats;
		for (size_t i = 0; i < jobs.size(); ++i)
		{
			const Job& job = jobs[i];
			FlatNode& fn = model.nodes[i];
			const size_t b = i / kGenerateBatch;
			fn.interp_offset = uint32_t(interp_base[b]) + job.interp_at + job.interp_program_at;
			fn.tree_offset = uint32_t(tree_base[b]) + job.tree_at;
			fn.flags = job.flags;
			fn.flops = job.flops;
			model.node_material[i] = job.material;
			if (job.max_slots > max_slots) max_slots = job.max_slots;
			s.nodes++;
			if (fn.terminus) s.leaves++;
			if (job.stack > s.max_stack) s.max_stack = job.stack;
			if (!reference_stats) continue;
			s.ref_words += job.ref_words;
			if (fn.terminus) s.ref_leaf_words += job.ref_words;
			if (job.ref_words > s.ref_max_words) s.ref_max_words = job.ref_words;
			s.hash = Fnv(s.hash, &job.node_hash, 8);
		}
		// room for what BuildFlatModel appends afterwards (the unpruned programs are about the root node's size)
		const size_t spare = jobs.empty() ? 0 : 2 * size_t(jobs[0].interp_size + jobs[0].tree_size) + 4096;
		model.interp.reserve(interp_words + spare);
		model.tree.reserve(tree_words + spare);
		// tens of MB of fresh memory: huge pages take a dozen faults where small ones take thousands (a third of this pass)
		AdviseHugePages(model.interp.data(), model.interp.capacity() * 4);
		AdviseHugePages(model.tree.data(), model.tree.capacity() * 4);
		model.interp.resize(interp_words);
		model.tree.resize(tree_words);
		auto copy = [&](size_t begin, size_t end)
		{
			for (size_t b = begin; b < end; ++b)
			{
				Batch& batch = batches[b];
				if (!batch.interp.empty()) std::memcpy(model.interp.data() + interp_base[b], batch.interp.data(), batch.interp.size() * 4);
				if (!batch.tree.empty()) std::memcpy(model.tree.data() + tree_base[b], batch.tree.data(), batch.tree.size() * 4);
				std::vector<uint32_t>().swap(batch.interp);
				std::vector<uint32_t>().swap(batch.tree);
			}
		};
		const size_t group = 8; // batches per task
		const size_t groups = (batches.size() + group - 1) / group;
		if (tasks && groups > 1)
		{
			std::atomic<int> pending{ int(groups) };
			for (size_t g = 0; g < groups; ++g)
			{
				tasks->Submit([&copy, &pending, g, group, this]()
				{
					copy(g * group, std::min(batches.size(), (g + 1) * group));
					pending.fetch_sub(1, std::memory_order_acq_rel);
				});
			}
			tasks->HelpUntil(pending);
		}
		else
		{
			copy(0, batches.size());
		}
	}
};

} // namespace

bool BuildFlatModel(const Tree& tree, float target_size, int threads, FlatModel& out, std::string& error, bool reference_stats, bool coalesce)
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
	builder.coalesce = coalesce;
	// One pool of workers per process, kept between builds: fresh threads start with cold allocator arenas and
	// unmapped stacks, which cost a first build more than the build itself.
	TaskPool* tasks = nullptr;
	if (threads > 1)
	{
		static std::mutex pools_lock;
		static std::vector<std::pair<int, std::unique_ptr<TaskPool>>> pools;
		std::lock_guard<std::mutex> guard(pools_lock);
		for (auto& p : pools)
		{
			if (p.first == threads) tasks = p.second.get();
		}
		if (!tasks)
		{
			pools.emplace_back(threads, std::unique_ptr<TaskPool>(new TaskPool(threads)));
			tasks = pools.back().second.get();
		}
		builder.tasks = tasks;
	}

	Subtree top;
	top.pool.Overlay(tree.pool); // prunes against the caller's tree without copying it
	top.root = builder.Construct(top, tree.root, cube, 1);
	if (builder.failed.load())
	{
		error = "out of memory while building the octree";
		return false;
	}
	if (top.nodes[top.root].evaluator == kNoNode)
	{
		error = "octree pruned the whole model away";
		return false;
	}

	out.live_octree = !coalesce;
	out.live_bounds = builder.live_any ? builder.live_box : cube;

	const double construct_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (std::getenv("TG_TRACE_HOST")) std::fprintf(stderr, "octree build: construct %.1f ms (threads %d)\n", construct_seconds * 1e3, threads);
	const bool trace_host = std::getenv("TG_TRACE_HOST") != nullptr;
	auto lap = [&](const char* what)
	{
		if (trace_host) std::fprintf(stderr, "octree build: %s at %.1f ms\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3);
	};
	out.stats.hash = 0xCBF29CE484222325ull;
	Flattener flattener{ out, {}, {}, {}, 0, reference_stats };
	out.stats.reference_done = reference_stats;
	flattener.inverse.resize(tree.pool.nodes.size());
	for (size_t i = 0; i < tree.pool.nodes.size(); ++i)
	{
		const Node& n = tree.pool.nodes[i];
		if (IsBrush(n.kind) && (!n.rotation.IsIdentity() || n.scalation != 1.0f)) flattener.inverse[i] = CompiledInverseMatrix(n);
	}
	{
		const float lo[3] = { -INFINITY, -INFINITY, -INFINITY }, hi[3] = { INFINITY, INFINITY, INFINITY };
		lap("inverse matrices");
		const size_t build_nodes = CountBuildNodes(top);
		out.nodes.reserve(build_nodes);
		flattener.jobs.reserve(build_nodes);
		out.regions.reserve(build_nodes * 3);
		out.leaf_nodes.reserve(build_nodes);
		out.leaf_span.reserve(build_nodes);
		flattener.Walk(top, top.root, lo, hi);
		lap("walk");
	}
	{
		// programs of all nodes, in batches on the task pool (or here, serially)
		std::vector<Flattener::Job>& jobs = flattener.jobs;
		const size_t batch = Flattener::kGenerateBatch;
		const size_t batches = (jobs.size() + batch - 1) / batch;
		flattener.batches.resize(batches);
		if (tasks)
		{
			std::atomic<int> pending{ int(batches) };
			std::atomic<bool> generate_failed{ false };
			for (size_t b = 0; b < batches; ++b)
			{
				tasks->Submit([&flattener, &jobs, &pending, &generate_failed, b, batch]()
				{
					try
					{
						for (size_t i = b * batch; i < std::min(jobs.size(), (b + 1) * batch); ++i) flattener.Generate(jobs[i], flattener.batches[b]);
					}
					catch (...)
					{
						generate_failed.store(true);
					}
					pending.fetch_sub(1, std::memory_order_acq_rel);
				});
			}
			tasks->HelpUntil(pending);
			if (generate_failed.load())
			{
				error = "out of memory while flattening the octree";
				return false;
			}
		}
		else
		{
			for (size_t i = 0; i < jobs.size(); ++i) flattener.Generate(jobs[i], flattener.batches[i / batch]);
		}
		lap("generate");
		flattener.Assemble(tasks);
		lap("assemble");
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
		// order by (~flops, index) ascending = costliest first, ties in node order: a stable radix sort of the indices by
		// ~flops, three passes of 11 bits (std::sort of the pairs was a tenth of a large octree's flattening)
		const size_t count = out.nodes.size();
		std::vector<uint32_t> order(count), other(count);
		for (size_t i = 0; i < count; ++i) order[i] = uint32_t(i);
		for (int pass = 0; pass < 3; ++pass)
		{
			const int shift = pass * 11;
			size_t buckets[2049] = { 0 };
			for (size_t i = 0; i < count; ++i) buckets[((~out.nodes[order[i]].flops >> shift) & 2047u) + 1]++;
			for (int b = 0; b < 2048; ++b) buckets[b + 1] += buckets[b];
			for (size_t i = 0; i < count; ++i) other[buckets[(~out.nodes[order[i]].flops >> shift) & 2047u]++] = order[i];
			order.swap(other);
		}
		out.node_rank.resize(count);
		for (size_t r = 0; r < count; ++r) out.node_rank[order[r]] = uint32_t(r);
	}

	lap("node ranks");
	if (trace_host)
	{
		size_t leaves = 0, uniform = 0;
		for (size_t i = 0; i < out.nodes.size(); ++i)
		{
			if (!out.nodes[i].terminus) continue;
			leaves++;
			if (out.node_material[i] != kMixedMaterial) uniform++;
		}
		std::fprintf(stderr, "octree build: %zu of %zu terminus nodes have one material\n", uniform, leaves);
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
	lap("unpruned programs");
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
