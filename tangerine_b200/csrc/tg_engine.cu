// CUDA engine for the SDF meshing hot path (sm_100a).  See DESIGN.md for the pipeline and data layout.
//
//   K0  CullRegionInit/Level/Resolve  Lipschitz culling of empty space driven by the octree's evaluation regions;
//       BrickOrder                     the active brick list, costliest first when the list is short
//   K1  MeshBricksKernel       per active 8^3 brick, one warp: box resolution (the 9^3 tile is split along the
//       (+K2 fused)            octree's pivot planes instead of descending sample by sample), one interpreter run per
//                              octree node (2 samples / lane), tile in shared memory, bit-parallel sign
//                              classification, surface-nets vertex, scan-compacted writes
//   --  PairSums/Scan/Prefix   one dual exclusive scan over the active-cell bitmap (vertex and quad numbering)
//   K3  FinalizeMeshKernel     final (k, j, i)-lexicographic vertex order, quads from the three lower neighbours
//                              (reference's active-only rule), node histogram for K4
//   K4  AttributesKernel       gradient-descent refinement + normal per vertex, vertices sorted by program cost
//       ColorsKernel           material walk -> colour bytes
//   K5  VoxelKernel / PointCloudKernel   dense centre sampling (MagicaVoxel / point-cloud export)
//
// There is no CPU fallback: every entry point fails with TG_ERR_NO_DEVICE / TG_ERR_CUDA when the device or
// the sm_100a kernel image is unavailable.
#include "tg_engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <cmath>

#include <cuda_runtime.h>

#include "tg_device.cuh"
#include "tg_bricks.cuh"

namespace tg
{

// tg_fast.cu: the brick and lattice kernels of the opt-in fast build
int LaunchMeshBricksFast(const void* mesh_params, unsigned blocks, void* stream);
int LaunchLatticeFast(const void* device_model, const void* device_grid, float* out, unsigned tiles_x, unsigned tiles_y, unsigned tile_count, unsigned long long* counters, unsigned live, void* stream);
int FastBrickBlocksPerSm();

#define TG_CUDA(call)                                                                                           \
	do                                                                                                          \
	{                                                                                                           \
		cudaError_t tg_err_ = (call);                                                                           \
		if (tg_err_ != cudaSuccess)                                                                             \
		{                                                                                                       \
			error = std::string(#call) + ": " + cudaGetErrorString(tg_err_);                                   \
			return (tg_err_ == cudaErrorNoDevice || tg_err_ == cudaErrorInsufficientDriver ||                  \
					   tg_err_ == cudaErrorNoKernelImageForDevice || tg_err_ == cudaErrorInvalidDevice)        \
				? TG_ERR_NO_DEVICE                                                                              \
				: (tg_err_ == cudaErrorMemoryAllocation ? TG_ERR_MEMORY : TG_ERR_CUDA);                         \
		}                                                                                                       \
	} while (0)

// (the brick kernel and its helpers live in tg_bricks.cuh, which is also compiled into the fast build, tg_fast.cu)

// ------------------------------------------------------------------------------------------------
// K0: empty-space culling, driven by the octree's evaluation regions (FlatRegion) instead of the grid.
//
// Every lattice sample belongs to exactly one region, and inside a region the reference evaluates ONE program.
// A work item is (region, brick at some level of the 8 / 16 / ... / 128-cell brick hierarchy): it evaluates the
// region's program once, at the centre of the part of the brick's sample box that lies in the region.  If the
// program is 1-Lipschitz (no Ellipsoid) and |d(centre)| exceeds that box's half diagonal, no sample of the region
// in this brick can change sign: the brick is flagged "empty, sign s" for this region at this level.  Otherwise the
// item splits into the child bricks the region touches, or, at the 8-cell level, flags the brick "evaluate".
// A brick is skipped when no region asked for it to be evaluated and all empty flags it inherits agree in sign.
// Work is proportional to the octree (a few 10^5 small-program evaluations), not to the grid volume, and the
// large programs of interior nodes are only run for the few coarse bricks of their empty octants.
// ------------------------------------------------------------------------------------------------

constexpr int kCullLevels = 5; // 8, 16, 32, 64, 128 cells
#ifndef TG_SEED_SPAN
#define TG_SEED_SPAN 8
#endif
constexpr uint32_t kSeedSpan = TG_SEED_SPAN; // a region is seeded at the finest level where it spans fewer bricks than this per axis
constexpr uint32_t kFlagPositive = 1u, kFlagNegative = 2u, kFlagEvaluate = 4u;
constexpr int kCostShift = 3;       // 8-cell level only: bits 3.. of a flag word accumulate the brick's work estimate
constexpr int kCostClasses = 32;    // half-octave classes of that estimate; the brick list is written longest class first

struct CullItem
{
	uint32_t region;
	uint32_t cell; // x | y << 10 | z << 20 at the item's level
};

struct RegionRange
{
	uint16_t a[3], b[3]; // inclusive lattice sample index range of the region inside the slab; a > b = empty
};

struct CullParams
{
	DeviceModel model;
	DeviceGrid grid;
	uint32_t sample_k_lo, sample_k_hi; // slab sample range in z (inclusive): owned cell layers + the halo layer below
	uint32_t cell_k_lo, cell_k_hi;     // slab cell range in z (hi exclusive)
	RegionRange* ranges;
	CullItem* lists[kCullLevels];
	uint32_t capacity[kCullLevels];
	uint32_t* counts;                  // kCullLevels item counters, then kCullLevels counters of the long-program items
	// Items whose region runs a long program (kNodeLong) are kept in a list of their own and handed to the first threads
	// of a level: one of them keeps its thread busy for ~100 us, which hides behind the level's other items only if it
	// starts with them instead of after them.
	CullItem* long_lists[kCullLevels];
	uint32_t long_capacity[kCullLevels];
	uint32_t* flags[kCullLevels];
	uint32_t dims[kCullLevels][3];     // bricks per axis at each level
	int level;
	int long_every_other_level;        // tuning: evaluate long programs at the 128 / 32 / 8-cell levels only
};

// First lattice index i in [0, last + 1] with LatticeCoord(origin, step, i) > p (last + 1 when there is none).
__device__ __forceinline__ uint32_t FirstIndexAbove(float origin, float step, uint32_t last, float p)
{
	if (!(p > -INFINITY)) return 0u; // -inf (or NaN: be conservative)
	float guess = floorf((p - origin) / step) - 2.0f;
	guess = fminf(fmaxf(guess, 0.0f), float(last));
	uint32_t i = uint32_t(guess);
	while (i > 0u && LatticeCoord(origin, step, i) > p) i--; // the estimate can only be off by rounding
	while (i <= last && !(LatticeCoord(origin, step, i) > p)) i++;
	return i;
}

// Reserves n slots per lane in a device list with one atomic per warp; every lane of the warp must call.
__device__ __forceinline__ uint32_t WarpAppend(uint32_t* counter, uint32_t n, bool& fits, uint32_t capacity)
{
	const int lane = threadIdx.x & 31;
	uint32_t incl = n;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += v;
	}
	const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	uint32_t base = 0;
	if (lane == 31 && total) base = atomicAdd(counter, total);
	base = __shfl_sync(0xFFFFFFFFu, base, 31);
	fits = base + total <= capacity;
	return base + incl - n;
}

// True when the sample box [lo, hi] lies within the ball around an empty octant's centre in which the octree build
// has shown the region's program to be positive (FlatRegion::known_value), with the same slack as the evaluated test.
__device__ __forceinline__ bool KnownEmpty(const DeviceModel& model, const FlatRegion& region, const DeviceGrid& g,
	uint32_t lo_i, uint32_t hi_i, uint32_t lo_j, uint32_t hi_j, uint32_t lo_k, uint32_t hi_k)
{
	if (!(region.known_value > 0.0f)) return false;
	if ((__ldg(&model.nodes[region.node].flags) & kNodeCullable) == 0u) return false;
	const float dx = fmaxf(fabsf(LatticeCoord(g.x, g.dx, lo_i) - region.center[0]), fabsf(LatticeCoord(g.x, g.dx, hi_i) - region.center[0]));
	const float dy = fmaxf(fabsf(LatticeCoord(g.y, g.dy, lo_j) - region.center[1]), fabsf(LatticeCoord(g.y, g.dy, hi_j) - region.center[1]));
	const float dz = fmaxf(fabsf(LatticeCoord(g.z, g.dz, lo_k) - region.center[2]), fabsf(LatticeCoord(g.z, g.dz, hi_k) - region.center[2]));
	const float reach = sqrtf(dx * dx + dy * dy + dz * dz);
	return region.known_value > reach * 1.001f + 1.0e-4f;
}

// One WARP per region: the lanes share the region's scalars and spread its seed items (up to kSeedSpan^3) between them.
__global__ void __launch_bounds__(128) CullRegionInitKernel(const CullParams p)
{
	const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31u;
	if (r >= p.model.region_count) return;
	const FlatRegion region = p.model.regions[r];
	const DeviceGrid& g = p.grid;
	const float origin[3] = { g.x, g.y, g.z }, step[3] = { g.dx, g.dy, g.dz };
	const uint32_t last[3] = { g.sx, g.sy, g.sz };
	uint32_t a[3], b[3];
	bool empty = false;
#pragma unroll
	for (int ax = 0; ax < 3; ++ax)
	{
		a[ax] = FirstIndexAbove(origin[ax], step[ax], last[ax], region.lo[ax]);                  // first sample with x > lo
		const uint32_t above = region.hi[ax] < INFINITY ? FirstIndexAbove(origin[ax], step[ax], last[ax], region.hi[ax]) : last[ax] + 1u;
		if (above == 0u) empty = true;
		b[ax] = above - 1u;                                                                        // last sample with x <= hi
	}
	a[2] = max(a[2], p.sample_k_lo);
	b[2] = empty ? 0u : min(b[2], p.sample_k_hi);
	empty = empty || a[0] > b[0] || a[1] > b[1] || a[2] > b[2];
	if (lane == 0)
	{
		RegionRange range;
#pragma unroll
		for (int ax = 0; ax < 3; ++ax)
		{
			range.a[ax] = uint16_t(empty ? 1u : a[ax]);
			range.b[ax] = uint16_t(empty ? 0u : b[ax]);
		}
		p.ranges[r] = range;
	}
	if (empty) return;
	// cells that use these samples: [a - 1, b], clipped to the grid / slab
	uint32_t clo[3], chi[3];
#pragma unroll
	for (int ax = 0; ax < 3; ++ax)
	{
		clo[ax] = a[ax] > 0u ? a[ax] - 1u : 0u;
		chi[ax] = min(b[ax], last[ax] - 1u);
	}
	clo[2] = max(clo[2], p.cell_k_lo);
	chi[2] = min(chi[2], p.cell_k_hi - 1u);
	if (clo[2] > chi[2]) return;
	int level = 0;
	for (; level < kCullLevels - 1; ++level)
	{
		const int sh = 3 + level;
		if ((chi[0] >> sh) - (clo[0] >> sh) < kSeedSpan && (chi[1] >> sh) - (clo[1] >> sh) < kSeedSpan && (chi[2] >> sh) - (clo[2] >> sh) < kSeedSpan) break;
	}
	const int sh = 3 + level;
	const uint32_t x0 = clo[0] >> sh, x1 = chi[0] >> sh, y0 = clo[1] >> sh, y1 = chi[1] >> sh, z0 = clo[2] >> sh, z1 = chi[2] >> sh;
	const uint32_t nx = x1 - x0 + 1u, ny = y1 - y0 + 1u;
	const uint32_t n = nx * ny * (z1 - z0 + 1u);
	auto cell_of = [&](uint32_t i, uint32_t& x, uint32_t& y, uint32_t& z)
	{
		x = x0 + i % nx;
		y = y0 + (i / nx) % ny;
		z = z0 + i / (nx * ny);
	};
	uint32_t flag_bits = 0u; // nonzero: nothing is queued, every brick under these cells gets these bits instead
	uint32_t base = 0;
	CullItem* list = nullptr;
	if (KnownEmpty(p.model, region, g, a[0], b[0], a[1], b[1], a[2], b[2]))
	{
		flag_bits = kFlagPositive; // the octree build already proved this whole region empty (outside)
	}
	else
	{
		const bool is_long = (__ldg(&p.model.nodes[region.node].flags) & kNodeLong) != 0u;
		list = is_long ? p.long_lists[level] : p.lists[level];
		if (lane == 0) base = atomicAdd(&p.counts[is_long ? kCullLevels + level : level], n);
		base = __shfl_sync(0xFFFFFFFFu, base, 0);
		if (base + n > (is_long ? p.long_capacity[level] : p.capacity[level])) flag_bits = kFlagEvaluate; // cannot queue
	}
	for (uint32_t i = lane; i < n; i += 32u)
	{
		uint32_t x, y, z;
		cell_of(i, x, y, z);
		if (flag_bits) atomicOr(&p.flags[level][(size_t(z) * p.dims[level][1] + y) * p.dims[level][0] + x], flag_bits);
		else
		{
			CullItem item = { r, x | (y << 10) | (z << 20) };
			list[base + i] = item;
		}
	}
}

// What both item kernels decide first: the sample box of the item's brick inside its region.
struct ItemBox
{
	uint32_t cx, cy, cz, lo_i, hi_i, lo_j, hi_j, lo_k, hi_k;
	bool live;
};

__device__ __forceinline__ ItemBox MakeItemBox(const CullParams& p, const CullItem& item, uint32_t width)
{
	const DeviceGrid& g = p.grid;
	ItemBox b;
	const RegionRange range = p.ranges[item.region];
	b.cx = item.cell & 1023u, b.cy = (item.cell >> 10) & 1023u, b.cz = (item.cell >> 20) & 1023u;
	// sample box of this brick (clipped to grid and slab) intersected with the region's sample box
	b.lo_i = max(b.cx * width, uint32_t(range.a[0])), b.hi_i = min(min((b.cx + 1u) * width, g.sx), uint32_t(range.b[0]));
	b.lo_j = max(b.cy * width, uint32_t(range.a[1])), b.hi_j = min(min((b.cy + 1u) * width, g.sy), uint32_t(range.b[1]));
	b.lo_k = max(b.cz * width, uint32_t(range.a[2])), b.hi_k = min(min((b.cz + 1u) * width, g.sz), uint32_t(range.b[2]));
	b.live = b.lo_i <= b.hi_i && b.lo_j <= b.hi_j && b.lo_k <= b.hi_k;
	return b;
}

// Centre and culling threshold of an item's sample box.
__device__ __forceinline__ void ItemProbe(const DeviceGrid& g, const ItemBox& b, float& x, float& y, float& z, float& threshold)
{
	const float lox = LatticeCoord(g.x, g.dx, b.lo_i), loy = LatticeCoord(g.y, g.dy, b.lo_j), loz = LatticeCoord(g.z, g.dz, b.lo_k);
	const float hix = LatticeCoord(g.x, g.dx, b.hi_i), hiy = LatticeCoord(g.y, g.dy, b.hi_j), hiz = LatticeCoord(g.z, g.dz, b.hi_k);
	const float ex = hix - lox, ey = hiy - loy, ez = hiz - loz;
	const float radius = 0.5f * sqrtf(ex * ex + ey * ey + ez * ez);
	threshold = radius * 1.001f + 1.0e-4f;
	x = 0.5f * (lox + hix);
	y = 0.5f * (loy + hiy);
	z = 0.5f * (loz + hiz);
}

// The child bricks (one level down) whose sample box still meets the region.
__device__ __forceinline__ uint32_t ItemChildren(const CullParams& p, const ItemBox& b, uint32_t width, uint32_t (&children)[8])
{
	const DeviceGrid& g = p.grid;
	const uint32_t half = width >> 1;
	uint32_t n = 0;
#pragma unroll
	for (int o = 0; o < 8; ++o)
	{
		const uint32_t x = b.cx * 2u + (o & 1), y = b.cy * 2u + ((o >> 1) & 1), z = b.cz * 2u + ((o >> 2) & 1);
		const bool hit = max(x * half, b.lo_i) <= min((x + 1u) * half, b.hi_i) && max(y * half, b.lo_j) <= min((y + 1u) * half, b.hi_j) &&
			max(z * half, b.lo_k) <= min((z + 1u) * half, b.hi_k) && x * half < g.sx && y * half < g.sy && z * half < p.cell_k_hi && (z + 1u) * half > p.cell_k_lo;
		if (hit) children[n++] = x | (y << 10) | (z << 20);
	}
	return n;
}

// Items of regions with short programs: one thread per item.
__global__ void __launch_bounds__(128) CullLevelKernel(const CullParams p)
{
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	const int level = p.level;
	const uint32_t count = min(p.counts[level], p.capacity[level]);
	if (blockIdx.x * blockDim.x >= count) return; // whole block beyond the list (warps stay intact below)
	const DeviceGrid& g = p.grid;
	const uint32_t width = uint32_t(kBrick) << level;
	CullItem item = { 0u, 0u };
	ItemBox b = {};
	if (index < count)
	{
		item = p.lists[level][index];
		b = MakeItemBox(p, item, width);
	}
	bool live = b.live;
	uint32_t* flag = &p.flags[level][(size_t(b.cz) * p.dims[level][1] + b.cy) * p.dims[level][0] + b.cx];
	const uint32_t node = p.model.regions[item.region].node;
	if (live && level == 0)
	{
		// work estimate of the brick, kept above the three flag bits of its word: samples of this region in the brick x
		// the cost of the region's program.  CullResolveKernel orders the brick list by it (longest first).
		const uint32_t samples = (b.hi_i - b.lo_i + 1u) * (b.hi_j - b.lo_j + 1u) * (b.hi_k - b.lo_k + 1u);
		const uint32_t flops = __ldg(&p.model.nodes[node].flops);
		atomicAdd(flag, (((samples * (min(flops, 1u << 20) + 64u)) >> 10) + 1u) << kCostShift);
	}
	if (live && KnownEmpty(p.model, p.model.regions[item.region], g, b.lo_i, b.hi_i, b.lo_j, b.hi_j, b.lo_k, b.hi_k))
	{
		atomicOr(flag, kFlagPositive);
		live = false; // decided without running the program
	}
	// An empty octant runs its parent's program: once the known ball has failed, do not pay for it at every level on
	// the way down -- split to the 8-cell bricks and evaluate there, once.
	const uint32_t node_flags = live ? __ldg(&p.model.nodes[node].flags) : 0u;
	const bool defer = level > 0 && p.model.regions[item.region].known_value > 0.0f;
	if (live && !defer && (node_flags & kNodeCullable))
	{
		float x, y, z, threshold;
		ItemProbe(g, b, x, y, z, threshold);
		const float d = EvalInterp1(p.model, __ldg(&p.model.nodes[node].interp_offset), x, y, z);
		if (fabsf(d) > threshold)
		{
			atomicOr(flag, d > 0.0f ? kFlagPositive : kFlagNegative);
			live = false; // decided
		}
	}
	if (level == 0)
	{
		if (live) atomicOr(flag, kFlagEvaluate);
		return;
	}
	uint32_t children[8];
	const uint32_t n = live ? ItemChildren(p, b, width, children) : 0u;
	bool fits;
	const uint32_t base = WarpAppend(&p.counts[level - 1], n, fits, p.capacity[level - 1]); // every lane takes part
	if (!live) return;
	if (!fits)
	{
		atomicOr(flag, kFlagEvaluate); // inherited by every brick below this one
		return;
	}
	CullItem* out = p.lists[level - 1] + base;
	for (uint32_t c = 0; c < n; ++c)
	{
		CullItem child = { item.region, children[c] };
		out[c] = child;
	}
}

// Items of regions with long programs (interior octree nodes keep hundreds of primitives: kNodeLong).  One thread
// walking such a program is a chain of dependent fetches at ~850 cycles per primitive -- 170 us for the 336 primitives
// of seaside_town's largest leaf, which used to be the whole duration of a level.  Here a GROUP of threads evaluates an
// item together (GroupEvalLong): a warp for programs of up to kLongWarpSteps instructions (pass A: persistent warps
// striding over the list), the whole block for the few larger ones (pass B: persistent blocks).
template <int GROUP>
__device__ __forceinline__ void CullLongItem(const CullParams& p, int level, const CullItem& item, LongScratch& scratch)
{
	const DeviceGrid& g = p.grid;
	const uint32_t width = uint32_t(kBrick) << level;
	const int tid = GROUP == 32 ? int(threadIdx.x & 31) : int(threadIdx.x);
	const int warp = int(threadIdx.x >> 5);
	const ItemBox b = MakeItemBox(p, item, width);
	if (!b.live) return;
	uint32_t* flag = &p.flags[level][(size_t(b.cz) * p.dims[level][1] + b.cy) * p.dims[level][0] + b.cx];
	const FlatRegion& region = p.model.regions[item.region];
	const uint32_t node = region.node;
	const uint32_t node_flags = __ldg(&p.model.nodes[node].flags);
	if (level == 0 && tid == 0)
	{
		const uint32_t samples = (b.hi_i - b.lo_i + 1u) * (b.hi_j - b.lo_j + 1u) * (b.hi_k - b.lo_k + 1u);
		const uint32_t flops = __ldg(&p.model.nodes[node].flops);
		atomicAdd(flag, (((samples * (min(flops, 1u << 20) + 64u)) >> 10) + 1u) << kCostShift);
	}
	if (KnownEmpty(p.model, region, g, b.lo_i, b.hi_i, b.lo_j, b.hi_j, b.lo_k, b.hi_k))
	{
		if (tid == 0) atomicOr(flag, kFlagPositive);
		return;
	}
	const bool defer = level > 0 && (region.known_value > 0.0f || (p.long_every_other_level && (level & 1) != 0));
	if (!defer && (node_flags & kNodeCullable))
	{
		float x, y, z, threshold;
		ItemProbe(g, b, x, y, z, threshold);
		const uint4* program = p.model.interp + (__ldg(&p.model.nodes[node].interp_offset) >> 2);
		const uint32_t count = node_flags >> kNodeCountShift;
		const float d = GROUP == 32
			? GroupEvalLong<32>(program, count, x, y, z, scratch.step + warp * kLongWarpSteps, scratch.param + warp * kLongWarpSteps, scratch.share + warp * 32,
				scratch.marks + warp * 32, scratch.result + warp, kLongWarpSteps)
			: GroupEvalLong<kLongThreads>(program, count, x, y, z, scratch.step, scratch.param, scratch.share, scratch.marks, scratch.result, kLongBlockSteps);
		if (fabsf(d) > threshold)
		{
			if (tid == 0) atomicOr(flag, d > 0.0f ? kFlagPositive : kFlagNegative);
			return;
		}
	}
	if (level == 0)
	{
		if (tid == 0) atomicOr(flag, kFlagEvaluate);
		return;
	}
	if (tid < 32) // one warp appends the children
	{
		uint32_t children[8];
		const uint32_t n = ItemChildren(p, b, width, children);
		uint32_t base = 0;
		if (tid == 0 && n) base = atomicAdd(&p.counts[kCullLevels + level - 1], n);
		base = __shfl_sync(0xFFFFFFFFu, base, 0);
		if (base + n > p.long_capacity[level - 1])
		{
			if (tid == 0) atomicOr(flag, kFlagEvaluate); // inherited by every brick below this one
		}
		else if (tid < int(n))
		{
			CullItem child = { item.region, children[0] };
#pragma unroll
			for (int c = 1; c < 8; ++c)
			{
				if (tid == c) child.cell = children[c];
			}
			p.long_lists[level - 1][base + tid] = child;
		}
	}
}

__global__ void __launch_bounds__(kLongThreads, 4) CullLongKernel(const CullParams p)
{
	static_assert(kLongWarpSteps * (kLongThreads / 32) <= kLongBlockSteps, "the warps' slices of the scratch");
	__shared__ LongScratch scratch;
	const int level = p.level;
	const uint32_t count = min(p.counts[kCullLevels + level], p.long_capacity[level]);
	// pass A: a warp per item
	const uint32_t warps = gridDim.x * (kLongThreads / 32);
	for (uint32_t index = blockIdx.x * (kLongThreads / 32) + (threadIdx.x >> 5); index < count; index += warps)
	{
		const CullItem item = p.long_lists[level][index];
		const uint32_t steps = __ldg(&p.model.nodes[p.model.regions[item.region].node].flags) >> kNodeCountShift;
		if (steps > uint32_t(kLongWarpSteps)) continue;
		CullLongItem<32>(p, level, item, scratch);
		__syncwarp();
	}
	__syncthreads();
	// pass B: the block per item, for the programs a warp would take too long over
	for (uint32_t index = blockIdx.x; index < count; index += gridDim.x)
	{
		const CullItem item = p.long_lists[level][index];
		const uint32_t steps = __ldg(&p.model.nodes[p.model.regions[item.region].node].flags) >> kNodeCountShift;
		if (steps <= uint32_t(kLongWarpSteps)) continue;
		CullLongItem<kLongThreads>(p, level, item, scratch);
		__syncthreads();
	}
}

// Self-check of the cooperative evaluation (tg_debug_check_long_programs): every long program at a few points around
// its node's pivot, by the whole block, by one warp (when it fits) and by one thread walking it; counts disagreements.
__global__ void __launch_bounds__(kLongThreads) LongEvalCheckKernel(const DeviceModel model, float reach, unsigned long long* __restrict__ out)
{
	__shared__ LongScratch scratch;
	__shared__ float reference;
	for (uint32_t node = blockIdx.x; node < model.node_count; node += gridDim.x)
	{
		const uint32_t flags = __ldg(&model.nodes[node].flags);
		if ((flags & kNodeLong) == 0u) continue;
		const uint32_t count = flags >> kNodeCountShift;
		const uint4* program = model.interp + (__ldg(&model.nodes[node].interp_offset) >> 2);
		const float4 head = __ldg(reinterpret_cast<const float4*>(&model.nodes[node]));
		for (int probe = 0; probe < 9; ++probe)
		{
			const float x = head.x + (probe == 0 ? 0.0f : ((probe & 1) ? reach : -reach) * float(probe) * 0.25f);
			const float y = head.y + (probe == 0 ? 0.0f : ((probe & 2) ? reach : -reach) * float(probe) * 0.25f);
			const float z = head.z + (probe == 0 ? 0.0f : ((probe & 4) ? reach : -reach) * float(probe) * 0.25f);
			if (threadIdx.x == 0) reference = EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), x, y, z);
			__syncthreads();
			const float by_block = GroupEvalLong<kLongThreads>(program, count, x, y, z, scratch.step, scratch.param, scratch.share, scratch.marks, scratch.result, kLongBlockSteps);
			__syncthreads();
			float by_warp = by_block;
			if (count <= uint32_t(kLongWarpSteps) && threadIdx.x < 32)
			{
				by_warp = GroupEvalLong<32>(program, count, x, y, z, scratch.step, scratch.param, scratch.share, scratch.marks, scratch.result, kLongWarpSteps);
			}
			__syncthreads();
			if (threadIdx.x == 0)
			{
				atomicAdd(&out[0], 1ull);
				const bool same_block = by_block == reference || (by_block != by_block && reference != reference);
				const bool same_warp = by_warp == reference || (by_warp != by_warp && reference != reference);
				if (!same_block) atomicAdd(&out[1], 1ull);
				if (!same_warp) atomicAdd(&out[2], 1ull);
			}
			__syncthreads();
		}
	}
}

// One thread per 8-cell brick of the slab (and of the halo row below it): combine the flags it inherits.
__global__ void __launch_bounds__(256) CullResolveKernel(const CullParams p, uint32_t bz_begin, uint32_t bz_end, int no_cull,
	uint32_t* __restrict__ out_list, unsigned long long* out_counter, uint32_t out_capacity, unsigned char* __restrict__ out_class, uint32_t* __restrict__ class_counts)
{
	const uint32_t nbx = p.dims[0][0], nby = p.dims[0][1];
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t rows = bz_end - bz_begin;
	bool active = false;
	uint32_t brick = 0, cost = 0;
	if (index < nbx * nby * rows)
	{
		const uint32_t x = index % nbx, y = (index / nbx) % nby, z = index / (nbx * nby) + bz_begin;
		uint32_t f = 0;
		if (no_cull) f = kFlagEvaluate;
		else
		{
			const uint32_t f0 = p.flags[0][(size_t(z) * p.dims[0][1] + y) * p.dims[0][0] + x];
			cost = f0 >> kCostShift;
			f = f0 & ((1u << kCostShift) - 1u);
#pragma unroll
			for (int level = 1; level < kCullLevels; ++level)
			{
				f |= p.flags[level][(size_t(z >> level) * p.dims[level][1] + (y >> level)) * p.dims[level][0] + (x >> level)];
			}
		}
		active = (f & kFlagEvaluate) != 0u || (f & 3u) == 3u || f == 0u;
		brick = x | (y << 10) | (z << 20);
	}
	const unsigned ballot = __ballot_sync(0xFFFFFFFFu, active);
	if (ballot == 0u) return;
	const int lane = threadIdx.x & 31;
	unsigned long long base = 0;
	if (lane == 0) base = atomicAdd(out_counter, (unsigned long long)__popc(ballot));
	base = __shfl_sync(0xFFFFFFFFu, base, 0);
	// half-octave class of the work estimate: 2 * log2 + the bit below the leading one
	const uint32_t lg = cost ? 31u - uint32_t(__clz(int(cost))) : 0u;
	const uint32_t cls = cost ? min(uint32_t(kCostClasses - 1), 2u * lg + (lg ? ((cost >> (lg - 1u)) & 1u) : 0u)) : 0u;
	if (active)
	{
		const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
		if (slot < out_capacity)
		{
			out_list[slot] = brick;
			out_class[slot] = (unsigned char)cls;
		}
	}
	const unsigned peers = __match_any_sync(0xFFFFFFFFu, active ? cls : 0xFFFFFFFFu);
	if (active && lane == __ffs(peers) - 1) atomicAdd(&class_counts[cls], uint32_t(__popc(peers)));
}

// Second half of the ordering: every list entry moves to its class's segment, the costliest class first, so that the
// persistent warps of the brick kernel meet the long bricks early and finish on short ones (longest-processing-time
// list scheduling: a brick keeps one warp busy for ~100 us, a slab of an 8-GPU run is only five bricks deep per warp).
__global__ void __launch_bounds__(256) BrickOrderKernel(const uint32_t* __restrict__ list, const unsigned char* __restrict__ classes, const unsigned long long* __restrict__ count_ptr,
	uint32_t capacity, const uint32_t* __restrict__ class_counts, uint32_t* __restrict__ class_cursor, uint32_t* __restrict__ ordered, uint32_t order_limit)
{
	__shared__ uint32_t class_base[kCostClasses];
	const uint32_t count = uint32_t(min(*count_ptr, (unsigned long long)capacity));
	// A long list (many bricks per persistent warp) has no tail to speak of, and then the spatial order of the list is
	// worth more -- neighbouring bricks share octree nodes and programs, and the vertex scatter that follows is more
	// local: everything goes to one class.
	const bool by_cost = count <= order_limit;
	if (threadIdx.x < kCostClasses)
	{
		uint32_t before = 0;
		for (int c = kCostClasses - 1; c > int(threadIdx.x); --c) before += class_counts[c];
		class_base[threadIdx.x] = before;
	}
	__syncthreads();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < count;
	if (!by_cost)
	{
		if (valid) ordered[i] = list[i];
		return;
	}
	const uint32_t cls = valid ? uint32_t(classes[i]) : 0xFFFFFFFFu;
	const unsigned peers = __match_any_sync(0xFFFFFFFFu, cls);
	if (!valid) return;
	const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
	uint32_t at = 0;
	if (lane == leader) at = atomicAdd(&class_cursor[cls], uint32_t(__popc(peers)));
	at = __shfl_sync(peers, at, leader);
	ordered[class_base[cls] + at + uint32_t(__popc(peers & ((1u << lane) - 1u)))] = list[i];
}

// Estimated work per brick layer: every active brick weighs in with the FLOPs of the program at its centre (evaluation
// and per-vertex attributes both scale with it) plus a constant for descent / classification.
__global__ void BrickLayerHistogramKernel(const DeviceModel model, const DeviceGrid grid, const uint32_t* __restrict__ list, uint32_t count, uint32_t* __restrict__ layers)
{
	const uint32_t index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= count) return;
	const uint32_t brick = list[index];
	const uint32_t bx = brick & 1023u, by = (brick >> 10) & 1023u, bz = (brick >> 20) & 1023u;
	const uint32_t node = Descend(model.nodes, 0, LatticeCoord(grid.x, grid.dx, bx * kBrick + kBrick / 2), LatticeCoord(grid.y, grid.dy, by * kBrick + kBrick / 2),
		LatticeCoord(grid.z, grid.dz, bz * kBrick + kBrick / 2));
	atomicAdd(&layers[bz], 64u + __ldg(&model.nodes[node].flops));
}

// ------------------------------------------------------------------------------------------------
// Device-wide exclusive scan (block sums -> scan of sums -> per-element prefix)
// ------------------------------------------------------------------------------------------------

constexpr int kScanBlock = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanBlock * kScanItems;

struct LoadPopcount
{
	const unsigned long long* words;
	__device__ uint32_t operator()(size_t i) const { return uint32_t(__popcll(words[i])); }
};
struct LoadU32
{
	const uint32_t* values;
	__device__ uint32_t operator()(size_t i) const { return values[i]; }
};

__device__ __forceinline__ uint32_t BlockExclusiveScan(uint32_t value, uint32_t* warp_sums, uint32_t& block_total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inclusive = value;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inclusive, o);
		if (lane >= o) inclusive += n;
	}
	if (lane == 31) warp_sums[warp] = inclusive;
	__syncthreads();
	if (warp == 0)
	{
		uint32_t w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += n;
		}
		warp_sums[lane] = w; // inclusive over warps
	}
	__syncthreads();
	block_total = warp_sums[(blockDim.x >> 5) - 1];
	const uint32_t warp_base = warp ? warp_sums[warp - 1] : 0u;
	__syncthreads();
	return warp_base + inclusive - value;
}

template <typename Load>
__global__ void __launch_bounds__(kScanBlock) ScanBlockSumsKernel(Load load, size_t count, uint32_t* block_sums)
{
	__shared__ uint32_t warp_sums[32];
	const size_t base = size_t(blockIdx.x) * kScanTile + size_t(threadIdx.x) * kScanItems;
	uint32_t sum = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i)
	{
		if (base + i < count) sum += load(base + i);
	}
	uint32_t total;
	BlockExclusiveScan(sum, warp_sums, total);
	if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// Single block: exclusive scan of block_sums in place, grand total to *total_out.
__global__ void __launch_bounds__(1024) ScanSumsKernel(uint32_t* block_sums, uint32_t count, unsigned long long* total_out)
{
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < count; base += blockDim.x)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < count ? block_sums[i] : 0u;
		uint32_t total;
		const uint32_t ex = BlockExclusiveScan(v, warp_sums, total);
		const uint32_t c = carry;
		if (i < count) block_sums[i] = c + ex;
		__syncthreads();
		if (threadIdx.x == 0) carry = c + total;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total_out = carry;
}

template <typename Load>
__global__ void __launch_bounds__(kScanBlock) ScanWriteKernel(Load load, size_t count, const uint32_t* block_sums, uint32_t* out)
{
	__shared__ uint32_t warp_sums[32];
	const size_t base = size_t(blockIdx.x) * kScanTile + size_t(threadIdx.x) * kScanItems;
	uint32_t v[kScanItems];
	uint32_t sum = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i)
	{
		v[i] = base + i < count ? load(base + i) : 0u;
		sum += v[i];
	}
	uint32_t total;
	uint32_t running = BlockExclusiveScan(sum, warp_sums, total) + block_sums[blockIdx.x];
#pragma unroll
	for (int i = 0; i < kScanItems; ++i)
	{
		if (base + i < count) out[base + i] = running;
		running += v[i];
	}
}

// ------------------------------------------------------------------------------------------------
// K3: vertex numbering, quad numbering and emission.
//
// The active-cell bitmap alone decides both numberings: a cell's vertex id is the number of active cells before it
// in (k, j, i) order, and -- because SecondLoopThunk emits a quad whenever the three neighbour cells of an edge are
// active, without testing the edge itself (surface_nets.cpp:1093-1096) -- the quads a cell owns are a bitwise
// function of seven bitmap words.  So one pass scans both popcounts per 64-cell word, and one kernel then writes
// positions and triangles at their final places.  No count travels to the host in between.
// ------------------------------------------------------------------------------------------------

// A device-side count that must fit its arrays.  When it does not, the export is going to be repeated with exact
// sizes (the host sees the same counters), and every kernel downstream of the overflow treats the count as zero
// instead of walking half-filled arrays.
__device__ __forceinline__ uint32_t BoundedCount(const unsigned long long* count, uint32_t capacity)
{
	const unsigned long long n = *count;
	return n > (unsigned long long)capacity ? 0u : uint32_t(n);
}

struct QuadWords
{
	unsigned long long z, y, x; // bit b set: cell b of the word owns the quad of its z / y / x edge
};

// Quads owned by the 64 cells of bitmap word `index` (row-major [layer][j][word]).  SecondLoopThunk
// (surface_nets.cpp:1016-1030, 1069, 1093-1096): cells with i, j or k = 0 emit nothing; edge z needs cells
// (i-1,j,k), (i-1,j-1,k), (i,j-1,k); edge y needs (i-1,j,k), (i-1,j,k-1), (i,j,k-1); edge x needs (i,j-1,k),
// (i,j-1,k-1), (i,j,k-1).  Layer 0 of the bitmap is either cell layer 0 or the halo layer owned by the slab below.
__device__ __forceinline__ QuadWords QuadMasks(const unsigned long long* __restrict__ bitmap, size_t index, unsigned long long self,
	uint32_t row_words, uint32_t sy)
{
	QuadWords q = { 0ull, 0ull, 0ull };
	if (self == 0ull) return q;
	const size_t row = index / row_words;
	const uint32_t wi = uint32_t(index - row * row_words);
	const uint32_t j = uint32_t(row % sy);
	const size_t layer = row / sy;
	if (layer == 0 || j == 0u) return q;
	const size_t below = index - size_t(sy) * row_words;
	const unsigned long long b = bitmap[index - row_words], c = bitmap[below], d = bitmap[below - row_words];
	unsigned long long wp = 0ull, bp = 0ull, cp = 0ull;
	if (wi != 0u)
	{
		wp = bitmap[index - 1];
		bp = bitmap[index - row_words - 1];
		cp = bitmap[below - 1];
	}
	const unsigned long long n0 = (self << 1) | (wp >> 63); // (i-1, j,   k)
	const unsigned long long n1 = (b << 1) | (bp >> 63);    // (i-1, j-1, k)
	const unsigned long long n5 = (c << 1) | (cp >> 63);    // (i-1, j,   k-1)
	const unsigned long long not_first = wi == 0u ? ~1ull : ~0ull; // i != 0
	q.z = self & n0 & n1 & b;
	q.y = self & n0 & n5 & c;
	q.x = self & b & d & c & not_first;
	return q;
}

struct LoadVertexQuadCounts
{
	const unsigned long long* bitmap;
	uint32_t row_words, sy;
	// low half: active cells of the word, high half: quads they own
	__device__ unsigned long long operator()(size_t i) const
	{
		const unsigned long long w = bitmap[i];
		if (w == 0ull) return 0ull;
		const QuadWords q = QuadMasks(bitmap, i, w, row_words, sy);
		return (unsigned long long)__popcll(w) | ((unsigned long long)(__popcll(q.z) + __popcll(q.y) + __popcll(q.x)) << 32);
	}
};

// The same count for a word KNOWN to hold cells (its flag is set) whose row position the caller already has: all seven
// bitmap loads are issued together instead of behind the test of the word itself -- one memory round trip, not two.
__device__ __forceinline__ unsigned long long CountFlaggedWord(const unsigned long long* __restrict__ bitmap, size_t index, uint32_t wi, bool inner,
	uint32_t row_words, uint32_t sy)
{
	// inner: layer != 0 && j != 0.  Outside of it (and for i - 1 at wi == 0) the loads fall back on the word itself.
	const size_t back = inner ? index - row_words : index;
	const size_t below = inner ? index - size_t(sy) * row_words : index;
	const size_t below_back = inner ? below - row_words : index;
	const size_t prev = (inner && wi != 0u) ? 1u : 0u;
	const unsigned long long self = bitmap[index], b = bitmap[back], c = bitmap[below], d = bitmap[below_back];
	unsigned long long wp = bitmap[index - prev], bp = bitmap[back - prev], cp = bitmap[below - prev];
	uint32_t quads = 0;
	if (inner)
	{
		if (wi == 0u) wp = bp = cp = 0ull;
		const unsigned long long n0 = (self << 1) | (wp >> 63); // (i-1, j,   k)
		const unsigned long long n1 = (b << 1) | (bp >> 63);    // (i-1, j-1, k)
		const unsigned long long n5 = (c << 1) | (cp >> 63);    // (i-1, j,   k-1)
		const unsigned long long not_first = wi == 0u ? ~1ull : ~0ull; // i != 0
		quads = uint32_t(__popcll(self & n0 & n1 & b) + __popcll(self & n0 & n5 & c) + __popcll(self & b & d & c & not_first));
	}
	return (unsigned long long)__popcll(self) | ((unsigned long long)quads << 32);
}

__device__ __forceinline__ unsigned long long BlockExclusiveScan64(unsigned long long value, unsigned long long* warp_sums, unsigned long long& block_total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long inclusive = value;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, inclusive, o);
		if (lane >= o) inclusive += n;
	}
	if (lane == 31) warp_sums[warp] = inclusive;
	__syncthreads();
	if (warp == 0)
	{
		unsigned long long w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0ull;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += n;
		}
		warp_sums[lane] = w;
	}
	__syncthreads();
	block_total = warp_sums[(blockDim.x >> 5) - 1];
	const unsigned long long warp_base = warp ? warp_sums[warp - 1] : 0ull;
	__syncthreads();
	return warp_base + inclusive - value;
}

// The dual (vertex, quad) scan is sparse: MeshBricksKernel raises one flag bit per bitmap word that holds active cells
// and words without a flag are neither read nor given a prefix (at seaside 1024^3 a few words in a hundred are flagged:
// 2 MB of flags stand in for 128 MB of bitmap and 128 MB of prefix).  A warp owns 32 flag words = 1024 consecutive bitmap
// words and walks the non-empty flag words one at a time, lane l on bitmap word l of the 32 -- coalesced like a dense pass.
constexpr int kPairWarpWords = 1024;
constexpr int kPairWarps = kScanBlock / 32;

// Dual scan, pass 1: (vertices, quads) of every flagged word, stashed packed (v | q << 16) in the word's prefix slot, and
// the totals of every warp's 1024 words, packed low / high.  The warp first lists its flagged words in shared memory and
// then counts them one per lane: the seven bitmap loads behind a count (QuadMasks) are independent across lanes.
__global__ void __launch_bounds__(kScanBlock) PairSumsKernel(LoadVertexQuadCounts load, const uint32_t* __restrict__ word_flags, size_t count, uint2* stash,
	unsigned long long* warp_sums)
{
	__shared__ uint16_t listed[kPairWarps][kPairWarpWords];
	const int lane = threadIdx.x & 31;
	uint16_t* list = listed[threadIdx.x >> 5];
	const size_t warp = size_t(blockIdx.x) * kPairWarps + (threadIdx.x >> 5);
	const size_t base = warp * kPairWarpWords;
	if (base >= count) return;
	uint32_t mine = base + size_t(lane) * 32 < count ? word_flags[(base >> 5) + lane] : 0u;
	const int own = __popc(mine);
	int inclusive = own;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const int n = __shfl_up_sync(0xFFFFFFFFu, inclusive, o);
		if (lane >= o) inclusive += n;
	}
	const int total = __shfl_sync(0xFFFFFFFFu, inclusive, 31);
	int at = inclusive - own;
	while (mine)
	{
		list[at++] = uint16_t(lane * 32 + __ffs(mine) - 1);
		mine &= mine - 1u;
	}
	__syncwarp();
	// row position of the warp's first word (64-bit divisions, once); the items then get by with 32-bit arithmetic
	const size_t row0 = base / load.row_words;
	const uint32_t wi0 = uint32_t(base - row0 * load.row_words), j0 = uint32_t(row0 % load.sy);
	const bool layer0 = row0 / load.sy == 0;
	uint32_t vertices = 0, quads = 0;
	for (int t = lane; t < total; t += 32)
	{
		const uint32_t along = wi0 + list[t], rows = along / load.row_words; // rows past row0
		const size_t word = base + list[t];
		const uint32_t wi = along - rows * load.row_words;
		const uint32_t j_long = j0 + rows; // j, or sy or more past it in the layers above
		const bool inner = !(layer0 && j_long < load.sy) && j_long % load.sy != 0u;
		const unsigned long long c = CountFlaggedWord(load.bitmap, word, wi, inner, load.row_words, load.sy);
		const uint32_t v = uint32_t(c), q = uint32_t(c >> 32);
		stash[word].x = v | (q << 16); // v <= 64, q <= 192
		vertices += v;
		quads += q;
	}
	vertices = __reduce_add_sync(0xFFFFFFFFu, vertices);
	quads = __reduce_add_sync(0xFFFFFFFFu, quads);
	if (lane == 0) warp_sums[warp] = (unsigned long long)vertices | ((unsigned long long)quads << 32);
}

// Pass 2 (one block): exclusive scan of the tile totals in place; grand totals to totals_out[0] (vertices) and [1] (quads).
__global__ void __launch_bounds__(1024) PairSumsScanKernel(unsigned long long* block_sums, uint32_t count, unsigned long long* totals_out)
{
	// 4096 block sums per pass (coalesced through shared memory, four consecutive entries per thread, one block scan)
	constexpr uint32_t kPer = 4;
	__shared__ unsigned long long tile[1024 * kPer];
	__shared__ unsigned long long warp_sums[32];
	unsigned long long carry = 0;
	for (uint32_t base = 0; base < count; base += 1024 * kPer)
	{
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			const uint32_t at = base + i * 1024 + threadIdx.x;
			tile[i * 1024 + threadIdx.x] = at < count ? block_sums[at] : 0ull;
		}
		__syncthreads();
		unsigned long long v[kPer], sum = 0;
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			v[i] = tile[threadIdx.x * kPer + i];
			sum += v[i];
		}
		unsigned long long total;
		unsigned long long running = carry + BlockExclusiveScan64(sum, warp_sums, total);
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			tile[threadIdx.x * kPer + i] = running;
			running += v[i];
		}
		__syncthreads();
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			const uint32_t at = base + i * 1024 + threadIdx.x;
			if (at < count) block_sums[at] = tile[i * 1024 + threadIdx.x];
		}
		carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals_out[0] = carry & 0xFFFFFFFFull;
		totals_out[1] = carry >> 32;
	}
}

// Pass 3: exclusive (vertex, quad) prefix of every flagged bitmap word, and of the first word of every cell layer
// (the layer profile and the halo count read those whether or not they hold cells).  Warps work alone: the scanned
// warp totals give each its start.
__global__ void __launch_bounds__(kScanBlock) PairPrefixKernel(const uint32_t* __restrict__ word_flags, size_t count, const unsigned long long* warp_sums, uint2* out,
	size_t layer_words)
{
	const int lane = threadIdx.x & 31;
	const size_t warp = size_t(blockIdx.x) * kPairWarps + (threadIdx.x >> 5);
	const size_t base = warp * kPairWarpWords;
	if (base >= count) return;
	const uint32_t mine = base + size_t(lane) * 32 < count ? word_flags[(base >> 5) + lane] : 0u;
	unsigned rounds = __ballot_sync(0xFFFFFFFFu, mine != 0u);
	// rounds that hold the first word of a layer run even when nothing in them is flagged
	size_t next_layer = (base + layer_words - 1) / layer_words * layer_words;
	const size_t end = base + kPairWarpWords < count ? base + kPairWarpWords : count;
	for (size_t at = next_layer; at < end; at += layer_words) rounds |= 1u << uint32_t((at - base) >> 5);
	const unsigned long long start = warp_sums[warp];
	uint32_t run_v = uint32_t(start), run_q = uint32_t(start >> 32);
	// eight flag words at a time: all their stash loads are in flight before the first in-warp scan needs one
#pragma unroll 1
	for (int batch = 0; batch < 4; ++batch)
	{
		const unsigned batch_rounds = (rounds >> (batch * 8)) & 0xFFu;
		if (batch_rounds == 0u) continue;
		uint32_t packed[8];
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			const int r = batch * 8 + i;
			const uint32_t flags = __shfl_sync(0xFFFFFFFFu, mine, r);
			packed[i] = ((flags >> lane) & 1u) ? out[base + size_t(r) * 32 + lane].x : 0u;
		}
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			if (!((batch_rounds >> i) & 1u)) continue;
			const int r = batch * 8 + i;
			const uint32_t flags = __shfl_sync(0xFFFFFFFFu, mine, r);
			const size_t round_base = base + size_t(r) * 32;
			uint32_t inclusive = packed[i]; // sums of a round stay below 2^16 in both halves (32 * 64, 32 * 192)
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inclusive, o);
				if (lane >= o) inclusive += n;
			}
			const uint32_t exclusive = inclusive - packed[i];
			const uint2 prefix = make_uint2(run_v + (exclusive & 0xFFFFu), run_q + (exclusive >> 16));
			if ((flags >> lane) & 1u) out[round_base + lane] = prefix;
			while (next_layer < round_base + 32 && next_layer < end)
			{
				if (next_layer >= round_base && size_t(lane) == next_layer - round_base) out[next_layer] = prefix;
				next_layer += layer_words;
			}
			const uint32_t total = __shfl_sync(0xFFFFFFFFu, inclusive, 31);
			run_v += total & 0xFFFFu;
			run_q += total >> 16;
		}
	}
}

struct FaceParams
{
	DeviceModel model;
	const unsigned long long* bitmap;
	const uint2* prefix;      // per bitmap word: exclusive (vertex, quad) counts
	uint32_t row_words;
	uint32_t sy;
	uint32_t k_base;
	uint32_t has_halo;        // layer 0 of the bitmap is the halo layer: its vertices come first locally and are not owned
	const float4* tmp_pos;
	const unsigned long long* tmp_key;
	const unsigned long long* counters; // kCntTmpVertices = entries in tmp_*
	uint32_t vertex_capacity;
	uint32_t quad_capacity;
	const unsigned long long* index_base; // device: added to every triangle index (vertices of the slabs below); null = 0
	float* positions;             // 3 per owned vertex
	uint32_t* triangles;          // 6 indices per quad
	uint32_t* vertex_node;        // octree node of every owned vertex (node-coherent attribute pass); may be null
	uint32_t* node_histogram;     // vertices per octree node; may be null
	// vertex numbering at the start of every brick layer of the slab (per-layer vertex profile)
	uint32_t* layer_starts;
	uint32_t profile_first_layer, profile_layers, layers_in_bitmap;
};

__device__ __forceinline__ uint32_t CellVertex(const FaceParams& p, uint32_t i, uint32_t j, uint32_t layer)
{
	const unsigned long long bit = ((unsigned long long)layer * p.sy + j) * ((unsigned long long)p.row_words * 64ull) + i;
	const unsigned long long word = p.bitmap[bit >> 6];
	return p.prefix[bit >> 6].x + uint32_t(__popcll(word & ((1ull << (bit & 63ull)) - 1ull)));
}

// One thread per extracted vertex (grid-stride over the device-side count): final position in (k, j, i) order, the
// vertex's octree node for the attribute pass, and the triangles of the quads its cell owns, at their final offsets.
__global__ void __launch_bounds__(256) FinalizeMeshKernel(const FaceParams p)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x;
	if (first < p.profile_layers)
	{
		const uint32_t k = (p.profile_first_layer + first) * kBrick; // first cell layer of that brick layer
		const uint32_t layer = k > p.k_base ? k - p.k_base : 0u;
		p.layer_starts[first] = layer < p.layers_in_bitmap ? p.prefix[size_t(layer) * p.sy * p.row_words].x : 0xFFFFFFFFu;
	}
	const uint32_t count = BoundedCount(p.counters + kCntTmpVertices, p.vertex_capacity);
	const uint32_t halo = p.has_halo ? p.prefix[size_t(p.sy) * p.row_words].x : 0u;
	const uint32_t base = p.index_base ? uint32_t(*p.index_base) : 0u;
	const unsigned long long row_bits = (unsigned long long)p.row_words * 64ull;
	for (uint32_t t = first; t < count; t += stride)
	{
		const unsigned long long key = p.tmp_key[t];
		const float4 pos = p.tmp_pos[t];
		const size_t word_index = size_t(key >> 6);
		const uint32_t bit = uint32_t(key & 63ull);
		const unsigned long long lower = (1ull << bit) - 1ull;
		const unsigned long long word = p.bitmap[word_index];
		const uint2 pre = p.prefix[word_index];
		const uint32_t id = pre.x + uint32_t(__popcll(word & lower)) - halo;
		if (id >= p.vertex_capacity) continue;
		p.positions[size_t(id) * 3 + 0] = pos.x;
		p.positions[size_t(id) * 3 + 1] = pos.y;
		p.positions[size_t(id) * 3 + 2] = pos.z;
		if (p.vertex_node)
		{
			// neighbouring vertices mostly share their node: one atomic per distinct node in the warp
			const uint32_t node = Descend(p.model.nodes, 0, pos.x, pos.y, pos.z);
			p.vertex_node[id] = node;
			const unsigned peers = __match_any_sync(__activemask(), node);
			if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.node_histogram[__ldg(&p.model.node_rank[node])], uint32_t(__popc(peers)));
		}
		const QuadWords q = QuadMasks(p.bitmap, word_index, word, p.row_words, p.sy);
		const uint32_t quads = uint32_t((q.z >> bit) & 1ull) | (uint32_t((q.y >> bit) & 1ull) << 1) | (uint32_t((q.x >> bit) & 1ull) << 2);
		if (quads == 0u) continue;
		uint32_t quad = pre.y + uint32_t(__popcll(q.z & lower) + __popcll(q.y & lower) + __popcll(q.x & lower));
		const uint32_t i = uint32_t(key % row_bits);
		const unsigned long long row = key / row_bits;
		const uint32_t j = uint32_t(row % p.sy), layer = uint32_t(row / p.sy);
		const uint32_t orient = __float_as_uint(pos.w) & 7u;
		const uint32_t shift = base - halo; // local number -> global index
		const uint32_t v = id + base;
#pragma unroll
		for (int e = 0; e < 3; ++e)
		{
			if (!(quads & (1u << e))) continue;
			uint32_t a, b, c;
			if (e == 0) // edge z: neighbours (i-1,j,k), (i-1,j-1,k), (i,j-1,k)
			{
				a = CellVertex(p, i - 1, j, layer);
				b = CellVertex(p, i - 1, j - 1, layer);
				c = CellVertex(p, i, j - 1, layer);
			}
			else if (e == 1) // edge y: (i-1,j,k), (i-1,j,k-1), (i,j,k-1)
			{
				a = CellVertex(p, i - 1, j, layer);
				b = CellVertex(p, i - 1, j, layer - 1);
				c = CellVertex(p, i, j, layer - 1);
			}
			else // edge x: (i,j-1,k), (i,j-1,k-1), (i,j,k-1)
			{
				a = CellVertex(p, i, j - 1, layer);
				b = CellVertex(p, i, j - 1, layer - 1);
				c = CellVertex(p, i, j, layer - 1);
			}
			if (quad < p.quad_capacity)
			{
				const bool forward = (orient >> e) & 1u; // :1103-1105
				const uint32_t v1 = (forward ? a : c) + shift, v2 = b + shift, v3 = (forward ? c : a) + shift;
				uint2* out = reinterpret_cast<uint2*>(p.triangles + size_t(quad) * 6u);
				out[0] = make_uint2(v, v1);
				out[1] = make_uint2(v2, v);
				out[2] = make_uint2(v2, v3);
			}
			quad++;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// MeshGenerator (tangerine/mesh_generators.cpp:20-80): welds a vertex stream -- Accumulate(vertex) for every entry in
// order.  A vertex equal to an earlier one (LessVec3's equivalence: numeric equality per component, so -0 is +0) gets
// that one's index; a new one gets the next index and is stored as (x, y, z, 1) with the bits of its FIRST occurrence.
// std::map there, here an open-addressing table keyed by the canonical bits: the first thread to claim a slot only
// fixes WHICH key the slot holds; the smallest index among the slot's vertices is the representative (atomicMin), and
// a prefix scan over "is its own representative" numbers the unique vertices in first-occurrence order.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t kWeldEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t WeldBits(float v)
{
	return v == 0.0f ? 0u : __float_as_uint(v);
}

__global__ void __launch_bounds__(256) WeldInsertKernel(const float* __restrict__ vertices, uint32_t count, uint32_t* owner, uint32_t* smallest, uint32_t mask,
	uint32_t* __restrict__ slot_of)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const uint32_t kx = WeldBits(vertices[size_t(i) * 3 + 0]), ky = WeldBits(vertices[size_t(i) * 3 + 1]), kz = WeldBits(vertices[size_t(i) * 3 + 2]);
	uint32_t h = kx * 0x9E3779B1u;
	h = (h ^ (h >> 15) ^ ky) * 0x85EBCA77u;
	h = (h ^ (h >> 13) ^ kz) * 0xC2B2AE3Du;
	h ^= h >> 16;
	for (uint32_t slot = h & mask;; slot = (slot + 1u) & mask)
	{
		uint32_t holder = atomicCAS(&owner[slot], kWeldEmpty, i);
		if (holder == kWeldEmpty) holder = i;
		if (holder == i || (WeldBits(vertices[size_t(holder) * 3 + 0]) == kx && WeldBits(vertices[size_t(holder) * 3 + 1]) == ky && WeldBits(vertices[size_t(holder) * 3 + 2]) == kz))
		{
			atomicMin(&smallest[slot], i);
			slot_of[i] = slot;
			return;
		}
	}
}

struct LoadWeldFirst
{
	const uint32_t* slot_of;
	const uint32_t* smallest;
	__device__ uint32_t operator()(size_t i) const { return smallest[slot_of[i]] == uint32_t(i) ? 1u : 0u; }
};

__global__ void __launch_bounds__(256) WeldEmitKernel(const float* __restrict__ vertices, uint32_t count, const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ smallest,
	const uint32_t* __restrict__ prefix, float4* __restrict__ out_vertices, uint32_t* __restrict__ out_indices)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const uint32_t representative = smallest[slot_of[i]];
	const uint32_t index = prefix[representative];
	out_indices[i] = index;
	if (representative == i) out_vertices[index] = make_float4(vertices[size_t(i) * 3 + 0], vertices[size_t(i) * 3 + 1], vertices[size_t(i) * 3 + 2], 1.0f);
}

// ------------------------------------------------------------------------------------------------
// K4: refinement + normal + colour
// ------------------------------------------------------------------------------------------------

struct AttributeParams
{
	DeviceModel model;
	float* positions;
	float* normals;      // may be null
	unsigned char* colors; // may be null
	const unsigned long long* count_ptr; // device-side vertex count ...
	uint32_t capacity;                   // ... clamped to the capacity of the arrays
	int refine_iterations;
	float half_x, half_y, half_z;
	float scale;
	// measured work profile for slab balancing: sum of program FLOPs of the vertices of each brick layer
	unsigned long long* layer_cost; // may be null
	float grid_z, grid_dz;
	uint32_t layer_count;
	// node-coherent execution order: thread t works on vertex perm[t], whose octree node is vertex_node[perm[t]]
	unsigned long long* cursor;  // device work cursor, zero at launch
	const uint32_t* perm;        // may be null (identity)
	const uint32_t* vertex_node; // may be null
	volatile uint32_t* progress; // page-locked host word, may be null
	uint32_t progress_base;
};

// Counting sort of the vertices by octree node, so that the lanes of a warp run the same tree program.  The
// histogram is filled by FinalizeMeshKernel (meshes) or VertexNodeKernel (point clouds).
__global__ void __launch_bounds__(256) VertexNodeKernel(const DeviceModel model, const float* __restrict__ positions, const unsigned long long* __restrict__ count_ptr,
	uint32_t capacity, uint32_t* __restrict__ vertex_node, uint32_t* __restrict__ histogram)
{
	const uint32_t count = BoundedCount(count_ptr, capacity);
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < count; v += stride)
	{
		const uint32_t node = Descend(model.nodes, 0, positions[size_t(v) * 3 + 0], positions[size_t(v) * 3 + 1], positions[size_t(v) * 3 + 2]);
		vertex_node[v] = node;
		const unsigned peers = __match_any_sync(__activemask(), node);
		if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&histogram[__ldg(&model.node_rank[node])], uint32_t(__popc(peers)));
	}
}

// One block: exclusive scan of the per-node histogram (a few tens of thousands of entries), 8192 entries per pass:
// coalesced into shared memory, eight consecutive entries per thread, one block scan, coalesced out.  (A fixed cost of
// every slab: 27 us when each pass covered 1024 entries.)
__global__ void __launch_bounds__(1024) NodeOffsetsKernel(const uint32_t* __restrict__ histogram, uint32_t node_count, uint32_t* __restrict__ node_offset)
{
	constexpr uint32_t kPer = 8;
	__shared__ uint32_t tile[1024 * kPer];
	__shared__ uint32_t warp_sums[32];
	uint32_t carry = 0;
	for (uint32_t base = 0; base < node_count; base += 1024 * kPer)
	{
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			const uint32_t at = base + i * 1024 + threadIdx.x;
			tile[i * 1024 + threadIdx.x] = at < node_count ? histogram[at] : 0u;
		}
		__syncthreads();
		uint32_t v[kPer], sum = 0;
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			v[i] = tile[threadIdx.x * kPer + i];
			sum += v[i];
		}
		uint32_t total;
		uint32_t running = carry + BlockExclusiveScan(sum, warp_sums, total);
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			tile[threadIdx.x * kPer + i] = running;
			running += v[i];
		}
		__syncthreads();
#pragma unroll
		for (uint32_t i = 0; i < kPer; ++i)
		{
			const uint32_t at = base + i * 1024 + threadIdx.x;
			if (at < node_count) node_offset[at] = tile[i * 1024 + threadIdx.x];
		}
		carry += total;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) VertexPermutationKernel(const uint32_t* __restrict__ vertex_node, const uint32_t* __restrict__ node_rank,
	const unsigned long long* __restrict__ count_ptr, uint32_t capacity, const uint32_t* __restrict__ node_offset, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm)
{
	const uint32_t count = BoundedCount(count_ptr, capacity);
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < count; v += stride)
	{
		const uint32_t node = __ldg(&node_rank[vertex_node[v]]); // the sort key: costliest programs first
		const unsigned peers = __match_any_sync(__activemask(), node);
		const int leader = __ffs(peers) - 1;
		uint32_t base = 0;
		if ((threadIdx.x & 31) == leader) base = atomicAdd(&cursor[node], uint32_t(__popc(peers)));
		base = __shfl_sync(peers, base, leader);
		perm[node_offset[node] + base + uint32_t(__popc(peers & ((1u << (threadIdx.x & 31)) - 1u)))] = v;
	}
}

__device__ __forceinline__ void ExportColor(const DeviceModel& model, uint32_t node, float x, float y, float z, unsigned char* out)
{
	// export.cpp:303-311: white unless the model is painted; `0xFF * c` truncated to a byte
	float r = 1.0f, g = 1.0f, b = 1.0f;
	if (model.has_paint)
	{
		// a node whose program can only return one material needs no walk (warps are node-coherent: no divergence here)
		uint32_t m[1] = { __ldg(&model.node_material[node]) };
		if (m[0] == kMixedMaterial) EvalTreeCentre(model.tree + __ldg(&model.nodes[node].tree_offset), x, y, z, &m[0]);
		const uint32_t id = m[0] == kNoMaterial || m[0] >= model.material_count ? model.material_count : m[0];
		r = __ldg(&model.material_rgb[id * 3 + 0]);
		g = __ldg(&model.material_rgb[id * 3 + 1]);
		b = __ldg(&model.material_rgb[id * 3 + 2]);
	}
	out[0] = (unsigned char)(255.0f * r);
	out[1] = (unsigned char)(255.0f * g);
	out[2] = (unsigned char)(255.0f * b);
}

__device__ __forceinline__ void VertexAttributes(const AttributeParams& p, uint32_t t);

// Persistent warps over the device-side vertex count: a warp takes the next 32 entries of the node-sorted order from
// a device-wide cursor (so it runs one tree program, and long and short programs balance themselves over the SMs).
__global__ void __launch_bounds__(128, 8) AttributesKernel(const AttributeParams p)
{
	const uint32_t count = BoundedCount(p.count_ptr, p.capacity);
	const int lane = threadIdx.x & 31;
	for (;;)
	{
		unsigned long long first = 0;
		if (lane == 0) first = atomicAdd(p.cursor, 32ull);
		first = __shfl_sync(0xFFFFFFFFu, first, 0);
		if (first >= count) break;
		if (p.progress && lane == 0 && (first & 2047ull) == 0ull) *p.progress = p.progress_base + uint32_t(first * 1023ull / count);
		const uint32_t t = uint32_t(first) + uint32_t(lane);
		if (t < count) VertexAttributes(p, t);
		__syncwarp();
	}
}

__device__ __forceinline__ void VertexAttributes(const AttributeParams& p, uint32_t t)
{
	const uint32_t v = p.perm ? p.perm[t] : t;
	float x = p.positions[size_t(v) * 3 + 0], y = p.positions[size_t(v) * 3 + 1], z = p.positions[size_t(v) * 3 + 2];
	const DeviceModel& model = p.model;
	const float vx = x, vy = y, vz = z;
	const int iterations = p.refine_iterations > 0 ? p.refine_iterations : 0;
	uint32_t node = 0;
	// One loop body serves the refinement steps (export.cpp:442-461 applied to a mesh vertex) and the final normal, so
	// the kernel holds a single copy of the gradient interpreter.
	for (int r = 0; r <= iterations; ++r)
	{
		const bool last = r == iterations;
		if (last && iterations > 0)
		{
			const float diagonal = sqrtf(p.half_x * p.half_x + p.half_y * p.half_y + p.half_z * p.half_z);
			x = sdf::gmin(sdf::gmax(x, vx - p.half_x), vx + p.half_x);
			y = sdf::gmin(sdf::gmax(y, vy - p.half_y), vy + p.half_y);
			z = sdf::gmin(sdf::gmax(z, vz - p.half_z), vz + p.half_z);
			const float mx = vx - x, my = vy - y, mz = vz - z;
			if (!(sqrtf(mx * mx + my * my + mz * mz) <= diagonal))
			{
				x = vx;
				y = vy;
				z = vz;
			}
		}
		node = (r == 0 && p.vertex_node) ? p.vertex_node[v] : Descend(model.nodes, 0, x, y, z);
		if (last && !p.normals) break;
		float gx, gy, gz;
		EvalGradient(model.tree + __ldg(&model.nodes[node].tree_offset), x, y, z, gx, gy, gz);
		if (last)
		{
			p.normals[size_t(v) * 3 + 0] = gx;
			p.normals[size_t(v) * 3 + 1] = gy;
			p.normals[size_t(v) * 3 + 2] = gz;
			break;
		}
		const float dist = -EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), x, y, z);
		x = x + gx * dist;
		y = y + gy * dist;
		z = z + gz * dist;
	}
	if (p.layer_cost)
	{
		// one atomic per distinct brick layer in the warp (vertices are in (k, j, i) order: usually one)
		const float cell = (z - p.grid_z) / p.grid_dz;
		const uint32_t layer = min(uint32_t(fmaxf(cell, 0.0f)) / uint32_t(kBrick), p.layer_count - 1u);
		const uint32_t flops = __ldg(&model.nodes[node].flops);
		const unsigned active = __activemask();
		const unsigned peers = __match_any_sync(active, layer);
		const uint32_t sum = __reduce_add_sync(peers, flops);
		if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.layer_cost[layer], (unsigned long long)sum);
	}
	p.positions[size_t(v) * 3 + 0] = x * p.scale;
	p.positions[size_t(v) * 3 + 1] = y * p.scale;
	p.positions[size_t(v) * 3 + 2] = z * p.scale;
}

// Colour pass, separate from the normal pass on purpose: together the two interpreters (4-tap gradient, material walk)
// exceed the instruction cache and the fused kernel spent most of its time waiting for instructions (ncu: stall
// no_instruction 7.8 per issue).  Reads the final (refined, unscaled) positions, writes colours and the scaled positions.
__global__ void __launch_bounds__(128) ColorsKernel(const AttributeParams p)
{
	const uint32_t count = BoundedCount(p.count_ptr, p.capacity);
	const int lane = threadIdx.x & 31;
	for (;;)
	{
		unsigned long long first = 0;
		if (lane == 0) first = atomicAdd(p.cursor, 32ull);
		first = __shfl_sync(0xFFFFFFFFu, first, 0);
		if (first >= count) break;
		const uint32_t t = uint32_t(first) + uint32_t(lane);
		if (t < count)
		{
			const uint32_t v = p.perm ? p.perm[t] : t;
			const float x = p.positions[size_t(v) * 3 + 0], y = p.positions[size_t(v) * 3 + 1], z = p.positions[size_t(v) * 3 + 2];
			const uint32_t node = (p.refine_iterations <= 0 && p.vertex_node) ? p.vertex_node[v] : Descend(p.model.nodes, 0, x, y, z);
			ExportColor(p.model, node, x, y, z, p.colors + size_t(v) * 3);
			if (p.scale != 1.0f)
			{
				p.positions[size_t(v) * 3 + 0] = x * p.scale;
				p.positions[size_t(v) * 3 + 1] = y * p.scale;
				p.positions[size_t(v) * 3 + 2] = z * p.scale;
			}
		}
		__syncwarp();
	}
}

// WriteSTL (export.cpp:130-140): gradient at the triangle centroid, before the vertices are scaled.
// *quad_count_ptr quads (two triangles each), clamped to quad_capacity.
__global__ void __launch_bounds__(128) FaceNormalsKernel(const DeviceModel model, const float* positions, const uint32_t* triangles,
	const unsigned long long* quad_count_ptr, uint32_t quad_capacity, float* out)
{
	const uint32_t triangle_count = BoundedCount(quad_count_ptr, quad_capacity) * 2u;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < triangle_count; t += stride)
	{
		const uint32_t a = triangles[size_t(t) * 3 + 0], b = triangles[size_t(t) * 3 + 1], c = triangles[size_t(t) * 3 + 2];
		const float cx = ((positions[size_t(a) * 3 + 0] + positions[size_t(b) * 3 + 0]) + positions[size_t(c) * 3 + 0]) / 3.0f;
		const float cy = ((positions[size_t(a) * 3 + 1] + positions[size_t(b) * 3 + 1]) + positions[size_t(c) * 3 + 1]) / 3.0f;
		const float cz = ((positions[size_t(a) * 3 + 2] + positions[size_t(b) * 3 + 2]) + positions[size_t(c) * 3 + 2]) / 3.0f;
		const uint32_t node = Descend(model.nodes, 0, cx, cy, cz);
		float gx, gy, gz;
		EvalGradient(model.tree + __ldg(&model.nodes[node].tree_offset), cx, cy, cz, gx, gy, gz);
		out[size_t(t) * 3 + 0] = gx;
		out[size_t(t) * 3 + 1] = gy;
		out[size_t(t) * 3 + 2] = gz;
	}
}

__global__ void ScalePositionsKernel(float* positions, const unsigned long long* vertex_count_ptr, uint32_t vertex_capacity, float scale)
{
	const size_t count = size_t(BoundedCount(vertex_count_ptr, vertex_capacity)) * 3;
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) positions[i] = positions[i] * scale;
}

// Delivers device-side words straight into page-locked host memory (the mailbox).  A kernel rather than a
// cudaMemcpyAsync: a copy on the compute stream would queue on the DMA engine behind the large result copies that
// run on the copy stream, and everything enqueued after it on the compute stream would wait with it.
__global__ void MailKernel(const uint32_t* __restrict__ a, uint32_t na, uint32_t* __restrict__ host_a, const uint32_t* __restrict__ b, uint32_t nb, uint32_t* __restrict__ host_b)
{
	for (uint32_t i = threadIdx.x; i < na; i += blockDim.x) host_a[i] = a[i];
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) host_b[i] = b[i];
	__threadfence_system();
}

// Slab pipeline: after a slab's triangles are out, the vertices it owned are added to the running index base.
__global__ void AdvanceIndexBaseKernel(unsigned long long* index_base, const unsigned long long* counters, uint32_t vertex_capacity)
{
	if (threadIdx.x == 0 && blockIdx.x == 0) *index_base += min(counters[kCntTmpVertices], (unsigned long long)vertex_capacity);
}

// ------------------------------------------------------------------------------------------------
// Point queries
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) EvalPointsKernel(const DeviceModel model, int mode, const float* __restrict__ points, uint32_t count, void* out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float x = points[size_t(i) * 3 + 0], y = points[size_t(i) * 3 + 1], z = points[size_t(i) * 3 + 2];
	if (mode == TG_EVAL_OCTREE)
	{
		const uint32_t node = Descend(model.nodes, 0, x, y, z);
		static_cast<float*>(out)[i] = EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), x, y, z);
	}
	else if (mode == TG_EVAL_LIVE)
	{
		const uint32_t node = DescendLive(model.nodes, x, y, z);
		static_cast<float*>(out)[i] = node == kLiveEmpty ? 100.0f : LiveClamp(EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), x, y, z));
	}
	else if (mode == TG_EVAL_INTERP)
	{
		static_cast<float*>(out)[i] = EvalInterp1(model, model.root_interp_offset, x, y, z);
	}
	else if (mode == TG_EVAL_TREE)
	{
		static_cast<float*>(out)[i] = EvalDistance1(model.tree + model.root_tree_offset, x, y, z);
	}
	else if (mode == TG_EVAL_GRADIENT)
	{
		const uint32_t node = Descend(model.nodes, 0, x, y, z);
		float gx, gy, gz;
		EvalGradient(model.tree + __ldg(&model.nodes[node].tree_offset), x, y, z, gx, gy, gz);
		float* o = static_cast<float*>(out) + size_t(i) * 3;
		o[0] = gx;
		o[1] = gy;
		o[2] = gz;
	}
	else
	{
		const uint32_t node = Descend(model.nodes, 0, x, y, z);
		ExportColor(model, node, x, y, z, static_cast<unsigned char*>(out) + size_t(i) * 3);
	}
}

// Slab planner input (tg_multi.inl): which quarter-span sub-cells of every terminus cell can hold surface.  64 threads
// per cell, one per sub-cell: the cell's own program at the sub-cell's centre against the sub-cell's half diagonal -- the
// test K0 applies per brick, at a granularity that does not depend on the export grid.  Runs once per model.
__global__ void __launch_bounds__(128) LeafProfileKernel(const DeviceModel model, const uint32_t* __restrict__ leaf_nodes, const float* __restrict__ leaf_span,
	uint32_t leaf_count, unsigned long long* __restrict__ masks)
{
	const uint32_t thread = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t leaf = thread >> 6, sub = thread & 63u;
	bool maybe = false;
	if (leaf < leaf_count)
	{
		const uint32_t node = leaf_nodes[leaf];
		const float span = leaf_span[leaf];
		const float4 head = __ldg(reinterpret_cast<const float4*>(&model.nodes[node])); // pivot = cell centre
		const float quarter = span * 0.25f;
		const float x = head.x + (float(sub & 3u) - 1.5f) * quarter;
		const float y = head.y + (float((sub >> 2) & 3u) - 1.5f) * quarter;
		const float z = head.z + (float(sub >> 4) - 1.5f) * quarter;
		const uint32_t flags = __ldg(&model.nodes[node].flags);
		if ((flags & kNodeCullable) == 0u) maybe = true; // not a distance bound: assume surface
		else
		{
			const float d = EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), x, y, z);
			maybe = !(fabsf(d) > quarter * 0.8660254f * 1.001f + 1.0e-4f);
		}
	}
	const unsigned ballot = __ballot_sync(0xFFFFFFFFu, maybe);
	if ((threadIdx.x & 31) == 0 && leaf < leaf_count)
	{
		atomicOr(&masks[leaf], (unsigned long long)ballot << (sub & 32u));
	}
}

// SDFNode::RayMarch (sdf_evaluator.cpp:336-354), one ray per thread, on the unpruned model's tree program -- what the
// Lua calls ray_cast / magnet run (lua_sdf.cpp:410-444; seaside_town.lua casts 1,764 of them while it builds itself).
__global__ void __launch_bounds__(128) RayMarchKernel(const DeviceModel model, const float* __restrict__ rays, uint32_t count, int max_iterations, float epsilon, int magnet,
	float* __restrict__ out5)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float sx = rays[size_t(i) * 6 + 0], sy = rays[size_t(i) * 6 + 1], sz = rays[size_t(i) * 6 + 2];
	float dx = rays[size_t(i) * 6 + 3], dy = rays[size_t(i) * 6 + 4], dz = rays[size_t(i) * 6 + 5];
	if (magnet) // Direction = normalize(Direction - Origin)
	{
		dx = dx - sx; dy = dy - sy; dz = dz - sz;
		const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
		dx = dx * inv; dy = dy * inv; dz = dz * inv;
	}
	{
		const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz); // glm::normalize = v * inversesqrt(dot(v, v))
		dx = dx * inv; dy = dy * inv; dz = dz * inv;
	}
	float px = sx, py = sy, pz = sz, travel = 0.0f;
	bool hit = false;
	const uint32_t* program = model.tree + model.root_tree_offset;
	for (int it = 0; it < max_iterations; ++it)
	{
		const float dist = EvalTreeCentre(program, px, py, pz);
		if (dist <= epsilon)
		{
			hit = true;
			break;
		}
		travel = travel + dist;
		// A ray that has left for infinity is a miss.  (The reference keeps evaluating the field at +-inf / NaN coordinates
		// until the iterations run out; whether a NaN survives its min / max there is a property of x86 glm, not of the model.)
		if (!(fabsf(travel) < INFINITY)) break;
		px = dx * travel + sx;
		py = dy * travel + sy;
		pz = dz * travel + sz;
	}
	if (!hit) travel = INFINITY;
	float* o = out5 + size_t(i) * 5;
	o[0] = hit ? 1.0f : 0.0f;
	o[1] = travel;
	o[2] = px;
	o[3] = py;
	o[4] = pz;
}

// ------------------------------------------------------------------------------------------------
// K5: dense centre sampling for MagicaVoxel and point-cloud export
// ------------------------------------------------------------------------------------------------

// VoxExport (magica.cpp:46-69): every lane runs the same (unpruned, tree-walk) program on 4 cells.
__global__ void __launch_bounds__(128) VoxelKernel(const DeviceModel model, float minx, float miny, float minz, float maxx, float maxy, float maxz,
	int sx, int sy, int sz, float radius, uint32_t* __restrict__ hit_words, unsigned long long total)
{
	const unsigned long long first = (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 4ull;
	float px[4], py[4], pz[4], d[4];
	const int slice = sx * sy;
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		const unsigned long long idx = first + q < total ? first + q : 0ull;
		const int i = int(idx);
		const int z = i / slice, y = (i % slice) / sx, x = i % sx;
		// Alpha = vec3(x + .5, y + .5, z + .5) / vec3(Size); Point = mix(Min, Max, Alpha) = Min * (1 - a) + Max * a
		const float ax = float(x + .5) / float(sx), ay = float(y + .5) / float(sy), az = float(z + .5) / float(sz);
		px[q] = minx * (1.0f - ax) + maxx * ax;
		py[q] = miny * (1.0f - ay) + maxy * ay;
		pz[q] = minz * (1.0f - az) + maxz * az;
	}
	EvalDistance<4>(model.tree + model.root_tree_offset, px, py, pz, d);
	unsigned bits = 0;
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		if (first + q < total && fabsf(d[q]) <= radius) bits |= 1u << q;
	}
	// 8 lanes x 4 cells = one 32-bit word
	const int lane = threadIdx.x & 31;
	unsigned word = bits << ((lane & 7) * 4);
	word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
	word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
	word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
	if ((lane & 7) == 0 && first < total) hit_words[first >> 5] = word;
}

// PointCloudExportThread generation pass (export.cpp:401-427): one cell centre per thread through the octree.
__global__ void __launch_bounds__(128) PointCloudKernel(const DeviceModel model, float startx, float starty, float startz, float stepx, float stepy, float stepz,
	int nx, int ny, int nz, float diagonal, uint32_t* __restrict__ hit_words, unsigned long long total)
{
	const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	bool hit = false;
	if (idx < total)
	{
		const int i = int(idx);
		const int slice = nx * ny;
		const float z = float(i / slice) * stepz + startz;
		const float y = float((i % slice) / nx) * stepy + starty;
		const float x = float(i % nx) * stepx + startx;
		const float cx = x + stepx / 2.0f, cy = y + stepy / 2.0f, cz = z + stepz / 2.0f;
		const uint32_t node = Descend(model.nodes, 0, cx, cy, cz);
		const float d = EvalInterp1(model, __ldg(&model.nodes[node].interp_offset), cx, cy, cz);
		hit = fabsf(d) < diagonal;
	}
	const unsigned ballot = __ballot_sync(0xFFFFFFFFu, hit);
	if ((threadIdx.x & 31) == 0 && idx < total) hit_words[idx >> 5] = ballot;
}

__global__ void __launch_bounds__(256) GatherCloudKernel(const uint32_t* __restrict__ hit_words, const uint32_t* __restrict__ prefix, unsigned long long total,
	float startx, float starty, float startz, float stepx, float stepy, float stepz, int nx, int ny, float* positions)
{
	const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= total) return;
	const uint32_t word = hit_words[idx >> 5];
	const uint32_t bit = uint32_t(idx & 31ull);
	if (!((word >> bit) & 1u)) return;
	const uint32_t id = prefix[idx >> 5] + uint32_t(__popc(word & ((1u << bit) - 1u)));
	const int i = int(idx);
	const int slice = nx * ny;
	const float z = float(i / slice) * stepz + startz;
	const float y = float((i % slice) / nx) * stepy + starty;
	const float x = float(i % nx) * stepx + startx;
	positions[size_t(id) * 3 + 0] = x + stepx / 2.0f;
	positions[size_t(id) * 3 + 1] = y + stepy / 2.0f;
	positions[size_t(id) * 3 + 2] = z + stepz / 2.0f;
}

struct LoadPopcount32
{
	const uint32_t* words;
	__device__ uint32_t operator()(size_t i) const { return uint32_t(__popc(words[i])); }
};

// Vertex numbering at the start of every brick layer of the slab (for the per-layer vertex profile).
__global__ void LayerStartsKernel(const uint32_t* __restrict__ prefix, size_t words_per_layer, uint32_t k_base, uint32_t first_layer, uint32_t layer_count,
	uint32_t layers_in_bitmap, uint32_t* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= layer_count) return;
	const uint32_t k = (first_layer + i) * kBrick; // first cell layer of brick layer first_layer + i
	const uint32_t layer = k > k_base ? k - k_base : 0u;
	out[i] = layer < layers_in_bitmap ? prefix[size_t(layer) * words_per_layer] : 0xFFFFFFFFu;
}

__global__ void RebaseIndicesKernel(uint32_t* __restrict__ indices, size_t count, uint32_t base)
{
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) indices[i] += base;
}

// ------------------------------------------------------------------------------------------------
// FP32 peak measurement and L2 flush (bench support)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) FmaChainKernel(float* out, int iterations)
{
	float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
	const float m = 0.999f, c = 1e-4f;
	for (int i = 0; i < iterations; ++i)
	{
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
			a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
			a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void FillKernel(uint32_t* data, size_t count, uint32_t value)
{
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) data[i] = value;
}

// ================================================================================================
// Host side
// ================================================================================================

static inline cudaStream_t StreamOf(Context* c) { return static_cast<cudaStream_t>(c->stream); }

Context* Context::Create(int device, std::string& error)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
	{
		error = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + "); tangerine_b200 has no CPU fallback";
		return nullptr;
	}
	if (device < 0 || device >= count)
	{
		error = "CUDA device index out of range";
		return nullptr;
	}
	if ((e = cudaSetDevice(device)) != cudaSuccess)
	{
		error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
		return nullptr;
	}
	// Fail loudly when the sm_100a image cannot run here.
	cudaFuncAttributes attr;
	if ((e = cudaFuncGetAttributes(&attr, MeshBricksKernel)) != cudaSuccess)
	{
		error = std::string("the sm_100a kernel image is not loadable on this device: ") + cudaGetErrorString(e);
		return nullptr;
	}
	Context* c = new Context();
	c->device = device;
	cudaStream_t stream;
	if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess)
	{
		error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
		delete c;
		return nullptr;
	}
	c->stream = stream;
	cudaStream_t copy_stream;
	if ((e = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
	{
		error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
		delete c;
		return nullptr;
	}
	c->copy_stream = copy_stream;
	cudaStream_t stream2;
	if ((e = cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking)) != cudaSuccess)
	{
		error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
		delete c;
		return nullptr;
	}
	c->stream2 = stream2;
	cudaEvent_t ce0;
	cudaEventCreateWithFlags(&ce0, cudaEventDisableTiming);
	c->copy_events[0] = ce0;
	for (void*& e : c->cull_events)
	{
		cudaEvent_t ce;
		cudaEventCreateWithFlags(&ce, cudaEventDisableTiming);
		e = ce;
	}
	cudaEvent_t ev0, ev1;
	cudaEventCreate(&ev0);
	cudaEventCreate(&ev1);
	c->timer_events[0] = ev0;
	c->timer_events[1] = ev1;
	cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->brick_blocks_per_sm, MeshBricksKernel, kBrickThreads, 0) != cudaSuccess || c->brick_blocks_per_sm < 1)
	{
		c->brick_blocks_per_sm = 1;
	}
	if (const char* env = std::getenv("TG_BRICK_BLOCKS")) // tuning: resident brick-kernel blocks per SM
	{
		const int n = std::atoi(env);
		if (n >= 1 && n < c->brick_blocks_per_sm) c->brick_blocks_per_sm = n;
	}
	// Keep freed scratch in the pool so steady-state exports do not hit the allocator.
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
	{
		uint64_t threshold = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
	}
	for (int i = 0; i < 4; ++i)
	{
		c->progress_done[i] = 0;
		c->progress_total[i] = 0;
	}
	void* words = nullptr;
	if (cudaMallocHost(&words, 64) == cudaSuccess)
	{
		std::memset(words, 0, 64);
		c->progress_words = static_cast<volatile uint32_t*>(words);
	}
	return c;
}

Context::~Context()
{
	cudaSetDevice(device);
	if (stream) cudaStreamSynchronize(StreamOf(this));
	for (PinnedBlock& b : pinned)
	{
		if (b.ptr) cudaFreeHost(b.ptr);
	}
	for (void* ev : timer_events)
	{
		if (ev) cudaEventDestroy(static_cast<cudaEvent_t>(ev));
	}
	if (progress_words) cudaFreeHost(const_cast<uint32_t*>(progress_words));
	if (arena) cudaFree(arena);
	if (arena2) cudaFree(arena2);
	for (PinnedBlock& b : device_blocks)
	{
		if (b.ptr) cudaFree(b.ptr);
	}
	if (stream2) cudaStreamDestroy(static_cast<cudaStream_t>(stream2));
	if (index_base) cudaFree(index_base);
	for (void* m : mailboxes) cudaFreeHost(m);
	if (copy_events[0]) cudaEventDestroy(static_cast<cudaEvent_t>(copy_events[0]));
	for (void* e : cull_events)
	{
		if (e) cudaEventDestroy(static_cast<cudaEvent_t>(e));
	}
	if (copy_stream) cudaStreamDestroy(static_cast<cudaStream_t>(copy_stream));
	if (stream) cudaStreamDestroy(StreamOf(this));
}

void* Context::AcquirePinned(size_t bytes, std::string& error)
{
	if (bytes == 0) bytes = 16;
	// best fit among free blocks
	int best = -1;
	for (size_t i = 0; i < pinned.size(); ++i)
	{
		// (a block more than twice the request stays free for what it was made for: a model's small tables must not sit in
		// the blocks of a 300 MB mesh)
		if (!pinned[i].in_use && pinned[i].bytes >= bytes && pinned[i].bytes <= 2 * bytes + (size_t(1) << 20) && (best < 0 || pinned[i].bytes < pinned[size_t(best)].bytes)) best = int(i);
	}
	if (best >= 0)
	{
		pinned[size_t(best)].in_use = true;
		return pinned[size_t(best)].ptr;
	}
	PinnedBlock b;
	const size_t rounded = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
	cudaError_t e = cudaMallocHost(&b.ptr, rounded);
	if (e != cudaSuccess)
	{
		error = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
		return nullptr;
	}
	b.bytes = rounded;
	b.in_use = true;
	pinned.push_back(b);
	return b.ptr;
}

void* Context::AcquireDevice(size_t bytes, std::string& error)
{
	if (bytes == 0) bytes = 256;
	int best = -1;
	for (size_t i = 0; i < device_blocks.size(); ++i)
	{
		const PinnedBlock& b = device_blocks[i];
		if (!b.in_use && b.bytes >= bytes && b.bytes <= 2 * bytes + (size_t(1) << 20) && (best < 0 || b.bytes < device_blocks[size_t(best)].bytes)) best = int(i);
	}
	if (best >= 0)
	{
		device_blocks[size_t(best)].in_use = true;
		return device_blocks[size_t(best)].ptr;
	}
	PinnedBlock b;
	const size_t rounded = (bytes + bytes / 8 + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
	cudaError_t e = cudaMalloc(&b.ptr, rounded);
	if (e != cudaSuccess)
	{
		// out of memory: give the cached blocks nobody uses back and try once more
		cudaGetLastError();
		for (size_t i = device_blocks.size(); i-- > 0;)
		{
			if (!device_blocks[i].in_use)
			{
				cudaFree(device_blocks[i].ptr);
				device_blocks.erase(device_blocks.begin() + long(i));
			}
		}
		e = cudaMalloc(&b.ptr, rounded);
	}
	if (e != cudaSuccess)
	{
		error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
		return nullptr;
	}
	b.bytes = rounded;
	b.in_use = true;
	device_blocks.push_back(b);
	return b.ptr;
}

void Context::ReleaseDevice(void* ptr)
{
	if (!ptr) return;
	for (PinnedBlock& b : device_blocks)
	{
		if (b.ptr == ptr) b.in_use = false;
	}
}

constexpr size_t kMailboxBytes = 16384;

void* Context::AcquireMailbox(std::string& error)
{
	if (!mailboxes.empty())
	{
		void* p = mailboxes.back();
		mailboxes.pop_back();
		return p;
	}
	void* p = nullptr;
	cudaError_t e = cudaMallocHost(&p, kMailboxBytes);
	if (e != cudaSuccess)
	{
		error = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
		return nullptr;
	}
	return p;
}

void Context::ReleaseMailbox(void* ptr)
{
	if (ptr) mailboxes.push_back(ptr);
}

void Context::ReleasePinned(void* ptr)
{
	for (PinnedBlock& b : pinned)
	{
		if (b.ptr == ptr) b.in_use = false;
	}
}

// Copies one table host -> device.  The device allocation and a page-locked staging copy of the host vector are
// made on first use and kept for the model's lifetime, so a repeated upload (tg_model_upload) is a plain async copy.
template <typename T>
static int UploadVector(Context* c, const std::vector<T>& v, void** device, void** staging, const void* shared_staging, uint64_t& bytes_total, std::string& error)
{
	const size_t payload = v.size() * sizeof(T);
	const size_t bytes = std::max<size_t>(payload, 16);
	if (!*device)
	{
		// blocks of the context's caches: the next model made on this context pays for no allocation (cudaMalloc and above all
		// cudaMallocHost of some 20 MB were a third of tg_model_create)
		*device = c->AcquireDevice(bytes, error);
		if (!*device) return TG_ERR_MEMORY;
		if (!shared_staging)
		{
			*staging = c->AcquirePinned(bytes, error);
			if (!*staging) return TG_ERR_MEMORY;
			std::memcpy(*staging, v.data(), payload);
		}
	}
	// a replica of a multi-GPU model copies from the primary's staging block (page-locked memory is visible to every device)
	TG_CUDA(cudaMemcpyAsync(*device, shared_staging ? shared_staging : *staging, payload, cudaMemcpyHostToDevice, StreamOf(c)));
	bytes_total += bytes;
	return TG_OK;
}

static void FreeModelTables(Model* m)
{
	void** tables[] = { &m->d_nodes, &m->d_interp, &m->d_tree, &m->d_materials, &m->d_regions, &m->d_node_rank, &m->d_node_material };
	if (m->context)
	{
		// the tables may still be read by kernels in flight: the blocks go back to the caches behind them
		cudaStreamSynchronize(StreamOf(m->context));
		cudaStreamSynchronize(static_cast<cudaStream_t>(m->context->stream2));
	}
	for (void** t : tables)
	{
		if (m->context) m->context->ReleaseDevice(*t);
		*t = nullptr;
	}
	for (void*& h : m->staging)
	{
		if (m->context && h) m->context->ReleasePinned(h);
		h = nullptr;
	}
	m->device_bytes = 0;
}

static int UploadModel(Model* m, std::string& error)
{
	Context* c = m->context;
	TG_CUDA(cudaSetDevice(c->device));
	m->device_bytes = 0;
	int rc;
	const Model* src = m->primary;
	if ((rc = UploadVector(c, m->flat.nodes, &m->d_nodes, &m->staging[0], src ? src->staging[0] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.interp, &m->d_interp, &m->staging[1], src ? src->staging[1] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.tree, &m->d_tree, &m->staging[2], src ? src->staging[2] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.material_rgb, &m->d_materials, &m->staging[3], src ? src->staging[3] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.regions, &m->d_regions, &m->staging[4], src ? src->staging[4] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.node_rank, &m->d_node_rank, &m->staging[5], src ? src->staging[5] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	if ((rc = UploadVector(c, m->flat.node_material, &m->d_node_material, &m->staging[6], src ? src->staging[6] : nullptr, m->device_bytes, error)) != TG_OK) return rc;
	TG_CUDA(cudaStreamSynchronize(StreamOf(c)));
	return TG_OK;
}

static DeviceModel MakeDeviceModel(const Model* m);

// Fills FlatModel::leaf_mask on the device (LeafProfileKernel): part of building a model, like the octree.
static int ProfileLeaves(Model* m, std::string& error)
{
	FlatModel& flat = m->flat;
	const uint32_t count = uint32_t(flat.leaf_nodes.size());
	flat.leaf_mask.assign(count, 0ull);
	if (count == 0) return TG_OK;
	cudaStream_t stream = StreamOf(m->context);
	uint32_t* d_nodes = nullptr;
	float* d_span = nullptr;
	unsigned long long* d_masks = nullptr;
	Context* ctx = m->context;
	d_nodes = static_cast<uint32_t*>(ctx->AcquireDevice(size_t(count) * 4, error));
	d_span = static_cast<float*>(ctx->AcquireDevice(size_t(count) * 4, error));
	d_masks = static_cast<unsigned long long*>(ctx->AcquireDevice(size_t(count) * 8, error));
	struct Return
	{
		Context* c;
		void* p[3];
		~Return() { for (void* q : p) c->ReleaseDevice(q); }
	} give_back{ ctx, { d_nodes, d_span, d_masks } };
	if (!d_nodes || !d_span || !d_masks) return TG_ERR_MEMORY;
	TG_CUDA(cudaMemcpyAsync(d_nodes, flat.leaf_nodes.data(), size_t(count) * 4, cudaMemcpyHostToDevice, stream));
	TG_CUDA(cudaMemcpyAsync(d_span, flat.leaf_span.data(), size_t(count) * 4, cudaMemcpyHostToDevice, stream));
	TG_CUDA(cudaMemsetAsync(d_masks, 0, size_t(count) * 8, stream));
	LeafProfileKernel<<<uint32_t((uint64_t(count) * 64 + 127) / 128), 128, 0, stream>>>(MakeDeviceModel(m), d_nodes, d_span, count, d_masks);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaMemcpyAsync(flat.leaf_mask.data(), d_masks, size_t(count) * 8, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	return TG_OK;
}

Model* Model::Create(Context* context, const Tree& tree, float target_size, int threads, std::string& error, bool live_octree)
{
	Model* m = new Model();
	m->context = context;
	context->live_results.fetch_add(1); // a model keeps its context alive like a mesh does (its tables sit in the context's caches)
	if (!BuildFlatModel(tree, target_size, threads, m->flat, error, false, !live_octree)) // reference statistics on demand (tg_model_get_stats)
	{
		delete m;
		return nullptr;
	}
	m->leaf_count = tree.LeafCount();
	m->source = std::make_shared<const Tree>(tree);
	m->source_target_size = target_size;
	const auto t0 = std::chrono::steady_clock::now();
	if (UploadModel(m, error) != TG_OK)
	{
		delete m;
		return nullptr;
	}
	m->upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (ProfileLeaves(m, error) != TG_OK)
	{
		delete m;
		return nullptr;
	}
	return m;
}

Model::~Model()
{
	if (!context) return;
	cudaSetDevice(context->device);
	FreeModelTables(this);
	Context* owner = context;
	context = nullptr;
	if (owner->live_results.fetch_sub(1) == 1 && owner->orphaned.load()) delete owner; // tg_context_destroy came first
}

int EngineUploadModel(Model* model, std::string& error)
{
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = UploadModel(model, error);
	model->upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return rc;
}

void EngineProgress(const std::vector<const Context*>& contexts, float out_ratios[4], int* out_stage)
{
	const Context* first = contexts[0];
	const int stage = first->stage.load();
	if (out_stage) *out_stage = stage;
	if (!out_ratios) return;
	double generation = 0.0, attributes = 0.0;
	for (const Context* c : contexts)
	{
		const double slabs = double(std::max<uint32_t>(c->progress_slabs, 1u)) * 1024.0;
		if (c->progress_words)
		{
			uint32_t seen[2];
			for (int w = 0; w < 2; ++w)
			{
				const uint32_t now = c->progress_words[w];
				uint32_t before = c->progress_seen[w].load();
				while (now > before && !c->progress_seen[w].compare_exchange_weak(before, now)) {}
				seen[w] = std::max(now, before);
			}
			generation += std::min(1.0, double(seen[0]) / slabs);
			attributes += std::min(1.0, double(seen[1]) / slabs);
		}
	}
	generation /= double(contexts.size());
	attributes /= double(contexts.size());
	// whole slabs that are home count for sure (the words only move while a kernel runs)
	const uint64_t total = first->progress_total[0].load();
	if (total) generation = std::max(generation, double(first->progress_done[0].load()) / double(total) * (stage == 0 ? 1.0 : 0.999));
	const bool refining = stage == 2;
	out_ratios[0] = float(generation);                  // GenerationProgress / GenerationEstimate
	out_ratios[1] = refining ? float(attributes) : 0.0f; // RefinementProgress / VertexCount
	out_ratios[2] = refining ? 0.0f : float(attributes); // SecondaryProgress / SecondaryCount (attribute loop)
	const uint64_t wtotal = first->progress_total[3].load();
	out_ratios[3] = wtotal ? float(double(first->progress_done[3].load()) / double(wtotal)) : 0.0f;
}

int EngineSynchronize(Context* ctx, std::string& error)
{
	TG_CUDA(cudaSetDevice(ctx->device));
	TG_CUDA(cudaStreamSynchronize(StreamOf(ctx)));
	return TG_OK;
}

static DeviceModel MakeDeviceModel(const Model* m)
{
	DeviceModel d;
	d.nodes = static_cast<const FlatNode*>(m->d_nodes);
	d.interp = static_cast<const uint4*>(m->d_interp);
	d.tree = static_cast<const uint32_t*>(m->d_tree);
	d.material_rgb = static_cast<const float*>(m->d_materials);
	d.regions = static_cast<const FlatRegion*>(m->d_regions);
	d.node_rank = static_cast<const uint32_t*>(m->d_node_rank);
	d.node_material = static_cast<const uint32_t*>(m->d_node_material);
	d.region_count = uint32_t(m->flat.regions.size());
	d.material_count = uint32_t(m->flat.material_rgb.size() / 3 - 1);
	d.root_interp_offset = m->flat.root_interp_offset;
	d.root_tree_offset = m->flat.root_tree_offset;
	d.node_count = uint32_t(m->flat.nodes.size());
	d.has_paint = m->flat.has_paint ? 1 : 0;
	return d;
}

static bool MakeDeviceGrid(const tg_grid& g, DeviceGrid& out, std::string& error)
{
	if (g.sx == 0 || g.sy == 0 || g.sz == 0 || g.sx > 8184 || g.sy > 8184 || g.sz > 8184)
	{
		error = "grid size must be 1..8184 cells per axis";
		return false;
	}
	if (!(g.dx > 0.0f) || !(g.dy > 0.0f) || !(g.dz > 0.0f))
	{
		error = "grid step must be positive";
		return false;
	}
	out.x = g.x; out.y = g.y; out.z = g.z;
	out.dx = g.dx; out.dy = g.dy; out.dz = g.dz;
	out.sx = uint32_t(g.sx); out.sy = uint32_t(g.sy); out.sz = uint32_t(g.sz);
	return true;
}

// RAII for stream-ordered scratch.
// Per-call scratch memory: bump allocation out of one grow-only arena the context keeps between calls (a context
// serves one call at a time and everything is ordered on its stream), so a steady-state export makes no allocator
// calls for scratch.  Requests the arena cannot hold fall back to the stream-ordered pool and make the arena grow
// for the next call.
struct Scratch
{
	Context* ctx;
	cudaStream_t stream;
	void*& arena;
	size_t& arena_bytes;
	std::vector<void*> blocks;
	size_t used = 0;
	size_t wanted = 0;
	// lane 0: the context's stream and arena; lane 1: the second compute lane of the slab pipeline
	explicit Scratch(Context* c, int lane = 0)
		: ctx(c), stream(static_cast<cudaStream_t>(lane ? c->stream2 : c->stream)), arena(lane ? c->arena2 : c->arena), arena_bytes(lane ? c->arena2_bytes : c->arena_bytes)
	{
	}
	~Scratch()
	{
		for (void* b : blocks) cudaFreeAsync(b, stream);
		if (wanted > arena_bytes)
		{
			if (arena) cudaFreeAsync(arena, stream);
			arena = nullptr;
			arena_bytes = 0;
			const size_t bytes = wanted + wanted / 4 + (size_t(1) << 20);
			void* p = nullptr;
			if (cudaMallocAsync(&p, bytes, stream) == cudaSuccess)
			{
				arena = p;
				arena_bytes = bytes;
			}
		}
	}
	template <typename T>
	cudaError_t Alloc(T** out, size_t count)
	{
		const size_t bytes = (std::max<size_t>(count * sizeof(T), 256) + 255) & ~size_t(255);
		wanted += bytes;
		if (used + bytes <= arena_bytes)
		{
			*out = reinterpret_cast<T*>(static_cast<char*>(arena) + used);
			used += bytes;
			return cudaSuccess;
		}
		void* p = nullptr;
		cudaError_t e = cudaMallocAsync(&p, bytes, stream);
		if (e == cudaSuccess) blocks.push_back(p);
		*out = static_cast<T*>(p);
		return e;
	}
};

struct MeshResultDevice
{
	Context* context = nullptr;
	explicit MeshResultDevice(Context* c) : context(c) { c->live_results.fetch_add(1); }
	~MeshResultDevice() { context->live_results.fetch_sub(1); }
	MeshResultDevice(const MeshResultDevice&) = delete;
	MeshResultDevice& operator=(const MeshResultDevice&) = delete;
	float* d_positions = nullptr;
	float* d_normals = nullptr;
	unsigned char* d_colors = nullptr;
	uint32_t* d_triangles = nullptr;
	float* d_face_normals = nullptr;
	std::vector<void*> pinned;
	// multi-GPU exports: the per-device results of a TG_MESH_DEVICE_ONLY export, and per-rank detail for tg_mesh_rank_info
	std::vector<MeshResultDevice*> parts;
	std::vector<tg_mesh_timings> rank_timings;
	std::vector<uint32_t> rank_cuts;
};

void EngineFreeMesh(tg_mesh* mesh)
{
	if (!mesh) return;
	MeshResultDevice* r = static_cast<MeshResultDevice*>(mesh->opaque);
	if (r)
	{
		for (MeshResultDevice* part : r->parts)
		{
			if (!part) continue;
			cudaSetDevice(part->context->device);
			part->context->ReleaseDevice(part->d_positions);
			part->context->ReleaseDevice(part->d_normals);
			part->context->ReleaseDevice(part->d_colors);
			part->context->ReleaseDevice(part->d_triangles);
			part->context->ReleaseDevice(part->d_face_normals);
			delete part;
		}
		cudaSetDevice(r->context->device);
		cudaStream_t s = StreamOf(r->context);
		(void)s;
		r->context->ReleaseDevice(r->d_positions);
		r->context->ReleaseDevice(r->d_normals);
		r->context->ReleaseDevice(r->d_colors);
		r->context->ReleaseDevice(r->d_triangles);
		r->context->ReleaseDevice(r->d_face_normals);
		for (void* p : r->pinned) r->context->ReleasePinned(p);
		Context* owner = r->context;
		delete r;
		if (owner->orphaned.load() && owner->live_results.load() == 0) delete owner; // tg_context_destroy came first
	}
	std::free(mesh->layer_vertices);
	std::free(mesh->layer_vertex_cost);
	std::memset(mesh, 0, sizeof(*mesh));
}

template <typename Load>
static int DeviceExclusiveScan(cudaStream_t stream, Scratch& scratch, Load load, size_t count, uint32_t* out, unsigned long long* total_out,
	uint64_t& launches, std::string& error)
{
	const uint32_t blocks = uint32_t((count + kScanTile - 1) / kScanTile);
	uint32_t* sums = nullptr;
	TG_CUDA(scratch.Alloc(&sums, std::max<uint32_t>(blocks, 1)));
	if (blocks == 0)
	{
		TG_CUDA(cudaMemsetAsync(total_out, 0, 8, stream));
		return TG_OK;
	}
	ScanBlockSumsKernel<Load><<<blocks, kScanBlock, 0, stream>>>(load, count, sums);
	ScanSumsKernel<<<1, 1024, 0, stream>>>(sums, blocks, total_out);
	ScanWriteKernel<Load><<<blocks, kScanBlock, 0, stream>>>(load, count, sums, out);
	launches += 3;
	TG_CUDA(cudaGetLastError());
	return TG_OK;
}

struct StageTimer
{
	cudaStream_t stream;
	std::vector<cudaEvent_t> events;
	explicit StageTimer(cudaStream_t s) : stream(s) {}
	~StageTimer()
	{
		for (cudaEvent_t e : events) cudaEventDestroy(e);
	}
	int Mark()
	{
		cudaEvent_t e;
		cudaEventCreate(&e);
		cudaEventRecord(e, stream);
		events.push_back(e);
		return int(events.size() - 1);
	}
	float Ms(int a, int b)
	{
		float ms = 0.f;
		cudaEventElapsedTime(&ms, events[size_t(a)], events[size_t(b)]);
		return ms;
	}
};

// Shared tail of mesh and point-cloud export: refinement / normals / colours.  The vertex count stays on the device
// (*count_ptr, clamped to `capacity`, which also sizes the attribute arrays), so nothing here waits for the host.
// `histogram_filled`: FinalizeMeshKernel already produced vertex_node[] and the per-node histogram.
struct AttributeScratch
{
	uint32_t* vertex_node = nullptr;
	uint32_t* histogram = nullptr; // node_count counters + node_count cursors
};

static int PrepareAttributeScratch(Model* model, cudaStream_t stream, Scratch& scratch, uint32_t capacity, AttributeScratch& as, std::string& error)
{
	const uint32_t node_count = uint32_t(model->flat.nodes.size());
	TG_CUDA(scratch.Alloc(&as.vertex_node, capacity));
	TG_CUDA(scratch.Alloc(&as.histogram, size_t(node_count) * 2));
	TG_CUDA(cudaMemsetAsync(as.histogram, 0, size_t(node_count) * 8, stream));
	return TG_OK;
}

static bool WantsAttributePass(const Model* model, const tg_mesh_options& options)
{
	const bool want_normals = (options.flags & TG_MESH_NORMALS) != 0;
	const bool want_colors = (options.flags & TG_MESH_COLORS) != 0 && model->flat.has_paint;
	const bool face_normals = (options.flags & TG_MESH_FACE_NORMALS) != 0;
	const float scale = options.scale == 0.0f ? 1.0f : options.scale;
	return want_normals || want_colors || options.refine_iterations > 0 || (scale != 1.0f && !face_normals);
}

// Resident blocks of `kernel` per SM times the SM count: the grid of a persistent kernel.
template <typename Kernel>
static uint32_t PersistentGrid(const Context* ctx, Kernel kernel, int threads)
{
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	return uint32_t(ctx->sm_count) * uint32_t(per_sm);
}

static int EnqueueAttributes(Model* model, Scratch& scratch, MeshResultDevice* result, const unsigned long long* count_ptr, unsigned long long* cursor, uint32_t capacity,
	const tg_mesh_options& options, const float half[3], AttributeScratch& as, bool histogram_filled, uint64_t& launches, std::string& error,
	unsigned long long* layer_cost = nullptr, float grid_z = 0.0f, float grid_dz = 1.0f, uint32_t layer_count = 1)
{
	Context* ctx = model->context;
	cudaStream_t stream = scratch.stream;
	const bool want_normals = (options.flags & TG_MESH_NORMALS) != 0;
	const bool want_colors = (options.flags & TG_MESH_COLORS) != 0 && model->flat.has_paint;
	const float scale = options.scale == 0.0f ? 1.0f : options.scale;
	if (capacity == 0) return TG_OK;
	if (want_normals) { void* p_ = ctx->AcquireDevice(size_t(capacity) * 12, error); if (!p_) return TG_ERR_MEMORY; result->d_normals = static_cast<decltype(result->d_normals)>(p_); }
	if (want_colors) { void* p_ = ctx->AcquireDevice(size_t(capacity) * 3, error); if (!p_) return TG_ERR_MEMORY; result->d_colors = static_cast<decltype(result->d_colors)>(p_); }
	const bool face_normals = (options.flags & TG_MESH_FACE_NORMALS) != 0;
	if (!WantsAttributePass(model, options)) return TG_OK;
	AttributeParams ap;
	ap.model = MakeDeviceModel(model);
	ap.positions = result->d_positions;
	ap.normals = result->d_normals;
	ap.colors = result->d_colors;
	ap.count_ptr = count_ptr;
	ap.capacity = capacity;
	ap.refine_iterations = options.refine_iterations;
	ap.half_x = half[0];
	ap.half_y = half[1];
	ap.half_z = half[2];
	ap.scale = face_normals ? 1.0f : scale; // STL scales after the centroid normals (export.cpp:142-145)
	ap.layer_cost = layer_cost;
	ap.grid_z = grid_z;
	ap.grid_dz = grid_dz;
	ap.layer_count = layer_count;
	ap.progress = ctx->progress_words ? ctx->progress_words + 1 : nullptr;
	ap.progress_base = ctx->progress_base;
	// node-coherent order: counting sort of the vertices by octree node
	const uint32_t node_count = uint32_t(model->flat.nodes.size());
	static const uint32_t wide_node = PersistentGrid(ctx, VertexNodeKernel, 256), wide_perm = PersistentGrid(ctx, VertexPermutationKernel, 256);
	static const uint32_t wide_attr = PersistentGrid(ctx, AttributesKernel, 128);
	uint32_t *node_offset = nullptr, *perm = nullptr;
	TG_CUDA(scratch.Alloc(&perm, capacity));
	TG_CUDA(scratch.Alloc(&node_offset, node_count));
	if (!histogram_filled)
	{
		VertexNodeKernel<<<wide_node, 256, 0, stream>>>(ap.model, result->d_positions, count_ptr, capacity, as.vertex_node, as.histogram);
		launches++;
	}
	NodeOffsetsKernel<<<1, 1024, 0, stream>>>(as.histogram, node_count, node_offset);
	VertexPermutationKernel<<<wide_perm, 256, 0, stream>>>(as.vertex_node, ap.model.node_rank, count_ptr, capacity, node_offset, as.histogram + node_count, perm);
	launches += 2;
	ap.perm = perm;
	ap.vertex_node = as.vertex_node;
	ap.cursor = cursor;
	const bool normal_pass = want_normals || options.refine_iterations > 0 || (!want_colors && ap.scale != 1.0f);
	if (normal_pass)
	{
		AttributeParams np = ap;
		if (want_colors)
		{
			np.colors = nullptr; // the colour pass follows and applies the scale
			np.scale = 1.0f;
		}
		AttributesKernel<<<uint32_t(std::min<uint64_t>((uint64_t(capacity) + 127) / 128, wide_attr)), 128, 0, stream>>>(np);
		launches++;
	}
	if (want_colors)
	{
		static const uint32_t wide_colors = PersistentGrid(ctx, ColorsKernel, 128);
		if (normal_pass) TG_CUDA(cudaMemsetAsync(cursor, 0, 8, stream));
		ColorsKernel<<<uint32_t(std::min<uint64_t>((uint64_t(capacity) + 127) / 128, wide_colors)), 128, 0, stream>>>(ap);
		launches++;
	}
	TG_CUDA(cudaGetLastError());
	return TG_OK;
}

// K0 driver, part 1: the cull flags of every level for cell layers [k_begin, k_end) (plus the halo layer below when
// has_halo).  `flags_storage` (device, FlagWords(grid) words) is zeroed here; the item lists are scratch.
static size_t FlagWords(const DeviceGrid& grid)
{
	const uint32_t nbx = (grid.sx + kBrick - 1) / kBrick, nby = (grid.sy + kBrick - 1) / kBrick, nbz = (grid.sz + kBrick - 1) / kBrick;
	size_t words = 0;
	for (int level = 0; level < kCullLevels; ++level)
	{
		words += size_t((nbx + (1u << level) - 1) >> level) * ((nby + (1u << level) - 1) >> level) * ((nbz + (1u << level) - 1) >> level);
	}
	return words;
}

static int BuildCullFlags(Model* model, cudaStream_t stream, Scratch& scratch, const DeviceGrid& grid, uint32_t k_begin, uint32_t k_end,
	bool has_halo, bool no_cull, uint32_t* flags_storage, CullParams& cp, uint64_t& launches, std::string& error)
{
	const uint32_t nbx = (grid.sx + kBrick - 1) / kBrick, nby = (grid.sy + kBrick - 1) / kBrick, nbz = (grid.sz + kBrick - 1) / kBrick;
	const uint32_t bz_end = (k_end + kBrick - 1) / kBrick;
	const uint32_t row_begin = (has_halo ? k_begin - 1 : k_begin) / kBrick;
	cp.model = MakeDeviceModel(model);
	cp.grid = grid;
	cp.cell_k_lo = has_halo ? k_begin - 1 : k_begin;
	cp.cell_k_hi = k_end;
	cp.sample_k_lo = cp.cell_k_lo;
	cp.sample_k_hi = k_end;
	cp.ranges = nullptr;
	cp.counts = nullptr;
	cp.level = 0;
	cp.long_every_other_level = 0;
	if (const char* env = std::getenv("TG_CULL_LONG_SKIP")) cp.long_every_other_level = std::atoi(env) > 0 ? 1 : 0;
	size_t flag_words = 0;
	for (int level = 0; level < kCullLevels; ++level)
	{
		cp.dims[level][0] = (nbx + (1u << level) - 1) >> level;
		cp.dims[level][1] = (nby + (1u << level) - 1) >> level;
		cp.dims[level][2] = (nbz + (1u << level) - 1) >> level;
		cp.flags[level] = no_cull ? nullptr : flags_storage + flag_words;
		cp.lists[level] = nullptr;
		cp.capacity[level] = 0;
		cp.long_lists[level] = nullptr;
		cp.long_capacity[level] = 0;
		flag_words += size_t(cp.dims[level][0]) * cp.dims[level][1] * cp.dims[level][2];
	}
	if (no_cull) return TG_OK;
	uint32_t* counts = nullptr;
	TG_CUDA(scratch.Alloc(&counts, 2 * kCullLevels));
	cp.counts = counts;
	TG_CUDA(cudaMemsetAsync(flags_storage, 0, flag_words * 4, stream));
	TG_CUDA(cudaMemsetAsync(counts, 0, 2 * kCullLevels * 4, stream));
	TG_CUDA(scratch.Alloc(&cp.ranges, cp.model.region_count));
	for (int level = 0; level < kCullLevels; ++level)
	{
		// every region can seed up to 27 items at its start level; splits add at most the bricks of the level below
		const size_t cells = size_t(cp.dims[level][0]) * cp.dims[level][1] * ((bz_end - row_begin + (1u << level) - 1) >> level);
		cp.capacity[level] = uint32_t(std::min<size_t>(size_t(cp.model.region_count) * 8 + cells * 4 + 65536, 0x7FFFFFFFu));
		TG_CUDA(scratch.Alloc(&cp.lists[level], cp.capacity[level]));
		cp.long_capacity[level] = cp.capacity[level]; // dense scenes (10k primitives) have long programs in most regions
		TG_CUDA(scratch.Alloc(&cp.long_lists[level], cp.long_capacity[level]));
	}
	CullRegionInitKernel<<<(cp.model.region_count + 3) / 4, 128, 0, stream>>>(cp); // a warp per region
	launches++;
	// Levels run top-down with grids sized for the level's capacity bound by what can actually arrive
	// (threads beyond the device-side count exit at once), so no count is read back between levels.
	uint64_t bound = 0;
	static const uint32_t wide_long = PersistentGrid(model->context, CullLongKernel, kLongThreads);
	for (int level = kCullLevels - 1; level >= 0; --level)
	{
		const uint64_t seeded = std::min<uint64_t>(uint64_t(cp.model.region_count) * kSeedSpan * kSeedSpan * kSeedSpan, cp.capacity[level]);
		bound = std::min<uint64_t>(bound * 8u + seeded, cp.capacity[level]);
		cp.level = level;
		// the two kernels of a level are independent (both only append to the next level's lists and OR flags): the
		// long-program items run beside the short ones on the context's second stream
		Context* ctx = model->context;
		cudaStream_t side = static_cast<cudaStream_t>(ctx->stream2);
		const bool fork = side != stream && ctx->cull_events[0] && ctx->cull_events[1];
		if (fork)
		{
			TG_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ctx->cull_events[0]), stream));
			TG_CUDA(cudaStreamWaitEvent(side, static_cast<cudaEvent_t>(ctx->cull_events[0]), 0));
		}
		CullLongKernel<<<wide_long, kLongThreads, 0, fork ? side : stream>>>(cp);
		CullLevelKernel<<<uint32_t((bound + 127) / 128), 128, 0, stream>>>(cp);
		if (fork)
		{
			TG_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ctx->cull_events[1]), side));
			TG_CUDA(cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(ctx->cull_events[1]), 0));
		}
		launches += 2;
	}
	TG_CUDA(cudaGetLastError());
	if (std::getenv("TG_TRACE_CULL"))
	{
		// diagnostics: items per level (short / long) and the sizes of the long programs
		uint32_t host_counts[2 * kCullLevels];
		TG_CUDA(cudaMemcpyAsync(host_counts, counts, sizeof(host_counts), cudaMemcpyDeviceToHost, stream));
		TG_CUDA(cudaStreamSynchronize(stream));
		for (int level = kCullLevels - 1; level >= 0; --level) std::fprintf(stderr, "cull level %d: %u short items, %u long items\n", level, host_counts[level], host_counts[kCullLevels + level]);
		uint32_t long_regions = 0, max_count = 0;
		uint64_t sum = 0;
		for (const FlatRegion& r : model->flat.regions)
		{
			const uint32_t f = model->flat.nodes[r.node].flags;
			if (f & kNodeLong)
			{
				long_regions++;
				sum += f >> kNodeCountShift;
				max_count = std::max(max_count, f >> kNodeCountShift);
			}
		}
		std::fprintf(stderr, "cull: %zu regions, %u with long programs (mean %.1f instructions, max %u)\n", model->flat.regions.size(), long_regions, long_regions ? double(sum) / long_regions : 0.0, max_count);
	}
	return TG_OK;
}

// K0 driver, part 2: the list of 8-cell bricks of cell layers [k_begin, k_end) that must be evaluated (plus, for slab
// runs, the halo bricks of the row below, flagged) in *out_list, its length on the device in counters[kCntListA] and
// its capacity in *out_count.  No host round trip at all: the brick kernel reads the length itself.
static int ResolveActiveList(cudaStream_t stream, Scratch& scratch, const CullParams& cp, const DeviceGrid& grid, uint32_t k_begin, uint32_t k_end,
	bool has_halo, bool no_cull, unsigned long long* counters, uint32_t** out_list, uint64_t* out_count, uint64_t& launches, std::string& error)
{
	const uint32_t order_warps = uint32_t(scratch.ctx->sm_count) * uint32_t(scratch.ctx->brick_blocks_per_sm) * uint32_t(kBrickWarps);
	const uint32_t nbx = (grid.sx + kBrick - 1) / kBrick, nby = (grid.sy + kBrick - 1) / kBrick;
	const uint32_t bz_end = (k_end + kBrick - 1) / kBrick;
	const uint32_t row_begin = (has_halo ? k_begin - 1 : k_begin) / kBrick; // the brick row that holds the halo layer
	const size_t slab_bricks = size_t(nbx) * nby * (bz_end - row_begin);
	const size_t list_capacity = slab_bricks + 8;
	uint32_t* active_list = nullptr;
	uint32_t* ordered_list = nullptr;
	unsigned char* classes = nullptr;
	uint32_t* class_counts = nullptr; // kCostClasses totals, then kCostClasses cursors
	TG_CUDA(scratch.Alloc(&active_list, list_capacity));
	TG_CUDA(scratch.Alloc(&ordered_list, list_capacity));
	TG_CUDA(scratch.Alloc(&classes, list_capacity));
	TG_CUDA(scratch.Alloc(&class_counts, 2 * kCostClasses));
	TG_CUDA(cudaMemsetAsync(class_counts, 0, 2 * kCostClasses * sizeof(uint32_t), stream));
	const uint32_t resolve_threads = uint32_t(slab_bricks);
	CullResolveKernel<<<(resolve_threads + 255) / 256, 256, 0, stream>>>(cp, row_begin, bz_end, no_cull ? 1 : 0,
		active_list, counters + kCntListA, uint32_t(list_capacity), classes, class_counts);
	launches++;
	if (!no_cull)
	{
		// ordered when a persistent warp of the brick kernel gets fewer than a dozen bricks (TG_BRICK_ORDER=0 / 1: never / always)
		uint32_t order_limit = order_warps * 12u;
		if (const char* env = std::getenv("TG_BRICK_ORDER")) order_limit = std::atoi(env) > 0 ? 0xFFFFFFFFu : 0u;
		BrickOrderKernel<<<uint32_t((list_capacity + 255) / 256), 256, 0, stream>>>(active_list, classes, counters + kCntListA, uint32_t(list_capacity),
			class_counts, class_counts + kCostClasses, ordered_list, order_limit);
		launches++;
		active_list = ordered_list;
	}
	TG_CUDA(cudaGetLastError());
	*out_list = active_list;
	*out_count = list_capacity; // capacity of the list; its length stays on the device in counters[kCntListA]
	return TG_OK;
}

static int BuildActiveList(Model* model, cudaStream_t stream, Scratch& scratch, const DeviceGrid& grid, uint32_t k_begin, uint32_t k_end,
	bool has_halo, bool no_cull, unsigned long long* counters, uint32_t** out_list, uint64_t* out_count, uint64_t& launches, std::string& error)
{
	CullParams cp;
	uint32_t* flags = nullptr;
	if (!no_cull) TG_CUDA(scratch.Alloc(&flags, FlagWords(grid)));
	int rc = BuildCullFlags(model, stream, scratch, grid, k_begin, k_end, has_halo, no_cull, flags, cp, launches, error);
	if (rc != TG_OK) return rc;
	return ResolveActiveList(stream, scratch, cp, grid, k_begin, k_end, has_halo, no_cull, counters, out_list, out_count, launches, error);
}

// ------------------------------------------------------------------------------------------------
// Mesh export.  A MeshJob is one slab's work enqueued on the context's stream with NO host round trip: every
// count (active bricks, vertices, quads) stays on the device, later kernels read it there, and the arrays are sized
// by a capacity.  The host learns the counts from a small pinned mailbox filled by two async copies and only then
// issues the device -> host copies of the exact sizes.  If a capacity was too small the export is repeated once with
// the exact sizes (the counts are exact even when the arrays overflowed).
// ------------------------------------------------------------------------------------------------

struct Mailbox
{
	unsigned long long counters[kCntCount];
	uint32_t halo_vertices;
	uint32_t pad[15];
	uint32_t layer_starts[1024];
	unsigned long long layer_cost[1024];
};

struct MeshJob
{
	Model* model = nullptr;
	DeviceGrid grid;
	tg_mesh_options options;
	uint32_t k_begin = 0, k_end = 0, k_base = 0, layers = 0, row_words = 0;
	uint32_t bz_begin = 0, bz_end = 0, nbz_all = 0, profile_layers = 0;
	bool has_halo = false;
	uint64_t own_bricks = 0, list_capacity = 0;
	uint32_t cap_v = 0, cap_q = 0;
	unsigned long long* counters = nullptr; // device
	MeshResultDevice* result = nullptr;
	Mailbox* mailbox = nullptr;             // pinned host
	cudaEvent_t marks[6] = {};              // start, cull, eval, scan, faces, attributes
	cudaEvent_t faces_ready = nullptr, all_ready = nullptr;
	cudaEvent_t bricks_done = nullptr, base_set = nullptr; // multi-GPU: around the all-gather of the per-slab vertex counts
	uint64_t launches = 0;
	bool enqueued = false;
	int lane = 0;

	~MeshJob()
	{
		for (cudaEvent_t e : marks)
		{
			if (e) cudaEventDestroy(e);
		}
		if (faces_ready) cudaEventDestroy(faces_ready);
		if (all_ready) cudaEventDestroy(all_ready);
		if (bricks_done) cudaEventDestroy(bricks_done);
		if (base_set) cudaEventDestroy(base_set);
		if (mailbox && model) model->context->ReleaseMailbox(mailbox);
	}
};

static void DefaultCapacities(uint64_t slab_cells, uint32_t& cap_v, uint32_t& cap_q)
{
	// A surface grows with the square of the resolution: the 10k-primitive scene has 46-50 vertices per cells^(2/3) from
	// 256^3 to 2048^3 (seaside_town 5.6), so 64 per cells^(2/3) bounds the first guess on large grids where an eighth of
	// the cells would be tens of GB (an overflow repeats the export with exact sizes).  Quads: about one per vertex.
	const double side = std::cbrt(double(slab_cells));
	const uint64_t by_area = uint64_t(64.0 * side * side);
	const uint64_t v = std::min<uint64_t>(slab_cells, std::max<uint64_t>(std::min<uint64_t>(slab_cells / 8, by_area), 1 << 22));
	cap_v = uint32_t(std::min<uint64_t>(v, 0xFFFFFFF0ull));
	cap_q = uint32_t(std::min<uint64_t>(std::min<uint64_t>(v * 3, std::max<uint64_t>(v + v / 2, 1 << 20)), 0x2AAAAAA0ull));
	if (const char* env = std::getenv("TG_TEST_CAPACITY")) // tests: force the overflow-and-repeat path
	{
		const long n = std::atol(env);
		if (n > 0) cap_v = cap_q = uint32_t(n);
	}
}

// Multi-GPU export (tg_multi.inl): what a slab's job needs to learn the vertex total of the slabs below it.
struct MultiHook
{
	void* comm = nullptr;                   // ncclComm_t of this rank
	int rank = 0, ranks = 1;
	unsigned long long* gathered = nullptr; // device: `ranks` entries, filled by the all-gather
	int (*all_gather)(void* comm, const void* send, void* recv, cudaStream_t stream, std::string& error) = nullptr;
};

// index_base = vertices owned by the slabs of the lower ranks.
__global__ void GatherPrefixKernel(const unsigned long long* __restrict__ gathered, int rank, unsigned long long* index_base)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		unsigned long long sum = 0;
		for (int r = 0; r < rank; ++r) sum += gathered[r];
		*index_base = sum;
	}
}

// Enqueues one slab: culling, brick evaluation, numbering, emission, attributes, mailbox copies.  Does not wait.
// index_base (device, may be null) is added to every triangle index and then advanced by the slab's vertex count.
static int EnqueueMesh(MeshJob& job, Model* model, const tg_grid& grid_in, const tg_mesh_options& options, uint32_t cap_v, uint32_t cap_q,
	unsigned long long* index_base, std::string& error, int lane = 0, cudaEvent_t base_ready = nullptr, const CullParams* shared_cull = nullptr,
	const MultiHook* hook = nullptr)
{
	Context* ctx = model->context;
	cudaStream_t stream = static_cast<cudaStream_t>(lane ? ctx->stream2 : ctx->stream);
	job.lane = lane;
	const bool trace = std::getenv("TG_TRACE_HOST") != nullptr;
	auto host_now = [] { return std::chrono::steady_clock::now(); };
	const auto h_begin = host_now();
	auto host_us = [&](const std::chrono::steady_clock::time_point& since) { return std::chrono::duration<double, std::micro>(host_now() - since).count(); };
	job.model = model;
	job.options = options;
	DeviceGrid& grid = job.grid;
	if (!MakeDeviceGrid(grid_in, grid, error)) return TG_ERR_INVALID;

	uint32_t k_begin = 0, k_end = grid.sz;
	if (options.slab_begin != 0 || options.slab_end != 0)
	{
		k_begin = uint32_t(options.slab_begin);
		k_end = uint32_t(std::min<uint64_t>(options.slab_end, grid.sz));
		if (k_begin >= k_end)
		{
			error = "slab is empty";
			return TG_ERR_INVALID;
		}
	}
	const bool has_halo = k_begin > 0;
	const bool face_normals = (options.flags & TG_MESH_FACE_NORMALS) != 0;
	if (face_normals && has_halo)
	{
		error = "face normals are not available for slab exports";
		return TG_ERR_UNSUPPORTED;
	}
	const uint32_t k_base = has_halo ? k_begin - 1 : k_begin;
	const uint32_t layers = k_end - k_base;
	const uint32_t row_words = (grid.sx + 63) / 64;
	const size_t bitmap_words = size_t(layers) * grid.sy * row_words;
	const uint32_t nbx = (grid.sx + kBrick - 1) / kBrick, nby = (grid.sy + kBrick - 1) / kBrick;
	const uint32_t bz_begin = k_begin / kBrick, bz_end = (k_end + kBrick - 1) / kBrick;
	const uint32_t nbz_all = (grid.sz + kBrick - 1) / kBrick;
	if (nbz_all > 1024)
	{
		error = "grids deeper than 8192 cell layers are not supported";
		return TG_ERR_UNSUPPORTED;
	}
	job.k_begin = k_begin;
	job.k_end = k_end;
	job.k_base = k_base;
	job.layers = layers;
	job.row_words = row_words;
	job.bz_begin = bz_begin;
	job.bz_end = bz_end;
	job.nbz_all = nbz_all;
	job.profile_layers = bz_end - bz_begin;
	job.has_halo = has_halo;
	job.own_bricks = uint64_t(nbx) * nby * (bz_end - bz_begin);
	const uint64_t slab_cells = uint64_t(grid.sx) * grid.sy * (k_end - k_begin);
	if (cap_v == 0)
	{
		DefaultCapacities(slab_cells, cap_v, cap_q);
		if (!std::getenv("TG_TEST_CAPACITY"))
		{
			for (const Context::ExportHint& h : ctx->hints)
			{
				if (h.model == model && h.sx == grid.sx && h.sy == grid.sy && h.sz == grid.sz && h.k_begin == k_begin && h.k_end == k_end &&
					h.flags == (options.flags & TG_MESH_NO_CULL))
				{
					// the same export ran before on this context: its counts (plus a little) are the capacities
					cap_v = uint32_t(std::min<uint64_t>(h.vertices + h.vertices / 64 + 1024, 0xFFFFFFF0ull));
					cap_q = uint32_t(std::min<uint64_t>(h.quads + h.quads / 64 + 1024, 0x2AAAAAA0ull));
					break;
				}
			}
		}
	}
	job.cap_v = cap_v;
	job.cap_q = cap_q;

	for (cudaEvent_t& e : job.marks) TG_CUDA(cudaEventCreate(&e));
	TG_CUDA(cudaEventCreateWithFlags(&job.faces_ready, cudaEventDisableTiming));
	TG_CUDA(cudaEventCreateWithFlags(&job.all_ready, cudaEventDisableTiming));
	static_assert(sizeof(Mailbox) <= kMailboxBytes, "mailbox block too small");
	job.mailbox = static_cast<Mailbox*>(ctx->AcquireMailbox(error));
	if (!job.mailbox) return TG_ERR_MEMORY;
	job.result = new MeshResultDevice(ctx);
	MeshResultDevice* result = job.result;

	Scratch scratch(ctx, lane);
	uint64_t& launches = job.launches;
	unsigned long long* counters = nullptr;
	uint32_t* active_list = nullptr;
	unsigned long long* bitmap = nullptr;
	uint2* prefix = nullptr;
	TG_CUDA(scratch.Alloc(&counters, kCntCount));
	TG_CUDA(scratch.Alloc(&bitmap, bitmap_words));
	TG_CUDA(scratch.Alloc(&prefix, bitmap_words));
	TG_CUDA(cudaMemsetAsync(counters, 0, kCntCount * 8, stream));
	TG_CUDA(cudaMemsetAsync(bitmap, 0, bitmap_words * 8, stream));
	uint32_t* word_flags = nullptr;
	TG_CUDA(scratch.Alloc(&word_flags, (bitmap_words + 31) / 32));
	TG_CUDA(cudaMemsetAsync(word_flags, 0, (bitmap_words + 31) / 32 * 4, stream));
	job.counters = counters;
	TG_CUDA(cudaEventRecord(job.marks[0], stream));
	const double h_setup = host_us(h_begin);

	// ---- K0: active brick list -------------------------------------------------------------------
	const bool no_cull = (options.flags & TG_MESH_NO_CULL) != 0;
	{
		// a pipelined export culls the whole grid once and every slab only resolves its rows of the shared flags
		const int rc0 = shared_cull
			? ResolveActiveList(stream, scratch, *shared_cull, grid, k_begin, k_end, has_halo, no_cull, counters, &active_list, &job.list_capacity, launches, error)
			: BuildActiveList(model, stream, scratch, grid, k_begin, k_end, has_halo, no_cull, counters, &active_list, &job.list_capacity, launches, error);
		if (rc0 != TG_OK) return rc0;
	}
	TG_CUDA(cudaEventRecord(job.marks[1], stream));

	// ---- K1 + K2: evaluate bricks, classify, extract vertices ------------------------------------
	// (result buffers first: an allocator call must not sit between this rank's collective and its peers')
	{ void* p_ = ctx->AcquireDevice(size_t(cap_v) * 12, error); if (!p_) return TG_ERR_MEMORY; result->d_positions = static_cast<decltype(result->d_positions)>(p_); }
	{ void* p_ = ctx->AcquireDevice(size_t(cap_q) * 24, error); if (!p_) return TG_ERR_MEMORY; result->d_triangles = static_cast<decltype(result->d_triangles)>(p_); }
	float4* tmp_pos = nullptr;
	unsigned long long* tmp_key = nullptr;
	TG_CUDA(scratch.Alloc(&tmp_pos, cap_v));
	TG_CUDA(scratch.Alloc(&tmp_key, cap_v));
	MeshParams mp;
	mp.model = MakeDeviceModel(model);
	mp.grid = grid;
	mp.bricks = active_list;
	mp.brick_count = counters + kCntListA;
	mp.brick_capacity = uint32_t(job.list_capacity);
	mp.bitmap = bitmap;
	mp.word_flags = word_flags;
	mp.row_words = row_words;
	mp.k_base = k_base;
	mp.k_own_begin = k_begin;
	mp.k_own_end = k_end;
	mp.counters = counters;
	mp.tmp_pos = tmp_pos;
	mp.tmp_key = tmp_key;
	mp.tmp_capacity = cap_v;
	mp.progress = ctx->progress_words;
	mp.progress_base = ctx->progress_base;
	mp.live = (options.flags & TG_MESH_LIVE_FIELD) ? 1u : 0u;
	// persistent warps: the grid fills every SM; warps beyond the list length leave at their first fetch
	if (options.flags & TG_MESH_FAST)
	{
		static const int fast_blocks_per_sm = FastBrickBlocksPerSm();
		TG_CUDA(cudaError_t(LaunchMeshBricksFast(&mp, uint32_t(ctx->sm_count) * uint32_t(fast_blocks_per_sm), stream)));
	}
	else
	{
		MeshBricksKernel<<<uint32_t(ctx->sm_count) * uint32_t(ctx->brick_blocks_per_sm), kBrickThreads, 0, stream>>>(mp);
		TG_CUDA(cudaGetLastError());
	}
	launches++;
	TG_CUDA(cudaEventRecord(job.marks[2], stream));
	if (hook)
	{
		// The one exchange of the path: every rank's owned vertex count, all-gathered over NVLink on the
		// side stream while this stream goes on numbering; FinalizeMeshKernel waits for the resulting index base.
		cudaStream_t side = static_cast<cudaStream_t>(ctx->stream2);
		TG_CUDA(cudaEventCreateWithFlags(&job.bricks_done, cudaEventDisableTiming));
		TG_CUDA(cudaEventCreateWithFlags(&job.base_set, cudaEventDisableTiming));
		TG_CUDA(cudaEventRecord(job.bricks_done, stream));
		TG_CUDA(cudaStreamWaitEvent(side, job.bricks_done, 0));
		const int rcg = hook->all_gather(hook->comm, counters + kCntTmpVertices, hook->gathered, side, error);
		if (rcg != TG_OK) return rcg;
		GatherPrefixKernel<<<1, 32, 0, side>>>(hook->gathered, hook->rank, index_base);
		launches += 2;
		TG_CUDA(cudaGetLastError());
		TG_CUDA(cudaEventRecord(job.base_set, side));
		base_ready = job.base_set;
	}

	// ---- vertex + quad numbering: one dual scan over the bitmap ----------------------------------
	{
		const LoadVertexQuadCounts load{ bitmap, row_words, grid.sy };
		const uint32_t warps = uint32_t((bitmap_words + kPairWarpWords - 1) / kPairWarpWords);
		const uint32_t blocks = (warps + kPairWarps - 1) / kPairWarps;
		unsigned long long* sums = nullptr;
		TG_CUDA(scratch.Alloc(&sums, std::max<uint32_t>(warps, 1)));
		PairSumsKernel<<<blocks, kScanBlock, 0, stream>>>(load, word_flags, bitmap_words, prefix, sums);
		PairSumsScanKernel<<<1, 1024, 0, stream>>>(sums, warps, counters + kCntTotalVertices);
		PairPrefixKernel<<<blocks, kScanBlock, 0, stream>>>(word_flags, bitmap_words, sums, prefix, size_t(grid.sy) * row_words);
		launches += 3;
		TG_CUDA(cudaGetLastError());
	}
	TG_CUDA(cudaEventRecord(job.marks[3], stream));
	const double h_scan = host_us(h_begin);

	// ---- K3: final vertex order, triangles -------------------------------------------------------
	uint32_t* layer_starts = nullptr;
	unsigned long long* layer_cost = nullptr;
	TG_CUDA(scratch.Alloc(&layer_starts, job.profile_layers + 2));
	TG_CUDA(scratch.Alloc(&layer_cost, nbz_all));
	TG_CUDA(cudaMemsetAsync(layer_cost, 0, size_t(nbz_all) * 8, stream));
	const double h_alloc = host_us(h_begin);
	const bool attribute_pass = WantsAttributePass(model, options);
	AttributeScratch as;
	if (attribute_pass)
	{
		const int rc1 = PrepareAttributeScratch(model, stream, scratch, cap_v, as, error);
		if (rc1 != TG_OK) return rc1;
	}
	FaceParams fp;
	fp.model = mp.model;
	fp.bitmap = bitmap;
	fp.prefix = prefix;
	fp.row_words = row_words;
	fp.sy = grid.sy;
	fp.k_base = k_base;
	fp.has_halo = has_halo ? 1u : 0u;
	fp.tmp_pos = tmp_pos;
	fp.tmp_key = tmp_key;
	fp.counters = counters;
	fp.vertex_capacity = cap_v;
	fp.quad_capacity = cap_q;
	fp.index_base = index_base;
	fp.positions = result->d_positions;
	fp.triangles = result->d_triangles;
	fp.vertex_node = as.vertex_node;
	fp.node_histogram = as.histogram;
	fp.layer_starts = layer_starts;
	fp.profile_first_layer = bz_begin;
	fp.profile_layers = job.profile_layers;
	fp.layers_in_bitmap = layers;
	static const uint32_t wide_finalize = PersistentGrid(ctx, FinalizeMeshKernel, 256);
	if (base_ready) TG_CUDA(cudaStreamWaitEvent(stream, base_ready, 0)); // the slab below has advanced the index base
	FinalizeMeshKernel<<<wide_finalize, 256, 0, stream>>>(fp);
	launches++;
	if (index_base && !hook)
	{
		AdvanceIndexBaseKernel<<<1, 32, 0, stream>>>(index_base, counters, cap_v);
		launches++;
	}
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaEventRecord(job.marks[4], stream));
	// first mailbox delivery: every count, and the halo total (vertices of bitmap layer 0)
	job.mailbox->halo_vertices = 0;
	MailKernel<<<1, 128, 0, stream>>>(reinterpret_cast<const uint32_t*>(counters), kCntCount * 2, reinterpret_cast<uint32_t*>(job.mailbox->counters),
		layer_starts, job.profile_layers, job.mailbox->layer_starts);
	if (has_halo) MailKernel<<<1, 32, 0, stream>>>(&prefix[size_t(grid.sy) * row_words].x, 1, &job.mailbox->halo_vertices, nullptr, 0, nullptr);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaEventRecord(job.faces_ready, stream));

	// ---- K4: attributes ----------------------------------------------------------------------
	ctx->stage.store(options.refine_iterations > 0 ? 2 : 3);
	const float half[3] = { grid.dx / 2.0f, grid.dy / 2.0f, grid.dz / 2.0f };
	int rc = EnqueueAttributes(model, scratch, result, counters + kCntTmpVertices, counters + kCntAttrCursor, cap_v, options, half, as, true, launches, error, layer_cost, grid.z, grid.dz, nbz_all);
	if (rc != TG_OK) return rc;
	if (face_normals)
	{
		{ void* p_ = ctx->AcquireDevice(size_t(cap_q) * 24, error); if (!p_) return TG_ERR_MEMORY; result->d_face_normals = static_cast<decltype(result->d_face_normals)>(p_); }
		FaceNormalsKernel<<<uint32_t(ctx->sm_count) * 16u, 128, 0, stream>>>(MakeDeviceModel(model), result->d_positions, result->d_triangles, counters + kCntTotalQuads, cap_q, result->d_face_normals);
		launches++;
		const float scale = options.scale == 0.0f ? 1.0f : options.scale;
		if (scale != 1.0f)
		{
			ScalePositionsKernel<<<uint32_t(ctx->sm_count) * 8u, 256, 0, stream>>>(result->d_positions, counters + kCntTmpVertices, cap_v, scale);
			launches++;
		}
	}
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaEventRecord(job.marks[5], stream));
	MailKernel<<<1, 128, 0, stream>>>(reinterpret_cast<const uint32_t*>(layer_cost), nbz_all * 2, reinterpret_cast<uint32_t*>(job.mailbox->layer_cost), nullptr, 0, nullptr);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaEventRecord(job.all_ready, stream));
	job.enqueued = true;
	if (trace) std::fprintf(stderr, "host enqueue us: setup %.0f  through scan %.0f  result allocs %.0f  total %.0f\n", h_setup, h_scan, h_alloc, host_us(h_begin));
	return TG_OK;
}

// What a finished job reports once its first mailbox delivery arrived.
struct MeshCounts
{
	uint64_t vertices = 0, quads = 0, halo = 0;
	bool overflow = false;
};

static int WaitCounts(MeshJob& job, MeshCounts& counts, std::string& error)
{
	TG_CUDA(cudaEventSynchronize(job.faces_ready));
	const Mailbox& mb = *job.mailbox;
	counts.vertices = mb.counters[kCntTmpVertices];
	counts.quads = mb.counters[kCntTotalQuads];
	counts.halo = mb.halo_vertices;
	counts.overflow = counts.vertices > job.cap_v || counts.quads > job.cap_q;
	// remember the counts for the next export of the same slab of the same model (most recent first, a few dozen kept)
	Context* ctx = job.model->context;
	Context::ExportHint hint = { job.model, job.grid.sx, job.grid.sy, job.grid.sz, job.k_begin, job.k_end, job.options.flags & TG_MESH_NO_CULL, counts.vertices, counts.quads };
	for (size_t i = 0; i < ctx->hints.size(); ++i)
	{
		const Context::ExportHint& h = ctx->hints[i];
		if (h.model == hint.model && h.sx == hint.sx && h.sy == hint.sy && h.sz == hint.sz && h.k_begin == hint.k_begin && h.k_end == hint.k_end && h.flags == hint.flags)
		{
			ctx->hints.erase(ctx->hints.begin() + long(i));
			break;
		}
	}
	ctx->hints.insert(ctx->hints.begin(), hint);
	if (ctx->hints.size() > 48) ctx->hints.pop_back();
	return TG_OK;
}

// Fills the host-side bookkeeping of *out from a completed job (everything but the arrays).
static int FinishJob(MeshJob& job, const MeshCounts& counts, tg_mesh* out, std::string& error)
{
	TG_CUDA(cudaEventSynchronize(job.all_ready));
	const Mailbox& mb = *job.mailbox;
	tg_mesh_timings& tm = out->timings;
	auto ms = [&](int a, int b) {
		float v = 0.f;
		cudaEventElapsedTime(&v, job.marks[a], job.marks[b]);
		return v;
	};
	tm.cull_ms = ms(0, 1);
	tm.evaluate_ms = ms(1, 2);
	tm.compact_ms = ms(2, 3);
	tm.faces_ms = ms(3, 4);
	tm.attributes_ms = ms(4, 5);
	tm.total_device_ms = ms(0, 5);
	tm.bricks_total = job.own_bricks;
	tm.bricks_evaluated = std::min<unsigned long long>(mb.counters[kCntListA], job.list_capacity);
	tm.samples_evaluated = tm.bricks_evaluated ? mb.counters[kCntSamples] : 0;
	tm.algorithmic_flops = tm.bricks_evaluated ? mb.counters[kCntFlops] : 0;
#ifdef TG_COUNT_SLOTS
	std::fprintf(stderr, "interpreter dispatches: %llu samples in %llu slots (%.1f %% used), %llu slots in dispatches of <= 32 samples\n", (unsigned long long)mb.counters[kCntSamples],
		(unsigned long long)mb.counters[kCntSlots], 100.0 * double(mb.counters[kCntSamples]) / double(std::max<unsigned long long>(mb.counters[kCntSlots], 1)), (unsigned long long)mb.counters[kCntTailSlots]);
#endif
	tm.kernel_launches = job.launches;
	out->vertex_count = counts.vertices;
	out->triangle_count = counts.quads * 2;
	out->halo_vertices = counts.halo;
	// vertices owned per brick layer (absolute layer index); layers outside the slab stay 0
	const uint32_t nbz = job.nbz_all;
	uint32_t* profile = static_cast<uint32_t*>(std::calloc(nbz, sizeof(uint32_t)));
	if (profile && counts.vertices > 0)
	{
		for (uint32_t i = 0; i < job.profile_layers; ++i)
		{
			const uint32_t begin = mb.layer_starts[i];
			const uint32_t end = (i + 1 < job.profile_layers && mb.layer_starts[i + 1] != 0xFFFFFFFFu) ? mb.layer_starts[i + 1] : uint32_t(counts.vertices + counts.halo);
			profile[job.bz_begin + i] = end >= begin ? end - begin : 0u;
		}
	}
	out->layer_vertices = profile;
	out->layer_count = nbz;
	double* cost = static_cast<double*>(std::calloc(nbz, sizeof(double)));
	if (cost)
	{
		for (uint32_t i = 0; i < nbz; ++i) cost[i] = double(mb.layer_cost[i]);
	}
	out->layer_vertex_cost = cost;
	return TG_OK;
}

static void FreeResultDevice(MeshResultDevice* r, cudaStream_t s)
{
	if (!r) return;
	(void)s; // every caller has synchronised the streams that used these buffers
	r->context->ReleaseDevice(r->d_positions);
	r->context->ReleaseDevice(r->d_normals);
	r->context->ReleaseDevice(r->d_colors);
	r->context->ReleaseDevice(r->d_triangles);
	r->context->ReleaseDevice(r->d_face_normals);
	r->d_positions = r->d_normals = r->d_face_normals = nullptr;
	r->d_colors = nullptr;
	r->d_triangles = nullptr;
}

// One slab (or the whole grid) in one shot.
static int ExportMeshOnce(Model* model, const tg_grid& grid_in, const tg_mesh_options& options, tg_mesh* out, std::string& error)
{
	Context* ctx = model->context;
	cudaStream_t stream = StreamOf(ctx);
	cudaStream_t copy_stream = static_cast<cudaStream_t>(ctx->copy_stream);
	const bool want_host = !(options.flags & TG_MESH_DEVICE_ONLY);
	uint32_t cap_v = 0, cap_q = 0;
	for (int attempt = 0; attempt < 2; ++attempt)
	{
		MeshJob job;
		int rc = EnqueueMesh(job, model, grid_in, options, cap_v, cap_q, nullptr, error);
		MeshCounts counts;
		if (rc == TG_OK) rc = WaitCounts(job, counts, error);
		if (rc != TG_OK)
		{
			cudaStreamSynchronize(stream);
			FreeResultDevice(job.result, stream);
			delete job.result;
			return rc;
		}
		if (counts.vertices > 0xFFFFFFF0ull || counts.quads * 2 > 0xFFFFFFF0ull)
		{
			cudaStreamSynchronize(stream);
			FreeResultDevice(job.result, stream);
			delete job.result;
			error = "mesh exceeds 2^32 vertices or triangles";
			return TG_ERR_UNSUPPORTED;
		}
		if (counts.overflow)
		{
			// exact sizes are known now; run again with them
			cudaStreamSynchronize(stream);
			FreeResultDevice(job.result, stream);
			delete job.result;
			if (attempt == 1)
			{
				error = "mesh capacities overflowed twice";
				return TG_ERR_CUDA;
			}
			cap_v = uint32_t(std::max<uint64_t>(counts.vertices, 1));
			cap_q = uint32_t(std::max<uint64_t>(counts.quads, 1));
			continue;
		}
		MeshResultDevice* result = job.result;
		out->opaque = result;
		const auto h0 = std::chrono::steady_clock::now();
		auto fetch = [&](cudaStream_t on, void* device, size_t bytes, void** host) -> int
		{
			if (!device || bytes == 0) return TG_OK;
			void* p = ctx->AcquirePinned(bytes, error);
			if (!p) return TG_ERR_MEMORY;
			result->pinned.push_back(p);
			TG_CUDA(cudaMemcpyAsync(p, device, bytes, cudaMemcpyDeviceToHost, on));
			*host = p;
			return TG_OK;
		};
		if (want_host && counts.quads > 0)
		{
			// the index buffer is final: bring it home on the copy stream while the attributes are computed
			TG_CUDA(cudaStreamWaitEvent(copy_stream, job.faces_ready, 0));
			if ((rc = fetch(copy_stream, result->d_triangles, size_t(counts.quads) * 24, reinterpret_cast<void**>(&out->triangles))) != TG_OK) return rc;
		}
		if ((rc = FinishJob(job, counts, out, error)) != TG_OK) return rc;
		if (want_host && counts.vertices > 0)
		{
			const size_t v = size_t(counts.vertices);
			if ((rc = fetch(stream, result->d_positions, v * 12, reinterpret_cast<void**>(&out->positions))) != TG_OK) return rc;
			if ((rc = fetch(stream, result->d_normals, v * 12, reinterpret_cast<void**>(&out->normals))) != TG_OK) return rc;
			if ((rc = fetch(stream, result->d_colors, v * 3, reinterpret_cast<void**>(&out->colors))) != TG_OK) return rc;
			if ((rc = fetch(stream, result->d_face_normals, size_t(counts.quads) * 24, reinterpret_cast<void**>(&out->face_normals))) != TG_OK) return rc;
		}
		if (want_host)
		{
			TG_CUDA(cudaStreamSynchronize(copy_stream));
			TG_CUDA(cudaStreamSynchronize(stream));
			// the arrays are home: the device copies are not needed any more
			FreeResultDevice(result, stream);
		}
		out->timings.download_ms = want_host ? float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count()) : 0.0f;
		return TG_OK;
	}
	return TG_ERR_CUDA;
}

// Host-side destination of a pipelined export: one page-locked array per attribute, filled slab by slab.
struct HostArrays
{
	Context* ctx = nullptr;
	float* positions = nullptr;
	float* normals = nullptr;
	uint8_t* colors = nullptr;
	uint32_t* triangles = nullptr;
	uint64_t cap_v = 0, cap_t = 0;
	bool want_normals = false, want_colors = false;

	void Release()
	{
		if (positions) ctx->ReleasePinned(positions);
		if (normals) ctx->ReleasePinned(normals);
		if (colors) ctx->ReleasePinned(colors);
		if (triangles) ctx->ReleasePinned(triangles);
		positions = normals = nullptr;
		colors = nullptr;
		triangles = nullptr;
		cap_v = cap_t = 0;
	}

	template <typename T>
	bool Grow(T*& array, uint64_t old_count, uint64_t new_count, size_t item_bytes, std::string& error)
	{
		T* bigger = static_cast<T*>(ctx->AcquirePinned(size_t(new_count) * item_bytes, error));
		if (!bigger) return false;
		if (array)
		{
			std::memcpy(bigger, array, size_t(old_count) * item_bytes);
			ctx->ReleasePinned(array);
		}
		array = bigger;
		return true;
	}

	// Makes room for v vertices and t triangles; `live_v` / `live_t` entries are already in place (copies drained by the caller).
	bool Reserve(uint64_t v, uint64_t t, uint64_t live_v, uint64_t live_t, std::string& error)
	{
		if (v > cap_v)
		{
			const uint64_t n = std::max<uint64_t>(v + v / 2, 1024);
			if (!Grow(positions, live_v, n, 12, error)) return false;
			if (want_normals && !Grow(normals, live_v, n, 12, error)) return false;
			if (want_colors && !Grow(colors, live_v, n, 3, error)) return false;
			cap_v = n;
		}
		if (t > cap_t)
		{
			const uint64_t n = std::max<uint64_t>(t + t / 2, 1024);
			if (!Grow(triangles, live_t, n, 12, error)) return false;
			cap_t = n;
		}
		return true;
	}
};

constexpr int TG_RETRY_ONE_SHOT = 1000; // internal: the pipelined export gave up (a slab overflowed its capacity)

// Whole-grid export with host results, software-pipelined over z-slabs on ONE device: while slab c is evaluated its
// predecessor's arrays travel device -> host on the copy stream, straight to their final place in the host arrays
// (vertex numbering is (k, j, i)-lexicographic, so slabs concatenate; triangle indices are made global on the device
// by the running index base).  Same kernels and the same halo rule as the multi-GPU partition.
// relative slab costs of the pipelined export, by slab count (see ExportMeshPipelined)
static const double kPipelineShares3[3] = { 1.0, 1.0, 1.0 };
static const double kPipelineShares4[4] = { 0.7, 1.3, 1.3, 0.7 };
static const double kPipelineShares5[5] = { 1.0, 1.0, 1.0, 1.0, 1.0 };

static int ExportMeshPipelined(Model* model, const tg_grid& grid_in, const tg_mesh_options& options, int chunks, tg_mesh* out, std::string& error)
{
	Context* ctx = model->context;
	cudaStream_t stream = StreamOf(ctx);
	cudaStream_t copy_stream = static_cast<cudaStream_t>(ctx->copy_stream);
	DeviceGrid grid;
	if (!MakeDeviceGrid(grid_in, grid, error)) return TG_ERR_INVALID;
	const uint32_t nbz = (grid.sz + kBrick - 1) / kBrick;
	chunks = std::max(1, std::min<int>(chunks, int(nbz)));
	if (!ctx->index_base) TG_CUDA(cudaMalloc(&ctx->index_base, 8));
	unsigned long long* index_base = static_cast<unsigned long long*>(ctx->index_base);
	TG_CUDA(cudaMemsetAsync(index_base, 0, 8, stream));
	// Optional second compute lane (TG_PIPELINE_LANES=2): even slabs on the context's stream, odd slabs on the second
	// one, each with its own scratch arena.  The lanes meet only at the running index base (slab c's triangles need
	// the vertex total of slabs < c).
	cudaStream_t lane1 = static_cast<cudaStream_t>(ctx->stream2);
	// K0 once for the whole grid; its flags outlive the slabs, its item lists only this block (they sit in the lane-0
	// arena, which slab 0 reuses later on the same stream)
	CullParams shared_cull;
	uint32_t* shared_flags = nullptr;
	uint64_t cull_launches = 0;
	const bool no_cull = (options.flags & TG_MESH_NO_CULL) != 0;
	cudaEvent_t cull_marks[2] = { nullptr, nullptr };
	TG_CUDA(cudaEventCreate(&cull_marks[0]));
	TG_CUDA(cudaEventCreate(&cull_marks[1]));
	TG_CUDA(cudaEventRecord(cull_marks[0], stream));
	{
		if (!no_cull) { void* p_ = ctx->AcquireDevice(FlagWords(grid) * 4, error); if (!p_) return TG_ERR_MEMORY; shared_flags = static_cast<decltype(shared_flags)>(p_); }
		Scratch cull_scratch(ctx, 0);
		const int rc = BuildCullFlags(model, stream, cull_scratch, grid, 0, grid.sz, false, no_cull, shared_flags, shared_cull, cull_launches, error);
		if (rc != TG_OK)
		{
			ctx->ReleaseDevice(shared_flags);
			cudaEventDestroy(cull_marks[0]);
			cudaEventDestroy(cull_marks[1]);
			return rc;
		}
	}
	TG_CUDA(cudaEventRecord(cull_marks[1], stream));
	TG_CUDA(cudaStreamWaitEvent(lane1, cull_marks[1], 0));

	HostArrays host;
	host.ctx = ctx;
	host.want_normals = (options.flags & TG_MESH_NORMALS) != 0;
	host.want_colors = (options.flags & TG_MESH_COLORS) != 0 && model->flat.has_paint;
	const uint32_t shape_flags = options.flags & (TG_MESH_NORMALS | TG_MESH_COLORS | TG_MESH_NO_CULL);
	{
		// exact when this context just exported the same model on the same grid, else surface ~ 6 cells^(2/3)
		const auto& last = ctx->last_export;
		const double cells = double(grid.sx) * grid.sy * grid.sz;
		uint64_t v = uint64_t(6.0 * std::pow(cells, 2.0 / 3.0)) + 65536, t = 2 * v;
		if (last.model == model && last.sx == grid.sx && last.sy == grid.sy && last.sz == grid.sz && last.flags == shape_flags)
		{
			v = last.vertices;
			t = last.triangles;
		}
		host.positions = static_cast<float*>(ctx->AcquirePinned(size_t(std::max<uint64_t>(v, 1)) * 12, error));
		if (host.want_normals) host.normals = static_cast<float*>(ctx->AcquirePinned(size_t(std::max<uint64_t>(v, 1)) * 12, error));
		if (host.want_colors) host.colors = static_cast<uint8_t*>(ctx->AcquirePinned(size_t(std::max<uint64_t>(v, 1)) * 3, error));
		host.triangles = static_cast<uint32_t*>(ctx->AcquirePinned(size_t(std::max<uint64_t>(t, 1)) * 12, error));
		if (!host.positions || !host.triangles || (host.want_normals && !host.normals) || (host.want_colors && !host.colors))
		{
			host.Release();
			return TG_ERR_MEMORY;
		}
		host.cap_v = v;
		host.cap_t = t;
	}

	std::vector<std::unique_ptr<MeshJob>> jobs;
	jobs.reserve(size_t(chunks));
	uint64_t v_done = 0, t_done = 0;
	tg_mesh_timings total;
	std::memset(&total, 0, sizeof(total));
	std::vector<uint32_t> layer_vertices(nbz, 0u);
	std::vector<double> layer_cost(nbz, 0.0);

	auto abandon = [&](int rc) {
		cudaStreamSynchronize(stream);
		cudaStreamSynchronize(lane1);
		cudaStreamSynchronize(copy_stream);
		ctx->ReleaseDevice(shared_flags);
		cudaEventDestroy(cull_marks[0]);
		cudaEventDestroy(cull_marks[1]);
		for (auto& j : jobs)
		{
			if (j && j->result)
			{
				FreeResultDevice(j->result, stream);
				delete j->result;
				j->result = nullptr;
			}
		}
		host.Release();
		return rc;
	};

	bool overflowed = false;
	auto collect = [&](int c) -> int {
		MeshJob& job = *jobs[size_t(c)];
		MeshCounts counts;
		int rc = WaitCounts(job, counts, error);
		if (rc != TG_OK) return rc;
		// A slab that outgrew its capacities spoils the export, but the remaining slabs are still collected (results
		// dropped): WaitCounts records every slab's exact counts, so the NEXT export of this grid pipelines cleanly
		// instead of tripping over the next slab.
		if (counts.overflow) overflowed = true;
		if (overflowed) return TG_OK;
		const uint64_t v = counts.vertices, t = counts.quads * 2;
		if (v_done + v > 0xFFFFFFF0ull || t_done + t > 0xFFFFFFF0ull)
		{
			error = "mesh exceeds 2^32 vertices or triangles";
			return TG_ERR_UNSUPPORTED;
		}
		if (v_done + v > host.cap_v || t_done + t > host.cap_t)
		{
			TG_CUDA(cudaStreamSynchronize(copy_stream)); // earlier slabs must have landed before the arrays move
			if (!host.Reserve(v_done + v, t_done + t, v_done, t_done, error)) return TG_ERR_MEMORY;
		}
		MeshResultDevice* r = job.result;
		if (t > 0)
		{
			TG_CUDA(cudaStreamWaitEvent(copy_stream, job.faces_ready, 0));
			TG_CUDA(cudaMemcpyAsync(host.triangles + t_done * 3, r->d_triangles, size_t(t) * 12, cudaMemcpyDeviceToHost, copy_stream));
		}
		tg_mesh part;
		std::memset(&part, 0, sizeof(part));
		if ((rc = FinishJob(job, counts, &part, error)) != TG_OK) return rc;
		if (v > 0)
		{
			TG_CUDA(cudaStreamWaitEvent(copy_stream, job.all_ready, 0));
			TG_CUDA(cudaMemcpyAsync(host.positions + v_done * 3, r->d_positions, size_t(v) * 12, cudaMemcpyDeviceToHost, copy_stream));
			if (host.normals && r->d_normals) TG_CUDA(cudaMemcpyAsync(host.normals + v_done * 3, r->d_normals, size_t(v) * 12, cudaMemcpyDeviceToHost, copy_stream));
			if (host.colors && r->d_colors) TG_CUDA(cudaMemcpyAsync(host.colors + v_done * 3, r->d_colors, size_t(v) * 3, cudaMemcpyDeviceToHost, copy_stream));
		}
		const tg_mesh_timings& tm = part.timings;
		total.cull_ms += tm.cull_ms;
		total.evaluate_ms += tm.evaluate_ms;
		total.compact_ms += tm.compact_ms;
		total.faces_ms += tm.faces_ms;
		total.attributes_ms += tm.attributes_ms;
		total.bricks_total += tm.bricks_total;
		total.bricks_evaluated += tm.bricks_evaluated;
		total.samples_evaluated += tm.samples_evaluated;
		total.algorithmic_flops += tm.algorithmic_flops;
		total.kernel_launches += tm.kernel_launches;
		for (uint32_t i = 0; i < nbz && i < part.layer_count; ++i)
		{
			if (part.layer_vertices) layer_vertices[i] += part.layer_vertices[i];
			if (part.layer_vertex_cost) layer_cost[i] += part.layer_vertex_cost[i];
		}
		std::free(part.layer_vertices);
		std::free(part.layer_vertex_cost);
		v_done += v;
		t_done += t;
		ctx->progress_done[0] = uint64_t(c + 1);
		return TG_OK;
	};

	// Slab cuts, in brick layers.  The copy of slab c runs while slab c + 1 is evaluated, and nothing can be copied before
	// the first slab is done, and the export ends one copy after the last slab: the slabs are cut by estimated COST (the
	// multi-GPU planner's host-side estimate, tg_multi.inl), the first and the last smaller than the middle ones.
	std::vector<uint32_t> cut(size_t(chunks) + 1, 0u);
	{
		for (int c = 0; c <= chunks; ++c) cut[size_t(c)] = uint32_t(uint64_t(nbz) * uint64_t(c) / uint64_t(chunks));
		std::vector<double> shares;
		if (const char* env = std::getenv("TG_PIPELINE_SHARES")) // tuning: comma-separated relative slab costs
		{
			for (const char* p = env; *p;)
			{
				char* end = nullptr;
				const double v = std::strtod(p, &end);
				if (end == p) break;
				shares.push_back(v);
				p = *end == ',' ? end + 1 : end;
			}
		}
		else if (chunks == 3) shares = { kPipelineShares3[0], kPipelineShares3[1], kPipelineShares3[2] };
		else if (chunks == 4) shares = { kPipelineShares4[0], kPipelineShares4[1], kPipelineShares4[2], kPipelineShares4[3] };
		else if (chunks == 5) shares = { kPipelineShares5[0], kPipelineShares5[1], kPipelineShares5[2], kPipelineShares5[3], kPipelineShares5[4] };
		if (shares.size() != size_t(chunks)) shares.assign(size_t(chunks), 1.0);
		auto& plan = model->pipeline_plan;
		if (!plan.valid || std::memcmp(&plan.grid, &grid_in, sizeof(tg_grid)) != 0)
		{
			plan.layer_cost = EstimateLayerCost(model->flat, grid_in);
			plan.grid = grid_in;
			plan.valid = true;
		}
		std::vector<double> cumulative(size_t(nbz) + 1, 0.0);
		for (uint32_t b = 0; b < nbz; ++b)
		{
			double sum = 0.0;
			for (uint32_t k = b * kBrick; k < std::min<uint32_t>((b + 1) * kBrick, grid.sz) && k < plan.layer_cost.size(); ++k) sum += plan.layer_cost[k];
			cumulative[size_t(b) + 1] = cumulative[b] + sum;
		}
		double share_total = 0.0;
		for (double v : shares) share_total += v;
		if (cumulative[nbz] > 0.0 && share_total > 0.0)
		{
			double wanted = 0.0;
			for (int c = 1; c < chunks; ++c)
			{
				wanted += shares[size_t(c) - 1] / share_total * cumulative[nbz];
				uint32_t b = cut[size_t(c) - 1] + 1;
				while (b < nbz && cumulative[b] < wanted) ++b;
				cut[size_t(c)] = std::min<uint32_t>(b, nbz - uint32_t(chunks - c)); // every later slab keeps at least one brick layer
			}
		}
	}
	ctx->progress_total[0] = uint64_t(chunks);
	ctx->progress_slabs = uint32_t(chunks);
	// One compute lane by default.  Two lanes (even / odd slabs on two streams with their own scratch arenas, so that a
	// slab's persistent brick kernel fills the SMs while its predecessor drains) measured no better on seaside 1024^3:
	// the small numbering / attribute kernels of slab c then queue behind the resident blocks of slab c + 1.
	int lanes = 1;
	if (const char* env = std::getenv("TG_PIPELINE_LANES")) lanes = std::atoi(env) >= 2 ? 2 : 1;
	const auto h0 = std::chrono::steady_clock::now();
	for (int c = 0; c < chunks; ++c)
	{
		if (ctx->Cancelled()) return abandon(TG_ERR_CANCELLED);
		tg_mesh_options slab = options;
		slab.flags |= TG_MESH_DEVICE_ONLY;
		slab.slab_begin = uint64_t(cut[size_t(c)]) * kBrick;
		slab.slab_end = c + 1 == chunks ? grid.sz : uint64_t(cut[size_t(c) + 1]) * kBrick;
		jobs.emplace_back(new MeshJob());
		ctx->progress_base = uint32_t(c) * 1024u;
		int rc = EnqueueMesh(*jobs.back(), model, grid_in, slab, 0, 0, index_base, error, lanes > 1 ? (c & 1) : 0, c > 0 ? jobs[size_t(c) - 1]->faces_ready : nullptr, &shared_cull);
		if (rc != TG_OK) return abandon(rc);
		if (c > 0 && (rc = collect(c - 1)) != TG_OK) return abandon(rc);
	}
	{
		const int rc = collect(chunks - 1);
		if (rc != TG_OK) return abandon(rc);
	}
	if (overflowed) return abandon(TG_RETRY_ONE_SHOT);
	TG_CUDA(cudaStreamSynchronize(copy_stream));
	TG_CUDA(cudaStreamSynchronize(lane1));
	TG_CUDA(cudaStreamSynchronize(stream));
	if (std::getenv("TG_TRACE"))
	{
		// poor man's timeline: when each stage of each slab ended, in ms after the cull started
		for (size_t c = 0; c < jobs.size(); ++c)
		{
			float t[6];
			for (int m = 0; m < 6; ++m) cudaEventElapsedTime(&t[m], cull_marks[0], jobs[c]->marks[m]);
			std::fprintf(stderr, "slab %2zu lane %d: start %.3f  resolve %.3f  eval %.3f  scan %.3f  finalize %.3f  attributes %.3f\n", c, jobs[c]->lane, t[0], t[1], t[2], t[3], t[4], t[5]);
		}
	}
	// the lanes overlap, so the device time of the export is the span from the cull's start to the last slab's end
	for (auto& j : jobs)
	{
		float span = 0.f;
		if (cudaEventElapsedTime(&span, cull_marks[0], j->marks[5]) == cudaSuccess) total.total_device_ms = std::max(total.total_device_ms, span);
	}
	{
		float cull_ms = 0.f;
		cudaEventElapsedTime(&cull_ms, cull_marks[0], cull_marks[1]);
		total.cull_ms += cull_ms;
		total.kernel_launches += cull_launches;
	}
	ctx->ReleaseDevice(shared_flags);
	cudaEventDestroy(cull_marks[0]);
	cudaEventDestroy(cull_marks[1]);
	for (auto& j : jobs)
	{
		FreeResultDevice(j->result, stream);
		delete j->result;
		j->result = nullptr;
	}

	MeshResultDevice* result = new MeshResultDevice(ctx);
	result->pinned.push_back(host.positions);
	if (host.normals) result->pinned.push_back(host.normals);
	if (host.colors) result->pinned.push_back(host.colors);
	result->pinned.push_back(host.triangles);
	out->opaque = result;
	out->positions = v_done ? host.positions : nullptr;
	out->normals = v_done ? host.normals : nullptr;
	out->colors = v_done ? host.colors : nullptr;
	out->triangles = t_done ? host.triangles : nullptr;
	out->vertex_count = v_done;
	out->triangle_count = t_done;
	out->halo_vertices = 0;
	out->timings = total;
	out->timings.download_ms = float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
	out->layer_count = nbz;
	out->layer_vertices = static_cast<uint32_t*>(std::calloc(nbz, sizeof(uint32_t)));
	out->layer_vertex_cost = static_cast<double*>(std::calloc(nbz, sizeof(double)));
	for (uint32_t i = 0; i < nbz; ++i)
	{
		if (out->layer_vertices) out->layer_vertices[i] = layer_vertices[i];
		if (out->layer_vertex_cost) out->layer_vertex_cost[i] = layer_cost[i];
	}
	auto& last = ctx->last_export;
	last.model = model;
	last.sx = grid.sx;
	last.sy = grid.sy;
	last.sz = grid.sz;
	last.flags = shape_flags;
	last.vertices = v_done;
	last.triangles = t_done;
	return TG_OK;
}

static int PipelineChunks(const tg_grid& g, const tg_mesh_options& options)
{
	if (options.flags & (TG_MESH_DEVICE_ONLY | TG_MESH_FACE_NORMALS)) return 1;
	if (options.slab_begin != 0 || options.slab_end != 0) return 1;
	if (const char* env = std::getenv("TG_PIPELINE_CHUNKS"))
	{
		const int n = std::atoi(env);
		if (n >= 1) return n;
	}
	// worth it once the result is tens of megabytes: below that the copies are short next to the launch overheads
	const double cells = double(g.sx) * double(g.sy) * double(g.sz);
	if (cells < double(1 << 24) || g.sz < 128) return 1;
	// measured on seaside_town 1024^3 (profiles/r2_pipeline_probe.txt), upload + export: one slab 9.3 ms, 2 equal-cost slabs
	// 8.1, 3: 7.7, 4: 7.5, 5: 7.7 -- every slab adds the tail of three persistent kernels and a dozen small launches, and
	// the kernels run slower while a copy is in flight, so the chain of kernels (3.9 ms alone) grows to 5.6 ms at four slabs
	// and the export ends one last-slab copy after it: few slabs, the first and the last smaller (kPipelineShares4)
	return 4;
}

int EngineExportMesh(Model* model, const tg_grid& grid_in, const tg_mesh_options& options, tg_mesh* out, std::string& error)
{
	std::memset(out, 0, sizeof(*out));
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	ctx->stage.store(1);
	ctx->progress_done[0] = 0;
	ctx->progress_total[0] = 1;
	ctx->progress_base = 0;
	ctx->progress_slabs = 1;
	if (ctx->progress_words) ctx->progress_words[0] = ctx->progress_words[1] = 0u;
	ctx->progress_seen[0] = ctx->progress_seen[1] = 0u;
	if (ctx->Cancelled()) return TG_ERR_CANCELLED;
	int rc = TG_RETRY_ONE_SHOT;
	const int chunks = PipelineChunks(grid_in, options);
	if (chunks > 1)
	{
		rc = ExportMeshPipelined(model, grid_in, options, chunks, out, error);
		if (rc == TG_RETRY_ONE_SHOT) std::memset(out, 0, sizeof(*out));
	}
	if (rc == TG_RETRY_ONE_SHOT) rc = ExportMeshOnce(model, grid_in, options, out, error);
	ctx->progress_done[0] = ctx->progress_total[0].load();
	ctx->stage.store(0);
	if (rc != TG_OK)
	{
		EngineFreeMesh(out);
		return rc;
	}
	if (ctx->Cancelled())
	{
		EngineFreeMesh(out);
		return TG_ERR_CANCELLED;
	}
	return TG_OK;
}

// Brings a TG_MESH_DEVICE_ONLY result to the host, adding index_base to every triangle index first (multi-GPU
// runs: the vertex total of the lower ranks, known only after the counts were exchanged).
int EngineMeshRankInfo(const tg_mesh* mesh, int rank, uint64_t* slab_begin, uint64_t* slab_end, tg_mesh_timings* timings)
{
	const MeshResultDevice* result = mesh ? static_cast<const MeshResultDevice*>(mesh->opaque) : nullptr;
	if (!result || rank < 0 || size_t(rank) >= result->rank_timings.size()) return -1;
	if (slab_begin) *slab_begin = result->rank_cuts[size_t(rank)];
	if (slab_end) *slab_end = result->rank_cuts[size_t(rank) + 1];
	if (timings) *timings = result->rank_timings[size_t(rank)];
	return int(result->rank_timings.size());
}

int EngineDownloadMesh(tg_mesh* mesh, uint32_t index_base, std::string& error)
{
	MeshResultDevice* result = static_cast<MeshResultDevice*>(mesh->opaque);
	if (!result)
	{
		error = "mesh has no device data";
		return TG_ERR_INVALID;
	}
	if (!result->parts.empty())
	{
		error = "tg_mesh_download does not apply to multi-GPU results: export without TG_MESH_DEVICE_ONLY";
		return TG_ERR_UNSUPPORTED;
	}
	Context* ctx = result->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	const size_t index_count = size_t(mesh->triangle_count) * 3;
	if (index_base != 0u && result->d_triangles && index_count)
	{
		RebaseIndicesKernel<<<ctx->sm_count * 8, 256, 0, stream>>>(result->d_triangles, index_count, index_base);
		TG_CUDA(cudaGetLastError());
		mesh->timings.kernel_launches++;
	}
	const auto h0 = std::chrono::steady_clock::now();
	auto fetch = [&](void* device, size_t bytes, void** host) -> int
	{
		if (!device || bytes == 0 || *host) return TG_OK;
		void* p = ctx->AcquirePinned(bytes, error);
		if (!p) return TG_ERR_MEMORY;
		result->pinned.push_back(p);
		TG_CUDA(cudaMemcpyAsync(p, device, bytes, cudaMemcpyDeviceToHost, stream));
		*host = p;
		return TG_OK;
	};
	int rc;
	if ((rc = fetch(result->d_triangles, index_count * 4, reinterpret_cast<void**>(&mesh->triangles))) != TG_OK) return rc;
	if ((rc = fetch(result->d_positions, size_t(mesh->vertex_count) * 12, reinterpret_cast<void**>(&mesh->positions))) != TG_OK) return rc;
	if ((rc = fetch(result->d_normals, size_t(mesh->vertex_count) * 12, reinterpret_cast<void**>(&mesh->normals))) != TG_OK) return rc;
	if ((rc = fetch(result->d_colors, size_t(mesh->vertex_count) * 3, reinterpret_cast<void**>(&mesh->colors))) != TG_OK) return rc;
	if ((rc = fetch(result->d_face_normals, index_count * 4, reinterpret_cast<void**>(&mesh->face_normals))) != TG_OK) return rc;
	TG_CUDA(cudaStreamSynchronize(stream));
	mesh->timings.download_ms = float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
	return TG_OK;
}

// Active 8-cell bricks per brick layer after culling: what bench.py / the multi-GPU driver balance z-slabs on.
int EngineBrickProfile(Model* model, const tg_grid& grid_in, uint32_t* out_layers, uint32_t layer_count, std::string& error)
{
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	DeviceGrid grid;
	if (!MakeDeviceGrid(grid_in, grid, error)) return TG_ERR_INVALID;
	const uint32_t nbz = (grid.sz + kBrick - 1) / kBrick;
	if (layer_count < nbz)
	{
		error = "brick profile needs ceil(sz / 8) entries";
		return TG_ERR_INVALID;
	}
	Scratch scratch(ctx);
	unsigned long long* counters = nullptr;
	uint32_t* layers = nullptr;
	TG_CUDA(scratch.Alloc(&counters, kCntCount));
	TG_CUDA(scratch.Alloc(&layers, 1024));
	TG_CUDA(cudaMemsetAsync(counters, 0, kCntCount * 8, stream));
	TG_CUDA(cudaMemsetAsync(layers, 0, 1024 * 4, stream));
	uint32_t* list = nullptr;
	uint64_t count = 0, launches = 0;
	uint64_t capacity = 0;
	const int rc = BuildActiveList(model, stream, scratch, grid, 0, grid.sz, false, false, counters, &list, &capacity, launches, error);
	if (rc != TG_OK) return rc;
	unsigned long long host_count = 0;
	TG_CUDA(cudaMemcpyAsync(&host_count, counters + kCntListA, 8, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	count = std::min<unsigned long long>(host_count, capacity);
	if (count) BrickLayerHistogramKernel<<<uint32_t((count + 255) / 256), 256, 0, stream>>>(MakeDeviceModel(model), grid, list, uint32_t(count), layers);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaMemcpyAsync(out_layers, layers, size_t(nbz) * 4, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	return TG_OK;
}

int EngineEvalLattice(Model* model, const tg_grid& grid_in, uint32_t flags, float* out, float* out_ms, std::string& error)
{
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	DeviceGrid grid;
	if (!MakeDeviceGrid(grid_in, grid, error)) return TG_ERR_INVALID;
	const uint32_t nx = grid.sx + 1, ny = grid.sy + 1, nz = grid.sz + 1;
	const uint32_t tx = (nx + kBrick - 1) / kBrick, ty = (ny + kBrick - 1) / kBrick, tz = (nz + kBrick - 1) / kBrick;
	const size_t total = size_t(nx) * ny * nz;
	Scratch scratch(ctx);
	float* d_out = nullptr;
	unsigned long long* counters = nullptr;
	TG_CUDA(scratch.Alloc(&counters, kCntCount));
	TG_CUDA(cudaMemsetAsync(counters, 0, kCntCount * 8, stream));
	if (out) TG_CUDA(scratch.Alloc(&d_out, total));
	StageTimer timer(stream);
	const int t0 = timer.Mark();
	const uint32_t tile_count = tx * ty * tz;
	const uint32_t live = (flags & TG_MESH_LIVE_FIELD) ? 1u : 0u;
	if (flags & TG_MESH_FAST)
	{
		const DeviceModel dm = MakeDeviceModel(model);
		TG_CUDA(cudaError_t(LaunchLatticeFast(&dm, &grid, d_out, tx, ty, tile_count, counters, live, stream)));
	}
	else LatticeKernel<<<(tile_count + kBrickWarps - 1) / kBrickWarps, kBrickThreads, 0, stream>>>(MakeDeviceModel(model), grid, d_out, tx, ty, tile_count, counters, live);
	const int t1 = timer.Mark();
	TG_CUDA(cudaGetLastError());
	if (out) TG_CUDA(cudaMemcpyAsync(out, d_out, total * 4, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	if (out_ms) *out_ms = timer.Ms(t0, t1);
	return TG_OK;
}

int EngineRayMarch(Model* model, const float* rays, uint64_t count, int max_iterations, float epsilon, int magnet, float* out5, std::string& error)
{
	if (count == 0) return TG_OK;
	if (count > 0x7FFFFFF0ull)
	{
		error = "too many rays for one call";
		return TG_ERR_INVALID;
	}
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	Scratch scratch(ctx);
	float *d_rays = nullptr, *d_out = nullptr;
	TG_CUDA(scratch.Alloc(&d_rays, size_t(count) * 6));
	TG_CUDA(scratch.Alloc(&d_out, size_t(count) * 5));
	TG_CUDA(cudaMemcpyAsync(d_rays, rays, size_t(count) * 24, cudaMemcpyHostToDevice, stream));
	RayMarchKernel<<<uint32_t((count + 127) / 128), 128, 0, stream>>>(MakeDeviceModel(model), d_rays, uint32_t(count), max_iterations, epsilon, magnet, d_out);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaMemcpyAsync(out5, d_out, size_t(count) * 20, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	return TG_OK;
}

int EngineWeld(Context* ctx, const float* vertices, uint64_t count, float* out_vertices4, uint32_t* out_indices, uint64_t* out_unique, std::string& error)
{
	*out_unique = 0;
	if (count == 0) return TG_OK;
	if (count > 0x3FFFFFF0ull)
	{
		error = "too many vertices for one weld";
		return TG_ERR_INVALID;
	}
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	Scratch scratch(ctx);
	uint32_t table = 1024;
	while (table < 2 * count) table <<= 1;
	float* d_vertices = nullptr;
	float4* d_out = nullptr;
	uint32_t *d_owner = nullptr, *d_smallest = nullptr, *d_slot = nullptr, *d_prefix = nullptr, *d_indices = nullptr;
	unsigned long long* d_total = nullptr;
	TG_CUDA(scratch.Alloc(&d_vertices, size_t(count) * 3));
	TG_CUDA(scratch.Alloc(&d_out, size_t(count)));
	TG_CUDA(scratch.Alloc(&d_owner, size_t(table)));
	TG_CUDA(scratch.Alloc(&d_smallest, size_t(table)));
	TG_CUDA(scratch.Alloc(&d_slot, size_t(count)));
	TG_CUDA(scratch.Alloc(&d_prefix, size_t(count)));
	TG_CUDA(scratch.Alloc(&d_indices, size_t(count)));
	TG_CUDA(scratch.Alloc(&d_total, 1));
	TG_CUDA(cudaMemcpyAsync(d_vertices, vertices, size_t(count) * 12, cudaMemcpyHostToDevice, stream));
	TG_CUDA(cudaMemsetAsync(d_owner, 0xFF, size_t(table) * 4, stream));
	TG_CUDA(cudaMemsetAsync(d_smallest, 0xFF, size_t(table) * 4, stream));
	const uint32_t n = uint32_t(count), blocks = (n + 255u) / 256u;
	WeldInsertKernel<<<blocks, 256, 0, stream>>>(d_vertices, n, d_owner, d_smallest, table - 1u, d_slot);
	TG_CUDA(cudaGetLastError());
	uint64_t launches = 0;
	const LoadWeldFirst first{ d_slot, d_smallest };
	const int rc = DeviceExclusiveScan(stream, scratch, first, size_t(count), d_prefix, d_total, launches, error);
	if (rc != TG_OK) return rc;
	WeldEmitKernel<<<blocks, 256, 0, stream>>>(d_vertices, n, d_slot, d_smallest, d_prefix, d_out, d_indices);
	TG_CUDA(cudaGetLastError());
	unsigned long long unique = 0;
	TG_CUDA(cudaMemcpyAsync(&unique, d_total, 8, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaMemcpyAsync(out_indices, d_indices, size_t(count) * 4, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	if (unique) TG_CUDA(cudaMemcpy(out_vertices4, d_out, size_t(unique) * 16, cudaMemcpyDeviceToHost));
	*out_unique = unique;
	return TG_OK;
}

int EngineCheckLongPrograms(Model* model, float reach, uint64_t out[3], std::string& error)
{
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	int debug = 0;
	if (const char* env = std::getenv("TG_LONG_DEBUG")) debug = std::atoi(env);
	TG_CUDA(cudaMemcpyToSymbolAsync(g_long_debug, &debug, sizeof(int), 0, cudaMemcpyHostToDevice, stream));
	Scratch scratch(ctx);
	unsigned long long* d = nullptr;
	TG_CUDA(scratch.Alloc(&d, 4));
	TG_CUDA(cudaMemsetAsync(d, 0, 32, stream));
	LongEvalCheckKernel<<<uint32_t(ctx->sm_count) * 4u, kLongThreads, 0, stream>>>(MakeDeviceModel(model), reach, d);
	TG_CUDA(cudaGetLastError());
	unsigned long long host[3] = { 0, 0, 0 };
	TG_CUDA(cudaMemcpyAsync(host, d, 24, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	for (int i = 0; i < 3; ++i) out[i] = host[i];
	debug = 0;
	TG_CUDA(cudaMemcpyToSymbol(g_long_debug, &debug, sizeof(int)));
	return TG_OK;
}

int EngineEvalPoints(Model* model, int mode, const float* points, uint64_t count, void* out, std::string& error)
{
	if (mode < TG_EVAL_OCTREE || mode > TG_EVAL_LIVE)
	{
		error = "unknown evaluation mode";
		return TG_ERR_INVALID;
	}
	if (count == 0) return TG_OK;
	if (count > 0xFFFFFFF0ull)
	{
		error = "too many points for one call";
		return TG_ERR_INVALID;
	}
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	Scratch scratch(ctx);
	float* d_points = nullptr;
	unsigned char* d_out = nullptr;
	const size_t out_bytes = size_t(count) * (mode == TG_EVAL_GRADIENT ? 12 : mode == TG_EVAL_COLOR ? 3 : 4);
	TG_CUDA(scratch.Alloc(&d_points, size_t(count) * 3));
	TG_CUDA(scratch.Alloc(&d_out, out_bytes));
	TG_CUDA(cudaMemcpyAsync(d_points, points, size_t(count) * 12, cudaMemcpyHostToDevice, stream));
	EvalPointsKernel<<<uint32_t((count + 127) / 128), 128, 0, stream>>>(MakeDeviceModel(model), mode, d_points, uint32_t(count), d_out);
	TG_CUDA(cudaGetLastError());
	TG_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	return TG_OK;
}

int EngineExportVoxels(Model* model, float grid_size, int32_t out_size[3], float* out_radius, int32_t** out_xyz, uint64_t* out_count, std::string& error)
{
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	// magica.cpp:29-36
	const Box3 b = model->flat.bounds;
	const Vec3 ext = b.max - b.min;
	const int sx = int(std::ceil(ext.x) * grid_size), sy = int(std::ceil(ext.y) * grid_size), sz = int(std::ceil(ext.z) * grid_size);
	if (sx <= 0 || sy <= 0 || sz <= 0 || double(sx) * sy * sz >= 2147483647.0)
	{
		error = "voxel grid is empty or exceeds the reference's int cell index (magica.cpp:38-39)";
		return TG_ERR_INVALID;
	}
	const Vec3 alpha = Vec3(.5f, .5f, .5f) / Vec3(float(sx), float(sy), float(sz));
	const float radius = Length(Mix(b.min, b.max, alpha) - b.min);
	const unsigned long long total = (unsigned long long)sx * sy * sz;
	const size_t words = size_t((total + 31) / 32);
	Scratch scratch(ctx);
	uint32_t* d_hits = nullptr;
	TG_CUDA(scratch.Alloc(&d_hits, words));
	TG_CUDA(cudaMemsetAsync(d_hits, 0, words * 4, stream));
	const unsigned long long threads = (total + 3) / 4;
	VoxelKernel<<<uint32_t((threads + 127) / 128), 128, 0, stream>>>(MakeDeviceModel(model), b.min.x, b.min.y, b.min.z, b.max.x, b.max.y, b.max.z,
		sx, sy, sz, radius, d_hits, total);
	TG_CUDA(cudaGetLastError());
	std::vector<uint32_t> hits(words);
	TG_CUDA(cudaMemcpyAsync(hits.data(), d_hits, words * 4, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	uint64_t count = 0;
	for (uint32_t w : hits) count += uint64_t(__builtin_popcount(w));
	int32_t* xyz = static_cast<int32_t*>(std::malloc((count + 1) * 12));
	if (!xyz)
	{
		error = "out of host memory";
		return TG_ERR_MEMORY;
	}
	uint64_t n = 0;
	const int slice = sx * sy;
	for (size_t w = 0; w < words; ++w)
	{
		uint32_t bits = hits[w];
		while (bits)
		{
			const int bit = __builtin_ctz(bits);
			bits &= bits - 1;
			const int i = int(w * 32 + size_t(bit));
			xyz[n * 3 + 0] = i % sx;
			xyz[n * 3 + 1] = (i % slice) / sx;
			xyz[n * 3 + 2] = i / slice;
			n++;
		}
	}
	out_size[0] = sx;
	out_size[1] = sy;
	out_size[2] = sz;
	*out_radius = radius;
	*out_xyz = xyz;
	*out_count = count;
	return TG_OK;
}

int EngineExportPoints(Model* model, const float mn[3], const float mx[3], const float step[3], int refine, uint32_t flags, float scale, tg_mesh* out, std::string& error)
{
	std::memset(out, 0, sizeof(*out));
	Context* ctx = model->context;
	TG_CUDA(cudaSetDevice(ctx->device));
	if (ctx->Cancelled()) return TG_ERR_CANCELLED;
	cudaStream_t stream = StreamOf(ctx);
	// export.cpp:394-399
	int n[3];
	for (int c = 0; c < 3; ++c)
	{
		if (!(step[c] > 0.0f))
		{
			error = "step must be positive";
			return TG_ERR_INVALID;
		}
		n[c] = int(std::ceil((mx[c] - mn[c]) / step[c]));
	}
	if (n[0] <= 0 || n[1] <= 0 || n[2] <= 0 || double(n[0]) * n[1] * n[2] >= 2147483647.0)
	{
		error = "point-cloud grid is empty or exceeds the reference's int cell index (export.cpp:396-399)";
		return TG_ERR_INVALID;
	}
	const unsigned long long total = (unsigned long long)n[0] * n[1] * n[2];
	const size_t words = size_t((total + 31) / 32);
	const Vec3 half(step[0] / 2.0f, step[1] / 2.0f, step[2] / 2.0f);
	const float diagonal = Length(half);
	Scratch scratch(ctx);
	uint64_t launches = 0;
	uint32_t *d_hits = nullptr, *d_prefix = nullptr;
	unsigned long long* counters = nullptr;
	TG_CUDA(scratch.Alloc(&d_hits, words));
	TG_CUDA(scratch.Alloc(&d_prefix, words));
	TG_CUDA(scratch.Alloc(&counters, kCntCount));
	TG_CUDA(cudaMemsetAsync(counters, 0, kCntCount * 8, stream));
	TG_CUDA(cudaMemsetAsync(d_hits, 0, words * 4, stream));
	ctx->stage.store(1);
	StageTimer timer(stream);
	const int t0 = timer.Mark();
	PointCloudKernel<<<uint32_t((total + 127) / 128), 128, 0, stream>>>(MakeDeviceModel(model), mn[0], mn[1], mn[2], step[0], step[1], step[2],
		n[0], n[1], n[2], diagonal, d_hits, total);
	launches++;
	int rc = DeviceExclusiveScan(stream, scratch, LoadPopcount32{ d_hits }, words, d_prefix, counters + kCntTotalVertices, launches, error);
	if (rc != TG_OK) return rc;
	unsigned long long host_total = 0;
	TG_CUDA(cudaMemcpyAsync(&host_total, counters + kCntTotalVertices, 8, cudaMemcpyDeviceToHost, stream));
	TG_CUDA(cudaStreamSynchronize(stream));
	const uint32_t count = uint32_t(host_total);
	MeshResultDevice* result = new MeshResultDevice(ctx);
	out->opaque = result;
	out->vertex_count = count;
	if (count > 0)
	{
		{ void* p_ = ctx->AcquireDevice(size_t(count) * 12, error); if (!p_) return TG_ERR_MEMORY; result->d_positions = static_cast<decltype(result->d_positions)>(p_); }
		GatherCloudKernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(d_hits, d_prefix, total, mn[0], mn[1], mn[2], step[0], step[1], step[2], n[0], n[1], result->d_positions);
		launches++;
		ctx->stage.store(refine > 0 ? 2 : 3);
		tg_mesh_options options;
		std::memset(&options, 0, sizeof(options));
		options.flags = flags;
		options.refine_iterations = refine;
		options.scale = scale == 0.0f ? 1.0f : scale; // WritePLY multiplies after sampling normal and colour (export.cpp:313, 476)
		const float halfv[3] = { half.x, half.y, half.z };
		AttributeScratch as;
		if (WantsAttributePass(model, options))
		{
			rc = PrepareAttributeScratch(model, stream, scratch, count, as, error);
			if (rc != TG_OK) return rc;
		}
		rc = EnqueueAttributes(model, scratch, result, counters + kCntTotalVertices, counters + kCntAttrCursor, count, options, halfv, as, false, launches, error);
		if (rc != TG_OK) return rc;
		const int t1 = timer.Mark();
		TG_CUDA(cudaStreamSynchronize(stream));
		out->timings.total_device_ms = timer.Ms(t0, t1);
		auto fetch = [&](void* device, size_t bytes, void** host) -> int
		{
			if (!device || bytes == 0) return TG_OK;
			void* p = ctx->AcquirePinned(bytes, error);
			if (!p) return TG_ERR_MEMORY;
			result->pinned.push_back(p);
			TG_CUDA(cudaMemcpyAsync(p, device, bytes, cudaMemcpyDeviceToHost, stream));
			*host = p;
			return TG_OK;
		};
		if ((rc = fetch(result->d_positions, size_t(count) * 12, reinterpret_cast<void**>(&out->positions))) != TG_OK) return rc;
		if ((rc = fetch(result->d_normals, size_t(count) * 12, reinterpret_cast<void**>(&out->normals))) != TG_OK) return rc;
		if ((rc = fetch(result->d_colors, size_t(count) * 3, reinterpret_cast<void**>(&out->colors))) != TG_OK) return rc;
		TG_CUDA(cudaStreamSynchronize(stream));
	}
	out->timings.kernel_launches = launches;
	ctx->stage.store(0);
	return TG_OK;
}

int EngineTimerBegin(Context* ctx, std::string& error)
{
	TG_CUDA(cudaSetDevice(ctx->device));
	TG_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ctx->timer_events[0]), StreamOf(ctx)));
	return TG_OK;
}

int EngineTimerEnd(Context* ctx, float* out_ms, std::string& error)
{
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaEvent_t e1 = static_cast<cudaEvent_t>(ctx->timer_events[1]);
	TG_CUDA(cudaEventRecord(e1, StreamOf(ctx)));
	TG_CUDA(cudaEventSynchronize(e1));
	TG_CUDA(cudaEventElapsedTime(out_ms, static_cast<cudaEvent_t>(ctx->timer_events[0]), e1));
	return TG_OK;
}

int EngineMeasureFp32Peak(Context* ctx, double* out_tflops, std::string& error)
{
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	Scratch scratch(ctx);
	const int blocks = ctx->sm_count * 8, threads = 256, iterations = 4096;
	float* d_out = nullptr;
	TG_CUDA(scratch.Alloc(&d_out, size_t(blocks) * threads));
	StageTimer timer(stream);
	double best = 0.0;
	for (int rep = 0; rep < 5; ++rep)
	{
		const int t0 = timer.Mark();
		FmaChainKernel<<<blocks, threads, 0, stream>>>(d_out, iterations);
		const int t1 = timer.Mark();
		TG_CUDA(cudaStreamSynchronize(stream));
		TG_CUDA(cudaGetLastError());
		const double flops = double(blocks) * threads * double(iterations) * 64.0 * 2.0;
		const double tf = flops / (double(timer.Ms(t0, t1)) * 1e-3) * 1e-12;
		if (rep > 0 && tf > best) best = tf;
	}
	*out_tflops = best;
	return TG_OK;
}

int EngineFlushL2(Context* ctx, std::string& error)
{
	TG_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t stream = StreamOf(ctx);
	Scratch scratch(ctx);
	const size_t count = size_t(256) << 18; // 256 MiB of u32 > 126 MB L2
	uint32_t* d = nullptr;
	TG_CUDA(scratch.Alloc(&d, count));
	FillKernel<<<ctx->sm_count * 8, 256, 0, stream>>>(d, count, 0u);
	TG_CUDA(cudaGetLastError());
	return TG_OK;
}

#include "tg_multi.inl"

} // namespace tg
