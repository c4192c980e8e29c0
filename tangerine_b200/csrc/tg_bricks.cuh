// K1 + K2: the brick kernel (lattice evaluation + sign classification + surface-nets vertices) and what it needs.
//
// This header is compiled twice.  tg_engine.cu holds the EXACT build (-fmad=false, IEEE sqrt / div: bit-identical to the
// reference).  tg_fast.cu holds the opt-in FAST build of the very same code (TG_MESH_FAST): FMA
// contraction, approximate sqrt / division, float instead of the reference's double promotions -- inside the north-star's
// tolerance (1e-5 relative / 4 ULP on samples) but not bit-identical.  The fast translation unit renames namespace tg
// (#define tg tg_fast) so that the two builds of every function stay apart.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tg_device.cuh"

namespace tg
{

constexpr int kBrick = 8;                 // cells per brick edge
constexpr int kTile = kBrick + 1;         // lattice samples per brick edge
constexpr int kTileSamples = kTile * kTile * kTile; // 729
constexpr int kTilePadded = 736;
constexpr int kBrickWarps = 4;            // warps per block of the brick kernels; every warp works alone
constexpr int kBrickThreads = kBrickWarps * 32;
constexpr int kMaxPending = 192;          // (node, box) pairs waiting in a warp's box resolution
constexpr int kMaxFinal = 32;             // resolved (node, box) pairs per evaluation batch (one per lane)
constexpr uint32_t kResolvedBit = 0x80000000u;
#ifndef TG_LANE_SAMPLES
#define TG_LANE_SAMPLES 2
#endif
constexpr int kLaneSamples = TG_LANE_SAMPLES; // samples interpreted per lane per dispatch

enum Counter
{
	kCntTmpVertices = 0,
	kCntSamples = 1,
	kCntFlops = 2,
	kCntListA = 3,
	kCntListB = 4,
	kCntTotalVertices = 5,
	kCntTotalQuads = 6,
	kCntHalo = 7,
	kCntBrickCursor = 8,
	kCntAttrCursor = 9,
	kCntSlots = 10,    // -DTG_COUNT_SLOTS: sample slots the interpreter dispatches covered (64 per dispatch)
	kCntTailSlots = 11, // ... of which in dispatches that held 32 samples or fewer
	kCntCount = 16
};

struct MeshParams
{
	DeviceModel model;
	DeviceGrid grid;
	const uint32_t* bricks;
	const unsigned long long* brick_count; // device: length of `bricks` (written by CullResolveKernel)
	uint32_t brick_capacity;
	unsigned long long* bitmap; // one bit per cell, rows padded to 64 cells; layer 0 = cell layer k_base
	uint32_t* word_flags;       // one bit per bitmap word: the word holds active cells (the numbering scans skip the rest unread)
	uint32_t row_words;
	uint32_t k_base;
	uint32_t k_own_begin, k_own_end;
	float4* tmp_pos;            // xyz + orientation bits
	unsigned long long* tmp_key; // bit index in the bitmap
	uint32_t tmp_capacity;
	unsigned long long* counters;
	volatile uint32_t* progress; // page-locked host word, may be null
	uint32_t progress_base;
	uint32_t live;               // TG_MESH_LIVE_FIELD: the live mesher's field (inexact descent, clamp +-100)
};

// ------------------------------------------------------------------------------------------------
// K1 + K2: evaluate one brick's 9^3 lattice tile and extract its surface-nets vertices.
// One WARP owns one brick from start to finish, so the kernel has no block-wide barrier at all: a warp that is
// waiting on an octree or instruction fetch is covered by the other resident warps, whatever phase they are in.
// ------------------------------------------------------------------------------------------------

struct WarpTile
{
	float tile[kTilePadded];          // sample values
	uint16_t order[kTilePadded];      // samples (li | lj << 4 | lk << 8) grouped by octree node; later the brick's active cell list
	uint32_t pend_node[kMaxPending];  // box resolution: (node to descend from, sample box) still to be resolved
	uint32_t pend_box[kMaxPending];
	uint32_t fin_node[kMaxFinal];     // resolved (octree node, sample box) pairs of the current evaluation batch
	uint32_t fin_box[kMaxFinal];
	uint32_t fin_start[kMaxFinal + 1]; // offset of each pair's samples in `order`
	uint16_t rows[kTile * kTile + 1];  // sign bits of the tile, one word per row of 9 samples
};

// A sample box of the tile: inclusive index ranges, four bits each.
__device__ __forceinline__ uint32_t PackBox(int x0, int x1, int y0, int y1, int z0, int z1)
{
	return uint32_t(x0) | (uint32_t(x1) << 4) | (uint32_t(y0) << 8) | (uint32_t(y1) << 12) | (uint32_t(z0) << 16) | (uint32_t(z1) << 20);
}
__device__ __forceinline__ int BoxSamples(uint32_t b)
{
	return (int((b >> 4) & 15u) - int(b & 15u) + 1) * (int((b >> 12) & 15u) - int((b >> 8) & 15u) + 1) * (int((b >> 20) & 15u) - int((b >> 16) & 15u) + 1);
}

// First index in [a, b + 1] whose lattice coordinate is > pivot (SDFOctree::Descend's strict test, :1806-1817);
// origin + float(i) * step is monotonic in i, so the samples below it take the lower octant and the rest the upper.
__device__ __forceinline__ int SplitIndex(float origin, float step, uint32_t base, int a, int b, float pivot)
{
	int i = a;
	while (i <= b && !(LatticeCoord(origin, step, base + uint32_t(i)) > pivot)) ++i;
	return i;
}

// Evaluates the lattice samples (li < ni, lj < nj, kmin <= lk < nk) of a tile whose corner sample has lattice
// index (i0, j0, k0) into w.tile.
//
// Which program a sample runs is decided by SDFOctree::Descend (sdf_evaluator.cpp:1801-1835), and the points that
// end at one node form a box.  So the tile is not descended sample by sample: the warp resolves BOXES.  A pending
// (node, box) pair follows the octree while the whole box stays in one octant; where a pivot plane cuts it, it
// splits into up to eight boxes (monotonic lattice coordinates: one split index per axis).  With leaves of >= 16
// cells an 8-cell tile is cut at most once per axis, so a brick resolves in two or three lane-parallel rounds; coarse
// grids (leaves smaller than a brick) just take more rounds.  Resolved pairs are sorted by node, their samples are
// written to `order`, and every distinct node gets ONE run of interpreter dispatches over all its samples,
// kLaneSamples per lane -- the only instantiation of the interpreter in the kernel (instruction-cache footprint).
// live: the field of the live mesher (sodapop.cpp:583-587) -- a box that ends in an empty octant is not evaluated with
// the parent's program (Exact = false, sdf_evaluator.cpp:1828-1834) but holds +infinity, and every value is clamped
// to +-100.
__device__ __forceinline__ void EvaluateTile(WarpTile& w, const DeviceModel& model, const DeviceGrid& grid,
	uint32_t i0, uint32_t j0, uint32_t k0, int ni, int nj, int nk, int kmin, unsigned long long* counters, bool live = false)
{
	const int lane = threadIdx.x & 31;
	const unsigned lanes_below = (1u << lane) - 1u;
	if (lane == 0)
	{
		w.pend_node[0] = 0u; // the octree root
		w.pend_box[0] = PackBox(0, ni - 1, 0, nj - 1, kmin, nk - 1);
	}
	__syncwarp();
	int pend = 1, fin = 0;
	for (;;)
	{
		// boxes resolved this round: as many as fit the lists (a split adds at most seven entries, a box at most one pair)
		// A nearly full list is worked depth-first, one box at a time: a box then adds at most 7 entries per octree level
		// below it, and 96 spare entries cover octrees 13 levels deep (the write below traps rather than overflow).
		const int take = min(min(32, pend), max(1, (kMaxPending - 96 - pend) / 7));
		if (fin > 0 && (pend == 0 || fin + take > kMaxFinal))
		{
			// ---- evaluation batch: sort the pairs by node, lay their samples out in `order`, run each node once ----
			uint32_t node_e = lane < fin ? w.fin_node[lane] : 0xFFFFFFFFu;
			uint32_t box_e = lane < fin ? w.fin_box[lane] : 0u;
			int rank = 0;
#pragma unroll 1
			for (int j = 0; j < fin; ++j)
			{
				const uint32_t other = __shfl_sync(0xFFFFFFFFu, node_e, j);
				rank += (other < node_e || (other == node_e && j < lane)) ? 1 : 0;
			}
			__syncwarp();
			if (lane < fin)
			{
				w.fin_node[rank] = node_e;
				w.fin_box[rank] = box_e;
			}
			__syncwarp();
			node_e = lane < fin ? w.fin_node[lane] : 0xFFFFFFFFu;
			box_e = lane < fin ? w.fin_box[lane] : 0u;
			const uint32_t size_e = lane < fin ? uint32_t(BoxSamples(box_e)) : 0u;
			uint32_t incl = size_e;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
				if (lane >= o) incl += v;
			}
			const uint32_t start_e = incl - size_e;
			if (lane < fin) w.fin_start[lane] = start_e;
			if (lane == 31) w.fin_start[fin] = incl; // the batch total closes the last run
			const uint32_t before = __shfl_up_sync(0xFFFFFFFFu, node_e, 1);
			unsigned heads = __ballot_sync(0xFFFFFFFFu, lane < fin && (lane == 0 || before != node_e));
			const uint32_t flops = (size_e && node_e != kLiveEmpty) ? size_e * __ldg(&model.nodes[node_e].flops) : 0u;
			const uint32_t flops_total = __reduce_add_sync(0xFFFFFFFFu, flops);
			if (lane == 31)
			{
				atomicAdd(&counters[kCntSamples], (unsigned long long)incl);
				atomicAdd(&counters[kCntFlops], (unsigned long long)flops_total);
			}
#pragma unroll 1
			for (int e = 0; e < fin; ++e)
			{
				const uint32_t b = __shfl_sync(0xFFFFFFFFu, box_e, e);
				const int first = int(__shfl_sync(0xFFFFFFFFu, start_e, e));
				const int x0 = int(b & 15u), y0 = int((b >> 8) & 15u), z0 = int((b >> 16) & 15u);
				const int dx = int((b >> 4) & 15u) - x0 + 1, dy = int((b >> 12) & 15u) - y0 + 1, dz = int((b >> 20) & 15u) - z0 + 1;
				const int layer = dx * dy, n = layer * dz;
				const float inv_layer = 1.0f / float(layer), inv_dx = 1.0f / float(dx);
				for (int u = lane; u < n; u += 32)
				{
					// u < 729 and the divisors are <= 81: (u + 0.5) / d is at least 0.5 / 81 away from an integer, far more than the rounding
					const int c = __float2int_rd((float(u) + 0.5f) * inv_layer);
					const int r = u - c * layer;
					const int q = __float2int_rd((float(r) + 0.5f) * inv_dx);
					w.order[first + u] = uint16_t((x0 + r - q * dx) | ((y0 + q) << 4) | ((z0 + c) << 8)); // li | lj << 4 | lk << 8
				}
			}
			__syncwarp();
			while (heads)
			{
				const int g = __ffs(heads) - 1;
				heads &= heads - 1u;
				const int first = int(w.fin_start[g]);
				const int total = int(w.fin_start[heads ? __ffs(heads) - 1 : fin]) - first;
				if (w.fin_node[g] == kLiveEmpty)
				{
					for (int u = lane; u < total; u += 32)
					{
						const uint32_t code = w.order[first + u];
						w.tile[((code >> 8) * kTile + ((code >> 4) & 15u)) * kTile + (code & 15u)] = 100.0f;
					}
					continue;
				}
				const uint4* program = model.interp + (__ldg(&model.nodes[w.fin_node[g]].interp_offset) >> 2);
				for (int done = 0; done < total; done += 32 * kLaneSamples)
				{
					const int count_here = min(total - done, 32 * kLaneSamples);
#ifdef TG_COUNT_SLOTS
					if (lane == 0)
					{
						atomicAdd(&counters[kCntSlots], (unsigned long long)(32 * kLaneSamples));
						if (count_here <= 32) atomicAdd(&counters[kCntTailSlots], (unsigned long long)(32 * kLaneSamples));
					}
#endif
					float px[kLaneSamples], py[kLaneSamples], pz[kLaneSamples], d[kLaneSamples];
					int sample[kLaneSamples];
#pragma unroll
					for (int q = 0; q < kLaneSamples; ++q)
					{
						const int idx = lane + 32 * q;
						const uint32_t code = w.order[first + done + (idx < count_here ? idx : 0)];
						const uint32_t li = code & 15u, lj = (code >> 4) & 15u, lk = code >> 8;
						sample[q] = idx < count_here ? int((lk * kTile + lj) * kTile + li) : -1;
						px[q] = LatticeCoord(grid.x, grid.dx, i0 + li);
						py[q] = LatticeCoord(grid.y, grid.dy, j0 + lj);
						pz[q] = LatticeCoord(grid.z, grid.dz, k0 + lk);
					}
					// square roots on the branch-free fast path; the (practically never taken) repeat keeps the result exact
					sdf::SqrtRange seen;
					EvalInterp<kLaneSamples, false, false, true>(program, px, py, pz, d, &seen);
					if (__any_sync(0xFFFFFFFFu, seen.Suspect())) EvalInterp<kLaneSamples>(program, px, py, pz, d);
#pragma unroll
					for (int q = 0; q < kLaneSamples; ++q)
					{
						if (sample[q] >= 0) w.tile[sample[q]] = live ? LiveClamp(d[q]) : d[q];
					}
				}
			}
			__syncwarp();
			fin = 0;
		}
		if (pend == 0) break;

		// ---- one round of box resolution: lane l takes the l-th pending pair from the top of the list ----
		const bool mine = lane < take;
		uint32_t node = 0u, box = 0u;
		if (mine)
		{
			node = w.pend_node[pend - 1 - lane];
			box = w.pend_box[pend - 1 - lane];
		}
		__syncwarp();
		pend -= take;
		bool resolved = false;
		int x0 = 0, x1 = 0, y0 = 0, y1 = 0, z0 = 0, z1 = 0, xs = 0, ys = 0, zs = 0, parts = 0;
		if (mine)
		{
			if (node & kResolvedBit)
			{
				node = live ? kLiveEmpty : node & ~kResolvedBit; // an empty octant met by a split: Descend stops at the parent (:1828-1834)
				resolved = true;
			}
			else
			{
				x0 = int(box & 15u), x1 = int((box >> 4) & 15u), y0 = int((box >> 8) & 15u), y1 = int((box >> 12) & 15u), z0 = int((box >> 16) & 15u), z1 = int((box >> 20) & 15u);
				const float lox = LatticeCoord(grid.x, grid.dx, i0 + x0), loy = LatticeCoord(grid.y, grid.dy, j0 + y0), loz = LatticeCoord(grid.z, grid.dz, k0 + z0);
				const float hix = LatticeCoord(grid.x, grid.dx, i0 + x1), hiy = LatticeCoord(grid.y, grid.dy, j0 + y1), hiz = LatticeCoord(grid.z, grid.dz, k0 + z1);
				for (;;)
				{
					const float4 head = __ldg(reinterpret_cast<const float4*>(&model.nodes[node]));
					if (__float_as_uint(head.w) != 0u)
					{
						resolved = true;
						break;
					}
					const int olo = (lox > head.x ? 1 : 0) | (loy > head.y ? 2 : 0) | (loz > head.z ? 4 : 0);
					const int ohi = (hix > head.x ? 1 : 0) | (hiy > head.y ? 2 : 0) | (hiz > head.z ? 4 : 0);
					if (olo != ohi)
					{
						// a pivot plane cuts the box: lower part [a, split - 1], upper part [split, b] on every axis (either may be empty)
						xs = ((olo ^ ohi) & 1) ? SplitIndex(grid.x, grid.dx, i0, x0, x1, head.x) : ((olo & 1) ? x0 : x1 + 1);
						ys = ((olo ^ ohi) & 2) ? SplitIndex(grid.y, grid.dy, j0, y0, y1, head.y) : ((olo & 2) ? y0 : y1 + 1);
						zs = ((olo ^ ohi) & 4) ? SplitIndex(grid.z, grid.dz, k0, z0, z1, head.z) : ((olo & 4) ? z0 : z1 + 1);
						parts = ((xs > x0 ? 1 : 0) + (xs <= x1 ? 1 : 0)) * ((ys > y0 ? 1 : 0) + (ys <= y1 ? 1 : 0)) * ((zs > z0 ? 1 : 0) + (zs <= z1 ? 1 : 0));
						break;
					}
					const int32_t child = __ldg(&model.nodes[node].children[olo]);
					if (child < 0)
					{
						if (live) node = kLiveEmpty;
						resolved = true;
						break;
					}
					node = uint32_t(child);
				}
			}
		}
		const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, resolved);
		if (resolved)
		{
			const int slot = fin + __popc(done_mask & lanes_below);
			w.fin_node[slot] = node;
			w.fin_box[slot] = box;
		}
		fin += __popc(done_mask);
		int incl = parts;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= o) incl += v;
		}
		const int added = __shfl_sync(0xFFFFFFFFu, incl, 31);
		if (parts)
		{
			int at = pend + incl - parts;
#pragma unroll 1
			for (int o = 0; o < 8; ++o)
			{
				const int ax = (o & 1) ? xs : x0, bx = (o & 1) ? x1 : xs - 1;
				const int ay = (o & 2) ? ys : y0, by = (o & 2) ? y1 : ys - 1;
				const int az = (o & 4) ? zs : z0, bz = (o & 4) ? z1 : zs - 1;
				if (ax > bx || ay > by || az > bz) continue;
				const int32_t child = __ldg(&model.nodes[node].children[o]);
				if (at >= kMaxPending) __trap();
				w.pend_node[at] = child < 0 ? (node | kResolvedBit) : uint32_t(child);
				w.pend_box[at] = PackBox(ax, bx, ay, by, az, bz);
				++at;
			}
		}
		pend += added;
		__syncwarp();
	}
}

// Loads the eight corner samples of cell c (0..511) of the brick in the reference's corner numbering
// (get_voxel_corner_grid_positions, surface_nets.cpp:632-646) and returns the `>= 0` mask.
__device__ __forceinline__ unsigned CellCorners(const WarpTile& w, int c, float (&v)[8])
{
	const int ci = c & 7, cj = (c >> 3) & 7, ck = c >> 6;
	const float* t = w.tile + (ck * kTile + cj) * kTile + ci;
	v[0] = t[0];
	v[1] = t[1];
	v[2] = t[kTile + 1];
	v[3] = t[kTile];
	v[4] = t[kTile * kTile];
	v[5] = t[kTile * kTile + 1];
	v[6] = t[kTile * kTile + kTile + 1];
	v[7] = t[kTile * kTile + kTile];
	// is_scalar_positive is `scalar >= isovalue` (:733-735): -0.0 is positive, NaN is negative
	unsigned signs = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) signs |= (v[k] >= 0.0f ? 1u : 0u) << k;
	return signs;
}

#ifndef TG_BRICK_MIN_BLOCKS
#define TG_BRICK_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(kBrickThreads, TG_BRICK_MIN_BLOCKS) MeshBricksKernel(const MeshParams p)
{
	__shared__ WarpTile tiles[kBrickWarps];
	WarpTile& w = tiles[threadIdx.x >> 5];
	const DeviceGrid& grid = p.grid;
	const int lane = threadIdx.x & 31;

	const float bbminx = grid.x, bbminy = grid.y, bbminz = grid.z;
	const float bbmaxx = __fadd_rn(grid.x, __fmul_rn(float(grid.sx), grid.dx));
	const float bbmaxy = __fadd_rn(grid.y, __fmul_rn(float(grid.sy), grid.dy));
	const float bbmaxz = __fadd_rn(grid.z, __fmul_rn(float(grid.sz), grid.dz));
	unsigned char* bitmap_bytes = reinterpret_cast<unsigned char*>(p.bitmap);
	const uint32_t brick_count = uint32_t(min(*p.brick_count, (unsigned long long)p.brick_capacity));

	for (;;)
	{
		// persistent warps pull bricks from one device-wide cursor
		uint32_t item = 0;
		if (lane == 0) item = uint32_t(atomicAdd(&p.counters[kCntBrickCursor], 1ull));
		item = __shfl_sync(0xFFFFFFFFu, item, 0);
		if (item >= brick_count) break;
		if (p.progress && lane == 0 && (item & 63u) == 0u) *p.progress = p.progress_base + uint32_t((unsigned long long)item * 1023ull / brick_count);

		const uint32_t brick = __ldg(&p.bricks[item]);
		const uint32_t bx = brick & 1023u, by = (brick >> 10) & 1023u, bz = (brick >> 20) & 1023u;
		const uint32_t i0 = bx * kBrick, j0 = by * kBrick, k0 = bz * kBrick;
		const int ni = int(min(uint32_t(kTile), grid.sx + 1 - i0));
		const int nj = int(min(uint32_t(kTile), grid.sy + 1 - j0));
		// The slab owns cell layers [k_own_begin, k_own_end) and also classifies the halo layer k_base below it; a
		// brick that sticks out of that window only evaluates the sample layers the window needs.
		const int nk = int(min(uint32_t(kTile), p.k_own_end + 1 - k0));
		const int kmin = p.k_base > k0 ? int(p.k_base - k0) : 0;

		EvaluateTile(w, p.model, grid, i0, j0, k0, ni, nj, nk, kmin, p.counters, p.live != 0u);

		// Classification: sign bits of FirstLoopInnerThunk (surface_nets.cpp:864-907).  is_scalar_positive is
		// `scalar >= isovalue` (:733-735: -0.0 is positive, NaN is negative), and a cell is active when its eight corners
		// do not agree (:903-907).  The signs of a sample row (9 samples along x) are packed into one word first; a lane
		// then classifies a whole row of 8 cells with a dozen bit operations on the four sample rows around it, and the
		// result IS the row's byte of the active-cell bitmap.
		for (int r = lane; r < kTile * kTile; r += 32)
		{
			const float* t = w.tile + r * kTile;
			uint32_t bits = 0;
#pragma unroll
			for (int i = 0; i < kTile; ++i) bits |= (t[i] >= 0.0f ? 1u : 0u) << i;
			w.rows[r] = uint16_t(bits);
		}
		__syncwarp();
		int emit_total = 0;
		const uint32_t cells_x = min(uint32_t(kBrick), grid.sx - i0);
#pragma unroll 1
		for (int round = 0; round < 2; ++round)
		{
			const int cr = round * 32 + lane; // cell row: cj = cr & 7, ck = cr >> 3
			const int cj = cr & 7, ck = cr >> 3;
			const uint32_t gj = j0 + cj, gk = k0 + ck;
			const uint32_t r00 = w.rows[ck * kTile + cj], r01 = w.rows[ck * kTile + cj + 1];
			const uint32_t r10 = w.rows[(ck + 1) * kTile + cj], r11 = w.rows[(ck + 1) * kTile + cj + 1];
			const uint32_t all = r00 & r01 & r10 & r11, any = r00 | r01 | r10 | r11;
			uint32_t active = ~((all & (all >> 1)) | ~(any | (any >> 1))) & ((1u << cells_x) - 1u);
			// the slab owns cell layers [k_own_begin, k_own_end) and classifies the halo layer k_base below it
			if (!(gj < grid.sy && gk < p.k_own_end && ck >= kmin)) active = 0u;
			if (active)
			{
				const size_t word = (size_t(gk - p.k_base) * grid.sy + gj) * size_t(p.row_words) + (bx >> 3);
				bitmap_bytes[word * 8u + (bx & 7u)] = (unsigned char)active;
				atomicOr(&p.word_flags[word >> 5], 1u << (word & 31u));
			}
			uint32_t emit = gk >= p.k_own_begin ? active : 0u; // the halo layer is classified but owned by the slab below
			const int count = __popc(emit);
			int incl = count;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
				if (lane >= o) incl += v;
			}
			int at = emit_total + incl - count;
			while (emit)
			{
				w.order[at++] = uint16_t(cr * kBrick + __ffs(emit) - 1);
				emit &= emit - 1u;
			}
			emit_total += __shfl_sync(0xFFFFFFFFu, incl, 31);
		}
		__syncwarp();
		if (emit_total == 0) continue;

		uint32_t out_base = 0;
		if (lane == 0) out_base = uint32_t(atomicAdd(&p.counters[kCntTmpVertices], (unsigned long long)emit_total));
		out_base = __shfl_sync(0xFFFFFFFFu, out_base, 0);

		// Vertices of the active cells, densely packed over the lanes: :920-965
		for (int e = lane; e < emit_total; e += 32)
		{
			const int c = w.order[e];
			const int ci = c & 7, cj = (c >> 3) & 7, ck = c >> 6;
			const uint32_t gi = i0 + ci, gj = j0 + cj, gk = k0 + ck;
			float v[8];
			CellCorners(w, c, v);
			const float fi = float(gi), fj = float(gj), fk = float(gk);
			const float gx[8] = { fi, fi + 1.f, fi + 1.f, fi, fi, fi + 1.f, fi + 1.f, fi };
			const float gy[8] = { fj, fj, fj + 1.f, fj + 1.f, fj, fj, fj + 1.f, fj + 1.f };
			const float gz[8] = { fk, fk, fk, fk, fk + 1.f, fk + 1.f, fk + 1.f, fk + 1.f };
			const int e0[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3 }; // edge table :889-901
			const int e1[12] = { 1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7 };
			float sx = 0.f, sy = 0.f, sz = 0.f;
			int n = 0;
#pragma unroll
			for (int ed = 0; ed < 12; ++ed)
			{
				const float s1 = v[e0[ed]], s2 = v[e1[ed]];
				if ((s1 >= 0.0f) != (s2 >= 0.0f))
				{
					const float t = (0.0f - s1) / (s2 - s1); // :937
					sx = sx + (gx[e0[ed]] + t * (gx[e1[ed]] - gx[e0[ed]]));
					sy = sy + (gy[e0[ed]] + t * (gy[e1[ed]] - gy[e0[ed]]));
					sz = sz + (gz[e0[ed]] + t * (gz[e1[ed]] - gz[e0[ed]]));
					n++;
				}
			}
			const float count = float(n);
			const float cx = sx / count, cy = sy / count, cz = sz / count;
			// :952-965  min + (max - min) * (centre - 0) / (size - 0)
			const float px = bbminx + (bbmaxx - bbminx) * (cx - 0.f) / (float(grid.sx) - 0.f);
			const float py = bbminy + (bbmaxy - bbminy) * (cy - 0.f) / (float(grid.sy) - 0.f);
			const float pz = bbminz + (bbmaxz - bbminz) * (cz - 0.f) / (float(grid.sz) - 0.f);
			// winding bits for SecondLoopThunk (:1041-1067, :1103-1105): edge (0,4), (3,0), (0,1)
			const uint32_t orient = (v[4] > v[0] ? 1u : 0u) | (v[0] > v[3] ? 2u : 0u) | (v[1] > v[0] ? 4u : 0u);
			const uint32_t dst = out_base + uint32_t(e);
			if (dst < p.tmp_capacity)
			{
				p.tmp_pos[dst] = make_float4(px, py, pz, __uint_as_float(orient));
				p.tmp_key[dst] = ((unsigned long long)(gk - p.k_base) * grid.sy + gj) * ((unsigned long long)p.row_words * 64ull) + gi;
			}
		}
		__syncwarp();
	}
}

// Dense lattice dump: one 8^3 tile of samples per warp, written to a (sz+1, sy+1, sx+1) array.
__global__ void __launch_bounds__(kBrickThreads) LatticeKernel(const DeviceModel model, const DeviceGrid grid, float* __restrict__ out,
	uint32_t tiles_x, uint32_t tiles_y, uint32_t tile_count, unsigned long long* counters, uint32_t live)
{
	__shared__ WarpTile tiles[kBrickWarps];
	WarpTile& w = tiles[threadIdx.x >> 5];
	const int lane = threadIdx.x & 31;
	const uint32_t tile = blockIdx.x * kBrickWarps + (threadIdx.x >> 5);
	if (tile >= tile_count) return;
	const uint32_t tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, tz = tile / (tiles_x * tiles_y);
	const uint32_t i0 = tx * kBrick, j0 = ty * kBrick, k0 = tz * kBrick;
	const uint32_t nx = grid.sx + 1, ny = grid.sy + 1, nz = grid.sz + 1;
	const int ni = int(min(uint32_t(kBrick), nx - i0)), nj = int(min(uint32_t(kBrick), ny - j0)), nk = int(min(uint32_t(kBrick), nz - k0));
	EvaluateTile(w, model, grid, i0, j0, k0, ni, nj, nk, 0, counters, live != 0u);
	if (out == nullptr) return;
	for (int s = lane; s < kBrick * kBrick * kBrick; s += 32)
	{
		const int li = s & 7, lj = (s >> 3) & 7, lk = s >> 6;
		if (li < ni && lj < nj && lk < nk)
		{
			out[(size_t(k0 + lk) * ny + (j0 + lj)) * nx + (i0 + li)] = w.tile[(lk * kTile + lj) * kTile + li];
		}
	}
}

} // namespace tg
