// Device-side SDF interpreter and octree descent (sm_100a).
//
// RunProgram executes one flat program (tg_program.h) for S sample points at once per thread: the
// operand "stack" is an accumulator register per sample plus statically numbered spill slots that only
// right-nested operands touch, so the common left-leaning CSG chain runs entirely in registers with one
// header decode per primitive, amortised over S samples of instruction-level parallelism.  When every
// lane of a warp was handed the same program (the brick kernels arrange that) the header and parameter
// loads are warp-uniform broadcasts and the opcode switch does not diverge.
//
// Arithmetic contract: this file is compiled with -fmad=false and IEEE sqrt/div so that every distance is
// bit-identical to the reference's CPU evaluation (see tg_sdf.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tg_program.h"
#include "tg_sdf.h"

namespace tg
{

struct DeviceModel
{
	const FlatNode* nodes;
	const uint4* interp;       // kStreamInterp, 16-byte quads; FlatNode::interp_offset / 4 indexes it
	const uint32_t* tree;
	const FlatRegion* regions;
	const uint32_t* node_rank; // node -> position by descending program cost (key of the attribute pass's counting sort)
	const uint32_t* node_material; // node -> the one material its program can return, or kMixedMaterial (FlatModel::node_material)
	uint32_t region_count;
	const float* material_rgb; // 3 floats per id
	uint32_t material_count;   // index of the trailing default-white entry
	uint32_t root_interp_offset;
	uint32_t root_tree_offset;
	uint32_t node_count;
	int has_paint;
};

struct DeviceGrid
{
	float x, y, z, dx, dy, dz;
	uint32_t sx, sy, sz; // cells
};

// get_voxel_corner_world_positions (surface_nets.cpp:648-687): origin + float(index) * step, no FMA
__device__ __forceinline__ float LatticeCoord(float origin, float step, uint32_t index)
{
	return __fadd_rn(origin, __fmul_rn(float(index), step));
}

// SDFOctree::Descend(Point, Exact = true) (sdf_evaluator.cpp:1801-1835), iteratively from `start`.
__device__ __forceinline__ uint32_t Descend(const FlatNode* __restrict__ nodes, uint32_t start, float px, float py, float pz)
{
	uint32_t n = start;
	for (;;)
	{
		// One 64-byte node = pivot + terminus and the eight child indices: three independent 128-bit loads, one
		// round trip per level instead of two dependent ones.
		const float4* q = reinterpret_cast<const float4*>(&nodes[n]);
		const float4 head = __ldg(q); // pivot.xyz, terminus bits
		const int4 lo = __ldg(reinterpret_cast<const int4*>(q + 1));
		const int4 hi = __ldg(reinterpret_cast<const int4*>(q + 2));
		if (__float_as_uint(head.w) != 0u)
		{
			return n;
		}
		const int4 zsel = pz > head.z ? hi : lo;
		const int c0 = py > head.y ? zsel.z : zsel.x;
		const int c1 = py > head.y ? zsel.w : zsel.y;
		const int32_t child = px > head.x ? c1 : c0;
		if (child < 0)
		{
			return n; // empty octant: the reference evaluates this (larger) node's program
		}
		n = uint32_t(child);
	}
}

// Returns v through an empty asm statement: the compiler can no longer tell that two comparisons test the same value.
__device__ __forceinline__ uint32_t Opaque(uint32_t& v)
{
	asm volatile("" : "+r"(v));
	return v;
}

__device__ __forceinline__ float4 AsFloat4(const uint4& v)
{
	return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

// kStreamInterp interpreter (tg_program.h): SDFInterpreter::Eval (sdf_evaluator.cpp:1386-1605) for S points per lane.
// Every fetch is a 128-bit load whose address does not depend on another fetch of the same instruction, and the
// first quad of the next instruction is requested before this instruction's arithmetic starts.  The operator
// switch sits outside the per-sample loops, so with a warp-uniform program there is one dispatch per brush.
// DEFER: square roots take the branch-free fast path (sdf::SqrtDeferred) and *seen (sdf::SqrtRange) records whether an argument fell
// outside its range -- the caller then evaluates again without DEFER.
// SDFOctree::Descend(Point, Exact = false) (sdf_evaluator.cpp:1801-1835): an empty octant ends the search with nothing
// found -- kLiveEmpty -- instead of falling back on the parent's program.
constexpr uint32_t kLiveEmpty = 0x7FFFFFFEu;
__device__ __forceinline__ uint32_t DescendLive(const FlatNode* __restrict__ nodes, float px, float py, float pz)
{
	uint32_t n = 0;
	for (;;)
	{
		const float4* q = reinterpret_cast<const float4*>(&nodes[n]);
		const float4 head = __ldg(q);
		const int4 lo = __ldg(reinterpret_cast<const int4*>(q + 1));
		const int4 hi = __ldg(reinterpret_cast<const int4*>(q + 2));
		if (__float_as_uint(head.w) != 0u) return n;
		const int4 zsel = pz > head.z ? hi : lo;
		const int c0 = py > head.y ? zsel.z : zsel.x;
		const int c1 = py > head.y ? zsel.w : zsel.y;
		const int32_t child = px > head.x ? c1 : c0;
		if (child < 0) return kLiveEmpty;
		n = uint32_t(child);
	}
}

// glm::clamp(x, -100.0f, 100.0f) = min(max(x, lo), hi) with glm's comparisons (func_common.inl): NaN passes through.
// The live mesher's implicit function wraps SDFOctree::Eval(Point, false) in it (sodapop.cpp:583-587); nothing found is
// +infinity there (sdf_evaluator.cpp:1978-1981), i.e. 100 after the clamp.
__device__ __forceinline__ float LiveClamp(float x)
{
	const float t = (x < -100.0f) ? -100.0f : x;
	return (100.0f < t) ? 100.0f : t;
}

template <bool DEFER> struct SqrtPolicy
{
	using type = sdf::SqrtExact;
	static __device__ __forceinline__ type Make(sdf::SqrtRange*) { return type(); }
};
template <> struct SqrtPolicy<true>
{
	using type = sdf::SqrtDeferred;
	static __device__ __forceinline__ type Make(sdf::SqrtRange* seen) { return type{ seen }; }
};

template <int S, bool PREFETCH = false, bool DEEP = false, bool DEFER = false>
__device__ __forceinline__ void EvalInterp(const uint4* __restrict__ pc, const float (&px)[S], const float (&py)[S], const float (&pz)[S], float (&result)[S],
	sdf::SqrtRange* seen = nullptr)
{
	const typename SqrtPolicy<DEFER>::type sq = SqrtPolicy<DEFER>::Make(seen);
	float acc[S];
	float stack[kMaxStackSlots][S];
#pragma unroll
	for (int s = 0; s < S; ++s) acc[s] = 0.0f;
	uint4 q = __ldg(pc);
	// DEEP (K0: one thread walks a long program with next to nothing else resident to hide its latency): the four quads
	// that can follow a header are requested a whole instruction ahead, whether or not the instruction turns out to
	// use them (programs are contiguous and the stream ends in padding, so reading past an instruction is harmless).
	uint4 m0 = q, m1 = q, m2 = q, m3 = q;
	if (DEEP)
	{
		m0 = __ldg(pc + 1);
		m1 = __ldg(pc + 2);
		m2 = __ldg(pc + 3);
		m3 = __ldg(pc + 4);
	}
	for (;;)
	{
		const uint32_t header = q.x;
		// fields are compared in place (mask, no shift): the decode runs once per primitive per 64 samples
		const uint32_t brush = header & kHdrBrushMask;
		const uint32_t op = header & (0xFu << kHdrOpShift);
		const uint4* next = pc + (header >> kHdrLenShift);
		uint4 nq = q, n0 = q, n1 = q, n2 = q, n3 = q;
		if (DEEP)
		{
			nq = __ldg(next);
			n0 = __ldg(next + 1);
			n1 = __ldg(next + 2);
			n2 = __ldg(next + 3);
			n3 = __ldg(next + 4);
		}
		if (PREFETCH) asm volatile("prefetch.global.L1 [%0];" ::"l"(pc + 32)); // 512 B ahead: long programs stream from L2 / HBM
		if (brush != kBrushNone)
		{
			const float p0 = __uint_as_float(q.y), p1 = __uint_as_float(q.z), p2 = __uint_as_float(q.w);
			const uint32_t xform = header & (3u << kHdrXformShift);
			float lx[S], ly[S], lz[S];
			float scale = 1.0f, threshold = 0.0f;
			if (xform == (kXformMatrix << kHdrXformShift))
			{
				const float4 a = DEEP ? AsFloat4(m0) : __ldg(reinterpret_cast<const float4*>(pc + 1));
				const float4 b = DEEP ? AsFloat4(m1) : __ldg(reinterpret_cast<const float4*>(pc + 2));
				const float4 c = DEEP ? AsFloat4(m2) : __ldg(reinterpret_cast<const float4*>(pc + 3));
				if (header & kHdrTailBit)
				{
					const float4 t = DEEP ? AsFloat4(m3) : __ldg(reinterpret_cast<const float4*>(pc + 4));
					scale = t.x;
					threshold = t.y;
				}
				if (!DEEP) q = __ldg(next);
				// glm mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*1)  (type_mat4x4.inl:561-572)
				// columns: m0 = (a.x a.y a.z), m1 = (a.w b.x b.y), m2 = (b.z b.w c.x), m3 = (c.y c.z c.w)
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = (a.x * px[s] + a.w * py[s]) + (b.z * pz[s] + c.y);
					ly[s] = (a.y * px[s] + b.x * py[s]) + (b.w * pz[s] + c.z);
					lz[s] = (a.z * px[s] + b.y * py[s]) + (c.x * pz[s] + c.w);
				}
			}
			else if (xform == (kXformOffset << kHdrXformShift))
			{
				const float4 o = DEEP ? AsFloat4(m0) : __ldg(reinterpret_cast<const float4*>(pc + 1));
				if (header & kHdrTailBit)
				{
					const float4 t = DEEP ? AsFloat4(m1) : __ldg(reinterpret_cast<const float4*>(pc + 2));
					scale = t.x;
					threshold = t.y;
				}
				if (!DEEP) q = __ldg(next);
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = px[s] + o.x;
					ly[s] = py[s] + o.y;
					lz[s] = pz[s] + o.z;
				}
			}
			else
			{
				if (header & kHdrTailBit)
				{
					const float4 t = DEEP ? AsFloat4(m0) : __ldg(reinterpret_cast<const float4*>(pc + 1));
					scale = t.x;
					threshold = t.y;
				}
				if (!DEEP) q = __ldg(next);
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = px[s];
					ly[s] = py[s];
					lz[s] = pz[s];
				}
			}
			pc = next;

			// Dispatch by comparison chains, commonest first: a jump table costs nine instructions per dispatch (clamp,
			// scale, constant-bank load, BRX), a taken comparison two, and these run once per primitive per 64 samples.
			float d[S];
			uint32_t kind = brush, oper = op; // laundered between comparisons: keeps the chain from being turned back into a table
			if (kind == kBrushBox)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Box(lx[s], ly[s], lz[s], p0, p1, p2, sq);
			}
			else if (Opaque(kind) == kBrushSphere)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Sphere(lx[s], ly[s], lz[s], p0, sq);
			}
			else if (Opaque(kind) == kBrushCylinder)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cylinder(lx[s], ly[s], lz[s], p0, p1, sq);
			}
			else if (Opaque(kind) == kBrushTorus)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Torus(lx[s], ly[s], lz[s], p0, p1, sq);
			}
			else if (Opaque(kind) == kBrushPlane)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Plane(lx[s], ly[s], lz[s], p0, p1, p2);
			}
			else if (Opaque(kind) == kBrushEllipsoid)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Ellipsoid(lx[s], ly[s], lz[s], p0, p1, p2, sq);
			}
			else if (Opaque(kind) == kBrushCone)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cone(lx[s], ly[s], lz[s], p0, p1, sq);
			}
			else
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Coninder(lx[s], ly[s], lz[s], p0, p1, p2, sq);
			}
			if (header & kHdrScaleBit)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = d[s] * scale; // ScaleField (:1587-1591)
			}
			if (oper == (kOpPush << kHdrOpShift))
			{
				if ((header & (0xFFu << kHdrSlotShift)) != (kNoSlot << kHdrSlotShift))
				{
					const uint32_t slot = (header >> kHdrSlotShift) & 0xFFu;
#pragma unroll
					for (int s = 0; s < S; ++s) stack[slot][s] = acc[s];
				}
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = d[s];
			}
			else if (Opaque(oper) == (kOpUnion << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::Union(acc[s], d[s]);
			}
			else if (Opaque(oper) == (kOpBlendUnion << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::BlendUnion(acc[s], d[s], threshold);
			}
			else if (Opaque(oper) == (kOpDiff << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::Diff(acc[s], d[s]);
			}
			else if (Opaque(oper) == (kOpInter << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::Inter(acc[s], d[s]);
			}
			else if (Opaque(oper) == (kOpBlendDiff << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::BlendDiff(acc[s], d[s], threshold);
			}
			else
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = sdf::BlendInter(acc[s], d[s], threshold);
			}
		}
		else if (op == (kOpStop << kHdrOpShift))
		{
#pragma unroll
			for (int s = 0; s < S; ++s) result[s] = acc[s];
			return;
		}
		else
		{
			const float param = __uint_as_float(q.y);
			if (!DEEP) q = __ldg(next);
			pc = next;
			if (op == (kOpFlate << kHdrOpShift))
			{
#pragma unroll
				for (int s = 0; s < S; ++s) acc[s] = acc[s] - param; // :1567-1571
			}
			else
			{
				// stack-form set operator: lhs was spilled, rhs is the accumulator
				const uint32_t slot = (header >> kHdrSlotShift) & 0xFFu;
				float lhs[S];
#pragma unroll
				for (int s = 0; s < S; ++s) lhs[s] = stack[slot][s];
				uint32_t oper = op;
				if (oper == (kOpUnion << kHdrOpShift))
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::Union(lhs[s], acc[s]);
				}
				else if (Opaque(oper) == (kOpDiff << kHdrOpShift))
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::Diff(lhs[s], acc[s]);
				}
				else if (Opaque(oper) == (kOpInter << kHdrOpShift))
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::Inter(lhs[s], acc[s]);
				}
				else if (Opaque(oper) == (kOpBlendUnion << kHdrOpShift))
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::BlendUnion(lhs[s], acc[s], param);
				}
				else if (Opaque(oper) == (kOpBlendInter << kHdrOpShift))
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::BlendInter(lhs[s], acc[s], param);
				}
				else
				{
#pragma unroll
					for (int s = 0; s < S; ++s) acc[s] = sdf::BlendDiff(lhs[s], acc[s], param);
				}
			}
		}
		if (DEEP)
		{
			q = nq;
			m0 = n0;
			m1 = n1;
			m2 = n2;
			m3 = n3;
		}
	}
}

template <bool DEEP = false>
__device__ __forceinline__ float EvalInterp1(const DeviceModel& model, uint32_t word_offset, float x, float y, float z)
{
	const float px[1] = { x }, py[1] = { y }, pz[1] = { z };
	float out[1];
	EvalInterp<1, true, DEEP>(model.interp + (word_offset >> 2), px, py, pz, out);
	return out[0];
}

// A GROUP of threads (a warp or a whole block) evaluates ONE point of a long program (FlatNode::flags & kNodeLong)
// cooperatively.  The program is preceded by the table of its instructions' quad offsets (tg_octree.cpp), so every thread
// fetches its own instructions directly -- no dependent fetch chain -- and the brush distances are computed in parallel
// (an instruction's operand quads sit at fixed offsets behind its header; reading past a short instruction is harmless:
// programs are contiguous and the stream ends in padding).  What is left is the operator chain over (code, value)
// steps parked in shared memory, and it is folded in parallel too:
//   1. every right-nested operand of the main chain -- [brush pushed with spill slot 0 ... stack-form operator with
//      slot 0] -- is folded by the thread that owns its first step, privately, and collapses into ONE step of the main
//      chain (min / max / blend with the operand's value);
//   2. the main chain is then a sequence of x -> min(x, v) / max(x, v) steps, i.e. of clamps x -> min(max(x, lo), hi),
//      and clamps compose associatively: every thread composes its contiguous share, thread 0 composes the shares in
//      order.  min and max select one of their arguments, so the result is the sequential value bit for bit.
// Blend or flate steps on the main chain (rare in long programs) fall back to one thread walking the collapsed chain.
// Same arithmetic as EvalInterp everywhere.
constexpr int kLongThreads = 256;     // block size of the kernel that evaluates long programs
constexpr int kLongWarpSteps = 256;   // a warp handles programs up to this many instructions on its own ...
constexpr int kLongBlockSteps = 2048; // ... the whole block the ones up to this many; beyond: sequential rounds

struct LongScratch
{
	uint2 step[kLongBlockSteps];  // x: fold code | slot << 8 | operator << 16, y: value (brush distance, negated for Diff; or operator parameter)
	float param[kLongBlockSteps]; // blend threshold of fused instructions
	float2 share[kLongThreads];   // clamp (lo, hi) of every thread's part of a chain; also the operand lists
	uint32_t marks[kLongThreads]; // scratch of the group-wide prefix and of FoldChain (3 words per warp)
	float result[kLongThreads / 32];
};

enum : uint32_t
{
	kFoldMin = 0, kFoldMax = 1, kFoldNop = 2, kFoldPush = 3, kFoldBlendUnion = 4, kFoldBlendInter = 5, kFoldBlendDiff = 6,
	kFoldFlate = 7, kFoldStackOp = 8
};

// One step of the chain applied to (acc, stack).
__device__ __forceinline__ void FoldStep(uint32_t word, float v, float param, float& acc, float (&stack)[kMaxStackSlots])
{
	const uint32_t code = word & 0xFFu;
	if (code == kFoldMin) acc = fminf(acc, v);
	else if (code == kFoldMax) acc = fmaxf(acc, v);
	else if (code == kFoldPush)
	{
		const uint32_t slot = (word >> 8) & 0xFFu;
		if (slot != kNoSlot) stack[slot] = acc;
		acc = v;
	}
	else if (code >= kFoldBlendUnion && code <= kFoldBlendDiff) acc = sdf::SetOp(kOpBlendUnion + (code - kFoldBlendUnion), acc, v, param);
	else if (code == kFoldFlate) acc = acc - v;
	else if (code == kFoldStackOp) acc = sdf::SetOp(word >> 16, stack[(word >> 8) & 0xFFu], acc, v);
}

// Brush distance and fold code of one instruction (header quad q, the four quads behind it) at (x, y, z).
__device__ __forceinline__ void LongStepLoaded(const uint4& q, const float4& m0, const float4& m1, const float4& m2, const float4& m3,
	float x, float y, float z, uint2& step, float& param)
{
	const uint32_t header = q.x;
	const uint32_t brush = header & kHdrBrushMask;
	const uint32_t op = (header >> kHdrOpShift) & 0xFu;
	const uint32_t slot = (header >> kHdrSlotShift) & 0xFFu;
	uint32_t code;
	float value;
	param = 0.0f;
	if (brush != kBrushNone)
	{
		const uint32_t xform = (header >> kHdrXformShift) & 3u;
		float lx = x, ly = y, lz = z;
		float4 tail = m0;
		if (xform == kXformMatrix)
		{
			lx = (m0.x * x + m0.w * y) + (m1.z * z + m2.y);
			ly = (m0.y * x + m1.x * y) + (m1.w * z + m2.z);
			lz = (m0.z * x + m1.y * y) + (m2.x * z + m2.w);
			tail = m3;
		}
		else if (xform == kXformOffset)
		{
			lx = x + m0.x;
			ly = y + m0.y;
			lz = z + m0.z;
			tail = m1;
		}
		const float p[3] = { __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w) };
		float d = sdf::Brush(brush, p, lx, ly, lz);
		if (header & kHdrTailBit)
		{
			if (header & kHdrScaleBit) d = d * tail.x;
			param = tail.y;
		}
		value = d;
		if (op == kOpPush) code = kFoldPush;
		else if (op == kOpUnion) code = kFoldMin;
		else if (op == kOpInter) code = kFoldMax;
		else if (op == kOpDiff) { code = kFoldMax; value = -d; } // Diff(l, r) = fmaxf(l, -r)
		else code = kFoldBlendUnion + (op - kOpBlendUnion);
	}
	else
	{
		value = __uint_as_float(q.y);
		code = op == kOpFlate ? kFoldFlate : op == kOpStop ? kFoldNop : kFoldStackOp;
	}
	step = make_uint2(code | (slot << 8) | (op << 16), __float_as_uint(value));
}

__device__ __forceinline__ void LongStep(const uint4* __restrict__ pc, float x, float y, float z, uint2& step, float& param)
{
	LongStepLoaded(__ldg(pc), __ldg(reinterpret_cast<const float4*>(pc + 1)), __ldg(reinterpret_cast<const float4*>(pc + 2)),
		__ldg(reinterpret_cast<const float4*>(pc + 3)), __ldg(reinterpret_cast<const float4*>(pc + 4)), x, y, z, step, param);
}

// Test hook (tg_debug_check_long_programs): bit 0 skips pass a, bit 1 pass b, bit 2 the batched loads of GroupEvalLong, so
// that a disagreement with the plain interpreter can be pinned on one of them.  0 in normal operation.
__device__ int g_long_debug = 0;

// Clamp x -> min(max(x, lo), hi) as a value; Then() composes "this first, g second".
struct Clamp
{
	float lo, hi;
	__device__ __forceinline__ void Then(float glo, float ghi)
	{
		lo = fminf(fmaxf(lo, glo), ghi);
		hi = fminf(fmaxf(hi, glo), ghi);
	}
};

// Composes the steps [begin, end) that this thread holds of a chain into `c`; returns false when a step is not a clamp
// (blend, flate, a push / stack operator that was not collapsed).  `first` is the chain's opening push (a constant).
__device__ __forceinline__ bool ComposeShare(const uint2* steps, uint32_t begin, uint32_t end, uint32_t first, Clamp& c)
{
	bool clamps_only = true;
	for (uint32_t i = begin; i < end; ++i)
	{
		const uint32_t code = steps[i].x & 0xFFu;
		const float v = __uint_as_float(steps[i].y);
		if (code == kFoldMin) c.Then(-INFINITY, v);
		else if (code == kFoldMax) c.Then(v, INFINITY);
		else if (code == kFoldPush && i == first) c.Then(v, v);
		else if (code != kFoldNop) clamps_only = false;
	}
	return clamps_only;
}

// The step a collapsed operand becomes in its parent chain: parent = op(parent, value).
__device__ __forceinline__ void CollapseInto(uint2* steps, float* params, uint32_t at, float value)
{
	const uint32_t op = steps[at].x >> 16;
	const float threshold = __uint_as_float(steps[at].y);
	uint32_t code;
	if (op == kOpUnion) code = kFoldMin;
	else if (op == kOpInter) code = kFoldMax;
	else if (op == kOpDiff) { code = kFoldMax; value = -value; }
	else code = kFoldBlendUnion + (op - kOpBlendUnion);
	steps[at] = make_uint2(code | (kNoSlot << 8) | (op << 16), __float_as_uint(value));
	params[at] = threshold;
}

// Folds the chain steps[begin, end) -- it opens with a constant at `first` (a push), or continues from acc = 0 -- with
// the whole group.  Every thread composes its contiguous share into a clamp; a share that holds anything but min / max
// steps (a blend, a flate, an operand that was not collapsed) is marked general.  A warp whose shares are all clamps
// composes them by shuffles (compose mine, then the partner's: the operator is associative, not commutative).  Thread 0
// then walks the chain in order: a whole warp's clamp at a time where it can, share by share where a general share sits
// in the warp, step by step (FoldStep) inside a general share.  Long programs have a handful of general steps (the
// blends of the first tree of seaside_town's town), so the walk is a few dozen operations instead of a thousand.
// Returns the chain's value in thread 0 (other threads: unspecified).  `marks` needs 3 words per warp of the group.
template <int GROUP>
__device__ __forceinline__ float FoldChain(const uint2* steps, const float* params, uint32_t begin, uint32_t end, uint32_t first,
	float2* share, uint32_t* marks)
{
	const int tid = GROUP == 32 ? int(threadIdx.x & 31) : int(threadIdx.x);
	const int lane = threadIdx.x & 31;
	const int warp = GROUP == 32 ? 0 : int(threadIdx.x >> 5);
	const uint32_t len = end - begin;
	const uint32_t part = (len + GROUP - 1) / GROUP;
	const uint32_t b0 = begin + min(len, uint32_t(tid) * part), b1 = begin + min(len, uint32_t(tid + 1) * part);
	Clamp mine = { -INFINITY, INFINITY };
	const bool general = !ComposeShare(steps, b0, b1, first, mine);
	share[tid] = make_float2(mine.lo, mine.hi);
	const unsigned general_mask = __ballot_sync(0xFFFFFFFFu, general);
	Clamp whole = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const float plo = __shfl_down_sync(0xFFFFFFFFu, whole.lo, o), phi = __shfl_down_sync(0xFFFFFFFFu, whole.hi, o);
		if (lane + o < 32) whole.Then(plo, phi);
	}
	if (lane == 0)
	{
		marks[warp * 3 + 0] = general_mask;
		marks[warp * 3 + 1] = __float_as_uint(whole.lo);
		marks[warp * 3 + 2] = __float_as_uint(whole.hi);
	}
	if (GROUP == 32) __syncwarp(); else __syncthreads();
	float acc = 0.0f;
	if (tid == 0)
	{
		float stack[kMaxStackSlots];
		for (int w = 0; w < GROUP / 32; ++w)
		{
			const uint32_t mask = marks[w * 3 + 0];
			if (mask == 0u)
			{
				acc = fminf(fmaxf(acc, __uint_as_float(marks[w * 3 + 1])), __uint_as_float(marks[w * 3 + 2]));
				continue;
			}
			for (int l = 0; l < 32; ++l)
			{
				const uint32_t t = uint32_t(w * 32 + l);
				if ((mask >> l) & 1u)
				{
					const uint32_t s0 = begin + min(len, t * part), s1 = begin + min(len, (t + 1u) * part);
					for (uint32_t i = s0; i < s1; ++i)
					{
						uint32_t word = steps[i].x;
						// the chain's opening push spills the accumulator of the chain ABOVE to a slot this walk does not own
						if (i == first && (word & 0xFFu) == kFoldPush) word = kFoldPush | (kNoSlot << 8);
						FoldStep(word, __uint_as_float(steps[i].y), params[i], acc, stack);
					}
				}
				else acc = fminf(fmaxf(acc, share[t].x), share[t].y);
			}
		}
	}
	if (GROUP == 32) __syncwarp(); else __syncthreads();
	return acc;
}

// Exclusive prefix of a per-thread count over the group, and the group total.
template <int GROUP>
__device__ __forceinline__ uint32_t PrefixGroup(uint32_t mine, uint32_t* marks, uint32_t& total)
{
	const int lane = threadIdx.x & 31;
	uint32_t incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += v;
	}
	if (GROUP == 32)
	{
		total = __shfl_sync(0xFFFFFFFFu, incl, 31);
		return incl - mine;
	}
	const int warp = threadIdx.x >> 5;
	if (lane == 31) marks[warp] = incl;
	__syncthreads();
	uint32_t base = 0;
	total = 0;
	for (int w = 0; w < GROUP / 32; ++w)
	{
		const uint32_t m = marks[w];
		if (w < warp) base += m;
		total += m;
	}
	__syncthreads();
	return base + incl - mine;
}

// GROUP = 32: the calling warp works alone (tid = lane; `steps`, `params`, `share`, `marks`, `result_out` are its own
// slices of the scratch); GROUP = kLongThreads: the whole block.  Every thread of the group must call; all get the value.
//
// The fold, after the brush distances are in `steps`:
//   a. operands nested two deep (pushed with spill slot 1; anything deeper sits inside them) are folded by the thread
//      that owns their first step and collapse into one step of the chain above;
//   b. the operands of the main chain (slot 0) are then plain chains.  Their k-th opening push pairs with the k-th
//      closing operator, so both are listed in order; short ones are folded by one thread each, long ones (a whole town
//      united into one right-hand operand, seaside_town.lua:106) by the whole group: min / max steps are clamps, clamps
//      compose associatively, every thread composes a contiguous share and thread 0 the shares in order;
//   c. the main chain is reduced the same way.
// min and max select one of their arguments, so every value is the sequential one bit for bit; chains with a blend or
// flate step are walked by one thread instead.
template <int GROUP>
__device__ __forceinline__ float GroupEvalLong(const uint4* __restrict__ program, uint32_t count, float x, float y, float z,
	uint2* steps, float* params, float2* share, uint32_t* marks, float* result_out, int capacity)
{
	const int tid = GROUP == 32 ? int(threadIdx.x & 31) : int(threadIdx.x);
	auto sync = [] { if (GROUP == 32) __syncwarp(); else __syncthreads(); };
	const uint32_t* __restrict__ table = reinterpret_cast<const uint32_t*>(program) - ((count + 3u) & ~3u);
	if (count > uint32_t(capacity))
	{
		// larger than the scratch: rounds of `capacity` steps, folded one after the other by the first thread
		float acc = 0.0f;
		float stack[kMaxStackSlots];
		for (uint32_t base = 0; base < count; base += uint32_t(capacity))
		{
			const uint32_t n = min(uint32_t(capacity), count - base);
			for (uint32_t i = tid; i < n; i += GROUP) LongStep(program + __ldg(&table[base + i]), x, y, z, steps[i], params[i]);
			sync();
			if (tid == 0)
			{
				for (uint32_t i = 0; i < n; ++i) FoldStep(steps[i].x, __uint_as_float(steps[i].y), params[i], acc, stack);
				*result_out = acc;
			}
			sync();
		}
		return *result_out;
	}
	const uint32_t n = count;
	// brush distances, four instructions per thread at a time: the four table entries travel together, then the twenty
	// operand quads (two memory round trips per batch instead of eight)
	for (uint32_t first = tid; first < n; first += GROUP * 4)
	{
		uint32_t at[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) at[u] = __ldg(&table[min(first + uint32_t(u) * GROUP, n - 1u)]);
		uint4 q[4];
		float4 m0[4], m1[4], m2[4], m3[4];
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const uint4* pc = program + at[u];
			q[u] = __ldg(pc);
			m0[u] = __ldg(reinterpret_cast<const float4*>(pc + 1));
			m1[u] = __ldg(reinterpret_cast<const float4*>(pc + 2));
			m2[u] = __ldg(reinterpret_cast<const float4*>(pc + 3));
			m3[u] = __ldg(reinterpret_cast<const float4*>(pc + 4));
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const uint32_t i = first + uint32_t(u) * GROUP;
			if (i < n) LongStepLoaded(q[u], m0[u], m1[u], m2[u], m3[u], x, y, z, steps[i], params[i]);
		}
	}
	sync();

	// a. operands two deep.  An owner first folds its operand reading only, the group meets, then the owners write: a
	// thread that tests a step inside somebody else's operand must not see it half rewritten.
	const int debug = g_long_debug;
	for (uint32_t base = 0; base < n && !(debug & 1); base += GROUP)
	{
		const uint32_t i = base + uint32_t(tid);
		const bool owner = i < n && (steps[i].x & 0xFFFFu) == (kFoldPush | (1u << 8));
		float acc = 0.0f;
		uint32_t j = n;
		if (owner)
		{
			acc = __uint_as_float(steps[i].y);
			float stack[kMaxStackSlots];
			for (j = i + 1; j < n; ++j)
			{
				const uint32_t word = steps[j].x;
				if ((word & 0xFFFFu) == (kFoldStackOp | (1u << 8))) break;
				FoldStep(word, __uint_as_float(steps[j].y), params[j], acc, stack);
			}
		}
		sync();
		if (owner && j < n)
		{
			for (uint32_t k = i; k < j; ++k) steps[k].x = kFoldNop | (kNoSlot << 8);
			CollapseInto(steps, params, j, acc);
		}
		sync();
	}

	// b. operands of the main chain: ordered lists of their opening pushes and closing operators.  `marks` holds, per
	// thread, the number of openings (low half) and closings (high half) in its contiguous share -> exclusive prefix.
	const uint32_t per = (n + GROUP - 1) / GROUP;
	const uint32_t lo_i = min(n, uint32_t(tid) * per), hi_i = min(n, uint32_t(tid + 1) * per);
	uint32_t mine = 0;
	for (uint32_t i = lo_i; i < hi_i; ++i)
	{
		const uint32_t key = steps[i].x & 0xFFFFu;
		mine += (key == (kFoldPush | (0u << 8)) ? 1u : 0u) + (key == (kFoldStackOp | (0u << 8)) ? 0x10000u : 0u);
	}
	uint32_t total = 0;
	const uint32_t before = PrefixGroup<GROUP>(mine, marks, total);
	// the lists reuse `share` (x: position of the k-th opening, y: of the k-th closing), as bit patterns
	const uint32_t operands = min(total & 0xFFFFu, total >> 16);
	{
		uint32_t open_rank = before & 0xFFFFu, close_rank = before >> 16;
		for (uint32_t i = lo_i; i < hi_i; ++i)
		{
			const uint32_t key = steps[i].x & 0xFFFFu;
			if (key == (kFoldPush | (0u << 8)) && open_rank < uint32_t(GROUP)) share[open_rank++].x = __uint_as_float(i);
			else if (key == (kFoldStackOp | (0u << 8)) && close_rank < uint32_t(GROUP)) share[close_rank++].y = __uint_as_float(i);
		}
	}
	sync();
	constexpr uint32_t kShortOperand = 48;
	bool walk_all = operands > uint32_t(GROUP) || (debug & 2) != 0; // more operands than list entries: one thread walks the whole program
	uint32_t my_open = 0, my_close = 0;
	if (!walk_all && uint32_t(tid) < operands)
	{
		my_open = __float_as_uint(share[tid].x);
		my_close = __float_as_uint(share[tid].y);
		if (my_close <= my_open) walk_all = true; // malformed pairing
	}
	walk_all = GROUP == 32 ? __any_sync(0xFFFFFFFFu, walk_all) : (__syncthreads_or(walk_all ? 1 : 0) != 0);
	if (!walk_all)
	{
		// short operands: one thread each
		if (uint32_t(tid) < operands && my_close - my_open <= kShortOperand)
		{
			float acc = __uint_as_float(steps[my_open].y);
			float stack[kMaxStackSlots];
			for (uint32_t j = my_open + 1; j < my_close; ++j) FoldStep(steps[j].x, __uint_as_float(steps[j].y), params[j], acc, stack);
			for (uint32_t k = my_open; k < my_close; ++k) steps[k].x = kFoldNop | (kNoSlot << 8);
			CollapseInto(steps, params, my_close, acc);
		}
		// long operands: the whole group, one after the other (the lists are read before `share` is reused)
		uint32_t long_open[4], long_close[4];
		uint32_t long_count = 0;
		for (uint32_t k = 0; k < operands; ++k)
		{
			const uint32_t o = __float_as_uint(share[k].x), c = __float_as_uint(share[k].y);
			if (c > o && c - o > kShortOperand)
			{
				if (long_count < 4u)
				{
					long_open[long_count] = o;
					long_close[long_count] = c;
				}
				long_count++;
			}
		}
		sync();
		for (uint32_t g = 0; g < min(long_count, 4u); ++g)
		{
			const uint32_t o = long_open[g], c = long_close[g], len = c - o;
			const uint32_t part = (len + GROUP - 1) / GROUP;
			const uint32_t b0 = o + min(len, uint32_t(tid) * part), b1 = o + min(len, uint32_t(tid + 1) * part);
			const float value = FoldChain<GROUP>(steps, params, o, c, o, share, marks);
			if (tid == 0) CollapseInto(steps, params, c, value); // (c itself lies outside every share)
			for (uint32_t k = b0; k < b1; ++k) steps[k].x = kFoldNop | (kNoSlot << 8);
			sync();
		}
		if (long_count > 4u) walk_all = true; // (uniform: every thread counted the same lists)
	}
	sync();

	// c. the main chain (after a walk_all its operands are still in place: the walk handles them step by step, every
	// share that holds one being general)
	const float value = FoldChain<GROUP>(steps, params, 0u, n, 0u, share, marks);
	if (tid == 0) *result_out = value;
	sync();
	return *result_out;
}

template <int S>
struct MaterialRegs
{
	uint32_t v[S];
};

// Executes `program` for S points.  MATERIAL selects the GetMaterial walk (tree stream only).
template <int S, bool MATERIAL, bool DEFER = false>
__device__ __forceinline__ void RunProgram(const uint32_t* __restrict__ program, const float (&px)[S], const float (&py)[S], const float (&pz)[S],
	float (&result)[S], uint32_t (&result_material)[S], sdf::SqrtRange* seen = nullptr)
{
	const typename SqrtPolicy<DEFER>::type sq = SqrtPolicy<DEFER>::Make(seen);
	float acc[S];
	uint32_t accm[S];
	float stack[kMaxStackSlots][S];
	uint32_t stackm[MATERIAL ? kMaxStackSlots : 1][S];
#pragma unroll
	for (int s = 0; s < S; ++s)
	{
		acc[s] = 0.0f;
		accm[s] = kNoMaterial;
	}
	const uint32_t* pc = program;
	for (;;)
	{
		const uint32_t header = __ldg(pc);
		const uint32_t brush = header & kHdrBrushMask;
		const uint32_t op = (header >> kHdrOpShift) & 0xFu;
		const uint32_t slot = (header >> kHdrSlotShift) & 0xFFu;
		const float* __restrict__ arg = reinterpret_cast<const float*>(pc + 1);
		pc += header >> kHdrLenShift;

		if (brush != kBrushNone)
		{
			float lx[S], ly[S], lz[S];
			const uint32_t xform = (header >> kHdrXformShift) & 3u;
			if (xform == kXformMatrix)
			{
				// glm mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*1)  (type_mat4x4.inl:561-572)
				const float m00 = __ldg(arg + 0), m01 = __ldg(arg + 1), m02 = __ldg(arg + 2);
				const float m10 = __ldg(arg + 3), m11 = __ldg(arg + 4), m12 = __ldg(arg + 5);
				const float m20 = __ldg(arg + 6), m21 = __ldg(arg + 7), m22 = __ldg(arg + 8);
				const float m30 = __ldg(arg + 9), m31 = __ldg(arg + 10), m32 = __ldg(arg + 11);
				arg += 12;
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = (m00 * px[s] + m10 * py[s]) + (m20 * pz[s] + m30);
					ly[s] = (m01 * px[s] + m11 * py[s]) + (m21 * pz[s] + m31);
					lz[s] = (m02 * px[s] + m12 * py[s]) + (m22 * pz[s] + m32);
				}
			}
			else if (xform == kXformOffset)
			{
				const float ox = __ldg(arg + 0), oy = __ldg(arg + 1), oz = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = px[s] + ox;
					ly[s] = py[s] + oy;
					lz[s] = pz[s] + oz;
				}
			}
			else if (xform == kXformQuat)
			{
				// Transform::ApplyInv (transform.cpp:64-67): rotate(inverse(q), p - t) / s, glm quat * vec3 (type_quat.inl:343-350)
				const float qw = __ldg(arg + 0), qx = __ldg(arg + 1), qy = __ldg(arg + 2), qz = __ldg(arg + 3);
				const float tx = __ldg(arg + 4), ty = __ldg(arg + 5), tz = __ldg(arg + 6), sc = __ldg(arg + 7);
				arg += 8;
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					const float vx = px[s] - tx, vy = py[s] - ty, vz = pz[s] - tz;
					const float uvx = qy * vz - vy * qz, uvy = qz * vx - vz * qx, uvz = qx * vy - vx * qy;
					const float uuvx = qy * uvz - uvy * qz, uuvy = qz * uvx - uvz * qx, uuvz = qx * uvy - uvx * qy;
					lx[s] = vx + ((uvx * qw) + uuvx) * 2.0f;
					ly[s] = vy + ((uvy * qw) + uuvy) * 2.0f;
					lz[s] = vz + ((uvz * qw) + uuvz) * 2.0f;
				}
				if (sc != 1.0f) // x / 1.0f is x, bit for bit: unscaled brushes (most) skip three IEEE divisions per sample
				{
#pragma unroll
					for (int s = 0; s < S; ++s)
					{
						lx[s] = lx[s] / sc;
						ly[s] = ly[s] / sc;
						lz[s] = lz[s] / sc;
					}
				}
			}
			else
			{
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					lx[s] = px[s];
					ly[s] = py[s];
					lz[s] = pz[s];
				}
			}

			float d[S];
			uint32_t kind = brush; // comparison chains instead of jump tables, as in EvalInterp
			if (kind == kBrushBox)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1), c = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Box(lx[s], ly[s], lz[s], a, b, c, sq);
			}
			else if (Opaque(kind) == kBrushSphere)
			{
				const float r = __ldg(arg);
				arg += 1;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Sphere(lx[s], ly[s], lz[s], r, sq);
			}
			else if (Opaque(kind) == kBrushCylinder)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cylinder(lx[s], ly[s], lz[s], a, b, sq);
			}
			else if (Opaque(kind) == kBrushTorus)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Torus(lx[s], ly[s], lz[s], a, b, sq);
			}
			else if (Opaque(kind) == kBrushPlane)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1), c = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Plane(lx[s], ly[s], lz[s], a, b, c);
			}
			else if (Opaque(kind) == kBrushEllipsoid)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1), c = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Ellipsoid(lx[s], ly[s], lz[s], a, b, c, sq);
			}
			else if (Opaque(kind) == kBrushCone)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cone(lx[s], ly[s], lz[s], a, b, sq);
			}
			else
			{
				const float a = __ldg(arg), b = __ldg(arg + 1), c = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Coninder(lx[s], ly[s], lz[s], a, b, c, sq);
			}
			if (header & kHdrScaleBit)
			{
				const float sc = __ldg(arg);
				arg += 1;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = d[s] * sc; // ScaleField (:1587-1591) / BrushNode::Eval `* Scalation` (:465)
			}
			uint32_t material = kNoMaterial;
			if (header & kHdrMaterialBit)
			{
				material = __ldg(reinterpret_cast<const uint32_t*>(arg));
				arg += 1;
			}

			if (op == kOpPush)
			{
				if (slot != kNoSlot)
				{
#pragma unroll
					for (int s = 0; s < S; ++s)
					{
						stack[slot][s] = acc[s];
						if (MATERIAL) stackm[slot][s] = accm[s];
					}
				}
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					acc[s] = d[s];
					accm[s] = material;
				}
			}
			else
			{
				const float threshold = (op >= kOpBlendUnion) ? __ldg(arg) : 0.0f;
				float dist[S];
				uint32_t oper = op; // one dispatch per instruction, not per sample
				if (oper == kOpUnion)
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::Union(acc[s], d[s]);
				}
				else if (Opaque(oper) == kOpBlendUnion)
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::BlendUnion(acc[s], d[s], threshold);
				}
				else if (Opaque(oper) == kOpDiff)
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::Diff(acc[s], d[s]);
				}
				else if (Opaque(oper) == kOpInter)
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::Inter(acc[s], d[s]);
				}
				else if (Opaque(oper) == kOpBlendDiff)
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::BlendDiff(acc[s], d[s], threshold);
				}
				else
				{
#pragma unroll
					for (int s = 0; s < S; ++s) dist[s] = sdf::BlendInter(acc[s], d[s], threshold);
				}
#pragma unroll
				for (int s = 0; s < S; ++s)
				{
					const float l = acc[s];
					const float r = d[s];
					if (MATERIAL)
					{
						// SetNode::GetMaterial (:957-1012)
						const bool take_left = (op >= kOpBlendUnion) ? (fabsf(l - dist[s]) <= fabsf(r - dist[s])) : (dist[s] == l);
						const uint32_t family = (op - 1u) % 3u; // 0 union, 1 inter, 2 diff
						uint32_t m;
						if (family == 2u) m = accm[s];
						else if (family == 0u) m = take_left ? accm[s] : material;
						else
						{
							const bool lv = (header & kHdrLhsPaintBit) != 0, rv = (header & kHdrRhsPaintBit) != 0;
							m = (lv && rv) ? (take_left ? accm[s] : material) : (lv ? accm[s] : material);
						}
						accm[s] = m;
					}
					acc[s] = dist[s];
				}
			}
		}
		else if (op == kOpStop)
		{
#pragma unroll
			for (int s = 0; s < S; ++s)
			{
				result[s] = acc[s];
				result_material[s] = accm[s];
			}
			return;
		}
		else if (op == kOpFlate)
		{
			const float radius = __ldg(arg);
#pragma unroll
			for (int s = 0; s < S; ++s) acc[s] = acc[s] - radius; // :1567-1571
		}
		else if (op == kOpStencil)
		{
			// StencilMaskNode::GetMaterial (:666-679): accumulator holds the mask distance, the child is on the stack
			const uint32_t material = __ldg(reinterpret_cast<const uint32_t*>(arg));
			const bool apply_to_negative = (header & kHdrStencilNegBit) != 0;
#pragma unroll
			for (int s = 0; s < S; ++s)
			{
				const bool interior = acc[s] < 0.0f;
				acc[s] = stack[slot][s];
				if (MATERIAL) accm[s] = (interior == apply_to_negative) ? material : stackm[slot][s];
			}
		}
		else
		{
			// stack-form set operator: lhs was spilled, rhs is the accumulator
			const float threshold = (op >= kOpBlendUnion) ? __ldg(arg) : 0.0f;
#pragma unroll
			for (int s = 0; s < S; ++s)
			{
				const float l = stack[slot][s];
				const float r = acc[s];
				const float dist = sdf::SetOp(op, l, r, threshold);
				if (MATERIAL)
				{
					const uint32_t ml = stackm[slot][s], mr = accm[s];
					const bool take_left = (op >= kOpBlendUnion) ? (fabsf(l - dist) <= fabsf(r - dist)) : (dist == l);
					const uint32_t family = (op - 1u) % 3u;
					uint32_t m;
					if (family == 2u) m = ml;
					else if (family == 0u) m = take_left ? ml : mr;
					else
					{
						const bool lv = (header & kHdrLhsPaintBit) != 0, rv = (header & kHdrRhsPaintBit) != 0;
						m = (lv && rv) ? (take_left ? ml : mr) : (lv ? ml : mr);
					}
					accm[s] = m;
				}
				acc[s] = dist;
			}
		}
	}
}

// Convenience wrappers -------------------------------------------------------------------------------

template <int S>
__device__ __forceinline__ void EvalDistance(const uint32_t* __restrict__ program, const float (&px)[S], const float (&py)[S], const float (&pz)[S], float (&out)[S])
{
	uint32_t unused[S];
	RunProgram<S, false>(program, px, py, pz, out, unused);
}

__device__ __forceinline__ float EvalDistance1(const uint32_t* __restrict__ program, float x, float y, float z)
{
	float px[1] = { x }, py[1] = { y }, pz[1] = { z }, out[1];
	uint32_t unused[1];
	RunProgram<1, false>(program, px, py, pz, out, unused);
	return out[0];
}

// Value (and, through *material, the GetMaterial result) of a tree-stream program at one point.  Deliberately not
// inlined: the attribute kernels call it from several places and must carry exactly one copy of the interpreter
// (their instruction footprint is what limits them).
__device__ __noinline__ float EvalTreeCentre(const uint32_t* __restrict__ tree_program, float x, float y, float z, uint32_t* material = nullptr)
{
	float px[1] = { x }, py[1] = { y }, pz[1] = { z }, d[1];
	uint32_t m[1];
	RunProgram<1, true>(tree_program, px, py, pz, d, m);
	if (material) *material = m[0];
	return d[0];
}

__device__ __noinline__ void EvalDistance4Exact(const uint32_t* __restrict__ program, const float (&px)[4], const float (&py)[4], const float (&pz)[4], float (&out)[4])
{
	EvalDistance<4>(program, px, py, pz, out);
}

// SDFNode::Gradient (sdf_evaluator.cpp:298-333) on a tree-stream program: the four tetrahedral taps are
// the four samples of one RunProgram<4> call.
__device__ __forceinline__ void EvalGradient(const uint32_t* __restrict__ tree_program, float x, float y, float z, float& gx, float& gy, float& gz)
{
	const float ox = 1.0f * 0.0001f;
	const float oy = -1.0f * 0.0001f;
	// taps: xyy, yyx, yxy, xxx
	const float px[4] = { x + ox, x + oy, x + oy, x + ox };
	const float py[4] = { y + oy, y + oy, y + ox, y + ox };
	const float pz[4] = { z + oy, z + ox, z + oy, z + ox };
	float d[4];
	{
		// branch-free square roots (sdf::SqrtDeferred); the exact repeat is practically never taken and lives out of line
		uint32_t unused[4];
		sdf::SqrtRange seen;
		RunProgram<4, false, true>(tree_program, px, py, pz, d, unused, &seen);
		if (seen.Suspect()) EvalDistance4Exact(tree_program, px, py, pz, d);
	}
	// Offset.xyy * d0 + Offset.yyx * d1 + Offset.yxy * d2 + Offset.xxx * d3, summed left to right
	float sx = ((ox * d[0] + oy * d[1]) + oy * d[2]) + ox * d[3];
	float sy = ((oy * d[0] + oy * d[1]) + ox * d[2]) + ox * d[3];
	float sz = ((oy * d[0] + ox * d[1]) + oy * d[2]) + ox * d[3];
	const float len_sq = sx * sx + sy * sy + sz * sz;
	if (len_sq == 0.0f)
	{
		// zero gradient: forward differences (:320-328); taps xyy, yxy, yyx are d[0], d[2], d[1]
		const float dist = EvalTreeCentre(tree_program, x, y, z);
		const float fx = d[0] - dist, fy = d[2] - dist, fz = d[1] - dist;
		const float inv = 1.0f / sqrtf(fx * fx + fy * fy + fz * fz); // glm::normalize = v * inversesqrt(dot(v, v))
		gx = fx * inv;
		gy = fy * inv;
		gz = fz * inv;
		return;
	}
	const float len = sqrtf(len_sq);
	gx = sx / len;
	gy = sy / len;
	gz = sz / len;
}

} // namespace tg
