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
template <int S, bool PREFETCH = false, bool DEEP = false>
__device__ __forceinline__ void EvalInterp(const uint4* __restrict__ pc, const float (&px)[S], const float (&py)[S], const float (&pz)[S], float (&result)[S])
{
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
				for (int s = 0; s < S; ++s) d[s] = sdf::Box(lx[s], ly[s], lz[s], p0, p1, p2);
			}
			else if (Opaque(kind) == kBrushSphere)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Sphere(lx[s], ly[s], lz[s], p0);
			}
			else if (Opaque(kind) == kBrushCylinder)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cylinder(lx[s], ly[s], lz[s], p0, p1);
			}
			else if (Opaque(kind) == kBrushTorus)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Torus(lx[s], ly[s], lz[s], p0, p1);
			}
			else if (Opaque(kind) == kBrushPlane)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Plane(lx[s], ly[s], lz[s], p0, p1, p2);
			}
			else if (Opaque(kind) == kBrushEllipsoid)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Ellipsoid(lx[s], ly[s], lz[s], p0, p1, p2);
			}
			else if (Opaque(kind) == kBrushCone)
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cone(lx[s], ly[s], lz[s], p0, p1);
			}
			else
			{
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Coninder(lx[s], ly[s], lz[s], p0, p1, p2);
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

template <int S>
struct MaterialRegs
{
	uint32_t v[S];
};

// Executes `program` for S points.  MATERIAL selects the GetMaterial walk (tree stream only).
template <int S, bool MATERIAL>
__device__ __forceinline__ void RunProgram(const uint32_t* __restrict__ program, const float (&px)[S], const float (&py)[S], const float (&pz)[S],
	float (&result)[S], uint32_t (&result_material)[S])
{
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
				for (int s = 0; s < S; ++s) d[s] = sdf::Box(lx[s], ly[s], lz[s], a, b, c);
			}
			else if (Opaque(kind) == kBrushSphere)
			{
				const float r = __ldg(arg);
				arg += 1;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Sphere(lx[s], ly[s], lz[s], r);
			}
			else if (Opaque(kind) == kBrushCylinder)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cylinder(lx[s], ly[s], lz[s], a, b);
			}
			else if (Opaque(kind) == kBrushTorus)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Torus(lx[s], ly[s], lz[s], a, b);
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
				for (int s = 0; s < S; ++s) d[s] = sdf::Ellipsoid(lx[s], ly[s], lz[s], a, b, c);
			}
			else if (Opaque(kind) == kBrushCone)
			{
				const float a = __ldg(arg), b = __ldg(arg + 1);
				arg += 2;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Cone(lx[s], ly[s], lz[s], a, b);
			}
			else
			{
				const float a = __ldg(arg), b = __ldg(arg + 1), c = __ldg(arg + 2);
				arg += 3;
#pragma unroll
				for (int s = 0; s < S; ++s) d[s] = sdf::Coninder(lx[s], ly[s], lz[s], a, b, c);
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
	EvalDistance<4>(tree_program, px, py, pz, d);
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
