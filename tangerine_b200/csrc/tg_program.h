// Device instruction stream ("flat program") shared by the host flattener and the CUDA interpreter.
//
// The reference evaluates each octree node through a postfix word stream, one opcode or float per
// 32-bit word (tangerine/sdf_evaluator.h:83-163, interpreter at sdf_evaluator.cpp:1386-1605).  For the
// GPU that stream is re-encoded into *accumulator-form superinstructions*: one header word selects a
// point transform, a brush, an optional field scale and the set operator that consumes the brush,
// so a left-leaning CSG chain (the common shape, sdf_evaluator.cpp:759-771) runs with one dispatch
// per primitive and no operand-stack traffic at all.  Only right-nested operands spill the
// accumulator to an explicit stack slot whose index is known at flatten time.
//
// Two streams exist per octree node:
//   * kStreamInterp  mirrors SDFInterpreter::Eval:   Matrix (3x4 of the compiled inverse) / Offset transforms
//   * kStreamTree    mirrors the virtual SDFNode::Eval tree walk used by Gradient / GetMaterial /
//                    VoxExport (Transform::ApplyInv quaternion path, transform.cpp:64-67), and carries
//                    material ids, paint flags and stencil masks so the same stream also evaluates
//                    GetMaterial (sdf_evaluator.cpp:537-547, 666-679, 957-1012).
//
// Word layout of a kStreamTree instruction:  [header] [transform params] [brush params] [scale] [material] [op param]
//
// kStreamInterp is the hot stream (every lattice sample runs it) and is laid out in 16-byte quads so that the
// interpreter fetches an instruction with warp-uniform 128-bit loads, all of them independent of each other:
//   quad 0            [header, p0, p1, p2]            brush parameters (unused ones are 0)
//   quads 1-3         matrix transform: 12 floats, column-major 3 rows x 4 columns      (kXformMatrix)
//   quad 1            [ox, oy, oz, 0]                                                   (kXformOffset)
//   tail quad         [scale, threshold, 0, 0]        only when kHdrTailBit is set (scaled brush or blend operator)
// Operator-only instructions are one quad [header, param, 0, 0].  The length field counts quads; programs start
// on quad boundaries (FlatNode::interp_offset is a word offset that is a multiple of 4).
#pragma once

#include <cstdint>

namespace tg
{

// Brush kinds keep the numbering of the reference's OpcodeT (sdf_evaluator.h:83-107).
enum : uint32_t
{
	kBrushNone = 0,
	kBrushSphere = 1,
	kBrushEllipsoid = 2,
	kBrushBox = 3,
	kBrushTorus = 4,
	kBrushCylinder = 5,
	kBrushCone = 6,
	kBrushConinder = 7,
	kBrushPlane = 8,
};

enum : uint32_t
{
	kXformNone = 0,
	kXformOffset = 1, // 3 floats: point = p + offset
	kXformMatrix = 2, // 12 floats, column-major 3 rows x 4 columns of the inverse matrix
	kXformQuat = 3,   // 8 floats: inverse quaternion (w x y z), translation (x y z), scale
};

enum : uint32_t
{
	kOpPush = 0, // brush value becomes the accumulator; previous accumulator spills to a stack slot
	kOpUnion = 1,
	kOpInter = 2,
	kOpDiff = 3,
	kOpBlendUnion = 4, // + 1 float threshold
	kOpBlendInter = 5,
	kOpBlendDiff = 6,
	kOpFlate = 7,   // + 1 float radius; accumulator -= radius
	kOpStencil = 8, // tree stream only: accumulator is the mask distance, the child is on the stack; + 1 word material
	kOpStop = 15,
};

// Header bit fields.
constexpr uint32_t kHdrBrushMask = 0xFu;         // bits 0-3
constexpr uint32_t kHdrXformShift = 4;           // bits 4-5
constexpr uint32_t kHdrScaleBit = 1u << 6;       // one float follows the brush params
constexpr uint32_t kHdrMaterialBit = 1u << 7;    // one word (material id) follows (tree stream brushes)
constexpr uint32_t kHdrOpShift = 8;              // bits 8-11
constexpr uint32_t kHdrLhsPaintBit = 1u << 12;   // HasPaint() of the left operand (material walk of Inter)
constexpr uint32_t kHdrRhsPaintBit = 1u << 13;   // HasPaint() of the right operand
constexpr uint32_t kHdrStencilNegBit = 1u << 14; // StencilMaskNode<ApplyToNegative = true>
constexpr uint32_t kHdrTailBit = 1u << 15;       // kStreamInterp: a [scale, threshold, 0, 0] quad closes the instruction
constexpr uint32_t kHdrSlotShift = 16;           // bits 16-23: stack slot to spill to / pop from, 0xFF = none
constexpr uint32_t kHdrLenShift = 24;            // bits 24-31: instruction length, header included (words; quads in kStreamInterp)
constexpr uint32_t kNoSlot = 0xFFu;

constexpr uint32_t kNoMaterial = 0xFFFFFFFFu; // the reference's default white material (sdf_evaluator.cpp:34-38)
constexpr int kMaxStackSlots = 16;            // deepest right-nesting the device interpreter accepts

inline constexpr uint32_t MakeHeader(uint32_t brush, uint32_t xform, uint32_t op, uint32_t slot, uint32_t flags, uint32_t len)
{
	return brush | (xform << kHdrXformShift) | (op << kHdrOpShift) | (slot << kHdrSlotShift) | (len << kHdrLenShift) | flags;
}

inline constexpr int BrushParamCount(uint32_t brush)
{
	return brush == kBrushSphere ? 1 : (brush == kBrushTorus || brush == kBrushCylinder || brush == kBrushCone) ? 2 : 3;
}

// Algorithmic FLOPs per instruction part -- the accounting convention frozen in SURVEY.md section 8(d):
// FMA = 2, sqrt / div = 1, abs / neg / compare-select = 1, counted from sdf_evaluator.cpp:165-295.
inline constexpr int BrushFlops(uint32_t brush)
{
	return brush == kBrushSphere ? 7 : brush == kBrushEllipsoid ? 24 : brush == kBrushBox ? 19 : brush == kBrushTorus ? 10
		: brush == kBrushCylinder ? 17 : brush == kBrushCone ? 46 : brush == kBrushConinder ? 40 : brush == kBrushPlane ? 5 : 0;
}
inline constexpr int XformFlops(uint32_t xform)
{
	return xform == kXformOffset ? 3 : xform == kXformMatrix ? 18 : xform == kXformQuat ? 36 : 0;
}
inline constexpr int OpFlops(uint32_t op)
{
	return (op == kOpUnion || op == kOpInter || op == kOpFlate) ? 1 : op == kOpDiff ? 2
		: (op == kOpBlendUnion || op == kOpBlendInter) ? 9 : op == kOpBlendDiff ? 10 : 0;
}

// One flattened octree node (device table entry, 64 bytes).
struct alignas(16) FlatNode
{
	float pivot[3];
	uint32_t terminus;   // 1: Descend stops here
	int32_t children[8]; // node index, or -1 when the octant is empty (Descend falls back to this node)
	uint32_t interp_offset; // word offset of this node's kStreamInterp program
	uint32_t tree_offset;   // word offset of this node's kStreamTree program
	uint32_t flags;         // kNodeCullable etc.
	uint32_t flops;         // algorithmic FLOPs of one evaluation of the interp program
};
static_assert(sizeof(FlatNode) == 64, "FlatNode layout");

// One evaluation region of the octree: the box of points whose SDFOctree::Descend ends at `node` -- the whole
// cell of a terminus node, or one empty octant of an interior node (the reference then evaluates the interior
// node's larger program there, sdf_evaluator.cpp:1828-1834).  A point belongs to the region when
// lo < p <= hi on every axis (the strict `>` of the pivot tests, :1806-1817); the octree root is unbounded.
// The regions partition space; empty-space culling (K0) walks them instead of the grid.
//
// For an empty octant the octree build has already evaluated the interior node's program at the octant's centre:
// SDFNode::Clip (sdf_evaluator.cpp:782-850, 468-478) returns null exactly when that value exceeds the octant's half
// diagonal.  `known_value` keeps it (0 when nothing is known), so culling can dismiss every brick within
// known_value of `center` without running the (large) program again.
struct FlatRegion
{
	float lo[3];
	float hi[3];
	float center[3];
	float known_value;
	uint32_t node;
	uint32_t pad;
};
static_assert(sizeof(FlatRegion) == 48, "FlatRegion layout");

constexpr uint32_t kNodeCullable = 1u; // every primitive in the program is a true distance bound (no Ellipsoid)
// Long programs (interior octree nodes keep hundreds of primitives) are flagged: one thread walking such a program
// is bound by the latency of a dependent fetch + sqrt chain per primitive, and K0 schedules around that.
constexpr uint32_t kNodeLong = 2u;       // FlatNode::flags: at least kLongProgram instructions
constexpr uint32_t kNodeCountShift = 8;  // FlatNode::flags >> 8: instruction count of the kStreamInterp program (Stop excluded)
constexpr uint32_t kLongProgram = 32;

} // namespace tg
