// Host-side CSG tree: the value the Lua / C front-ends build and the export path consumes.
//
// Mirrors the reference's SDFNode hierarchy (tangerine/sdf_evaluator.h:171-264, .cpp:432-1372) as a
// flat, index-linked node array instead of a shared_ptr class tree: brushes are immutable leaves,
// operators refer to children by index, and pruning (Clip) only ever appends new operator nodes.
// That makes a pruned subtree a single index, lets octree workers share the brush table read-only,
// and serialises to disk (.tgm) with a memcpy.
//
// .tgm file layout (little endian), written by oracle/ref_tool.cpp from the reference's own trees:
//   u32 magic "TGM1", u32 node_count, u32 material_count, u32 root
//   material_count x f32[3]   sampled sRGB colour, i.e. SampleColor(Material->GuessColor()) (export.cpp:303-307)
//   node_count x TgmRecord    post-order (children before parents)
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "tg_math.h"
#include "tg_program.h"

namespace tg
{

// Node kinds: brushes and set operators use the reference's OpcodeT numbers.
enum : uint32_t
{
	kKindSphere = 1, kKindEllipsoid, kKindBox, kKindTorus, kKindCylinder, kKindCone, kKindConinder, kKindPlane,
	kKindUnion = 9, kKindInter, kKindDiff, kKindBlendUnion, kKindBlendInter, kKindBlendDiff, kKindFlate,
	kKindStencilPos = 100, // StencilMaskNode<false>: override where the mask is >= 0
	kKindStencilNeg = 101, // StencilMaskNode<true>:  override where the mask is < 0
};
constexpr uint32_t kNoNode = 0xFFFFFFFFu;

struct TgmRecord
{
	uint32_t kind, a, b, material;
	float params[4];
	float quat[4]; // w x y z
	float trans[3];
	float scale;
	float bounds_min[3], bounds_max[3]; // brush-local AABB
};
static_assert(sizeof(TgmRecord) == 88, "TgmRecord layout");

struct Node
{
	uint32_t kind = 0;
	uint32_t a = kNoNode; // lhs / child
	uint32_t b = kNoNode; // rhs / stencil mask
	uint32_t material = kNoMaterial;
	float params[4] = { 0, 0, 0, 0 }; // brush params | blend threshold | flate radius
	Quat rotation;
	Vec3 translation;
	float scalation = 1.0f;
	Box3 local_bounds;
	// Derived on insertion.
	uint32_t stack_size = 0; // SDFNode::StackSize
	int32_t leaf_count = 0;  // LeafCount()
	bool has_paint = false;  // HasPaint()
	bool finite = true;      // HasFiniteBounds()
};

inline bool IsBrush(uint32_t k) { return k >= kKindSphere && k <= kKindPlane; }
inline bool IsSet(uint32_t k) { return k >= kKindUnion && k <= kKindBlendDiff; }
inline bool IsBlend(uint32_t k) { return k >= kKindBlendUnion && k <= kKindBlendDiff; }
inline bool IsStencil(uint32_t k) { return k == kKindStencilPos || k == kKindStencilNeg; }
enum class Family { Union, Inter, Diff };
inline Family SetFamily(uint32_t k)
{
	return (k == kKindUnion || k == kKindBlendUnion) ? Family::Union : (k == kKindInter || k == kKindBlendInter) ? Family::Inter : Family::Diff;
}

// Process-wide material table: id -> sampled sRGB colour.
uint32_t RegisterMaterial(float r, float g, float b);
bool MaterialColor(uint32_t id, float out_rgb[3]);
uint32_t MaterialCount();
void SnapshotMaterials(std::vector<float>& out_rgb);

// Transform::ToMatrix followed by glm::inverse: the matrix EvaluatorTransform::Compile emits (sdf_evaluator.cpp:409-429).
Mat4 CompiledInverseMatrix(const struct Node& n);

// Node storage of a pool: its own nodes, optionally BEHIND the first base_count nodes of another pool's storage, which
// must outlive this one and not grow while this one is in use.  Octree workers prune against their parent's nodes
// this way without copying them (node indices mean the same thing in the parent and in every pool laid over it).
class NodeStore
{
public:
	const NodeStore* base = nullptr;
	uint32_t base_count = 0;
	std::vector<Node> own;

	size_t size() const { return size_t(base_count) + own.size(); }
	const Node& operator[](size_t i) const
	{
		const NodeStore* s = this;
		while (i < s->base_count) s = s->base;
		return s->own[i - s->base_count];
	}
	// (nodes of the base are never modified through an overlay; the reference is mutable for the pool's own nodes)
	Node& operator[](size_t i) { return const_cast<Node&>(static_cast<const NodeStore&>(*this)[i]); }
	// iteration covers the pool's own nodes: whole trees (no base) are the only things iterated
	std::vector<Node>::iterator begin() { return own.begin(); }
	std::vector<Node>::iterator end() { return own.end(); }
	std::vector<Node>::const_iterator begin() const { return own.begin(); }
	std::vector<Node>::const_iterator end() const { return own.end(); }
	void push_back(const Node& n) { own.push_back(n); }
	void reserve(size_t n) { own.reserve(n > base_count ? n - base_count : 0); }
};

// Growable node pool.  Trees own one; octree workers lay theirs over their parent's (Overlay) and grow as they prune.
struct NodePool
{
	NodeStore nodes;

	// Makes this (empty) pool a view of `parent`'s nodes plus whatever gets added here.
	void Overlay(const NodePool& parent)
	{
		nodes.base = &parent.nodes;
		nodes.base_count = uint32_t(parent.nodes.size());
		nodes.own.clear();
	}

	uint32_t Add(const Node& n);
	// SetNode constructor (sdf_evaluator.cpp:737-772) with its left-leaning operand swap.
	uint32_t AddSet(uint32_t kind, uint32_t lhs, uint32_t rhs, float threshold);
	uint32_t AddFlate(uint32_t child, float radius);
	uint32_t AddStencil(uint32_t kind, uint32_t child, uint32_t mask, uint32_t material);

	float Eval(uint32_t index, Vec3 point) const;                 // virtual Eval
	// top_value (optional): receives this node's own value at `point`, which the clip computes anyway
	uint32_t Clip(uint32_t index, Vec3 point, float radius, float* top_value = nullptr);      // virtual Clip; kNoNode when pruned away
	bool Equal(uint32_t x, uint32_t y) const;                     // operator==
	Box3 Bounds(uint32_t index) const;                            // Bounds()
	Box3 InnerBounds(uint32_t index) const;                       // InnerBounds()
	uint32_t Material(uint32_t index, Vec3 point) const;          // GetMaterial -> material id
	Vec3 Gradient(uint32_t index, Vec3 point) const;              // SDFNode::Gradient
	// `inverse` (optional): CompiledInverseMatrix of every brush node, indexed like `nodes` (brushes are immutable and
	// shared by all pruned programs, so the octree flattener inverts each of them once instead of once per occurrence)
	void CompileReference(uint32_t index, std::vector<uint32_t>& words, const Mat4* inverse = nullptr) const; // Compile(ProgramBuffer&)

private:
	void Derive(uint32_t index);
	// One Clip call evaluates the same point at every set node it visits, and the reference's recursion re-evaluates the
	// whole subtree each time (quadratic in the length of a left-leaning chain: 500k brush evaluations for one clip of
	// a 1000-primitive tree).  The values are the same numbers whichever call computes them, so a clip keeps them.
	uint32_t ClipRec(uint32_t index, Vec3 point, float radius, float* top_value);
	float EvalMemo(uint32_t index, Vec3 point);
};

class Tree
{
public:
	NodePool pool;
	uint32_t root = kNoNode;

	bool Valid() const { return root != kNoNode; }

	// SDF:: brush constructors (sdf_evaluator.cpp:1206-1317); same argument meaning.
	static Tree Sphere(float radius);
	static Tree Ellipsoid(float rx, float ry, float rz);
	static Tree Box(float ex, float ey, float ez);
	static Tree Torus(float major_radius, float minor_radius);
	static Tree Cylinder(float radius, float extent);
	static Tree Plane(float nx, float ny, float nz);
	static Tree Cone(float radius, float height);
	static Tree Coninder(float radius_l, float radius_h, float height);
	// SDF:: set operators (sdf_evaluator.cpp:1320-1371); kind is one of kKindUnion..kKindBlendDiff.
	static Tree Combine(uint32_t kind, const Tree& lhs, const Tree& rhs, float threshold);
	static Tree Flate(const Tree& child, float radius);
	static Tree Stencil(const Tree& child, const Tree& mask, uint32_t material, bool apply_to_negative);
	// this = kind(this, rhs) without re-copying this tree: linear-time left folds (Lua variadic operators, lua_sdf.cpp:354-359).
	void Fold(uint32_t kind, const Tree& rhs, float threshold);

	// In-place modifiers (SDFNode::Move/Rotate/Scale/ApplyMaterial, SDF::Align/RotateX/Y/Z).
	void Move(Vec3 offset);
	void Rotate(Quat rotation);
	void RotateX(float degrees);
	void RotateY(float degrees);
	void RotateZ(float degrees);
	void Scale(float scale);
	void Paint(uint32_t material, bool force);
	void Align(Vec3 anchors);

	float Eval(Vec3 p) const { return pool.Eval(root, p); }
	Box3 Bounds() const { return pool.Bounds(root); }
	bool HasPaint() const { return pool.nodes[root].has_paint; }
	bool HasFiniteBounds() const { return pool.nodes[root].finite; }
	int LeafCount() const { return pool.nodes[root].leaf_count; }
	uint32_t StackSize() const { return pool.nodes[root].stack_size; }

	static bool LoadTgm(const std::string& path, Tree& out, std::string& error);
	bool SaveTgm(const std::string& path, std::string& error) const;

private:
	static Tree Brush(uint32_t kind, const float* params, int count, Box3 bounds);
	uint32_t Append(const Tree& other); // copies other's nodes, returns its new root index
};

} // namespace tg
