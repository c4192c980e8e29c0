// Host CSG tree: construction, evaluation, Lipschitz pruning, bounds, reference-format compilation.
// See tg_tree.h.  Line references are to the reference's tangerine/sdf_evaluator.cpp unless noted.
#include "tg_tree.h"

#include <cstdio>
#include <mutex>

#include "tg_sdf.h"

namespace tg
{

// ------------------------------------------------------------------------------------------------
// Material table
// ------------------------------------------------------------------------------------------------

namespace
{
std::mutex g_material_lock;
std::vector<float> g_material_rgb;
}

uint32_t RegisterMaterial(float r, float g, float b)
{
	std::lock_guard<std::mutex> lock(g_material_lock);
	g_material_rgb.push_back(r);
	g_material_rgb.push_back(g);
	g_material_rgb.push_back(b);
	return uint32_t(g_material_rgb.size() / 3 - 1);
}

bool MaterialColor(uint32_t id, float out_rgb[3])
{
	std::lock_guard<std::mutex> lock(g_material_lock);
	if (size_t(id) * 3 + 2 >= g_material_rgb.size())
	{
		return false;
	}
	for (int i = 0; i < 3; ++i)
	{
		out_rgb[i] = g_material_rgb[size_t(id) * 3 + i];
	}
	return true;
}

uint32_t MaterialCount()
{
	std::lock_guard<std::mutex> lock(g_material_lock);
	return uint32_t(g_material_rgb.size() / 3);
}

void SnapshotMaterials(std::vector<float>& out_rgb)
{
	std::lock_guard<std::mutex> lock(g_material_lock);
	out_rgb = g_material_rgb;
}

// ------------------------------------------------------------------------------------------------
// NodePool
// ------------------------------------------------------------------------------------------------

// StackSize (:448, :595, :771, :1056), LeafCount, HasPaint, HasFiniteBounds (:549-562, :681-694, :1014-1027, :1135-1148)
void NodePool::Derive(uint32_t index)
{
	Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		n.stack_size = 1;
		n.leaf_count = 1;
		n.has_paint = n.material != kNoMaterial;
		n.finite = true;
		for (int i = 0; i < 3; ++i)
		{
			if (std::isinf(n.local_bounds.min[i]) || std::isinf(n.local_bounds.max[i]))
			{
				n.finite = false;
			}
		}
	}
	else if (IsSet(n.kind))
	{
		const Node& l = nodes[n.a];
		const Node& r = nodes[n.b];
		n.stack_size = l.stack_size > r.stack_size + 1 ? l.stack_size : r.stack_size + 1;
		n.leaf_count = l.leaf_count + r.leaf_count;
		n.has_paint = l.has_paint || r.has_paint;
		n.finite = l.finite || r.finite;
	}
	else
	{
		const Node& c = nodes[n.a];
		n.stack_size = c.stack_size;
		n.leaf_count = c.leaf_count;
		n.has_paint = IsStencil(n.kind) ? true : c.has_paint;
		n.finite = c.finite;
	}
}

uint32_t NodePool::Add(const Node& n)
{
	nodes.push_back(n);
	uint32_t index = uint32_t(nodes.size() - 1);
	Derive(index);
	return index;
}

uint32_t NodePool::AddSet(uint32_t kind, uint32_t lhs, uint32_t rhs, float threshold)
{
	// :761-769 -- keep the tree left leaning so the interpreter's stack stays shallow
	if (SetFamily(kind) != Family::Diff && nodes[rhs].stack_size > nodes[lhs].stack_size)
	{
		uint32_t t = lhs;
		lhs = rhs;
		rhs = t;
	}
	Node n;
	n.kind = kind;
	n.a = lhs;
	n.b = rhs;
	n.params[0] = threshold;
	return Add(n);
}

uint32_t NodePool::AddFlate(uint32_t child, float radius)
{
	Node n;
	n.kind = kKindFlate;
	n.a = child;
	n.params[0] = radius;
	return Add(n);
}

uint32_t NodePool::AddStencil(uint32_t kind, uint32_t child, uint32_t mask, uint32_t material)
{
	Node n;
	n.kind = kind;
	n.a = child;
	n.b = mask;
	n.material = material;
	return Add(n);
}

// Transform::ApplyInv (tangerine/transform.cpp:64-67)
static inline Vec3 ApplyInv(const Node& n, Vec3 p)
{
	return Rotate(Inverse(n.rotation), p - n.translation) / n.scalation;
}

// Transform::Apply (tangerine/transform.cpp:58-61)
static inline Vec3 ApplyFwd(const Node& n, Vec3 p)
{
	return Rotate(n.rotation, p * n.scalation) + n.translation;
}

// BrushNode::Eval :463-466, SetNode::Eval :774-780, FlateNode::Eval :1059-1062, StencilMaskNode::Eval :598-601
float NodePool::Eval(uint32_t index, Vec3 point) const
{
	const Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		Vec3 local = ApplyInv(n, point);
		return sdf::Brush(n.kind, n.params, local.x, local.y, local.z) * n.scalation;
	}
	if (IsSet(n.kind))
	{
		float l = Eval(n.a, point);
		float r = Eval(n.b, point);
		return sdf::SetOp(n.kind - 8, l, r, n.params[0]);
	}
	if (n.kind == kKindFlate)
	{
		return Eval(n.a, point) - n.params[0];
	}
	return Eval(n.a, point);
}

// Clip: brush :468-478, set :782-850, flate :1064-1075, stencil :603-615.  Brushes are immutable, so
// the reference's Copy() of a surviving brush is the brush's own index here.
// The values a clip keeps live in per-THREAD arrays indexed by node: one clip runs on one thread from start to end, and
// the octree build makes hundreds of thousands of short-lived pools (one per task), which must not each allocate and
// zero arrays as long as the whole tree.
namespace
{
thread_local std::vector<float> memo_value;
thread_local std::vector<uint32_t> memo_stamp;
thread_local uint32_t memo_epoch = 0;
} // namespace

uint32_t NodePool::Clip(uint32_t index, Vec3 point, float radius, float* top_value)
{
	// a new epoch invalidates the values kept by the previous clip (another point, possibly another pool)
	if (++memo_epoch == 0)
	{
		std::fill(memo_stamp.begin(), memo_stamp.end(), 0u);
		memo_epoch = 1;
	}
	return ClipRec(index, point, radius, top_value);
}

// NodePool::Eval with the values of this clip's point kept per node: the same arithmetic in the same order, once.
float NodePool::EvalMemo(uint32_t index, Vec3 point)
{
	if (index < memo_stamp.size() && memo_stamp[index] == memo_epoch)
	{
		return memo_value[index];
	}
	const Node& n = nodes[index];
	float value;
	if (IsBrush(n.kind))
	{
		Vec3 local = ApplyInv(n, point);
		value = sdf::Brush(n.kind, n.params, local.x, local.y, local.z) * n.scalation;
	}
	else if (IsSet(n.kind))
	{
		const uint32_t a = n.a, b = n.b, op = n.kind - 8;
		const float threshold = n.params[0];
		float l = EvalMemo(a, point);
		float r = EvalMemo(b, point);
		value = sdf::SetOp(op, l, r, threshold);
	}
	else if (n.kind == kKindFlate)
	{
		const float radius = n.params[0];
		value = EvalMemo(n.a, point) - radius;
	}
	else
	{
		value = EvalMemo(n.a, point);
	}
	if (index >= memo_stamp.size())
	{
		const size_t size = std::max<size_t>(nodes.size() + nodes.size() / 2, size_t(index) + 1);
		memo_stamp.resize(size, 0u);
		memo_value.resize(size, 0.0f);
	}
	memo_stamp[index] = memo_epoch;
	memo_value[index] = value;
	return value;
}

uint32_t NodePool::ClipRec(uint32_t index, Vec3 point, float radius, float* top_value)
{
	const uint32_t kind = nodes[index].kind;
	if (IsBrush(kind))
	{
		const float value = EvalMemo(index, point);
		if (top_value) *top_value = value;
		return value <= radius ? index : kNoNode;
	}
	if (IsSet(kind))
	{
		const float value = EvalMemo(index, point);
		if (top_value) *top_value = value;
		if (!(value <= radius))
		{
			return kNoNode;
		}
		const uint32_t lhs = nodes[index].a;
		const uint32_t rhs = nodes[index].b;
		const float threshold = nodes[index].params[0];
		const Family family = SetFamily(kind);
		if (IsBlend(kind))
		{
			// Inside the blending region both operands must survive a clip widened by the threshold.
			uint32_t nl = ClipRec(lhs, point, radius + threshold, nullptr);
			uint32_t nr = ClipRec(rhs, point, radius + threshold, nullptr);
			if (nl != kNoNode && nr != kNoNode)
			{
				// (both operands came through whole: the node the reference would build anew is this very node)
				return (nl == lhs && nr == rhs) ? index : AddSet(kind, nl, nr, threshold);
			}
			if (family == Family::Inter)
			{
				return kNoNode;
			}
		}
		uint32_t nl = ClipRec(lhs, point, radius, nullptr);
		uint32_t nr = ClipRec(rhs, point, radius, nullptr);
		if (nl != kNoNode && nr != kNoNode)
		{
			return (nl == lhs && nr == rhs) ? index : AddSet(kind, nl, nr, threshold);
		}
		if (family == Family::Union)
		{
			return nl != kNoNode ? nl : nr;
		}
		if (family == Family::Diff)
		{
			return nl;
		}
		return kNoNode;
	}
	if (kind == kKindFlate)
	{
		const float value = EvalMemo(index, point);
		if (top_value) *top_value = value;
		if (!(value <= radius))
		{
			return kNoNode;
		}
		const float flate = nodes[index].params[0];
		uint32_t child = ClipRec(nodes[index].a, point, radius + flate, nullptr);
		return child == kNoNode ? kNoNode : (child == nodes[index].a ? index : AddFlate(child, flate));
	}
	if (top_value) *top_value = EvalMemo(index, point);
	uint32_t child = ClipRec(nodes[index].a, point, radius, nullptr);
	return child == kNoNode ? kNoNode : (child == nodes[index].a ? index : AddStencil(kind, child, nodes[index].b, nodes[index].material));
}

// operator== of the node classes (:564-579, :696-709, :1029-1037, :1150-1154)
bool NodePool::Equal(uint32_t x, uint32_t y) const
{
	if (x == y)
	{
		return true;
	}
	const Node& p = nodes[x];
	const Node& q = nodes[y];
	if (p.kind != q.kind)
	{
		return false;
	}
	if (IsBrush(p.kind))
	{
		if (p.material != q.material || !(p.rotation == q.rotation) || !(p.translation == q.translation) || p.scalation != q.scalation)
		{
			return false;
		}
		for (int i = 0; i < BrushParamCount(p.kind); ++i)
		{
			if (p.params[i] != q.params[i])
			{
				return false;
			}
		}
		return true;
	}
	if (IsSet(p.kind))
	{
		return p.params[0] == q.params[0] && Equal(p.a, q.a) && Equal(p.b, q.b);
	}
	if (p.kind == kKindFlate)
	{
		return p.params[0] == q.params[0] && Equal(p.a, q.a);
	}
	return Equal(p.a, q.a) && Equal(p.b, q.b) && p.material == q.material;
}

// EvaluatorTransform::Apply(AABB) :367-406
static Box3 BrushBounds(const Node& n)
{
	const Vec3 a = n.local_bounds.min;
	const Vec3 b = n.local_bounds.max;
	if (n.rotation.IsIdentity())
	{
		return { (a * n.scalation) + n.translation, (b * n.scalation) + n.translation };
	}
	const Vec3 points[7] = { b, Vec3(b.x, a.y, a.z), Vec3(a.x, b.y, a.z), Vec3(a.x, a.y, b.z), Vec3(a.x, b.y, b.z), Vec3(b.x, a.y, b.z), Vec3(b.x, b.y, a.z) };
	Box3 out;
	out.min = ApplyFwd(n, a);
	out.max = out.min;
	for (const Vec3& p : points)
	{
		Vec3 t = ApplyFwd(n, p);
		out.min = GlmMin(out.min, t);
		out.max = GlmMax(out.max, t);
	}
	return out;
}

static Box3 CombineBounds(Family family, const Box3& l, const Box3& r)
{
	Box3 c;
	if (family == Family::Union)
	{
		c.min = GlmMin(l.min, r.min);
		c.max = GlmMax(l.max, r.max);
	}
	else if (family == Family::Diff)
	{
		c = l;
	}
	else
	{
		c.min = GlmMax(l.min, r.min);
		c.max = GlmMin(l.max, r.max);
	}
	return c;
}

// Bounds(): :485-488, :622-625, :857-889, :1082-1088
Box3 NodePool::Bounds(uint32_t index) const
{
	const Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		return BrushBounds(n);
	}
	if (IsSet(n.kind))
	{
		Box3 l = Bounds(n.a);
		Box3 r = Bounds(n.b);
		Box3 c = CombineBounds(SetFamily(n.kind), l, r);
		if (IsBlend(n.kind))
		{
			Vec3 t(n.params[0]);
			Box3 liminal = { GlmMax(l.min, r.min) - t, GlmMin(l.max, r.max) + t };
			c.min = GlmMin(c.min, liminal.min);
			c.max = GlmMax(c.max, liminal.max);
		}
		return c;
	}
	if (n.kind == kKindFlate)
	{
		Box3 c = Bounds(n.a);
		Vec3 pad(n.params[0] * 2);
		c.max = c.max + pad;
		c.min = c.min - pad;
		return c;
	}
	return Bounds(n.a);
}

// InnerBounds(): :490-493, :627-630, :891-913, :1090-1096
Box3 NodePool::InnerBounds(uint32_t index) const
{
	const Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		return BrushBounds(n);
	}
	if (IsSet(n.kind))
	{
		return CombineBounds(SetFamily(n.kind), InnerBounds(n.a), InnerBounds(n.b));
	}
	if (n.kind == kKindFlate)
	{
		Box3 c = InnerBounds(n.a);
		Vec3 pad(n.params[0] * 2);
		c.max = c.max + pad;
		c.min = c.min - pad;
		return c;
	}
	return InnerBounds(n.a);
}

// GetMaterial: brush :537-547, stencil :666-679, set :957-1012, flate :1130-1133
uint32_t NodePool::Material(uint32_t index, Vec3 p) const
{
	const Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		return n.material;
	}
	if (n.kind == kKindFlate)
	{
		return Material(n.a, p);
	}
	if (IsStencil(n.kind))
	{
		const bool interior = Eval(n.b, p) < 0.0f;
		return interior == (n.kind == kKindStencilNeg) ? n.material : Material(n.a, p);
	}
	const Family family = SetFamily(n.kind);
	if (family == Family::Diff)
	{
		return Material(n.a, p);
	}
	const float el = Eval(n.a, p);
	const float er = Eval(n.b, p);
	const float dist = sdf::SetOp(n.kind - 8, el, er, n.params[0]);
	const bool take_left = IsBlend(n.kind) ? (std::fabs(el - dist) <= std::fabs(er - dist)) : (dist == el);
	if (family == Family::Union)
	{
		return take_left ? Material(n.a, p) : Material(n.b, p);
	}
	const uint32_t sl = Material(n.a, p);
	const uint32_t sr = Material(n.b, p);
	const bool lv = nodes[n.a].has_paint;
	const bool rv = nodes[n.b].has_paint;
	if (lv && rv)
	{
		return take_left ? sl : sr;
	}
	return lv ? sl : sr;
}

// SDFNode::Gradient :298-333 (tetrahedral taps, forward-difference fallback)
Vec3 NodePool::Gradient(uint32_t index, Vec3 p) const
{
	const float almost_zero = 0.0001f;
	const float ox = 1.0f * almost_zero;
	const float oy = -1.0f * almost_zero;
	const Vec3 xyy(ox, oy, oy), yyx(oy, oy, ox), yxy(oy, ox, oy), xxx(ox, ox, ox);
	Vec3 g = xyy * Eval(index, p + xyy) + yyx * Eval(index, p + yyx) + yxy * Eval(index, p + yxy) + xxx * Eval(index, p + xxx);
	float len_sq = Dot(g, g);
	if (len_sq == 0.0f)
	{
		float d = Eval(index, p);
		Vec3 f(Eval(index, p + xyy) - d, Eval(index, p + yxy) - d, Eval(index, p + yyx) - d);
		return f * (1.0f / std::sqrt(Dot(f, f)));
	}
	return g / std::sqrt(len_sq);
}

static inline void PushFloat(std::vector<uint32_t>& words, float f)
{
	words.push_back(FloatBits(f));
}

// Transform::ToMatrix (tangerine/transform.cpp:49-55) followed by glm::inverse, as EvaluatorTransform::Compile does (:409-429)
Mat4 CompiledInverseMatrix(const Node& n)
{
	Mat4 rotation = ToMat4(n.rotation);
	Mat4 translation = Translate(n.translation);
	Mat4 scalation = ScaleSlow(translation, Vec3(n.scalation));
	return Inverse(scalation * rotation);
}

// Compile(ProgramBuffer&): brush :495-507, set :915-924, flate :1098-1103, stencil :632-635.
// Produces the reference's own word stream (opcode or float per word); used for hashing / parity only.
void NodePool::CompileReference(uint32_t index, std::vector<uint32_t>& words, const Mat4* inverse) const
{
	const Node& n = nodes[index];
	if (IsBrush(n.kind))
	{
		const bool has_rotation = !n.rotation.IsIdentity();
		const bool has_scalation = n.scalation != 1.0f;
		const bool has_translation = !(n.translation == Vec3(0.0f, 0.0f, 0.0f));
		if (has_rotation || has_scalation)
		{
			Mat4 inv = inverse ? inverse[index] : CompiledInverseMatrix(n);
			words.push_back(17); // OpcodeT::Matrix
			for (int c = 0; c < 4; ++c)
			{
				for (int r = 0; r < 4; ++r)
				{
					PushFloat(words, inv.m[c][r]);
				}
			}
		}
		else if (has_translation)
		{
			words.push_back(16); // OpcodeT::Offset
			PushFloat(words, -n.translation.x);
			PushFloat(words, -n.translation.y);
			PushFloat(words, -n.translation.z);
		}
		words.push_back(n.kind);
		for (int i = 0; i < BrushParamCount(n.kind); ++i)
		{
			PushFloat(words, n.params[i]);
		}
		if (has_scalation)
		{
			words.push_back(18); // OpcodeT::ScaleField
			PushFloat(words, n.scalation);
		}
	}
	else if (IsSet(n.kind))
	{
		CompileReference(n.a, words, inverse);
		CompileReference(n.b, words, inverse);
		words.push_back(n.kind);
		if (IsBlend(n.kind))
		{
			PushFloat(words, n.params[0]);
		}
	}
	else if (n.kind == kKindFlate)
	{
		CompileReference(n.a, words, inverse);
		words.push_back(kKindFlate);
		PushFloat(words, n.params[0]);
	}
	else
	{
		CompileReference(n.a, words, inverse);
	}
}

// ------------------------------------------------------------------------------------------------
// Tree
// ------------------------------------------------------------------------------------------------

static Box3 Symmetrical(Vec3 high) // SymmetricalBounds :1163-1166
{
	return { high * Vec3(-1.0f), high };
}

Tree Tree::Brush(uint32_t kind, const float* params, int count, Box3 bounds)
{
	Tree t;
	Node n;
	n.kind = kind;
	for (int i = 0; i < count; ++i)
	{
		n.params[i] = params[i];
	}
	n.local_bounds = bounds;
	t.root = t.pool.Add(n);
	return t;
}

Tree Tree::Sphere(float radius) // :1206-1214
{
	return Brush(kKindSphere, &radius, 1, Symmetrical(Vec3(radius)));
}

Tree Tree::Ellipsoid(float rx, float ry, float rz) // :1216-1225
{
	float p[3] = { rx, ry, rz };
	return Brush(kKindEllipsoid, p, 3, Symmetrical(Vec3(rx, ry, rz)));
}

Tree Tree::Box(float ex, float ey, float ez) // :1227-1236
{
	float p[3] = { ex, ey, ez };
	return Brush(kKindBox, p, 3, Symmetrical(Vec3(ex, ey, ez)));
}

Tree Tree::Torus(float major_radius, float minor_radius) // :1238-1248
{
	float p[2] = { major_radius, minor_radius };
	float radius = major_radius + minor_radius;
	return Brush(kKindTorus, p, 2, Symmetrical(Vec3(radius, radius, minor_radius)));
}

Tree Tree::Cylinder(float radius, float extent) // :1250-1258
{
	float p[2] = { radius, extent };
	return Brush(kKindCylinder, p, 2, Symmetrical(Vec3(radius, radius, extent)));
}

Tree Tree::Plane(float nx, float ny, float nz) // :1260-1294
{
	Vec3 v(nx, ny, nz);
	Vec3 normal = v * (1.0f / std::sqrt(Dot(v, v))); // glm::normalize
	float p[3] = { normal.x, normal.y, normal.z };
	const float inf = INFINITY;
	Box3 unbound = Symmetrical(Vec3(inf, inf, inf));
	if (normal.x == -1.0f) unbound.min.x = 0.0f;
	else if (normal.x == 1.0f) unbound.max.x = 0.0f;
	else if (normal.y == -1.0f) unbound.min.y = 0.0f;
	else if (normal.y == 1.0f) unbound.max.y = 0.0f;
	else if (normal.z == -1.0f) unbound.min.z = 0.0f;
	else if (normal.z == 1.0f) unbound.max.z = 0.0f;
	return Brush(kKindPlane, p, 3, unbound);
}

Tree Tree::Cone(float radius, float height) // :1296-1305
{
	float tangent = radius / height;
	float p[2] = { tangent, height };
	return Brush(kKindCone, p, 2, Symmetrical(Vec3(radius, radius, float(height * .5))));
}

Tree Tree::Coninder(float radius_l, float radius_h, float height) // :1307-1317
{
	float half_height = float(height * .5);
	float p[3] = { radius_l, radius_h, half_height };
	float max_radius = std::fmax(radius_l, radius_h);
	return Brush(kKindConinder, p, 3, Symmetrical(Vec3(max_radius, max_radius, half_height)));
}

uint32_t Tree::Append(const Tree& other)
{
	const uint32_t base = uint32_t(pool.nodes.size());
	pool.nodes.reserve(pool.nodes.size() + other.pool.nodes.size());
	for (Node n : other.pool.nodes)
	{
		if (n.a != kNoNode) n.a += base;
		if (n.b != kNoNode) n.b += base;
		pool.nodes.push_back(n);
	}
	return other.root + base;
}

Tree Tree::Combine(uint32_t kind, const Tree& lhs, const Tree& rhs, float threshold)
{
	Tree t;
	uint32_t l = t.Append(lhs);
	uint32_t r = t.Append(rhs);
	t.root = t.pool.AddSet(kind, l, r, IsBlend(kind) ? threshold : 0.0f);
	return t;
}

void Tree::Fold(uint32_t kind, const Tree& rhs, float threshold)
{
	const uint32_t lhs_root = root;
	const uint32_t rhs_root = Append(rhs);
	root = pool.AddSet(kind, lhs_root, rhs_root, IsBlend(kind) ? threshold : 0.0f);
}

Tree Tree::Flate(const Tree& child, float radius)
{
	Tree t;
	uint32_t c = t.Append(child);
	t.root = t.pool.AddFlate(c, radius);
	return t;
}

Tree Tree::Stencil(const Tree& child, const Tree& mask, uint32_t material, bool apply_to_negative)
{
	Tree t;
	uint32_t c = t.Append(child);
	uint32_t m = t.Append(mask);
	t.root = t.pool.AddStencil(apply_to_negative ? kKindStencilNeg : kKindStencilPos, c, m, material);
	return t;
}

// Every node of a Tree is reachable from its root (trees only grow by Append), and the reference's
// Move / Rotate / Scale recurse into every child including stencil masks (:637-653, :926-943, :1105-1118),
// so the modifiers can sweep the node array.
void Tree::Move(Vec3 offset)
{
	for (Node& n : pool.nodes)
	{
		if (IsBrush(n.kind))
		{
			n.translation = n.translation + offset; // Transform::Move
		}
	}
}

void Tree::Rotate(Quat rotation)
{
	for (Node& n : pool.nodes)
	{
		if (IsBrush(n.kind))
		{
			n.translation = tg::Rotate(rotation, n.translation); // Transform::Rotate
			n.rotation = rotation * n.rotation;
		}
	}
}

void Tree::Scale(float scale)
{
	for (Node& n : pool.nodes)
	{
		if (IsBrush(n.kind))
		{
			n.translation = n.translation * scale; // Transform::Scale
			n.scalation *= scale;
		}
		else if (IsSet(n.kind))
		{
			n.params[0] *= scale; // SetNode::Scale :938-943
		}
	}
}

static Quat AxisQuat(float degrees, int axis) // SDF::RotateX/Y/Z :1180-1202
{
	float r = float((degrees * 0.01745329251994329576923690768489f) * .5);
	float s = std::sin(r);
	float c = std::cos(r);
	Quat q;
	q.w = c;
	q.x = axis == 0 ? s : 0.0f;
	q.y = axis == 1 ? s : 0.0f;
	q.z = axis == 2 ? s : 0.0f;
	return q;
}

void Tree::RotateX(float degrees) { Rotate(AxisQuat(degrees, 0)); }
void Tree::RotateY(float degrees) { Rotate(AxisQuat(degrees, 1)); }
void Tree::RotateZ(float degrees) { Rotate(AxisQuat(degrees, 2)); }

static void PaintNode(NodePool& pool, uint32_t index, uint32_t material, bool force)
{
	Node& n = pool.nodes[index];
	if (IsBrush(n.kind))
	{
		if (n.material == kNoMaterial || force) // BrushNode::ApplyMaterial :524-530
		{
			n.material = material;
		}
	}
	else if (IsSet(n.kind))
	{
		PaintNode(pool, n.a, material, force);
		PaintNode(pool, n.b, material, force);
	}
	else if (n.kind == kKindFlate)
	{
		PaintNode(pool, n.a, material, force);
	}
	// StencilMaskNode::ApplyMaterial is a no-op (:655-658)
}

void Tree::Paint(uint32_t material, bool force)
{
	PaintNode(pool, root, material, force);
	// has_paint is derived bottom-up; recompute in storage (post-)order
	for (uint32_t i = 0; i < pool.nodes.size(); ++i)
	{
		Node& n = pool.nodes[i];
		if (IsBrush(n.kind)) n.has_paint = n.material != kNoMaterial;
		else if (IsSet(n.kind)) n.has_paint = pool.nodes[n.a].has_paint || pool.nodes[n.b].has_paint;
		else if (n.kind == kKindFlate) n.has_paint = pool.nodes[n.a].has_paint;
	}
}

void Tree::Align(Vec3 anchors) // SDF::Align :1171-1177
{
	Vec3 alignment = anchors * Vec3(0.5f) + Vec3(0.5f);
	Box3 bounds = pool.InnerBounds(root);
	Vec3 offset = Mix(bounds.min, bounds.max, alignment) * Vec3(-1.0f);
	Move(offset);
}

bool Tree::LoadTgm(const std::string& path, Tree& out, std::string& error)
{
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f)
	{
		error = "cannot open " + path;
		return false;
	}
	uint32_t header[4];
	bool ok = std::fread(header, 4, 4, f) == 4 && header[0] == 0x314D4754u;
	if (ok)
	{
		// the counts come from the file: they must account for its size exactly before anything is sized by them
		long size = -1;
		if (std::fseek(f, 0, SEEK_END) == 0) size = std::ftell(f);
		ok = size >= 16 && uint64_t(size) == 16ull + uint64_t(header[2]) * 12ull + uint64_t(header[1]) * sizeof(TgmRecord) && std::fseek(f, 16, SEEK_SET) == 0;
	}
	std::vector<float> colors;
	std::vector<TgmRecord> records;
	if (ok)
	{
		colors.resize(size_t(header[2]) * 3);
		records.resize(header[1]);
		ok = std::fread(colors.data(), 12, header[2], f) == header[2] && std::fread(records.data(), sizeof(TgmRecord), header[1], f) == header[1];
	}
	std::fclose(f);
	if (!ok || header[3] >= header[1])
	{
		error = "malformed .tgm file " + path;
		return false;
	}
	std::vector<uint32_t> material_ids(header[2]);
	for (uint32_t i = 0; i < header[2]; ++i)
	{
		material_ids[i] = RegisterMaterial(colors[i * 3], colors[i * 3 + 1], colors[i * 3 + 2]);
	}
	out = Tree();
	for (uint32_t i = 0; i < header[1]; ++i)
	{
		const TgmRecord& r = records[i];
		Node n;
		n.kind = r.kind;
		n.a = r.a;
		n.b = r.b;
		const bool valid_kind = IsBrush(r.kind) || IsSet(r.kind) || r.kind == kKindFlate || IsStencil(r.kind);
		const bool children_ok = IsBrush(r.kind) || (r.a < i && (r.kind == kKindFlate || r.b < i));
		if (!valid_kind || !children_ok || (r.material != kNoMaterial && r.material >= header[2]))
		{
			error = "malformed node in " + path;
			return false;
		}
		n.material = r.material == kNoMaterial ? kNoMaterial : material_ids[r.material];
		for (int k = 0; k < 4; ++k) n.params[k] = r.params[k];
		n.rotation.w = r.quat[0];
		n.rotation.x = r.quat[1];
		n.rotation.y = r.quat[2];
		n.rotation.z = r.quat[3];
		n.translation = Vec3(r.trans[0], r.trans[1], r.trans[2]);
		n.scalation = r.scale;
		n.local_bounds.min = Vec3(r.bounds_min[0], r.bounds_min[1], r.bounds_min[2]);
		n.local_bounds.max = Vec3(r.bounds_max[0], r.bounds_max[1], r.bounds_max[2]);
		out.pool.Add(n); // records are post-order, so children are already derived
	}
	out.root = header[3];
	return true;
}

bool Tree::SaveTgm(const std::string& path, std::string& error) const
{
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f)
	{
		error = "cannot open " + path;
		return false;
	}
	// Materials are re-indexed densely in first-use order.
	std::vector<uint32_t> used;
	std::vector<TgmRecord> records(pool.nodes.size());
	for (size_t i = 0; i < pool.nodes.size(); ++i)
	{
		const Node& n = pool.nodes[i];
		TgmRecord& r = records[i];
		std::memset(&r, 0, sizeof(r));
		r.kind = n.kind;
		r.a = n.a;
		r.b = n.b;
		r.material = kNoMaterial;
		if (n.material != kNoMaterial)
		{
			size_t slot = 0;
			while (slot < used.size() && used[slot] != n.material) ++slot;
			if (slot == used.size()) used.push_back(n.material);
			r.material = uint32_t(slot);
		}
		for (int k = 0; k < 4; ++k) r.params[k] = n.params[k];
		r.quat[0] = n.rotation.w;
		r.quat[1] = n.rotation.x;
		r.quat[2] = n.rotation.y;
		r.quat[3] = n.rotation.z;
		for (int k = 0; k < 3; ++k)
		{
			r.trans[k] = n.translation[k];
			r.bounds_min[k] = n.local_bounds.min[k];
			r.bounds_max[k] = n.local_bounds.max[k];
		}
		r.scale = n.scalation;
	}
	uint32_t header[4] = { 0x314D4754u, uint32_t(records.size()), uint32_t(used.size()), root };
	std::fwrite(header, 4, 4, f);
	for (uint32_t id : used)
	{
		float rgb[3] = { 1.0f, 1.0f, 1.0f };
		MaterialColor(id, rgb);
		std::fwrite(rgb, 4, 3, f);
	}
	std::fwrite(records.data(), sizeof(TgmRecord), records.size(), f);
	std::fclose(f);
	return true;
}

} // namespace tg
