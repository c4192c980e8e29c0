/*
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.  See tg_oracle.h.
 *
 * Plain-C restatement of the reference CPU algorithm.  Paths below are relative to the reference
 * repository root.  Arithmetic is written so that every float/double promotion of the C++
 * original is preserved (x86-64 SSE2, FLT_EVAL_METHOD 0, no FMA contraction: built with
 * -ffp-contract=off), which makes the results bit-identical to the compiled reference.
 */
#define _GNU_SOURCE
#include "tg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* Small vector helpers mirroring the glm 0.9.9.8 scalar code paths the reference compiles to.  */
/* ------------------------------------------------------------------------------------------- */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float w, x, y, z; } quat;

static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 divs3(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
/* glm/detail/func_geometric.inl:48-55 -- tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot2(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline float len2(v2 a) { return sqrtf(dot2(a, a)); }
/* glm::min / glm::max (func_common.inl): min = (y < x) ? y : x, max = (x < y) ? y : x */
static inline float gmin(float x, float y) { return (y < x) ? y : x; }
static inline float gmax(float x, float y) { return (x < y) ? y : x; }
static inline v3 gmin3(v3 a, v3 b) { return V3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
static inline v3 gmax3(v3 a, v3 b) { return V3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
/* glm::sign for floats (func_common.inl:144-150) */
static inline float gsign(float x) { return (float)(0.0f < x) - (float)(x < 0.0f); }
/* func_geometric.inl:68-79 */
static inline v3 cross3(v3 x, v3 y)
{
	return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
/* detail/type_quat.inl:343-350 -- q * v */
static inline v3 quat_rotate(quat q, v3 v)
{
	v3 qv = V3(q.x, q.y, q.z);
	v3 uv = cross3(qv, v);
	v3 uuv = cross3(qv, uv);
	return add3(v, muls3(add3(muls3(uv, q.w), uuv), 2.0f));
}
/* ext/quaternion_common.inl:113-122 + type_quat.inl:16-23, 384-388 */
static inline quat quat_inverse(quat q)
{
	float d = (q.w * q.w + q.x * q.x) + (q.y * q.y + q.z * q.z);
	quat r = { q.w / d, -q.x / d, -q.y / d, -q.z / d };
	return r;
}
static inline int quat_is_identity(quat q) { return q.w == 1.0f && q.x == 0.0f && q.y == 0.0f && q.z == 0.0f; }
/* compute_mix_vector: x * (1 - a) + y * a */
static inline v3 mix3(v3 x, v3 y, v3 a)
{
	return V3(x.x * (1.0f - a.x) + y.x * a.x, x.y * (1.0f - a.y) + y.y * a.y, x.z * (1.0f - a.z) + y.z * a.z);
}

/* ------------------------------------------------------------------------------------------- */
/* Opcodes (tangerine/sdf_evaluator.h:83-107) and tree nodes                                    */
/* ------------------------------------------------------------------------------------------- */

enum
{
	OP_STOP = 0,
	OP_SPHERE, OP_ELLIPSOID, OP_BOX, OP_TORUS, OP_CYLINDER, OP_CONE, OP_CONINDER, OP_PLANE,
	OP_UNION, OP_INTER, OP_DIFF, OP_BLEND_UNION, OP_BLEND_INTER, OP_BLEND_DIFF, OP_FLATE,
	OP_OFFSET, OP_MATRIX, OP_SCALE_FIELD,
	KIND_STENCIL_POS = 100, /* StencilMaskNode<false> */
	KIND_STENCIL_NEG = 101  /* StencilMaskNode<true>  */
};
#define NONE 0xFFFFFFFFu

typedef struct
{
	uint32_t kind, a, b, material;
	float params[4];
	float quat[4]; /* w x y z */
	float trans[3];
	float scale;
	float bmin[3], bmax[3];
} TgmRecord; /* on-disk layout, 88 bytes */

typedef struct
{
	TgmRecord r;
	uint32_t stack_size; /* SDFNode::StackSize */
	int32_t leaf_count;  /* LeafCount()        */
	uint8_t has_paint;   /* HasPaint()         */
	uint8_t finite;      /* HasFiniteBounds()  */
} Node;

typedef struct
{
	Node* nodes;
	size_t count, capacity;
} Arena;

struct TgoModel
{
	Arena arena;
	uint32_t root;
	float (*materials)[3];
	uint32_t material_count;
};

static int is_brush(uint32_t k) { return k >= OP_SPHERE && k <= OP_PLANE; }
static int is_set(uint32_t k) { return k >= OP_UNION && k <= OP_BLEND_DIFF; }
static int is_blend(uint32_t k) { return k >= OP_BLEND_UNION && k <= OP_BLEND_DIFF; }
static int is_stencil(uint32_t k) { return k == KIND_STENCIL_POS || k == KIND_STENCIL_NEG; }
enum { FAM_UNION, FAM_INTER, FAM_DIFF };
static int set_family(uint32_t k)
{
	if (k == OP_UNION || k == OP_BLEND_UNION) return FAM_UNION;
	if (k == OP_INTER || k == OP_BLEND_INTER) return FAM_INTER;
	return FAM_DIFF;
}

static uint32_t arena_push(Arena* a, const Node* n)
{
	if (a->count == a->capacity)
	{
		a->capacity = a->capacity ? a->capacity * 2 : 1024;
		a->nodes = (Node*)realloc(a->nodes, a->capacity * sizeof(Node));
	}
	a->nodes[a->count] = *n;
	return (uint32_t)a->count++;
}

/* Derived per-node facts: StackSize (sdf_evaluator.cpp:448,595,771,1056), LeafCount, HasPaint,
 * HasFiniteBounds (:549-562, 681-694, 1014-1027, 1135-1148). */
static void node_derive(Arena* a, uint32_t index)
{
	Node* n = &a->nodes[index];
	uint32_t k = n->r.kind;
	if (is_brush(k))
	{
		n->stack_size = 1;
		n->leaf_count = 1;
		n->has_paint = n->r.material != NONE;
		n->finite = 1;
		for (int i = 0; i < 3; ++i)
		{
			if (isinf(n->r.bmin[i]) || isinf(n->r.bmax[i])) n->finite = 0;
		}
	}
	else if (is_set(k))
	{
		const Node* l = &a->nodes[n->r.a];
		const Node* r = &a->nodes[n->r.b];
		uint32_t rs = r->stack_size + 1;
		n->stack_size = l->stack_size > rs ? l->stack_size : rs;
		n->leaf_count = l->leaf_count + r->leaf_count;
		n->has_paint = l->has_paint || r->has_paint;
		n->finite = l->finite || r->finite;
	}
	else if (k == OP_FLATE)
	{
		const Node* c = &a->nodes[n->r.a];
		n->stack_size = c->stack_size;
		n->leaf_count = c->leaf_count;
		n->has_paint = c->has_paint;
		n->finite = c->finite;
	}
	else /* stencil */
	{
		const Node* c = &a->nodes[n->r.a];
		n->stack_size = c->stack_size;
		n->leaf_count = c->leaf_count;
		n->has_paint = 1;
		n->finite = c->finite;
	}
}

/* SetNode constructor (sdf_evaluator.cpp:737-772), including the left-leaning operand swap. */
static uint32_t make_set(Arena* a, uint32_t kind, uint32_t lhs, uint32_t rhs, float threshold)
{
	Node n;
	memset(&n, 0, sizeof(n));
	if (set_family(kind) != FAM_DIFF && a->nodes[rhs].stack_size > a->nodes[lhs].stack_size)
	{
		uint32_t t = lhs;
		lhs = rhs;
		rhs = t;
	}
	n.r.kind = kind;
	n.r.a = lhs;
	n.r.b = rhs;
	n.r.material = NONE;
	n.r.params[0] = threshold;
	n.r.quat[0] = 1.0f;
	n.r.scale = 1.0f;
	uint32_t index = arena_push(a, &n);
	node_derive(a, index);
	return index;
}

static uint32_t make_unary(Arena* a, uint32_t kind, uint32_t child, uint32_t other, uint32_t material, float param)
{
	Node n;
	memset(&n, 0, sizeof(n));
	n.r.kind = kind;
	n.r.a = child;
	n.r.b = other;
	n.r.material = material;
	n.r.params[0] = param;
	n.r.quat[0] = 1.0f;
	n.r.scale = 1.0f;
	uint32_t index = arena_push(a, &n);
	node_derive(a, index);
	return index;
}

/* ------------------------------------------------------------------------------------------- */
/* SDFMath (tangerine/sdf_evaluator.cpp:165-295)                                                */
/* ------------------------------------------------------------------------------------------- */

static float sdf_sphere(v3 p, float radius) /* :167-170 */
{
	return len3(p) - radius;
}

static float sdf_ellipsoid(v3 p, v3 r) /* :173-178; K0 - 1.0 promotes the product and quotient to double */
{
	float k0 = len3(V3(p.x / r.x, p.y / r.y, p.z / r.z));
	float k1 = len3(V3(p.x / (r.x * r.x), p.y / (r.y * r.y), p.z / (r.z * r.z)));
	return (float)(k0 * (k0 - 1.0) / k1);
}

static float sdf_box(v3 p, v3 e) /* :188-192 */
{
	v3 a = V3(fabsf(p.x) - e.x, fabsf(p.y) - e.y, fabsf(p.z) - e.z);
	v3 m = gmax3(a, V3(0.0f, 0.0f, 0.0f));
	return len3(m) + fminf(fmaxf(fmaxf(a.x, a.y), a.z), 0.0f);
}

static float sdf_torus(v3 p, float major, float minor) /* :202-205 */
{
	v2 xy = { p.x, p.y };
	v2 q = { len2(xy) - major, p.z };
	return len2(q) - minor;
}

static float sdf_cylinder(v3 p, float radius, float extent) /* :208-212 */
{
	v2 xy = { p.x, p.y };
	v2 d = { fabsf(len2(xy)) - radius, fabsf(p.z) - extent };
	v2 m = { gmax(d.x, 0.0f), gmax(d.y, 0.0f) };
	return fminf(fmaxf(d.x, d.y), 0.0f) + len2(m);
}

static float sdf_plane(v3 p, v3 n) /* :215-218 */
{
	return dot3(p, n);
}

static float sdf_cone(v3 p, float tangent, float height) /* :227-237 */
{
	v2 q = { height * tangent, height * -1.0f };
	v2 xy = { p.x, p.y };
	/* Height * -.5 + Point.z is evaluated in double, then narrowed by the vec2 constructor. */
	v2 w = { len2(xy), (float)(height * -.5 + p.z) };
	float ta = gclamp(dot2(w, q) / dot2(q, q), 0.0f, 1.0f);
	v2 a = { w.x - q.x * ta, w.y - q.y * ta };
	float tb = gclamp(w.x / q.x, 0.0f, 1.0f);
	v2 b = { w.x - q.x * tb, w.y - q.y * 1.0f };
	float k = gsign(q.y);
	float d = fminf(dot2(a, a), dot2(b, b));
	float s = fmaxf(k * (w.x * q.y - w.y * q.x), k * (w.y - q.y));
	return sqrtf(d) * gsign(s);
}

static float sdf_coninder(v3 p, float radius_l, float radius_h, float height) /* :240-249 */
{
	v2 xy = { p.x, p.y };
	v2 q = { len2(xy), p.z };
	v2 k1 = { radius_h, height };
	v2 k2 = { radius_h - radius_l, (float)(2.0 * height) };
	v2 ca = { q.x - fminf(q.x, (q.y < 0.0) ? radius_l : radius_h), fabsf(q.y) - height };
	v2 k1q = { k1.x - q.x, k1.y - q.y };
	float t = gclamp(dot2(k1q, k2) / dot2(k2, k2), 0.0f, 1.0f);
	v2 cb = { q.x - k1.x + k2.x * t, q.y - k1.y + k2.y * t };
	float s = (cb.x < 0.0 && ca.y < 0.0) ? -1.0f : 1.0f;
	return s * sqrtf(fminf(dot2(ca, ca), dot2(cb, cb)));
}

/* :252-288.  H * H * 0.25 / Threshold is a double expression; so is the subtraction/addition. */
static float sdf_union(float l, float r) { return fminf(l, r); }
static float sdf_inter(float l, float r) { return fmaxf(l, r); }
static float sdf_diff(float l, float r) { return fmaxf(l, -r); }
static float sdf_blend_union(float l, float r, float t)
{
	float h = fmaxf(t - fabsf(l - r), 0.0f);
	return (float)(fminf(l, r) - h * h * 0.25 / t);
}
static float sdf_blend_inter(float l, float r, float t)
{
	float h = fmaxf(t - fabsf(l - r), 0.0f);
	return (float)(fmaxf(l, r) + h * h * 0.25 / t);
}
static float sdf_blend_diff(float l, float r, float t)
{
	float h = fmaxf(t - fabsf(l + r), 0.0f);
	return (float)(fmaxf(l, -r) + h * h * 0.25 / t);
}

static float set_fn(uint32_t kind, float l, float r, float t)
{
	switch (kind)
	{
	case OP_UNION: return sdf_union(l, r);
	case OP_INTER: return sdf_inter(l, r);
	case OP_DIFF: return sdf_diff(l, r);
	case OP_BLEND_UNION: return sdf_blend_union(l, r, t);
	case OP_BLEND_INTER: return sdf_blend_inter(l, r, t);
	default: return sdf_blend_diff(l, r, t);
	}
}

static float brush_fn(uint32_t kind, const float* p, v3 point)
{
	switch (kind)
	{
	case OP_SPHERE: return sdf_sphere(point, p[0]);
	case OP_ELLIPSOID: return sdf_ellipsoid(point, V3(p[0], p[1], p[2]));
	case OP_BOX: return sdf_box(point, V3(p[0], p[1], p[2]));
	case OP_TORUS: return sdf_torus(point, p[0], p[1]);
	case OP_CYLINDER: return sdf_cylinder(point, p[0], p[1]);
	case OP_CONE: return sdf_cone(point, p[0], p[1]);
	case OP_CONINDER: return sdf_coninder(point, p[0], p[1], p[2]);
	default: return sdf_plane(point, V3(p[0], p[1], p[2]));
	}
}

static int brush_param_count(uint32_t kind)
{
	switch (kind)
	{
	case OP_SPHERE: return 1;
	case OP_TORUS: case OP_CYLINDER: case OP_CONE: return 2;
	default: return 3;
	}
}

/* ------------------------------------------------------------------------------------------- */
/* Tree evaluation (virtual Eval of the node classes)                                           */
/* ------------------------------------------------------------------------------------------- */

/* Transform::ApplyInv (tangerine/transform.cpp:64-67) */
static v3 apply_inv(const TgmRecord* r, v3 point)
{
	quat q = { r->quat[0], r->quat[1], r->quat[2], r->quat[3] };
	v3 t = V3(r->trans[0], r->trans[1], r->trans[2]);
	return divs3(quat_rotate(quat_inverse(q), sub3(point, t)), r->scale);
}

/* Transform::Apply (transform.cpp:58-61) */
static v3 apply_fwd(const TgmRecord* r, v3 point)
{
	quat q = { r->quat[0], r->quat[1], r->quat[2], r->quat[3] };
	v3 t = V3(r->trans[0], r->trans[1], r->trans[2]);
	return add3(quat_rotate(q, muls3(point, r->scale)), t);
}

/* BrushNode::Eval :463-466, SetNode::Eval :774-780, FlateNode::Eval :1059-1062, Stencil :598-601 */
static float tree_eval(const Arena* a, uint32_t index, v3 point)
{
	const Node* n = &a->nodes[index];
	uint32_t k = n->r.kind;
	if (is_brush(k))
	{
		return brush_fn(k, n->r.params, apply_inv(&n->r, point)) * n->r.scale;
	}
	if (is_set(k))
	{
		float l = tree_eval(a, n->r.a, point);
		float r = tree_eval(a, n->r.b, point);
		return set_fn(k, l, r, n->r.params[0]);
	}
	if (k == OP_FLATE)
	{
		return tree_eval(a, n->r.a, point) - n->r.params[0];
	}
	return tree_eval(a, n->r.a, point);
}

/* operator== of the node classes (:564-579, 696-709, 1029-1037, 1150-1154) */
static int tree_equal(const Arena* a, uint32_t x, uint32_t y)
{
	if (x == y) return 1;
	const Node* p = &a->nodes[x];
	const Node* q = &a->nodes[y];
	if (p->r.kind != q->r.kind) return 0;
	uint32_t k = p->r.kind;
	if (is_brush(k))
	{
		if (p->r.material != q->r.material) return 0;
		if (memcmp(p->r.quat, q->r.quat, 16) != 0 && !(p->r.quat[0] == q->r.quat[0] && p->r.quat[1] == q->r.quat[1] && p->r.quat[2] == q->r.quat[2] && p->r.quat[3] == q->r.quat[3])) return 0;
		for (int i = 0; i < 3; ++i) if (p->r.trans[i] != q->r.trans[i]) return 0;
		if (p->r.scale != q->r.scale) return 0;
		for (int i = 0; i < brush_param_count(k); ++i) if (p->r.params[i] != q->r.params[i]) return 0;
		return 1;
	}
	if (is_set(k))
	{
		return p->r.params[0] == q->r.params[0] && tree_equal(a, p->r.a, q->r.a) && tree_equal(a, p->r.b, q->r.b);
	}
	if (k == OP_FLATE)
	{
		return p->r.params[0] == q->r.params[0] && tree_equal(a, p->r.a, q->r.a);
	}
	return tree_equal(a, p->r.a, q->r.a) && tree_equal(a, p->r.b, q->r.b) && p->r.material == q->r.material;
}

/* Clip (:468-478 brush, :782-850 set, :1064-1075 flate, :603-615 stencil).  Brushes are immutable
 * here, so "Copy()" of a brush is the brush itself; new operator nodes are appended to the arena. */
static uint32_t tree_clip(Arena* a, uint32_t index, v3 point, float radius)
{
	uint32_t k = a->nodes[index].r.kind;
	if (is_brush(k))
	{
		return (tree_eval(a, index, point) <= radius) ? index : NONE;
	}
	if (is_set(k))
	{
		if (!(tree_eval(a, index, point) <= radius)) return NONE;
		uint32_t lhs = a->nodes[index].r.a;
		uint32_t rhs = a->nodes[index].r.b;
		float threshold = a->nodes[index].r.params[0];
		int family = set_family(k);
		if (is_blend(k))
		{
			uint32_t nl = tree_clip(a, lhs, point, radius + threshold);
			uint32_t nr = tree_clip(a, rhs, point, radius + threshold);
			if (nl != NONE && nr != NONE)
			{
				return make_set(a, k, nl, nr, threshold);
			}
			if (family == FAM_INTER)
			{
				return NONE;
			}
		}
		uint32_t nl = tree_clip(a, lhs, point, radius);
		uint32_t nr = tree_clip(a, rhs, point, radius);
		if (nl != NONE && nr != NONE)
		{
			return make_set(a, k, nl, nr, threshold);
		}
		if (family == FAM_UNION)
		{
			return nl != NONE ? nl : nr;
		}
		if (family == FAM_DIFF)
		{
			return nl;
		}
		return NONE;
	}
	if (k == OP_FLATE)
	{
		if (!(tree_eval(a, index, point) <= radius)) return NONE;
		float flate = a->nodes[index].r.params[0];
		uint32_t child = tree_clip(a, a->nodes[index].r.a, point, radius + flate);
		if (child == NONE) return NONE; /* the reference would dereference null here */
		return make_unary(a, OP_FLATE, child, NONE, NONE, flate);
	}
	/* stencil: the mask is kept whole */
	{
		uint32_t child = tree_clip(a, a->nodes[index].r.a, point, radius);
		if (child == NONE) return NONE;
		return make_unary(a, k, child, a->nodes[index].r.b, a->nodes[index].r.material, 0.0f);
	}
}

typedef struct { v3 min, max; } AABB;

/* EvaluatorTransform::Apply(AABB) :367-406 */
static AABB brush_bounds(const Node* n)
{
	const TgmRecord* r = &n->r;
	quat q = { r->quat[0], r->quat[1], r->quat[2], r->quat[3] };
	v3 a = V3(r->bmin[0], r->bmin[1], r->bmin[2]);
	v3 b = V3(r->bmax[0], r->bmax[1], r->bmax[2]);
	v3 t = V3(r->trans[0], r->trans[1], r->trans[2]);
	AABB out;
	if (quat_is_identity(q))
	{
		out.min = add3(muls3(a, r->scale), t);
		out.max = add3(muls3(b, r->scale), t);
		return out;
	}
	v3 pts[7] = { b, V3(b.x, a.y, a.z), V3(a.x, b.y, a.z), V3(a.x, a.y, b.z), V3(a.x, b.y, b.z), V3(b.x, a.y, b.z), V3(b.x, b.y, a.z) };
	out.min = apply_fwd(r, a);
	out.max = out.min;
	for (int i = 0; i < 7; ++i)
	{
		v3 tmp = apply_fwd(r, pts[i]);
		out.min = gmin3(out.min, tmp);
		out.max = gmax3(out.max, tmp);
	}
	return out;
}

/* Bounds() :485-488, 622-625, 857-889, 1082-1088 */
static AABB tree_bounds(const Arena* a, uint32_t index)
{
	const Node* n = &a->nodes[index];
	uint32_t k = n->r.kind;
	if (is_brush(k)) return brush_bounds(n);
	if (is_set(k))
	{
		AABB l = tree_bounds(a, n->r.a);
		AABB r = tree_bounds(a, n->r.b);
		AABB c;
		int family = set_family(k);
		if (family == FAM_UNION)
		{
			c.min = gmin3(l.min, r.min);
			c.max = gmax3(l.max, r.max);
		}
		else if (family == FAM_DIFF)
		{
			c = l;
		}
		else
		{
			c.min = gmax3(l.min, r.min);
			c.max = gmin3(l.max, r.max);
		}
		if (is_blend(k))
		{
			float t = n->r.params[0];
			AABB lim;
			lim.min = sub3(gmax3(l.min, r.min), V3(t, t, t));
			lim.max = add3(gmin3(l.max, r.max), V3(t, t, t));
			c.min = gmin3(c.min, lim.min);
			c.max = gmax3(c.max, lim.max);
		}
		return c;
	}
	if (k == OP_FLATE)
	{
		AABB c = tree_bounds(a, n->r.a);
		float pad = n->r.params[0] * 2;
		c.max = add3(c.max, V3(pad, pad, pad));
		c.min = sub3(c.min, V3(pad, pad, pad));
		return c;
	}
	return tree_bounds(a, n->r.a);
}

/* ------------------------------------------------------------------------------------------- */
/* Program compilation (Compile :409-429, 495-507, 632-635, 915-924, 1098-1103)                 */
/* ------------------------------------------------------------------------------------------- */

typedef struct
{
	uint32_t* words;
	size_t count, capacity;
} Program;

static void prog_push_u(Program* p, uint32_t w)
{
	if (p->count == p->capacity)
	{
		p->capacity = p->capacity ? p->capacity * 2 : 64;
		p->words = (uint32_t*)realloc(p->words, p->capacity * 4);
	}
	p->words[p->count++] = w;
}
static void prog_push_f(Program* p, float f)
{
	uint32_t w;
	memcpy(&w, &f, 4);
	prog_push_u(p, w);
}

typedef struct { float m[4][4]; } mat4; /* column-major: m[col][row] */

/* glm type_mat4x4.inl:634-653: Result[c] = A0*B[c][0] + A1*B[c][1] + A2*B[c][2] + A3*B[c][3] */
static mat4 mat4_mul(const mat4* a, const mat4* b)
{
	mat4 r;
	for (int c = 0; c < 4; ++c)
	{
		for (int row = 0; row < 4; ++row)
		{
			r.m[c][row] = ((a->m[0][row] * b->m[c][0] + a->m[1][row] * b->m[c][1]) + a->m[2][row] * b->m[c][2]) + a->m[3][row] * b->m[c][3];
		}
	}
	return r;
}

static mat4 mat4_identity(void)
{
	mat4 r;
	memset(&r, 0, sizeof(r));
	r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
	return r;
}

/* Transform::ToMatrix (transform.cpp:49-55): toMat4 (gtc/quaternion.inl:41-66), translate
 * (ext/matrix_transform.inl:10-15), scale_slow (:89-96), then ScalationMatrix * RotationMatrix. */
static mat4 transform_to_matrix(const TgmRecord* r)
{
	float qw = r->quat[0], qx = r->quat[1], qy = r->quat[2], qz = r->quat[3];
	mat4 rot = mat4_identity();
	float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz;
	float qxz = qx * qz, qxy = qx * qy, qyz = qy * qz;
	float qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
	rot.m[0][0] = 1.0f - 2.0f * (qyy + qzz);
	rot.m[0][1] = 2.0f * (qxy + qwz);
	rot.m[0][2] = 2.0f * (qxz - qwy);
	rot.m[1][0] = 2.0f * (qxy - qwz);
	rot.m[1][1] = 1.0f - 2.0f * (qxx + qzz);
	rot.m[1][2] = 2.0f * (qyz + qwx);
	rot.m[2][0] = 2.0f * (qxz + qwy);
	rot.m[2][1] = 2.0f * (qyz - qwx);
	rot.m[2][2] = 1.0f - 2.0f * (qxx + qyy);

	mat4 id = mat4_identity();
	mat4 tr = id;
	for (int row = 0; row < 4; ++row)
	{
		tr.m[3][row] = ((id.m[0][row] * r->trans[0] + id.m[1][row] * r->trans[1]) + id.m[2][row] * r->trans[2]) + id.m[3][row];
	}
	mat4 sc = mat4_identity();
	sc.m[0][0] = sc.m[1][1] = sc.m[2][2] = r->scale;
	mat4 scalation = mat4_mul(&tr, &sc);
	return mat4_mul(&scalation, &rot);
}

/* glm detail/func_matrix.inl:294-352 */
static mat4 mat4_inverse(const mat4* mp)
{
	const float (*m)[4] = mp->m;
	float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	float f0[4] = { c00, c00, c02, c03 };
	float f1[4] = { c04, c04, c06, c07 };
	float f2[4] = { c08, c08, c10, c11 };
	float f3[4] = { c12, c12, c14, c15 };
	float f4[4] = { c16, c16, c18, c19 };
	float f5[4] = { c20, c20, c22, c23 };
	float v0[4] = { m[1][0], m[0][0], m[0][0], m[0][0] };
	float v1[4] = { m[1][1], m[0][1], m[0][1], m[0][1] };
	float v2[4] = { m[1][2], m[0][2], m[0][2], m[0][2] };
	float v3_[4] = { m[1][3], m[0][3], m[0][3], m[0][3] };
	static const float sa[4] = { +1, -1, +1, -1 };
	static const float sb[4] = { -1, +1, -1, +1 };
	mat4 inv;
	for (int i = 0; i < 4; ++i)
	{
		float i0 = v1[i] * f0[i] - v2[i] * f1[i] + v3_[i] * f2[i];
		float i1 = v0[i] * f0[i] - v2[i] * f3[i] + v3_[i] * f4[i];
		float i2 = v0[i] * f1[i] - v1[i] * f3[i] + v3_[i] * f5[i];
		float i3 = v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i];
		inv.m[0][i] = i0 * sa[i];
		inv.m[1][i] = i1 * sb[i];
		inv.m[2][i] = i2 * sa[i];
		inv.m[3][i] = i3 * sb[i];
	}
	float d0 = m[0][0] * inv.m[0][0], d1 = m[0][1] * inv.m[1][0], d2 = m[0][2] * inv.m[2][0], d3 = m[0][3] * inv.m[3][0];
	float det = (d0 + d1) + (d2 + d3);
	float ood = 1.0f / det;
	for (int c = 0; c < 4; ++c) for (int row = 0; row < 4; ++row) inv.m[c][row] = inv.m[c][row] * ood;
	return inv;
}

static void tree_compile(const Arena* a, uint32_t index, Program* p)
{
	const Node* n = &a->nodes[index];
	uint32_t k = n->r.kind;
	if (is_brush(k))
	{
		const TgmRecord* r = &n->r;
		quat q = { r->quat[0], r->quat[1], r->quat[2], r->quat[3] };
		int has_rotation = !quat_is_identity(q);
		int has_scalation = r->scale != 1.0;
		int has_translation = !(r->trans[0] == 0.0f && r->trans[1] == 0.0f && r->trans[2] == 0.0f);
		if (has_rotation || has_scalation)
		{
			mat4 fwd = transform_to_matrix(r);
			mat4 inv = mat4_inverse(&fwd);
			prog_push_u(p, OP_MATRIX);
			for (int c = 0; c < 4; ++c) for (int row = 0; row < 4; ++row) prog_push_f(p, inv.m[c][row]);
		}
		else if (has_translation)
		{
			prog_push_u(p, OP_OFFSET);
			for (int i = 0; i < 3; ++i) prog_push_f(p, -r->trans[i]);
		}
		prog_push_u(p, k);
		for (int i = 0; i < brush_param_count(k); ++i) prog_push_f(p, r->params[i]);
		if (r->scale != 1.0)
		{
			prog_push_u(p, OP_SCALE_FIELD);
			prog_push_f(p, r->scale);
		}
	}
	else if (is_set(k))
	{
		tree_compile(a, n->r.a, p);
		tree_compile(a, n->r.b, p);
		prog_push_u(p, k);
		if (is_blend(k)) prog_push_f(p, n->r.params[0]);
	}
	else if (k == OP_FLATE)
	{
		tree_compile(a, n->r.a, p);
		prog_push_u(p, OP_FLATE);
		prog_push_f(p, n->r.params[0]);
	}
	else
	{
		tree_compile(a, n->r.a, p);
	}
}

/* SDFInterpreter::Eval (:1386-1605) */
static float interp_eval(const uint32_t* words, size_t count, v3 eval_point)
{
	float stack[64];
	int sp = 0;
	size_t pc = 0;
	v3 point = eval_point;
#define RDF(i) (((const float*)words)[(i)])
	while (pc < count)
	{
		uint32_t op = words[pc++];
		switch (op)
		{
		case OP_STOP:
			return stack[sp - 1];
		case OP_SPHERE:
			stack[sp++] = sdf_sphere(point, RDF(pc));
			pc += 1;
			point = eval_point;
			break;
		case OP_ELLIPSOID:
			stack[sp++] = sdf_ellipsoid(point, V3(RDF(pc), RDF(pc + 1), RDF(pc + 2)));
			pc += 3;
			point = eval_point;
			break;
		case OP_BOX:
			stack[sp++] = sdf_box(point, V3(RDF(pc), RDF(pc + 1), RDF(pc + 2)));
			pc += 3;
			point = eval_point;
			break;
		case OP_TORUS:
			stack[sp++] = sdf_torus(point, RDF(pc), RDF(pc + 1));
			pc += 2;
			point = eval_point;
			break;
		case OP_CYLINDER:
			stack[sp++] = sdf_cylinder(point, RDF(pc), RDF(pc + 1));
			pc += 2;
			point = eval_point;
			break;
		case OP_CONE:
			stack[sp++] = sdf_cone(point, RDF(pc), RDF(pc + 1));
			pc += 2;
			point = eval_point;
			break;
		case OP_CONINDER:
			stack[sp++] = sdf_coninder(point, RDF(pc), RDF(pc + 1), RDF(pc + 2));
			pc += 3;
			point = eval_point;
			break;
		case OP_PLANE:
			stack[sp++] = sdf_plane(point, V3(RDF(pc), RDF(pc + 1), RDF(pc + 2)));
			pc += 3;
			point = eval_point;
			break;
		case OP_UNION: case OP_INTER: case OP_DIFF:
		{
			float r = stack[--sp];
			float l = stack[--sp];
			stack[sp++] = set_fn(op, l, r, 0.0f);
			break;
		}
		case OP_BLEND_UNION: case OP_BLEND_INTER: case OP_BLEND_DIFF:
		{
			float r = stack[--sp];
			float l = stack[--sp];
			stack[sp++] = set_fn(op, l, r, RDF(pc));
			pc += 1;
			break;
		}
		case OP_FLATE:
			stack[sp - 1] -= RDF(pc);
			pc += 1;
			break;
		case OP_OFFSET:
			point = add3(eval_point, V3(RDF(pc), RDF(pc + 1), RDF(pc + 2)));
			pc += 3;
			break;
		case OP_MATRIX:
		{
			/* mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*1)  (type_mat4x4.inl:561-572) */
			const float* m = &RDF(pc);
			float x = eval_point.x, y = eval_point.y, z = eval_point.z;
			point.x = (m[0] * x + m[4] * y) + (m[8] * z + m[12] * 1.0f);
			point.y = (m[1] * x + m[5] * y) + (m[9] * z + m[13] * 1.0f);
			point.z = (m[2] * x + m[6] * y) + (m[10] * z + m[14] * 1.0f);
			pc += 16;
			break;
		}
		case OP_SCALE_FIELD:
			stack[sp - 1] *= RDF(pc);
			pc += 1;
			break;
		default:
			return 0.0f;
		}
	}
#undef RDF
	return 0.0f;
}

/* ------------------------------------------------------------------------------------------- */
/* SDFOctree (:1609-1835)                                                                       */
/* ------------------------------------------------------------------------------------------- */

typedef struct
{
	AABB bounds;
	v3 pivot;
	int terminus;
	int incomplete;
	uint32_t evaluator; /* arena index or NONE */
	int evaluator_leaves;
	int32_t children[8];
	size_t prog_offset, prog_count;
	uint32_t stack_size;
} OctNode;

struct TgoOctree
{
	Arena arena; /* private copy of the model tree + pruned operator nodes */
	float (*materials)[3];
	uint32_t material_count;
	OctNode* nodes;
	size_t count, capacity;
	Program programs; /* all node programs, concatenated */
	float target_size;
	int32_t root;
	int coalesce;   /* SDFOctree::Create's Coalesce argument (1 for the export path) */
	int max_depth;  /* ... and MaxDepth (-1 = no limit) */
	int live_field; /* oct_eval is the live mesher's implicit function (sodapop.cpp:583-587) */
};

static int32_t oct_alloc(TgoOctree* o)
{
	if (o->count == o->capacity)
	{
		o->capacity = o->capacity ? o->capacity * 2 : 1024;
		o->nodes = (OctNode*)realloc(o->nodes, o->capacity * sizeof(OctNode));
	}
	memset(&o->nodes[o->count], 0, sizeof(OctNode));
	for (int i = 0; i < 8; ++i) o->nodes[o->count].children[i] = -1;
	return (int32_t)o->count++;
}

static void oct_populate(TgoOctree* o, int32_t self, int depth);

/* SDFOctree::SDFOctree :1641-1700 (the export path asks for Coalesce = true, MaxDepth = -1; the live mesher for
 * Coalesce = false, MaxDepth = 3 and populates the incomplete nodes afterwards, see tgo_octree_create_live) */
static int32_t oct_construct(TgoOctree* o, uint32_t in_evaluator, AABB bounds, int depth)
{
	int32_t self = oct_alloc(o);
	o->nodes[self].bounds = bounds;
	v3 extent = sub3(bounds.max, bounds.min);
	float span = fmaxf(fmaxf(extent.x, extent.y), extent.z);
	float half_span = (float)(span * 0.5);
	v3 pivot = add3(V3(half_span, half_span, half_span), bounds.min);
	o->nodes[self].pivot = pivot;
	float radius = (float)(len3(V3(span, span, span)) * 0.5);
	uint32_t evaluator = tree_clip(&o->arena, in_evaluator, pivot, radius);
	o->nodes[self].evaluator = evaluator;
	o->nodes[self].evaluator_leaves = evaluator != NONE ? o->arena.nodes[evaluator].leaf_count : 0;
	o->nodes[self].terminus = span <= o->target_size || evaluator == NONE;
	if (!o->nodes[self].terminus)
	{
		o->nodes[self].incomplete = 1;
		if (o->coalesce || o->max_depth == -1 || depth < o->max_depth) /* :1672-1682 */
		{
			oct_populate(o, self, depth);
		}
	}
	if (o->nodes[self].evaluator != NONE)
	{
		/* SDFInterpreter ctor :1376-1383 */
		size_t offset = o->programs.count;
		tree_compile(&o->arena, o->nodes[self].evaluator, &o->programs);
		prog_push_u(&o->programs, OP_STOP);
		o->nodes[self].prog_offset = offset;
		o->nodes[self].prog_count = o->programs.count - offset;
		o->nodes[self].stack_size = o->arena.nodes[o->nodes[self].evaluator].stack_size;
	}
	return self;
}

/* SDFOctree::Populate :1703-1783 */
static void oct_populate(TgoOctree* o, int32_t self, int depth)
{
	if (!o->nodes[self].incomplete) return;
	o->nodes[self].incomplete = 0;
	int uniform = 1;
	int penultimate = 1;
	int live = 0;
	AABB bounds = o->nodes[self].bounds;
	v3 pivot = o->nodes[self].pivot;
	uint32_t evaluator = o->nodes[self].evaluator;
	for (int i = 0; i < 8; ++i)
	{
		AABB cb = bounds;
		if (i & 1) cb.min.x = pivot.x; else cb.max.x = pivot.x;
		if (i & 2) cb.min.y = pivot.y; else cb.max.y = pivot.y;
		if (i & 4) cb.min.z = pivot.z; else cb.max.z = pivot.z;
		int32_t child = oct_construct(o, evaluator, cb, depth + 1);
		if (o->nodes[child].evaluator == NONE)
		{
			o->nodes[self].children[i] = -1;
		}
		else
		{
			o->nodes[self].children[i] = child;
			uniform &= tree_equal(&o->arena, evaluator, o->nodes[child].evaluator);
			penultimate &= o->nodes[child].terminus;
			live++;
		}
	}
	if (live == 0)
	{
		o->nodes[self].evaluator = NONE;
		o->nodes[self].terminus = 1;
	}
	else
	{
		/* :1761-1767 Bounds = union of the live children's Bounds (as they are at this moment) */
		int first = 1;
		for (int i = 0; i < 8; ++i)
		{
			int32_t c = o->nodes[self].children[i];
			if (c < 0) continue;
			AABB cb = o->nodes[c].bounds;
			if (first) o->nodes[self].bounds = cb;
			else
			{
				AABB* b = &o->nodes[self].bounds;
				b->min = V3(fminf(b->min.x, cb.min.x), fminf(b->min.y, cb.min.y), fminf(b->min.z, cb.min.z));
				b->max = V3(fmaxf(b->max.x, cb.max.x), fmaxf(b->max.y, cb.max.y), fmaxf(b->max.z, cb.max.z));
			}
			first = 0;
		}
		int limit = depth > 3 ? depth : 3;
		if (o->coalesce && ((penultimate && uniform) || o->nodes[self].evaluator_leaves <= limit))
		{
			for (int i = 0; i < 8; ++i) o->nodes[self].children[i] = -1;
			o->nodes[self].terminus = 1;
		}
	}
}

/* AABB::BoundingCube / Volume / Degenerate (tangerine/aabb.cpp) */
static int aabb_degenerate(AABB b)
{
	const float* lo = &b.min.x;
	const float* hi = &b.max.x;
	for (int i = 0; i < 3; ++i)
	{
		if (isinf(lo[i]) || isinf(hi[i]) || isnan(lo[i]) || isnan(hi[i]) || hi[i] <= lo[i]) return 1;
	}
	return 0;
}

static void arena_copy(Arena* dst, const Arena* src)
{
	dst->capacity = src->count + 1024;
	dst->count = src->count;
	dst->nodes = (Node*)malloc(dst->capacity * sizeof(Node));
	memcpy(dst->nodes, src->nodes, src->count * sizeof(Node));
}

static TgoOctree* octree_create(const TgoModel* model, float target_size, int coalesce, int max_depth);

/* SDFOctree::Create :1609-1638 with the export path's arguments (export.cpp:322) */
TgoOctree* tgo_octree_create(const TgoModel* model, float target_size)
{
	return octree_create(model, target_size, 1, -1);
}

/* SDFOctree::Walk :1950-1966: terminus or incomplete nodes */
static void oct_collect_incomplete(const TgoOctree* o, int32_t index, int32_t* list, size_t* count)
{
	const OctNode* n = &o->nodes[index];
	if (n->terminus || n->incomplete)
	{
		if (n->incomplete) list[(*count)++] = index;
		return;
	}
	for (int i = 0; i < 8; ++i)
	{
		if (n->children[i] >= 0) oct_collect_incomplete(o, n->children[i], list, count);
	}
}

/* The live mesher's octree: MeshingJob::Run (sodapop.cpp:240) creates it with Coalesce = false, MaxDepth = 3, Margin = 0;
 * MeshingScratch's constructor (:108-120) collects the incomplete nodes and MeshingOctreeTask (:568-571) calls
 * Populate(false, 3, -1) on each.  oct_eval on the result is the implicit function of :583-587. */
TgoOctree* tgo_octree_create_live(const TgoModel* model, float target_size)
{
	TgoOctree* o = octree_create(model, target_size, 0, 3);
	if (!o) return NULL;
	int32_t* list = (int32_t*)malloc(sizeof(int32_t) * (o->count + 1));
	size_t count = 0;
	oct_collect_incomplete(o, o->root, list, &count);
	o->max_depth = -1;
	for (size_t i = 0; i < count; ++i) oct_populate(o, list[i], 3);
	free(list);
	o->live_field = 1;
	return o;
}

/* NaiveSurfaceNetsScratch's constructor, sodapop.cpp:153-179 */
void tgo_live_grid(const TgoOctree* o, float meshing_density, TgoGrid* g)
{
	const AABB b = o->nodes[o->root].bounds;
	const float density = floorf(meshing_density);
	const v3 extent = sub3(b.max, b.min);
	const v3 samples = V3(fmaxf(extent.x * density, 8.0f), fmaxf(extent.y * density, 8.0f), fmaxf(extent.z * density, 8.0f));
	g->x = b.min.x;
	g->y = b.min.y;
	g->z = b.min.z;
	g->sx = (uint64_t)ceilf(samples.x);
	g->sy = (uint64_t)ceilf(samples.y);
	g->sz = (uint64_t)ceilf(samples.z);
	g->dx = extent.x / (float)g->sx;
	g->dy = extent.y / (float)g->sy;
	g->dz = extent.z / (float)g->sz;
	g->x -= g->dx * 2;
	g->y -= g->dy * 2;
	g->z -= g->dz * 2;
	g->sx += 3;
	g->sy += 3;
	g->sz += 3;
}

static TgoOctree* octree_create(const TgoModel* model, float target_size, int coalesce, int max_depth)
{
	if (!model->arena.nodes[model->root].finite) return NULL;
	AABB bounds = tree_bounds(&model->arena, model->root);
	if (aabb_degenerate(bounds)) return NULL;
	v3 extent = sub3(bounds.max, bounds.min);
	float longest = fmaxf(fmaxf(extent.x, extent.y), extent.z);
	v3 padding = mul3(sub3(V3(longest, longest, longest), extent), V3(0.5f, 0.5f, 0.5f));
	AABB cube = { sub3(bounds.min, padding), add3(bounds.max, padding) };
	if (aabb_degenerate(cube)) return NULL;
	/* operator+(float Margin = 0.0) */
	cube.min = sub3(cube.min, V3(0.0f, 0.0f, 0.0f));
	cube.max = add3(cube.max, V3(0.0f, 0.0f, 0.0f));
	if (aabb_degenerate(cube)) return NULL;

	TgoOctree* o = (TgoOctree*)calloc(1, sizeof(TgoOctree));
	arena_copy(&o->arena, &model->arena);
	o->material_count = model->material_count;
	o->materials = (float(*)[3])malloc(sizeof(float[3]) * (model->material_count + 1));
	memcpy(o->materials, model->materials, sizeof(float[3]) * model->material_count);
	o->target_size = target_size;
	o->coalesce = coalesce;
	o->max_depth = max_depth;
	o->root = oct_construct(o, model->root, cube, 1);
	if (o->nodes[o->root].evaluator == NONE)
	{
		tgo_octree_free(o);
		return NULL;
	}
	return o;
}

void tgo_octree_free(TgoOctree* o)
{
	if (!o) return;
	free(o->arena.nodes);
	free(o->materials);
	free(o->nodes);
	free(o->programs.words);
	free(o);
}

/* SDFOctree::Descend(Point, Exact) :1801-1835, as written: a child that finds nothing (possible only in a live
 * octree, where a node populated after its parent may have lost its evaluator) hands the search back to its parent
 * when Exact, and ends it when not. */
static const OctNode* oct_descend_from(const TgoOctree* o, const OctNode* node, v3 point, int exact)
{
	if (!node->terminus)
	{
		int i = 0;
		if (point.x > node->pivot.x) i |= 1;
		if (point.y > node->pivot.y) i |= 2;
		if (point.z > node->pivot.z) i |= 4;
		int32_t child = node->children[i];
		if (child >= 0)
		{
			const OctNode* found = oct_descend_from(o, &o->nodes[child], point, exact);
			if (found || !exact) return found;
		}
		else if (!exact)
		{
			return NULL; /* empty octant, and empty regions need no evaluation */
		}
	}
	return node->evaluator != NONE ? node : NULL;
}

static const OctNode* oct_descend(const TgoOctree* o, v3 point)
{
	return oct_descend_from(o, &o->nodes[o->root], point, 1);
}

static uint64_t fnv(uint64_t hash, const void* data, size_t bytes)
{
	const uint8_t* c = (const uint8_t*)data;
	for (size_t i = 0; i < bytes; ++i)
	{
		hash ^= c[i];
		hash *= 0x100000001B3ull;
	}
	return hash;
}

/* Same pre-order walk and hash as oracle/ref_tool.cpp WalkOctree. */
static void oct_walk(const TgoOctree* o, int32_t index, TgoOctreeStats* s)
{
	const OctNode* n = &o->nodes[index];
	uint32_t mask = 0;
	/* (a child without evaluator -- possible only in a live octree -- is as good as absent, as in ref_tool.cpp WalkOctree) */
	for (int i = 0; i < 8; ++i) if (n->children[i] >= 0 && o->nodes[n->children[i]].evaluator != NONE) mask |= 1u << i;
	uint32_t terminus = n->terminus ? 1 : 0;
	uint32_t words = (uint32_t)n->prog_count;
	s->nodes++;
	s->words += words;
	if (terminus)
	{
		s->leaves++;
		s->leaf_words += words;
	}
	if (words > s->max_words) s->max_words = words;
	if (n->stack_size > s->max_stack) s->max_stack = n->stack_size;
	/* per node FNV-1a over (pivot, terminus, child mask, words); the octree hash is FNV-1a over those in pre-order */
	uint64_t node_hash = 0xCBF29CE484222325ull;
	node_hash = fnv(node_hash, &n->pivot, 12);
	node_hash = fnv(node_hash, &terminus, 4);
	node_hash = fnv(node_hash, &mask, 4);
	node_hash = fnv(node_hash, o->programs.words + n->prog_offset, words * 4);
	s->hash = fnv(s->hash, &node_hash, 8);
	for (int i = 0; i < 8; ++i) if (n->children[i] >= 0 && o->nodes[n->children[i]].evaluator != NONE) oct_walk(o, n->children[i], s);
}

/* SDFOctree::Bounds of the root (what the live mesher's grid is made from) */
void tgo_octree_bounds(const TgoOctree* o, float out_min[3], float out_max[3])
{
	const AABB b = o->nodes[o->root].bounds;
	out_min[0] = b.min.x; out_min[1] = b.min.y; out_min[2] = b.min.z;
	out_max[0] = b.max.x; out_max[1] = b.max.y; out_max[2] = b.max.z;
}

void tgo_octree_stats(const TgoOctree* o, TgoOctreeStats* out)
{
	memset(out, 0, sizeof(*out));
	out->hash = 0xCBF29CE484222325ull;
	oct_walk(o, o->root, out);
}

/* ------------------------------------------------------------------------------------------- */
/* Model loading                                                                                */
/* ------------------------------------------------------------------------------------------- */

TgoModel* tgo_model_load(const char* path)
{
	FILE* f = fopen(path, "rb");
	if (!f) return NULL;
	uint32_t header[4];
	if (fread(header, 4, 4, f) != 4 || header[0] != 0x314D4754u)
	{
		fclose(f);
		return NULL;
	}
	TgoModel* m = (TgoModel*)calloc(1, sizeof(TgoModel));
	m->material_count = header[2];
	m->materials = (float(*)[3])malloc(sizeof(float[3]) * (header[2] + 1));
	if (fread(m->materials, 12, header[2], f) != header[2]) goto fail;
	for (uint32_t i = 0; i < header[1]; ++i)
	{
		Node n;
		memset(&n, 0, sizeof(n));
		if (fread(&n.r, sizeof(TgmRecord), 1, f) != 1) goto fail;
		arena_push(&m->arena, &n);
		node_derive(&m->arena, i); /* records are in post-order: children precede parents */
	}
	m->root = header[3];
	fclose(f);
	return m;
fail:
	fclose(f);
	tgo_model_free(m);
	return NULL;
}

void tgo_model_free(TgoModel* m)
{
	if (!m) return;
	free(m->arena.nodes);
	free(m->materials);
	free(m);
}

void tgo_model_bounds(const TgoModel* m, float out_min[3], float out_max[3])
{
	AABB b = tree_bounds(&m->arena, m->root);
	memcpy(out_min, &b.min, 12);
	memcpy(out_max, &b.max, 12);
}

int tgo_model_has_paint(const TgoModel* m) { return m->arena.nodes[m->root].has_paint; }
int tgo_model_leaf_count(const TgoModel* m) { return m->arena.nodes[m->root].leaf_count; }

uint64_t tgo_model_root_program(const TgoModel* m, uint32_t* out_words, uint64_t capacity)
{
	Program p = { 0 };
	tree_compile(&m->arena, m->root, &p);
	prog_push_u(&p, OP_STOP);
	uint64_t count = p.count;
	if (out_words && capacity >= count) memcpy(out_words, p.words, count * 4);
	free(p.words);
	return count;
}

void tgo_free(void* pointer) { free(pointer); }

/* ------------------------------------------------------------------------------------------- */
/* Thread helper: static split of [0, count) over `threads` pthreads                            */
/* ------------------------------------------------------------------------------------------- */

typedef void (*RangeFn)(void* ctx, uint64_t begin, uint64_t end);
typedef struct { RangeFn fn; void* ctx; uint64_t begin, end; } RangeJob;
static void* range_thunk(void* arg)
{
	RangeJob* j = (RangeJob*)arg;
	j->fn(j->ctx, j->begin, j->end);
	return NULL;
}
static void parallel_for(uint64_t count, int threads, RangeFn fn, void* ctx)
{
	if (threads <= 1 || count < 1024)
	{
		fn(ctx, 0, count);
		return;
	}
	if (threads > 256) threads = 256;
	pthread_t tids[256];
	RangeJob jobs[256];
	/* interleave chunks so that uneven cost spreads across threads */
	uint64_t chunk = (count + threads - 1) / threads;
	for (int t = 0; t < threads; ++t)
	{
		jobs[t].fn = fn;
		jobs[t].ctx = ctx;
		jobs[t].begin = chunk * t < count ? chunk * t : count;
		jobs[t].end = chunk * (t + 1) < count ? chunk * (t + 1) : count;
		pthread_create(&tids[t], NULL, range_thunk, &jobs[t]);
	}
	for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
}

/* ------------------------------------------------------------------------------------------- */
/* Point queries                                                                                */
/* ------------------------------------------------------------------------------------------- */

static const OctNode* oct_descend_inexact(const TgoOctree* o, v3 point)
{
	return oct_descend_from(o, &o->nodes[o->root], point, 0);
}

/* glm::clamp = min(max(x, lo), hi) with glm's comparisons (detail/func_common.inl) */
static float glm_clamp(float x, float lo, float hi)
{
	float t = (x < lo) ? lo : x;
	return (hi < t) ? hi : t;
}

/* SDFOctree::Eval(Point, Exact = true) :1969-1989; on a live octree the live mesher's implicit function
 * clamp(Eval(Point, false), -100, 100) (sodapop.cpp:583-587; nothing found = +infinity, :1978-1981) */
static float oct_eval(const TgoOctree* o, v3 p)
{
	if (o->live_field)
	{
		const OctNode* found = oct_descend_inexact(o, p);
		float d = found ? interp_eval(o->programs.words + found->prog_offset, found->prog_count, p) : INFINITY;
		return glm_clamp(d, -100.0f, 100.0f);
	}
	const OctNode* n = oct_descend(o, p);
	return interp_eval(o->programs.words + n->prog_offset, n->prog_count, p);
}

typedef struct { const TgoOctree* o; const TgoModel* m; const float* pts; float* out; } EvalCtx;
static void eval_octree_range(void* c, uint64_t b, uint64_t e)
{
	EvalCtx* x = (EvalCtx*)c;
	for (uint64_t i = b; i < e; ++i) x->out[i] = oct_eval(x->o, V3(x->pts[i * 3], x->pts[i * 3 + 1], x->pts[i * 3 + 2]));
}
static void eval_tree_range(void* c, uint64_t b, uint64_t e)
{
	EvalCtx* x = (EvalCtx*)c;
	for (uint64_t i = b; i < e; ++i) x->out[i] = tree_eval(&x->m->arena, x->m->root, V3(x->pts[i * 3], x->pts[i * 3 + 1], x->pts[i * 3 + 2]));
}

void tgo_eval_octree(const TgoOctree* o, const float* points, uint64_t count, float* out, int threads)
{
	EvalCtx c = { o, NULL, points, out };
	parallel_for(count, threads, eval_octree_range, &c);
}

void tgo_eval_tree(const TgoModel* m, const float* points, uint64_t count, float* out, int threads)
{
	EvalCtx c = { NULL, m, points, out };
	parallel_for(count, threads, eval_tree_range, &c);
}

/* SDFNode::RayMarch (sdf_evaluator.cpp:336-354) on the unpruned tree, as the Lua ray_cast / magnet calls run it
 * (lua_sdf.cpp:410-444).  rays: 6 floats each (origin, direction); out5: hit (1.0 / 0.0), travel, position xyz. */
void tgo_ray_march(const TgoModel* m, const float* rays, uint64_t count, int max_iterations, float epsilon, int magnet, float* out5)
{
	for (uint64_t i = 0; i < count; ++i)
	{
		v3 start = V3(rays[i * 6], rays[i * 6 + 1], rays[i * 6 + 2]);
		v3 dir = V3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
		if (magnet) /* Direction = normalize(Direction - Origin), lua_sdf.cpp:419-422 */
		{
			v3 d = sub3(dir, start);
			dir = muls3(d, 1.0f / sqrtf(dot3(d, d)));
		}
		dir = muls3(dir, 1.0f / sqrtf(dot3(dir, dir))); /* glm::normalize = v * inversesqrt(dot(v, v)) */
		v3 position = start;
		float travel = 0.0f;
		int hit = 0;
		for (int it = 0; it < max_iterations; ++it)
		{
			float dist = tree_eval(&m->arena, m->root, position);
			if (dist <= epsilon)
			{
				hit = 1;
				break;
			}
			travel += dist;
			position = add3(muls3(dir, travel), start);
		}
		if (!hit) travel = INFINITY;
		out5[i * 5 + 0] = hit ? 1.0f : 0.0f;
		out5[i * 5 + 1] = travel;
		out5[i * 5 + 2] = position.x;
		out5[i * 5 + 3] = position.y;
		out5[i * 5 + 4] = position.z;
	}
}

void tgo_eval_interp(const TgoModel* m, const float* points, uint64_t count, float* out)
{
	Program p = { 0 };
	tree_compile(&m->arena, m->root, &p);
	prog_push_u(&p, OP_STOP);
	for (uint64_t i = 0; i < count; ++i) out[i] = interp_eval(p.words, p.count, V3(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]));
	free(p.words);
}

/* SDFNode::Gradient :298-333 on the tree `index` */
static v3 tree_gradient(const Arena* a, uint32_t index, v3 p)
{
	float almost_zero = 0.0001f;
	float ox = 1.0f * almost_zero;  /* Offset.x */
	float oy = -1.0f * almost_zero; /* Offset.y */
	v3 xyy = V3(ox, oy, oy), yyx = V3(oy, oy, ox), yxy = V3(oy, ox, oy), xxx = V3(ox, ox, ox);
	v3 g = add3(add3(add3(
		muls3(xyy, tree_eval(a, index, add3(p, xyy))),
		muls3(yyx, tree_eval(a, index, add3(p, yyx)))),
		muls3(yxy, tree_eval(a, index, add3(p, yxy)))),
		muls3(xxx, tree_eval(a, index, add3(p, xxx))));
	float len_sq = dot3(g, g);
	if (len_sq == 0.0)
	{
		float d = tree_eval(a, index, p);
		v3 f = V3(tree_eval(a, index, add3(p, xyy)) - d, tree_eval(a, index, add3(p, yxy)) - d, tree_eval(a, index, add3(p, yyx)) - d);
		float inv = 1.0f / sqrtf(dot3(f, f)); /* glm::normalize = v * inversesqrt(dot(v, v)) */
		return muls3(f, inv);
	}
	return divs3(g, sqrtf(len_sq));
}

/* SDFOctree::Gradient (tangerine/sdf_evaluator.h:327-331) */
static v3 oct_gradient(const TgoOctree* o, v3 p)
{
	const OctNode* n = oct_descend(o, p);
	return tree_gradient(&o->arena, n->evaluator, p);
}

void tgo_gradient(const TgoOctree* o, const float* points, uint64_t count, float* out3)
{
	for (uint64_t i = 0; i < count; ++i)
	{
		v3 g = oct_gradient(o, V3(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]));
		out3[i * 3] = g.x;
		out3[i * 3 + 1] = g.y;
		out3[i * 3 + 2] = g.z;
	}
}

/* GetMaterial (:537-547 brush, :666-679 stencil, :957-1012 set, :1130-1133 flate).
 * Returns a material id, or NONE for the default (white) material. */
static uint32_t tree_material(const Arena* a, uint32_t index, v3 p)
{
	const Node* n = &a->nodes[index];
	uint32_t k = n->r.kind;
	if (is_brush(k)) return n->r.material;
	if (k == OP_FLATE) return tree_material(a, n->r.a, p);
	if (is_stencil(k))
	{
		int interior = tree_eval(a, n->r.b, p) < 0.0;
		int apply_to_negative = (k == KIND_STENCIL_NEG);
		if (interior == apply_to_negative) return n->r.material;
		return tree_material(a, n->r.a, p);
	}
	int family = set_family(k);
	if (family == FAM_DIFF) return tree_material(a, n->r.a, p);
	float el = tree_eval(a, n->r.a, p);
	float er = tree_eval(a, n->r.b, p);
	float dist = set_fn(k, el, er, n->r.params[0]);
	int take_left;
	if (is_blend(k)) take_left = fabsf(el - dist) <= fabsf(er - dist);
	else take_left = (dist == el);
	if (family == FAM_UNION)
	{
		return take_left ? tree_material(a, n->r.a, p) : tree_material(a, n->r.b, p);
	}
	uint32_t sl = tree_material(a, n->r.a, p);
	uint32_t sr = tree_material(a, n->r.b, p);
	int lv = a->nodes[n->r.a].has_paint;
	int rv = a->nodes[n->r.b].has_paint;
	if (lv && rv) return take_left ? sl : sr;
	if (lv) return sl;
	return sr;
}

/* tangerine/export.cpp:297-312 colour bytes */
void tgo_color(const TgoOctree* o, const float* points, uint64_t count, uint8_t* out3)
{
	int export_color = o->arena.nodes[o->nodes[o->root].evaluator].has_paint;
	for (uint64_t i = 0; i < count; ++i)
	{
		v3 p = V3(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]);
		float c[3] = { 1.0f, 1.0f, 1.0f };
		if (export_color)
		{
			const OctNode* n = oct_descend(o, p);
			uint32_t material = tree_material(&o->arena, n->evaluator, p);
			if (material != NONE) memcpy(c, o->materials[material], 12);
		}
		for (int ch = 0; ch < 3; ++ch) out3[i * 3 + ch] = (uint8_t)(0xFF * c[ch]);
	}
}

/* ------------------------------------------------------------------------------------------- */
/* Surface nets (third_party/naive-surface-nets/src/surface_nets.cpp)                           */
/* ------------------------------------------------------------------------------------------- */

/* MeshExportThread grid set-up (tangerine/export.cpp:324-337) */
void tgo_export_grid(const float model_min[3], const float model_max[3], const float step[3], TgoGrid* g)
{
	float mn[3];
	for (int i = 0; i < 3; ++i) mn[i] = model_min[i] - step[i] * 2.0f;
	g->x = mn[0]; g->y = mn[1]; g->z = mn[2];
	g->dx = step[0]; g->dy = step[1]; g->dz = step[2];
	g->sx = (uint64_t)(int32_t)ceilf((model_max[0] - mn[0]) / step[0]);
	g->sy = (uint64_t)(int32_t)ceilf((model_max[1] - mn[1]) / step[1]);
	g->sz = (uint64_t)(int32_t)ceilf((model_max[2] - mn[2]) / step[2]);
}

typedef struct { const TgoOctree* o; const TgoGrid* g; float* out; } LatticeCtx;
static void lattice_range(void* c, uint64_t b, uint64_t e)
{
	LatticeCtx* x = (LatticeCtx*)c;
	const TgoGrid* g = x->g;
	uint64_t nx = g->sx + 1, ny = g->sy + 1;
	for (uint64_t k = b; k < e; ++k)
	{
		for (uint64_t j = 0; j < ny; ++j)
		{
			for (uint64_t i = 0; i < nx; ++i)
			{
				/* get_voxel_corner_world_positions :648-687: origin + float(index) * step */
				v3 p = V3(g->x + (float)i * g->dx, g->y + (float)j * g->dy, g->z + (float)k * g->dz);
				x->out[(k * ny + j) * nx + i] = oct_eval(x->o, p);
			}
		}
	}
}

void tgo_lattice_samples(const TgoOctree* o, const TgoGrid* g, float* out, int threads)
{
	LatticeCtx c = { o, g, out };
	/* split by z planes; small grids still thread because planes are expensive */
	uint64_t planes = g->sz + 1;
	if (threads > 1 && planes >= 2)
	{
		int t = threads > (int)planes ? (int)planes : threads;
		pthread_t tids[256];
		RangeJob jobs[256];
		if (t > 256) t = 256;
		for (int i = 0; i < t; ++i)
		{
			jobs[i].fn = lattice_range;
			jobs[i].ctx = &c;
			jobs[i].begin = planes * i / t;
			jobs[i].end = planes * (i + 1) / t;
			pthread_create(&tids[i], NULL, range_thunk, &jobs[i]);
		}
		for (int i = 0; i < t; ++i) pthread_join(tids[i], NULL);
	}
	else
	{
		lattice_range(&c, 0, planes);
	}
}

int tgo_surface_nets(const TgoOctree* o, const TgoGrid* g, TgoMesh* mesh, int threads)
{
	memset(mesh, 0, sizeof(*mesh));
	uint64_t sx = g->sx, sy = g->sy, sz = g->sz;
	uint64_t nx = sx + 1, ny = sy + 1, nz = sz + 1;
	float* s = (float*)malloc(nx * ny * nz * sizeof(float));
	int32_t* cell_vertex = (int32_t*)malloc(sx * sy * sz * sizeof(int32_t));
	if (!s || !cell_vertex) return -1;
	tgo_lattice_samples(o, g, s, threads);
#define S(i, j, k) s[((uint64_t)(k) * ny + (j)) * nx + (i)]

	size_t vcap = 1 << 16, vcount = 0;
	float* verts = (float*)malloc(vcap * 12);
	int64_t* cells = (int64_t*)malloc(vcap * 8);

	/* mesh bounding box :838-840 */
	float bbmin[3] = { g->x, g->y, g->z };
	float bbmax[3] = { g->x + sx * g->dx, g->y + sy * g->dy, g->z + sz * g->dz };
	static const uint8_t edges[12][2] = { {0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7} }; /* :889-901 */
	const float iso = 0.0f;

	/* Loop 1 (FirstLoopInnerThunk :864-974), serial k, j, i order as in the PSTL-serial reference */
	for (uint64_t k = 0; k < sz; ++k)
	for (uint64_t j = 0; j < sy; ++j)
	for (uint64_t i = 0; i < sx; ++i)
	{
		float fi = (float)i, fj = (float)j, fk = (float)k;
		float gp[8][3] = { /* get_voxel_corner_grid_positions :632-646 */
			{ fi, fj, fk }, { fi + 1.f, fj, fk }, { fi + 1.f, fj + 1.f, fk }, { fi, fj + 1.f, fk },
			{ fi, fj, fk + 1.f }, { fi + 1.f, fj, fk + 1.f }, { fi + 1.f, fj + 1.f, fk + 1.f }, { fi, fj + 1.f, fk + 1.f } };
		float cv[8] = { S(i, j, k), S(i + 1, j, k), S(i + 1, j + 1, k), S(i, j + 1, k),
			S(i, j, k + 1), S(i + 1, j, k + 1), S(i + 1, j + 1, k + 1), S(i, j + 1, k + 1) };
		int bipolar[12];
		int active = 0;
		for (int e = 0; e < 12; ++e)
		{
			/* is_scalar_positive is `scalar >= isovalue` :733-740 */
			bipolar[e] = (cv[edges[e][0]] >= iso) != (cv[edges[e][1]] >= iso);
			active |= bipolar[e];
		}
		cell_vertex[(k * sy + j) * sx + i] = -1;
		if (!active) continue;
		float sum[3] = { 0.f, 0.f, 0.f };
		int n = 0;
		for (int e = 0; e < 12; ++e)
		{
			if (!bipolar[e]) continue;
			const float* p1 = gp[edges[e][0]];
			const float* p2 = gp[edges[e][1]];
			float s1 = cv[edges[e][0]], s2 = cv[edges[e][1]];
			float t = (iso - s1) / (s2 - s1);
			for (int c = 0; c < 3; ++c) sum[c] = sum[c] + (p1[c] + t * (p2[c] - p1[c])); /* :937-938, accumulate :944-947 */
			n++;
		}
		float count = (float)n;
		float gc[3] = { sum[0] / count, sum[1] / count, sum[2] / count };
		float fs[3] = { (float)sx, (float)sy, (float)sz };
		if (vcount == vcap)
		{
			vcap *= 2;
			verts = (float*)realloc(verts, vcap * 12);
			cells = (int64_t*)realloc(cells, vcap * 8);
		}
		for (int c = 0; c < 3; ++c)
		{
			/* :952-965 */
			verts[vcount * 3 + c] = bbmin[c] + (bbmax[c] - bbmin[c]) * (gc[c] - 0.f) / (fs[c] - 0.f);
		}
		cells[vcount] = (int64_t)((k * sy + j) * sx + i);
		cell_vertex[(k * sy + j) * sx + i] = (int32_t)vcount;
		vcount++;
	}

	/* Loop 2 (SecondLoopThunk :1001-1121), in vertex order */
	size_t tcap = vcount * 2 + 16, tcount = 0;
	uint32_t* tris = (uint32_t*)malloc(tcap * 12);
#define CV(i, j, k) cell_vertex[((uint64_t)(k) * sy + (j)) * sx + (i)]
	for (size_t v = 0; v < vcount; ++v)
	{
		uint64_t idx = (uint64_t)cells[v];
		uint64_t i = idx % sx, j = (idx / sx) % sy, k = idx / (sx * sy);
		if (i == 0 || j == 0 || k == 0) continue; /* :1016-1022 */
		uint64_t nb[6][3] = { { i - 1, j, k }, { i - 1, j - 1, k }, { i, j - 1, k }, { i, j - 1, k - 1 }, { i, j, k - 1 }, { i - 1, j, k - 1 } };
		float c0 = S(i, j, k), c4 = S(i, j, k + 1), c3 = S(i, j + 1, k), c1 = S(i + 1, j, k);
		float esv[3][2] = { { c0, c4 }, { c3, c0 }, { c0, c1 } }; /* :1041-1067 */
		static const int qn[3][3] = { { 0, 1, 2 }, { 0, 5, 4 }, { 2, 3, 4 } }; /* :1069 */
		for (int e = 0; e < 3; ++e)
		{
			int32_t n0 = CV(nb[qn[e][0]][0], nb[qn[e][0]][1], nb[qn[e][0]][2]);
			int32_t n1 = CV(nb[qn[e][1]][0], nb[qn[e][1]][1], nb[qn[e][1]][2]);
			int32_t n2 = CV(nb[qn[e][2]][0], nb[qn[e][2]][1], nb[qn[e][2]][2]);
			if (n0 < 0 || n1 < 0 || n2 < 0) continue; /* :1093-1096: no bipolarity test */
			int forward = esv[e][1] > esv[e][0];      /* :1103-1105 */
			uint32_t v0 = (uint32_t)v;
			uint32_t v1 = (uint32_t)(forward ? n0 : n2);
			uint32_t v2 = (uint32_t)n1;
			uint32_t v3i = (uint32_t)(forward ? n2 : n0);
			if (tcount + 2 > tcap)
			{
				tcap *= 2;
				tris = (uint32_t*)realloc(tris, tcap * 12);
			}
			tris[tcount * 3 + 0] = v0; tris[tcount * 3 + 1] = v1; tris[tcount * 3 + 2] = v2; tcount++;
			tris[tcount * 3 + 0] = v0; tris[tcount * 3 + 1] = v2; tris[tcount * 3 + 2] = v3i; tcount++;
		}
	}
#undef CV
#undef S
	free(s);
	free(cell_vertex);
	mesh->vertices = verts;
	mesh->cells = cells;
	mesh->triangles = tris;
	mesh->vertex_count = vcount;
	mesh->triangle_count = tcount;
	return 0;
}

void tgo_mesh_free(TgoMesh* mesh)
{
	free(mesh->vertices);
	free(mesh->cells);
	free(mesh->triangles);
	memset(mesh, 0, sizeof(*mesh));
}

/* ------------------------------------------------------------------------------------------- */
/* Refinement + point cloud (tangerine/export.cpp:384-469)                                      */
/* ------------------------------------------------------------------------------------------- */

void tgo_refine(const TgoOctree* o, float* points, uint64_t count, const float half_[3], int iterations)
{
	v3 half = V3(half_[0], half_[1], half_[2]);
	float diagonal = len3(half);
	for (uint64_t i = 0; i < count; ++i)
	{
		v3 vertex = V3(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]);
		v3 low = sub3(vertex, half);
		v3 high = add3(vertex, half);
		v3 cursor = vertex;
		for (int r = 0; r < iterations; ++r)
		{
			v3 dir = oct_gradient(o, cursor);
			float dist = (float)(oct_eval(o, cursor) * -1.0);
			cursor = add3(cursor, muls3(dir, dist));
		}
		cursor = gmin3(gmax3(cursor, low), high); /* glm::clamp */
		if (len3(sub3(vertex, cursor)) <= diagonal) /* distance(Cursor, Vertex) = length(Vertex - Cursor) */
		{
			points[i * 3] = cursor.x;
			points[i * 3 + 1] = cursor.y;
			points[i * 3 + 2] = cursor.z;
		}
	}
}

uint64_t tgo_point_cloud(const TgoOctree* o, const float mn[3], const float mx[3], const float step[3], float** out_points)
{
	v3 half = V3(step[0] / 2.0f, step[1] / 2.0f, step[2] / 2.0f);
	float diagonal = len3(half);
	int32_t it[3];
	for (int c = 0; c < 3; ++c) it[c] = (int32_t)ceilf((mx[c] - mn[c]) / step[c]);
	int32_t slice = it[0] * it[1];
	int32_t total = it[0] * it[1] * it[2];
	size_t cap = 1 << 16, count = 0;
	float* pts = (float*)malloc(cap * 12);
	for (int32_t i = 0; i < total; ++i)
	{
		float z = (float)(i / slice) * step[2] + mn[2];
		float y = (float)((i % slice) / it[0]) * step[1] + mn[1];
		float x = (float)(i % it[0]) * step[0] + mn[0];
		v3 cursor = add3(V3(x, y, z), half);
		float dist = oct_eval(o, cursor);
		if (fabsf(dist) < diagonal)
		{
			if (count == cap)
			{
				cap *= 2;
				pts = (float*)realloc(pts, cap * 12);
			}
			pts[count * 3] = cursor.x;
			pts[count * 3 + 1] = cursor.y;
			pts[count * 3 + 2] = cursor.z;
			count++;
		}
	}
	*out_points = pts;
	return count;
}

/* ------------------------------------------------------------------------------------------- */
/* VoxExport occupancy (tangerine/magica.cpp:27-69)                                             */
/* ------------------------------------------------------------------------------------------- */

typedef struct { const TgoModel* m; AABB b; int32_t size[3]; float radius; uint8_t* hit; } VoxCtx;
static void vox_range(void* c, uint64_t begin, uint64_t end)
{
	VoxCtx* x = (VoxCtx*)c;
	int32_t slice = x->size[0] * x->size[1];
	v3 fsize = V3((float)x->size[0], (float)x->size[1], (float)x->size[2]);
	for (uint64_t idx = begin; idx < end; ++idx)
	{
		int32_t i = (int32_t)idx;
		int32_t z = i / slice;
		int32_t y = (i % slice) / x->size[0];
		int32_t xx = i % x->size[0];
		v3 alpha = V3((float)(xx + .5) / fsize.x, (float)(y + .5) / fsize.y, (float)(z + .5) / fsize.z);
		v3 point = mix3(x->b.min, x->b.max, alpha);
		float dist = tree_eval(&x->m->arena, x->m->root, point);
		x->hit[idx] = fabsf(dist) <= x->radius;
	}
}

uint64_t tgo_voxels(const TgoModel* m, float grid_size, int32_t out_size[3], float* out_radius, int32_t** out_xyz, int threads)
{
	VoxCtx c;
	c.m = m;
	c.b = tree_bounds(&m->arena, m->root);
	v3 ext = sub3(c.b.max, c.b.min);
	c.size[0] = (int32_t)(ceilf(ext.x) * grid_size);
	c.size[1] = (int32_t)(ceilf(ext.y) * grid_size);
	c.size[2] = (int32_t)(ceilf(ext.z) * grid_size);
	v3 alpha = V3(.5f / (float)c.size[0], .5f / (float)c.size[1], .5f / (float)c.size[2]);
	c.radius = len3(sub3(mix3(c.b.min, c.b.max, alpha), c.b.min)); /* distance(p0, p1) = length(p1 - p0) */
	uint64_t total = (uint64_t)c.size[0] * c.size[1] * c.size[2];
	c.hit = (uint8_t*)calloc(total, 1);
	parallel_for(total, threads, vox_range, &c);
	uint64_t count = 0;
	for (uint64_t i = 0; i < total; ++i) count += c.hit[i];
	int32_t* xyz = (int32_t*)malloc((count + 1) * 12);
	uint64_t w = 0;
	int32_t slice = c.size[0] * c.size[1];
	for (uint64_t i = 0; i < total; ++i)
	{
		if (!c.hit[i]) continue;
		xyz[w * 3] = (int32_t)(i % c.size[0]);
		xyz[w * 3 + 1] = (int32_t)((i % slice) / c.size[0]);
		xyz[w * 3 + 2] = (int32_t)(i / slice);
		w++;
	}
	free(c.hit);
	memcpy(out_size, c.size, 12);
	*out_radius = c.radius;
	*out_xyz = xyz;
	return count;
}

/* ------------------------------------------------------------------------------------------- */
/* MeshGenerator (tangerine/mesh_generators.cpp)                                                */
/* ------------------------------------------------------------------------------------------- */

/* LessVec3 :20-39: z, then y, then x */
static int weld_less(const float* l, const float* r)
{
	if (l[2] < r[2]) return 1;
	if (l[2] == r[2])
	{
		if (l[1] < r[1]) return 1;
		if (l[1] == r[1]) return l[0] < r[0];
	}
	return 0;
}

typedef struct { const float* v; uint32_t index; } WeldEntry;
static int weld_compare(const void* a, const void* b)
{
	const WeldEntry* x = (const WeldEntry*)a;
	const WeldEntry* y = (const WeldEntry*)b;
	if (weld_less(x->v, y->v)) return -1;
	if (weld_less(y->v, x->v)) return 1;
	return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0); /* equivalent keys: earliest first */
}

/* Accumulate :42-50 for every vertex in order.  std::map::insert keeps the first key of an equivalence class and its
 * value (the index the vertex got then); sorting by (key, position) and walking the classes gives the same answer. */
uint64_t tgo_weld(const float* vertices, uint64_t count, float* out_vertices4, uint32_t* out_indices)
{
	if (count == 0) return 0;
	WeldEntry* entries = (WeldEntry*)malloc(sizeof(WeldEntry) * count);
	uint32_t* first_of = (uint32_t*)malloc(sizeof(uint32_t) * count); /* per vertex: position of its class's first vertex */
	for (uint64_t i = 0; i < count; ++i)
	{
		entries[i].v = vertices + i * 3;
		entries[i].index = (uint32_t)i;
	}
	qsort(entries, count, sizeof(WeldEntry), weld_compare);
	for (uint64_t i = 0; i < count;)
	{
		uint64_t j = i + 1;
		while (j < count && !weld_less(entries[i].v, entries[j].v) && !weld_less(entries[j].v, entries[i].v)) ++j;
		for (uint64_t k = i; k < j; ++k) first_of[entries[k].index] = entries[i].index;
		i = j;
	}
	uint64_t distinct = 0;
	uint32_t* number = (uint32_t*)malloc(sizeof(uint32_t) * count);
	for (uint64_t i = 0; i < count; ++i)
	{
		if (first_of[i] == i)
		{
			number[i] = (uint32_t)distinct;
			out_vertices4[distinct * 4 + 0] = vertices[i * 3 + 0];
			out_vertices4[distinct * 4 + 1] = vertices[i * 3 + 1];
			out_vertices4[distinct * 4 + 2] = vertices[i * 3 + 2];
			out_vertices4[distinct * 4 + 3] = 1.0f;
			distinct++;
		}
		out_indices[i] = number[first_of[i]];
	}
	free(number);
	free(first_of);
	free(entries);
	return distinct;
}
