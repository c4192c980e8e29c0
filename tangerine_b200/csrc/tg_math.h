// Host-side float vector / quaternion / matrix helpers.
//
// The octree that the GPU path consumes has to be *bit-identical* to the one the reference builds
// (SURVEY.md section 7 "hard parts"): every Lipschitz clip decision and every compiled matrix must
// round the same way.  The reference does its arithmetic with glm 0.9.9.8's scalar code paths, so the
// helpers here spell out the same operation order (file:line cited per function, relative to the
// reference's third_party/glm-0.9.9.8/glm).  This translation unit must be built without FMA
// contraction (-ffp-contract=off) and without fast-math.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace tg
{

struct Vec3
{
	float x = 0.0f, y = 0.0f, z = 0.0f;
	Vec3() = default;
	Vec3(float ix, float iy, float iz) : x(ix), y(iy), z(iz) {}
	explicit Vec3(float s) : x(s), y(s), z(s) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};

inline Vec3 operator+(Vec3 a, Vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline Vec3 operator*(Vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline Vec3 operator/(Vec3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline Vec3 operator/(Vec3 a, Vec3 b) { return { a.x / b.x, a.y / b.y, a.z / b.z }; }
inline bool operator==(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct Vec2
{
	float x, y;
};

// detail/func_geometric.inl:38-55 -- dot is "tmp = a * b; tmp.x + tmp.y (+ tmp.z)"
inline float Dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float Dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
inline float Length(Vec3 a) { return std::sqrt(Dot(a, a)); }
inline float Length(Vec2 a) { return std::sqrt(Dot(a, a)); }

// detail/func_common.inl: glm::min(x, y) = (y < x) ? y : x,  glm::max(x, y) = (x < y) ? y : x
inline float GlmMin(float x, float y) { return (y < x) ? y : x; }
inline float GlmMax(float x, float y) { return (x < y) ? y : x; }
inline Vec3 GlmMin(Vec3 a, Vec3 b) { return { GlmMin(a.x, b.x), GlmMin(a.y, b.y), GlmMin(a.z, b.z) }; }
inline Vec3 GlmMax(Vec3 a, Vec3 b) { return { GlmMax(a.x, b.x), GlmMax(a.y, b.y), GlmMax(a.z, b.z) }; }
inline float GlmClamp(float x, float lo, float hi) { return GlmMin(GlmMax(x, lo), hi); }
// detail/func_common.inl:144-150
inline float GlmSign(float x) { return float(0.0f < x) - float(x < 0.0f); }

// detail/func_geometric.inl:68-79
inline Vec3 Cross(Vec3 x, Vec3 y)
{
	return { x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y };
}

// compute_mix_vector (detail/func_common.inl:81-89): x * (1 - a) + y * a
inline Vec3 Mix(Vec3 x, Vec3 y, Vec3 a)
{
	return { x.x * (1.0f - a.x) + y.x * a.x, x.y * (1.0f - a.y) + y.y * a.y, x.z * (1.0f - a.z) + y.z * a.z };
}

struct Quat
{
	float w = 1.0f, x = 0.0f, y = 0.0f, z = 0.0f;
	bool IsIdentity() const { return w == 1.0f && x == 0.0f && y == 0.0f && z == 0.0f; }
};
inline bool operator==(Quat a, Quat b) { return a.w == b.w && a.x == b.x && a.y == b.y && a.z == b.z; }

// detail/type_quat.inl:282-293 (operator*=)
inline Quat operator*(Quat p, Quat q)
{
	Quat r;
	r.w = p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z;
	r.x = p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y;
	r.y = p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z;
	r.z = p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x;
	return r;
}

// detail/type_quat.inl:343-350 (quat * vec3), which is what gtx/quaternion rotate() calls
inline Vec3 Rotate(Quat q, Vec3 v)
{
	Vec3 qv(q.x, q.y, q.z);
	Vec3 uv = Cross(qv, v);
	Vec3 uuv = Cross(qv, uv);
	return v + ((uv * q.w) + uuv) * 2.0f;
}

// ext/quaternion_common.inl:113-122, dot from detail/type_quat.inl:16-23
inline Quat Inverse(Quat q)
{
	float d = (q.w * q.w + q.x * q.x) + (q.y * q.y + q.z * q.z);
	Quat r;
	r.w = q.w / d;
	r.x = -q.x / d;
	r.y = -q.y / d;
	r.z = -q.z / d;
	return r;
}

// Column-major 4x4, m[col][row] like glm::mat4.
struct Mat4
{
	float m[4][4];

	static Mat4 Identity()
	{
		Mat4 r;
		std::memset(&r, 0, sizeof(r));
		r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
		return r;
	}
};

// detail/type_mat4x4.inl:634-653
inline Mat4 operator*(const Mat4& a, const Mat4& b)
{
	Mat4 r;
	for (int c = 0; c < 4; ++c)
	{
		for (int row = 0; row < 4; ++row)
		{
			r.m[c][row] = ((a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1]) + a.m[2][row] * b.m[c][2]) + a.m[3][row] * b.m[c][3];
		}
	}
	return r;
}

// gtc/quaternion.inl:41-72 (mat4_cast)
inline Mat4 ToMat4(Quat q)
{
	Mat4 r = Mat4::Identity();
	float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
	float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
	float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
	r.m[0][0] = 1.0f - 2.0f * (qyy + qzz);
	r.m[0][1] = 2.0f * (qxy + qwz);
	r.m[0][2] = 2.0f * (qxz - qwy);
	r.m[1][0] = 2.0f * (qxy - qwz);
	r.m[1][1] = 1.0f - 2.0f * (qxx + qzz);
	r.m[1][2] = 2.0f * (qyz + qwx);
	r.m[2][0] = 2.0f * (qxz + qwy);
	r.m[2][1] = 2.0f * (qyz - qwx);
	r.m[2][2] = 1.0f - 2.0f * (qxx + qyy);
	return r;
}

// ext/matrix_transform.inl:10-15 on the identity matrix
inline Mat4 Translate(Vec3 v)
{
	Mat4 id = Mat4::Identity();
	Mat4 r = id;
	for (int row = 0; row < 4; ++row)
	{
		r.m[3][row] = ((id.m[0][row] * v.x + id.m[1][row] * v.y) + id.m[2][row] * v.z) + id.m[3][row];
	}
	return r;
}

// ext/matrix_transform.inl:89-96
inline Mat4 ScaleSlow(const Mat4& m, Vec3 v)
{
	Mat4 s = Mat4::Identity();
	s.m[0][0] = v.x;
	s.m[1][1] = v.y;
	s.m[2][2] = v.z;
	return m * s;
}

// detail/func_matrix.inl:294-352 (cofactor expansion, generic path)
inline Mat4 Inverse(const Mat4& in)
{
	const float (*m)[4] = in.m;
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
	const float f0[4] = { c00, c00, c02, c03 };
	const float f1[4] = { c04, c04, c06, c07 };
	const float f2[4] = { c08, c08, c10, c11 };
	const float f3[4] = { c12, c12, c14, c15 };
	const float f4[4] = { c16, c16, c18, c19 };
	const float f5[4] = { c20, c20, c22, c23 };
	const float v0[4] = { m[1][0], m[0][0], m[0][0], m[0][0] };
	const float v1[4] = { m[1][1], m[0][1], m[0][1], m[0][1] };
	const float v2[4] = { m[1][2], m[0][2], m[0][2], m[0][2] };
	const float v3[4] = { m[1][3], m[0][3], m[0][3], m[0][3] };
	const float sign_a[4] = { +1.0f, -1.0f, +1.0f, -1.0f };
	const float sign_b[4] = { -1.0f, +1.0f, -1.0f, +1.0f };
	Mat4 inv;
	for (int i = 0; i < 4; ++i)
	{
		float i0 = v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i];
		float i1 = v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i];
		float i2 = v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i];
		float i3 = v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i];
		inv.m[0][i] = i0 * sign_a[i];
		inv.m[1][i] = i1 * sign_b[i];
		inv.m[2][i] = i2 * sign_a[i];
		inv.m[3][i] = i3 * sign_b[i];
	}
	float d0 = m[0][0] * inv.m[0][0];
	float d1 = m[0][1] * inv.m[1][0];
	float d2 = m[0][2] * inv.m[2][0];
	float d3 = m[0][3] * inv.m[3][0];
	float one_over_det = 1.0f / ((d0 + d1) + (d2 + d3));
	for (int c = 0; c < 4; ++c)
	{
		for (int row = 0; row < 4; ++row)
		{
			inv.m[c][row] = inv.m[c][row] * one_over_det;
		}
	}
	return inv;
}

struct Box3
{
	Vec3 min, max;
};

inline uint32_t FloatBits(float f)
{
	uint32_t u;
	std::memcpy(&u, &f, 4);
	return u;
}

} // namespace tg
