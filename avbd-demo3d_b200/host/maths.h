// maths.h — value types of the host API (vec3 / quat / mat3 and their free functions).
//
// API-compatible with alxspiker/avbd-demo3d source/maths.h (same type names, members and function names, so
// scene code and user code written against the reference compiles unchanged), implemented as a thin layer over
// the __host__ __device__ primitives the CUDA kernels use (../csrc/avbd_math.cuh) — one definition of every
// operation for host and device.
#pragma once
#include <cfloat>
#include <cmath>
#include "../csrc/avbd_math.cuh"

const float VEC_EPSILON = avbd::kVecEps;

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(const avbd::V3& v) : x(v.x), y(v.y), z(v.z) {}
    operator avbd::V3() const { return avbd::mk3(x, y, z); }
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3 operator-() const { return vec3(-x, -y, -z); }
    vec3 operator+(const vec3& r) const { return vec3(x + r.x, y + r.y, z + r.z); }
    vec3 operator-(const vec3& r) const { return vec3(x - r.x, y - r.y, z - r.z); }
    vec3 operator*(float s) const { return vec3(x * s, y * s, z * s); }
    vec3 operator/(float s) const { return vec3(x / s, y / s, z / s); }
    vec3& operator+=(const vec3& r) { x += r.x; y += r.y; z += r.z; return *this; }
    vec3& operator-=(const vec3& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};
inline vec3 operator*(float s, const vec3& v) { return v * s; }
inline float dot(const vec3& a, const vec3& b) { return avbd::dot(a, b); }
inline float lengthSq(const vec3& v) { return avbd::len2(v); }
inline float length(const vec3& v) { return avbd::len(v); }
inline vec3 normalize(const vec3& v) { float l = length(v); return l < VEC_EPSILON ? vec3() : v / l; }
inline vec3 cross(const vec3& a, const vec3& b) { return avbd::cross(a, b); }
inline vec3 abs(const vec3& v) { return avbd::vabs(v); }

struct quat {
    float x, y, z, w;
    quat() : x(0), y(0), z(0), w(1) {}
    quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    quat(const vec3& axis, float angle) { float h = angle * 0.5f, s = sinf(h); w = cosf(h); x = axis.x * s; y = axis.y * s; z = axis.z * s; }
    quat(const avbd::Q4& q) : x(q.x), y(q.y), z(q.z), w(q.w) {}
    operator avbd::Q4() const { return avbd::qmk(x, y, z, w); }
    quat operator+(const quat& r) const { return quat(x + r.x, y + r.y, z + r.z, w + r.w); }
    quat operator-(const quat& r) const { return quat(x - r.x, y - r.y, z - r.z, w - r.w); }
};
inline quat operator*(const quat& q, float s) { return avbd::qscl(q, s); }
inline quat normalize(const quat& q) { return avbd::qunit(q); }
inline quat conjugate(const quat& q) { return avbd::qconj(q); }
inline quat operator*(const quat& a, const quat& b) { return avbd::qmul(a, b); }
inline vec3 rotate(const quat& q, const vec3& v) { return avbd::qrot(q, v); }

struct mat3 {
    vec3 cols[3];
    mat3() { cols[0] = vec3(1, 0, 0); cols[1] = vec3(0, 1, 0); cols[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& c) { cols[0] = a; cols[1] = b; cols[2] = c; }
    mat3(const avbd::M3& m) { cols[0] = m.c[0]; cols[1] = m.c[1]; cols[2] = m.c[2]; }
    operator avbd::M3() const { return avbd::m3(cols[0], cols[1], cols[2]); }
    static mat3 diagonal(float d) { return mat3(vec3(d, 0, 0), vec3(0, d, 0), vec3(0, 0, d)); }
    static mat3 diagonal(const vec3& v) { return mat3(vec3(v.x, 0, 0), vec3(0, v.y, 0), vec3(0, 0, v.z)); }
};
inline mat3 transpose(const mat3& m) {
    return mat3(vec3(m.cols[0].x, m.cols[1].x, m.cols[2].x), vec3(m.cols[0].y, m.cols[1].y, m.cols[2].y), vec3(m.cols[0].z, m.cols[1].z, m.cols[2].z));
}
inline vec3 operator*(const mat3& m, const vec3& v) { return avbd::mv(m, v); }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.cols[0], a * b.cols[1], a * b.cols[2]); }
inline mat3 operator+(const mat3& a, const mat3& b) { return mat3(a.cols[0] + b.cols[0], a.cols[1] + b.cols[1], a.cols[2] + b.cols[2]); }
inline mat3 operator-(const mat3& a, const mat3& b) { return mat3(a.cols[0] - b.cols[0], a.cols[1] - b.cols[1], a.cols[2] - b.cols[2]); }
inline mat3 operator*(const mat3& m, float s) { return mat3(m.cols[0] * s, m.cols[1] * s, m.cols[2] * s); }
inline mat3 operator/(const mat3& m, float s) { return mat3(m.cols[0] / s, m.cols[1] / s, m.cols[2] / s); }
inline mat3& operator+=(mat3& a, const mat3& b) { a = a + b; return a; }
inline mat3 outer_product(const vec3& a, const vec3& b) { return mat3(b * a.x, b * a.y, b * a.z); }
inline mat3 mat3_from_quat(const quat& q) { return avbd::qmat(q); }

inline float min(float a, float b) { return avbd::fmin2(a, b); }
inline float max(float a, float b) { return avbd::fmax2(a, b); }
inline float clamp(float x, float a, float b) { return avbd::clampf(x, a, b); }
inline vec3 solve(const mat3& A, const vec3& b) { return avbd::ldl3(A, b); }
