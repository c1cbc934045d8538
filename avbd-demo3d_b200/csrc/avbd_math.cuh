// avbd_math.cuh — FP32 vector / quaternion / 3x3 helpers and solver constants
// shared by every kernel.  Operand order follows the reference's maths.h
// (alxspiker/avbd-demo3d source/maths.h:21-104) expression by expression, and
// every TU is compiled with -fmad=false, so per-pair narrowphase results and
// feature ids are bit-identical to the CPU reference.  min/max/clamp are the
// reference's NaN-propagating ternaries (maths.h:101-103), NOT fminf/fmaxf.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#define AVBD_HD __host__ __device__ __forceinline__

namespace avbd {

// solver.h:25-36, manifold.cpp:17-23, collision.cpp:18-23, solver.cpp:29,87,98,434-435
constexpr float kPenaltyMin = 20000.0f;
constexpr float kPenaltyMax = 1000000000.0f;
constexpr float kCollisionMargin = 0.02f;
constexpr float kStickThresh = 0.02f;
constexpr float kPenetrationSlop = 0.005f;
constexpr float kManifoldPenaltyCap = 2000000.0f;
constexpr float kNormalContactMargin = 0.01f;
constexpr float kStickAnchorMaxDrift = 0.015f;
constexpr float kStickNormalMinDot = 0.995f;
constexpr float kWarmMaxDrift = 0.08f;
constexpr float kWarmNormalMinDot = 0.9f;
constexpr float kNormalForceCap = 5000.0f;
constexpr float kSatEps = 1.0e-6f;
constexpr float kPlaneEps = 1.0e-5f;
constexpr float kMergeDistSq = 1.0e-6f;
constexpr float kVecEps = 1e-6f;
constexpr float kMaxAngularSpeed = 80.0f;
constexpr float kAngularBetaScale = 0.01f;
constexpr float kLinearDamping = 0.995f;
constexpr float kAngularDamping = 0.97f;
constexpr float kKineticFrictionScale = 0.9f;
constexpr float kEdgeRelTol = 0.95f;
constexpr float kEdgeAbsTol = 0.01f;

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
struct M3 { V3 c[3]; };   // columns

AVBD_HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
AVBD_HD V3 zero3() { return mk3(0.0f, 0.0f, 0.0f); }
AVBD_HD V3 xyz(const float4& f) { return mk3(f.x, f.y, f.z); }
AVBD_HD Q4 quat(const float4& f) { Q4 q; q.x = f.x; q.y = f.y; q.z = f.z; q.w = f.w; return q; }
AVBD_HD float4 f4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
AVBD_HD float4 f4(Q4 q) { return make_float4(q.x, q.y, q.z, q.w); }
AVBD_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
AVBD_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
AVBD_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
AVBD_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
AVBD_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
AVBD_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
AVBD_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
AVBD_HD float len2(V3 a) { return dot(a, a); }
AVBD_HD float len(V3 a) { return sqrtf(len2(a)); }
AVBD_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
AVBD_HD V3 vabs(V3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
AVBD_HD float fmin2(float a, float b) { return a < b ? a : b; }
AVBD_HD float fmax2(float a, float b) { return a > b ? a : b; }
AVBD_HD float clampf(float x, float lo, float hi) { return fmax2(lo, fmin2(hi, x)); }
AVBD_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; __builtin_memcpy(&i, &f, 4); return i;
#endif
}
AVBD_HD float i2f(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; __builtin_memcpy(&f, &i, 4); return f;
#endif
}
AVBD_HD bool finite1(float x) { return fabsf(x) <= FLT_MAX; }   // false for NaN and +-inf
AVBD_HD bool finite3(V3 v) { return finite1(v.x) && finite1(v.y) && finite1(v.z); }
AVBD_HD bool finite4(Q4 q) { return finite1(q.x) && finite1(q.y) && finite1(q.z) && finite1(q.w); }
// manifold.cpp:30-37
AVBD_HD V3 unit_or(V3 v, V3 fb) { float l2 = len2(v); if (l2 < kVecEps) return fb; return v / sqrtf(l2); }

AVBD_HD Q4 qid() { Q4 q; q.x = 0.0f; q.y = 0.0f; q.z = 0.0f; q.w = 1.0f; return q; }
AVBD_HD Q4 qmk(float x, float y, float z, float w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
AVBD_HD Q4 qadd(Q4 a, Q4 b) { return qmk(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
AVBD_HD Q4 qsub(Q4 a, Q4 b) { return qmk(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
AVBD_HD Q4 qscl(Q4 q, float s) { return qmk(q.x * s, q.y * s, q.z * s, q.w * s); }
AVBD_HD Q4 qconj(Q4 q) { return qmk(-q.x, -q.y, -q.z, q.w); }
AVBD_HD Q4 qunit(Q4 q) {                                      // maths.h:65
    float m = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (m < kVecEps) return qid();
    return qscl(q, 1.0f / sqrtf(m));
}
AVBD_HD Q4 qmul(Q4 a, Q4 b) {                                 // maths.h:67
    return qmk(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
               a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
               a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
AVBD_HD V3 qrot(Q4 q, V3 v) {                                 // maths.h:68
    V3 u = mk3(q.x, q.y, q.z);
    V3 t = cross(u, v) * 2.0f;
    return (v + t * q.w) + cross(u, t);
}

AVBD_HD M3 m3(V3 a, V3 b, V3 c) { M3 m; m.c[0] = a; m.c[1] = b; m.c[2] = c; return m; }
AVBD_HD V3 mv(const M3& m, V3 v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }   // maths.h:82
AVBD_HD M3 qmat(Q4 q) {                                       // maths.h:88
    float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
    float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    return m3(mk3(1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy)),
              mk3(2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx)),
              mk3(2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy)));
}
// R * diag(d) * R^T evaluated in the reference's association (rigid.cpp:51-59):
// (R*D) has columns R.c[k]*d[k]; element (r,j) = ((M0[r]*R0[j]) + (M1[r]*R1[j])) + (M2[r]*R2[j]).
AVBD_HD M3 rot_diag(const M3& R, V3 d) {
    V3 m0 = R.c[0] * d.x, m1 = R.c[1] * d.y, m2 = R.c[2] * d.z;
    M3 o;
    o.c[0] = (m0 * R.c[0].x + m1 * R.c[1].x) + m2 * R.c[2].x;
    o.c[1] = (m0 * R.c[0].y + m1 * R.c[1].y) + m2 * R.c[2].y;
    o.c[2] = (m0 * R.c[0].z + m1 * R.c[1].z) + m2 * R.c[2].z;
    return o;
}

// Symmetric 3x3 LDL^T reading the lower triangle of a column-major matrix,
// zero solution on a pivot below FLT_EPSILON.  maths.h:104
AVBD_HD V3 ldl3(const M3& A, V3 b) {
    V3 L0 = A.c[0];
    if (fabsf(L0.x) < FLT_EPSILON) return zero3();
    float d0 = L0.x, l10 = L0.y / d0, l20 = L0.z / d0;
    V3 L1 = A.c[1] - L0 * l10;
    if (fabsf(L1.y) < FLT_EPSILON) return zero3();
    float d1 = L1.y, l21 = L1.z / d1;
    V3 L2 = (A.c[2] - L0 * l20) - L1 * l21;
    if (fabsf(L2.z) < FLT_EPSILON) return zero3();
    float d2 = L2.z;
    V3 y; y.x = b.x; y.y = b.y - l10 * y.x; y.z = b.z - l20 * y.x - l21 * y.y;
    V3 z; z.x = y.x / d0; z.y = y.y / d1; z.z = y.z / d2;
    V3 x; x.z = z.z; x.y = z.y - l21 * x.z; x.x = z.x - l10 * x.y - l20 * x.z;
    return x;
}

// The same LDL^T split into factor + substitute: the factor depends on A only, so solving several right-hand
// sides against one factor gives bit-identical results to calling ldl3 once per right-hand side.
struct Ldl3 { float d0, d1, d2, l10, l20, l21; bool ok; };
// Reciprocal-pivot variant for the solver kernels (tolerance path): 3 divisions per factor, none per solve.
struct Ldl3R { float i0, i1, i2, l10, l20, l21; bool ok; };
AVBD_HD Ldl3R ldl3r_factor(const M3& A) {
    Ldl3R f; f.ok = false; f.i0 = f.i1 = f.i2 = 0.0f; f.l10 = f.l20 = f.l21 = 0.0f;
    V3 L0 = A.c[0];
    if (fabsf(L0.x) < FLT_EPSILON) return f;
    f.i0 = 1.0f / L0.x; f.l10 = L0.y * f.i0; f.l20 = L0.z * f.i0;
    float d1 = A.c[1].y - L0.y * f.l10, l1z = A.c[1].z - L0.z * f.l10;
    if (fabsf(d1) < FLT_EPSILON) return f;
    f.i1 = 1.0f / d1; f.l21 = l1z * f.i1;
    float d2 = (A.c[2].z - L0.z * f.l20) - l1z * f.l21;
    if (fabsf(d2) < FLT_EPSILON) return f;
    f.i2 = 1.0f / d2; f.ok = true;
    return f;
}
AVBD_HD V3 ldl3r_solve(const Ldl3R& f, V3 b) {
    if (!f.ok) return zero3();
    float y0 = b.x, y1 = b.y - f.l10 * y0, y2 = b.z - f.l20 * y0 - f.l21 * y1;
    V3 x; x.z = y2 * f.i2; x.y = y1 * f.i1 - f.l21 * x.z; x.x = y0 * f.i0 - f.l10 * x.y - f.l20 * x.z;
    return x;
}
AVBD_HD Ldl3 ldl3_factor(const M3& A) {
    Ldl3 f; f.ok = false; f.d0 = f.d1 = f.d2 = 1.0f; f.l10 = f.l20 = f.l21 = 0.0f;
    V3 L0 = A.c[0];
    if (fabsf(L0.x) < FLT_EPSILON) return f;
    f.d0 = L0.x; f.l10 = L0.y / f.d0; f.l20 = L0.z / f.d0;
    V3 L1 = A.c[1] - L0 * f.l10;
    if (fabsf(L1.y) < FLT_EPSILON) return f;
    f.d1 = L1.y; f.l21 = L1.z / f.d1;
    V3 L2 = (A.c[2] - L0 * f.l20) - L1 * f.l21;
    if (fabsf(L2.z) < FLT_EPSILON) return f;
    f.d2 = L2.z; f.ok = true;
    return f;
}
AVBD_HD V3 ldl3_solve(const Ldl3& f, V3 b) {
    if (!f.ok) return zero3();
    V3 y; y.x = b.x; y.y = b.y - f.l10 * y.x; y.z = b.z - f.l20 * y.x - f.l21 * y.y;
    V3 z; z.x = y.x / f.d0; z.y = y.y / f.d1; z.z = y.z / f.d2;
    V3 x; x.z = z.z; x.y = z.y - f.l21 * x.z; x.x = z.x - f.l10 * x.y - f.l20 * x.z;
    return x;
}

// Contact frame of an ALREADY-UNIT normal (stored normals are normalised once by Manifold::initialize,
// manifold.cpp:157-159).  Same axis choice as manifold.cpp:39-50; the reference's re-normalisation of n
// and t2 is idempotent up to 1 ulp, so it is skipped here (solver-side tolerance, not the bit-exact path).
AVBD_HD void contact_basis_unit(V3 n, V3& t1, V3& t2) {
    float a, b, l2;
    bool useX = fabsf(n.x) >= fabsf(n.z);
    if (useX) { a = -n.y; b = n.x; } else { a = -n.z; b = n.y; }
    l2 = a * a + b * b;
    if (l2 < kVecEps) { t1 = mk3(1.0f, 0.0f, 0.0f); }
    else {
        float inv = 1.0f / sqrtf(l2);
        t1 = useX ? mk3(a * inv, b * inv, 0.0f) : mk3(0.0f, a * inv, b * inv);
    }
    t2 = cross(n, t1);
}

// Contact frame from a stored normal.  manifold.cpp:39-50
AVBD_HD void contact_basis(V3 nin, V3& n, V3& t1, V3& t2) {
    n = unit_or(nin, mk3(0.0f, 1.0f, 0.0f));
    if (fabsf(n.x) >= fabsf(n.z)) t1 = mk3(-n.y, n.x, 0.0f); else t1 = mk3(0.0f, -n.z, n.y);
    t1 = unit_or(t1, mk3(1.0f, 0.0f, 0.0f));
    t2 = unit_or(cross(n, t1), mk3(0.0f, 0.0f, 1.0f));
}

} // namespace avbd
