// oracle/avbd_oracle.cpp — TEST INFRASTRUCTURE ONLY (see avbd_oracle.h).
//
// CPU restatement of the AVBD step loop of alxspiker/avbd-demo3d, written
// against the reference's behaviour, not its text: bodies and forces live in
// index-addressed arrays (no intrusive lists), one tagged Force record serves
// all four constraint kinds, and the step is split into callable stages.
// Arithmetic follows the reference expression by expression (operand order
// included) so results are BIT-identical when both are built with
// -ffp-contract=off; tests/test_oracle_pin.py pins that.
//
// Citations are reference file:line (relative to /root/reference/source).

#include "avbd_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_set>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------- constants
// solver.h:25-36, manifold.cpp:17-23, collision.cpp:18-23, solver.cpp:29,87,98,434-435
constexpr int   kMaxRows = 12;
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
constexpr int   kMaxContacts = 4;
constexpr int   kMaxPoly = 16;
constexpr float kSatEps = 1.0e-6f;
constexpr float kPlaneEps = 1.0e-5f;
constexpr float kMergeDistSq = 1.0e-6f;
constexpr float kVecEps = 1e-6f;

// ----------------------------------------------------------------------------- math (maths.h)
struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
struct M3 { V3 c[3]; };   // column-major like the reference's mat3

static inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 zero3() { return V3{0.0f, 0.0f, 0.0f}; }
static inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
static inline V3 scl(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline V3 dvd(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
static inline float comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                    // maths.h:42
static inline float len2(V3 a) { return dot(a, a); }
static inline float len(V3 a) { return sqrtf(len2(a)); }
static inline V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline V3 vabs(V3 a) { return mk(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline V3 unit(V3 v) { float l = len(v); if (l < kVecEps) return zero3(); return dvd(v, l); }  // maths.h:45
// NaN-propagating ternaries, maths.h:101-103 (NOT fminf/fmaxf)
static inline float fmin2(float a, float b) { return a < b ? a : b; }
static inline float fmax2(float a, float b) { return a > b ? a : b; }
static inline float clampf(float x, float lo, float hi) { return fmax2(lo, fmin2(hi, x)); }

static inline Q4 qid() { return Q4{0.0f, 0.0f, 0.0f, 1.0f}; }
static inline Q4 qadd(Q4 a, Q4 b) { return Q4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline Q4 qsub(Q4 a, Q4 b) { return Q4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
static inline Q4 qscl(Q4 q, float s) { return Q4{q.x * s, q.y * s, q.z * s, q.w * s}; }
static inline Q4 qconj(Q4 q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
static inline Q4 qunit(Q4 q) {                                                                         // maths.h:65
    float m = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (m < kVecEps) return qid();
    return qscl(q, 1.0f / sqrtf(m));
}
static inline Q4 qmul(Q4 a, Q4 b) {                                                                    // maths.h:67
    return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
static inline V3 qrot(Q4 q, V3 v) {                                                                    // maths.h:68
    V3 u = mk(q.x, q.y, q.z);
    V3 t = scl(cross(u, v), 2.0f);
    return add(add(v, scl(t, q.w)), cross(u, t));
}
static inline Q4 qaxis(V3 axis, float angle) {                                                         // maths.h:60
    float h = angle * 0.5f; float s = sinf(h);
    Q4 q; q.w = cosf(h); q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s; return q;
}

static inline M3 m3(V3 a, V3 b, V3 c) { M3 m; m.c[0] = a; m.c[1] = b; m.c[2] = c; return m; }
static inline M3 m3zero() { return m3(zero3(), zero3(), zero3()); }
static inline M3 m3diag(V3 d) { return m3(mk(d.x, 0, 0), mk(0, d.y, 0), mk(0, 0, d.z)); }
static inline M3 m3T(const M3& m) {
    return m3(mk(m.c[0].x, m.c[1].x, m.c[2].x), mk(m.c[0].y, m.c[1].y, m.c[2].y), mk(m.c[0].z, m.c[1].z, m.c[2].z));
}
static inline V3 mv(const M3& m, V3 v) { return add(add(scl(m.c[0], v.x), scl(m.c[1], v.y)), scl(m.c[2], v.z)); }  // maths.h:82
static inline M3 mm(const M3& a, const M3& b) { return m3(mv(a, b.c[0]), mv(a, b.c[1]), mv(a, b.c[2])); }
static inline M3 madd(const M3& a, const M3& b) { return m3(add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2])); }
static inline M3 msub(const M3& a, const M3& b) { return m3(sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2])); }
static inline M3 mscl(const M3& a, float s) { return m3(scl(a.c[0], s), scl(a.c[1], s), scl(a.c[2], s)); }
static inline M3 qmat(Q4 q) {                                                                          // maths.h:88
    float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
    float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    return m3(mk(1 - 2 * (yy + zz), 2 * (xy + wz), 2 * (xz - wy)),
              mk(2 * (xy - wz), 1 - 2 * (xx + zz), 2 * (yz + wx)),
              mk(2 * (xz + wy), 2 * (yz - wx), 1 - 2 * (xx + yy)));
}
// solver.cpp:36-39 (its own `outer`, element(r,c) = a[r]*b[c])
static inline M3 outerp(V3 a, V3 b) { return m3(scl(a, b.x), scl(a, b.y), scl(a, b.z)); }

// Symmetric 3x3 LDL^T, zero on a tiny pivot.  maths.h:104
static V3 ldl3(const M3& A, V3 b) {
    V3 L0 = A.c[0];
    if (fabsf(L0.x) < FLT_EPSILON) return zero3();
    float d0 = L0.x, l10 = L0.y / d0, l20 = L0.z / d0;
    V3 L1 = sub(A.c[1], scl(L0, l10));
    if (fabsf(L1.y) < FLT_EPSILON) return zero3();
    float d1 = L1.y, l21 = L1.z / d1;
    V3 L2 = sub(sub(A.c[2], scl(L0, l20)), scl(L1, l21));
    if (fabsf(L2.z) < FLT_EPSILON) return zero3();
    float d2 = L2.z;
    V3 y; y.x = b.x; y.y = b.y - l10 * y.x; y.z = b.z - l20 * y.x - l21 * y.y;
    V3 z; z.x = y.x / d0; z.y = y.y / d1; z.z = y.z / d2;
    V3 x; x.z = z.z; x.y = z.y - l21 * x.z; x.x = z.x - l10 * x.y - l20 * x.z;
    return x;
}

struct V6 { V3 l, a; };
struct M66 { M3 ll, la, al, aa; };

// Block elimination on the linear block.  solver.cpp:68-83
static V6 schur6(const M66& A, const V6& b) {
    M3 W = m3(ldl3(A.ll, A.la.c[0]), ldl3(A.ll, A.la.c[1]), ldl3(A.ll, A.la.c[2]));
    V3 x0 = ldl3(A.ll, b.l);
    M3 S = msub(A.aa, mm(A.al, W));
    V3 rs = sub(b.a, mv(A.al, x0));
    V3 y = ldl3(S, rs);
    V3 x = sub(x0, mv(W, y));
    return V6{x, y};
}

// ----------------------------------------------------------------------------- data model
struct Body {                 // solver.h:48-82
    int id;
    V3 pos; Q4 rot;
    V3 lin, ang, prevLin, prevAng;
    V3 pos0; Q4 rot0;         // initialPosition / initialOrientation
    V3 posI; Q4 rotI;         // inertialPosition / inertialOrientation
    V3 size; float mass, invMass;
    M3 inertia, invInertia;
    float friction, radius;
    std::vector<int> forces;  // force slots touching this body, creation order (walk backwards = reference list order)
};

struct Contact {              // solver.h:119-129
    int feature;
    V3 rA, rB, normal;
    float penetration, C0n;
    V3 C0t;
    bool stick;
};

enum Kind { MANIFOLD = 0, JOINT = 1, SPRING = 2, IGNORE = 3 };

struct Force {                // solver.h:85-109 plus the subclass payloads
    Kind kind;
    int a, b;                 // body indices; a == -1 is the world (joint.cpp:41-45)
    bool alive;
    float C[kMaxRows], fmin[kMaxRows], fmax[kMaxRows], lambda[kMaxRows], penalty[kMaxRows], motor[kMaxRows],
        stiffness[kMaxRows];
    // manifold
    Contact ct[4]; int nct; float mu;
    // joint / spring
    V3 rA, rB; Q4 rel0; float rest;
};

static int g_nextId = 1;      // rigid.cpp:10 — process-global, never reset

struct World {
    float dt; V3 g; int iterations; float alpha, beta, gamma; bool postStab;
    bool logOn = false; int logFreq = 60; int stepIndex = 0;
    std::vector<Body> bodies;
    std::vector<Force> forces;          // slots; creation order; dead slots are compacted at step end
    std::unordered_set<unsigned long long> linked;   // pair keys with any force between them (rigid.cpp:61-69)
    float diagF[5] = {0, 0, 0, 0, 0}; int diagI[3] = {0, 0, 0};
    int dynCount = 0;
    float statMaxLin = 0, statMaxAng = 0;
};

static unsigned long long pairKey(int a, int b) {
    unsigned hi = (unsigned)(a > b ? a : b) + 1u, lo = (unsigned)(a > b ? b : a) + 1u;   // +1: world = -1 -> 0
    return ((unsigned long long)hi << 32) | lo;
}

static void defaults(World& w) {        // solver.cpp:240-253
    w.dt = 1.0f / 60.0f; w.g = mk(0.0f, -10.0f, 0.0f); w.iterations = 10;
    w.alpha = 0.95f; w.beta = 100000.0f; w.gamma = 0.99f; w.postStab = false;
    if (w.logFreq <= 0) w.logFreq = 60;
    for (float& f : w.diagF) f = 0; for (int& i : w.diagI) i = 0;
}

static void clearWorld(World& w) {      // solver.cpp:230-238
    w.forces.clear(); w.bodies.clear(); w.linked.clear(); w.stepIndex = 0;
    for (float& f : w.diagF) f = 0; for (int& i : w.diagI) i = 0;
}

static int addBody(World& w, V3 size, float density, float friction, V3 pos, Q4 rot, V3 lin, V3 ang) {  // rigid.cpp:12-41
    Body b{};
    b.id = g_nextId++;
    b.pos = pos; b.rot = rot; b.lin = lin; b.ang = ang; b.prevLin = lin; b.prevAng = ang;
    b.pos0 = zero3(); b.rot0 = qid(); b.posI = zero3(); b.rotI = qid();
    b.size = size; b.friction = friction;
    b.mass = size.x * size.y * size.z * density;
    b.invMass = (b.mass > 0.0f) ? 1.0f / b.mass : 0.0f;
    b.radius = len(size) * 0.5f;
    if (b.invMass > 0.0f) {
        float ixx = (1.0f / 12.0f) * b.mass * (size.y * size.y + size.z * size.z);
        float iyy = (1.0f / 12.0f) * b.mass * (size.x * size.x + size.z * size.z);
        float izz = (1.0f / 12.0f) * b.mass * (size.x * size.x + size.y * size.y);
        b.inertia = m3diag(mk(ixx, iyy, izz));
        b.invInertia = m3diag(mk(1.0f / ixx, 1.0f / iyy, 1.0f / izz));
    } else {
        b.inertia = m3zero(); b.invInertia = m3zero();
    }
    w.bodies.push_back(b);
    return (int)w.bodies.size() - 1;
}

static M3 worldInertia(const Body& b) { M3 R = qmat(b.rot); return mm(mm(R, b.inertia), m3T(R)); }        // rigid.cpp:56-59
static M3 worldInvInertia(const Body& b) { M3 R = qmat(b.rot); return mm(mm(R, b.invInertia), m3T(R)); }  // rigid.cpp:51-54

static int newForce(World& w, Kind k, int a, int b) {   // force.cpp:12-41
    Force f{};
    f.kind = k; f.a = a; f.b = b; f.alive = true; f.nct = 0; f.mu = 0.0f;
    for (int i = 0; i < kMaxRows; ++i) {
        f.stiffness[i] = 0.0f; f.lambda[i] = 0.0f; f.penalty[i] = 0.0f; f.motor[i] = 0.0f;
        f.fmin[i] = -FLT_MAX; f.fmax[i] = FLT_MAX; f.C[i] = 0.0f;
    }
    for (int i = 0; i < 4; ++i) { f.ct[i] = Contact{}; }
    f.rA = f.rB = zero3(); f.rel0 = qid(); f.rest = 0.0f;
    w.forces.push_back(f);
    int slot = (int)w.forces.size() - 1;
    if (a >= 0) w.bodies[a].forces.push_back(slot);
    if (b >= 0) w.bodies[b].forces.push_back(slot);
    w.linked.insert(pairKey(a, b));
    return slot;
}

static void killForce(World& w, int slot) {             // force.cpp:43-68 (slot is reclaimed by compactForces)
    Force& f = w.forces[slot];
    f.alive = false;
    for (int side = 0; side < 2; ++side) {
        int bi = side == 0 ? f.a : f.b;
        if (bi < 0) continue;
        std::vector<int>& v = w.bodies[bi].forces;
        v.erase(std::find(v.begin(), v.end(), slot));
    }
    w.linked.erase(pairKey(f.a, f.b));
}

static void compactForces(World& w) {
    std::vector<int> remap(w.forces.size(), -1);
    size_t n = 0;
    for (size_t i = 0; i < w.forces.size(); ++i)
        if (w.forces[i].alive) { remap[i] = (int)n; if (n != i) w.forces[n] = w.forces[i]; ++n; }
    if (n == w.forces.size()) return;
    w.forces.resize(n);
    for (Body& b : w.bodies) for (int& s : b.forces) s = remap[s];
}

// ----------------------------------------------------------------------------- narrowphase (collision.cpp)
struct Box { V3 c, h, ax[3]; };
struct Sat { int type, ia, ib; float sep; V3 n; bool valid; };   // type: 0 face-of-A, 1 face-of-B, 2 edge (collision.cpp:25-29)

static Box boxOf(V3 pos, Q4 rot, V3 size) {              // collision.cpp:56-66
    Box b; b.c = pos; b.h = scl(size, 0.5f);
    M3 R = qmat(rot); b.ax[0] = R.c[0]; b.ax[1] = R.c[1]; b.ax[2] = R.c[2];
    return b;
}
static float adot(V3 a, V3 b) { return fabsf(dot(a, b)); }

static void faceAxes(const Box& b, int k, V3& u, V3& v, float& eu, float& ev) {   // collision.cpp:73-92
    if (k == 0) { u = b.ax[1]; v = b.ax[2]; eu = b.h.y; ev = b.h.z; }
    else if (k == 1) { u = b.ax[0]; v = b.ax[2]; eu = b.h.x; ev = b.h.z; }
    else { u = b.ax[0]; v = b.ax[1]; eu = b.h.x; ev = b.h.y; }
}

// One SAT axis; false = separated beyond the persistence margin.  collision.cpp:208-247
static bool satAxis(const Box& A, const Box& B, V3 d, V3 axis, int type, int ia, int ib, Sat& best) {
    float l2 = len2(axis);
    if (l2 < kSatEps) return true;
    V3 n = dvd(axis, sqrtf(l2));
    if (dot(n, d) < 0.0f) n = neg(n);
    float dist = fabsf(dot(d, n));
    float ra = A.h.x * adot(n, A.ax[0]) + A.h.y * adot(n, A.ax[1]) + A.h.z * adot(n, A.ax[2]);
    float rb = B.h.x * adot(n, B.ax[0]) + B.h.y * adot(n, B.ax[1]) + B.h.z * adot(n, B.ax[2]);
    float sep = dist - (ra + rb);
    if (sep > kCollisionMargin) return false;
    if (!best.valid || sep > best.sep) { best.valid = true; best.type = type; best.ia = ia; best.ib = ib; best.sep = sep; best.n = n; }
    return true;
}

// Sutherland-Hodgman against dot(n,p) <= off.  collision.cpp:136-174
static int clipPoly(const V3* in, int nin, V3 n, float off, V3* out) {
    if (nin <= 0) return 0;
    int no = 0;
    V3 a = in[nin - 1];
    float da = dot(n, a) - off;
    for (int i = 0; i < nin; ++i) {
        V3 b = in[i];
        float db = dot(n, b) - off;
        bool ain = da <= kPlaneEps, bin = db <= kPlaneEps;
        if (ain != bin) {
            float t = 0.0f, den = da - db;
            if (fabsf(den) > kSatEps) t = clampf(da / den, 0.0f, 1.0f);
            if (no < kMaxPoly) out[no++] = add(a, scl(sub(b, a), t));
        }
        if (bin && no < kMaxPoly) out[no++] = b;
        a = b; da = db;
    }
    return no;
}

struct PoseRef { V3 pos; Q4 rot; };

// collision.cpp:176-206
static bool pushContact(const PoseRef& pa, const PoseRef& pb, Contact* out, int& n, V3* mids, V3 xA, V3 xB, int key, V3 nBA) {
    V3 mid = scl(add(xA, xB), 0.5f);
    for (int i = 0; i < n; ++i) if (len2(sub(mid, mids[i])) < kMergeDistSq) return false;
    if (n >= kMaxContacts) return false;
    Contact& c = out[n];
    c.feature = key;
    c.rA = qrot(qconj(pa.rot), sub(xA, pa.pos));
    c.rB = qrot(qconj(pb.rot), sub(xB, pb.pos));
    c.normal = nBA;
    c.penetration = fmax2(0.0f, -dot(sub(xA, xB), nBA));
    c.C0n = 0.0f; c.C0t = zero3(); c.stick = false;
    mids[n] = mid; ++n;
    return true;
}

// Reference face + clipped incident face.  collision.cpp:313-394 (helpers :94-134)
static int faceManifold(const PoseRef& pa, const PoseRef& pb, const Box& A, const Box& B, bool refIsA, int refAxis, V3 nAB, Contact* out) {
    const Box& R = refIsA ? A : B;
    const Box& I = refIsA ? B : A;
    V3 outward = refIsA ? nAB : neg(nAB);
    V3 nBA = neg(nAB);
    // reference face frame (:94-101)
    float sgn = dot(outward, R.ax[refAxis]) >= 0.0f ? 1.0f : -1.0f;
    V3 fn = scl(R.ax[refAxis], sgn);
    V3 fc = add(R.c, scl(fn, comp(R.h, refAxis)));
    V3 fu, fv; float eu, ev;
    faceAxes(R, refAxis, fu, fv, eu, ev);
    // incident face = most anti-parallel (:103-117)
    int inc = 0; float bestd = -FLT_MAX;
    for (int i = 0; i < 3; ++i) { float d = adot(I.ax[i], fn); if (d > bestd) { bestd = d; inc = i; } }
    // incident quad (:119-134)
    float isg = dot(I.ax[inc], fn) > 0.0f ? -1.0f : 1.0f;
    V3 inrm = scl(I.ax[inc], isg);
    V3 icen = add(I.c, scl(inrm, comp(I.h, inc)));
    V3 iu, iv; float ieu, iev;
    faceAxes(I, inc, iu, iv, ieu, iev);
    V3 p0[kMaxPoly], p1[kMaxPoly];
    p0[0] = add(add(icen, scl(iu, ieu)), scl(iv, iev));
    p0[1] = add(sub(icen, scl(iu, ieu)), scl(iv, iev));
    p0[2] = sub(sub(icen, scl(iu, ieu)), scl(iv, iev));
    p0[3] = sub(add(icen, scl(iu, ieu)), scl(iv, iev));
    int cnt = 4;
    cnt = clipPoly(p0, cnt, fu, dot(fu, fc) + eu, p1); if (!cnt) return 0;
    V3 nu = neg(fu);
    cnt = clipPoly(p1, cnt, nu, dot(nu, fc) + eu, p0); if (!cnt) return 0;
    cnt = clipPoly(p0, cnt, fv, dot(fv, fc) + ev, p1); if (!cnt) return 0;
    V3 nv = neg(fv);
    cnt = clipPoly(p1, cnt, nv, dot(nv, fc) + ev, p0); if (!cnt) return 0;

    int n = 0; V3 mids[kMaxContacts];
    int prefix = ((refIsA ? 0 : 1) << 24) | ((refAxis & 0xFF) << 16) | ((inc & 0xFF) << 8);
    for (int i = 0; i < cnt && n < kMaxContacts; ++i) {
        V3 pi = p0[i];
        float dist = dot(sub(pi, fc), fn);
        if (dist > kCollisionMargin) continue;
        V3 pr = sub(pi, scl(fn, dist));
        V3 xA = refIsA ? pr : pi, xB = refIsA ? pi : pr;
        V3 rel = sub(pr, fc);
        float uc = dot(rel, fu), vc = dot(rel, fv);
        float un = (eu > kSatEps) ? (uc / eu) : 0.0f;
        float vn = (ev > kSatEps) ? (vc / ev) : 0.0f;
        int qu = (int)floorf(clampf((un + 1.0f) * 7.5f, 0.0f, 15.0f));
        int qv = (int)floorf(clampf((vn + 1.0f) * 7.5f, 0.0f, 15.0f));
        int key = prefix | ((qu & 0x0F) << 4) | (qv & 0x0F);
        pushContact(pa, pb, out, n, mids, xA, xB, key, nBA);
    }
    return n;
}

static void supportEdge(const Box& b, int k, V3 dir, V3& e0, V3& e1) {      // collision.cpp:249-263
    int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    float s1 = dot(dir, b.ax[k1]) >= 0.0f ? 1.0f : -1.0f;
    float s2 = dot(dir, b.ax[k2]) >= 0.0f ? 1.0f : -1.0f;
    V3 ec = add(add(b.c, scl(b.ax[k1], comp(b.h, k1) * s1)), scl(b.ax[k2], comp(b.h, k2) * s2));
    e0 = sub(ec, scl(b.ax[k], comp(b.h, k)));
    e1 = add(ec, scl(b.ax[k], comp(b.h, k)));
}

static void segClosest(V3 p0, V3 p1, V3 q0, V3 q1, V3& c0, V3& c1) {        // collision.cpp:265-311
    V3 d1 = sub(p1, p0), d2 = sub(q1, q0), r = sub(p0, q0);
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    float s = 0.0f, t = 0.0f;
    if (a <= kSatEps && e <= kSatEps) { c0 = p0; c1 = q0; return; }
    if (a <= kSatEps) {
        t = clampf(f / e, 0.0f, 1.0f);
    } else {
        float c = dot(d1, r);
        if (e <= kSatEps) {
            s = clampf(-c / a, 0.0f, 1.0f);
        } else {
            float b = dot(d1, d2);
            float den = a * e - b * b;
            if (fabsf(den) > kSatEps) s = clampf((b * f - c * e) / den, 0.0f, 1.0f);
            t = (b * s + f) / e;
            if (t < 0.0f) { t = 0.0f; s = clampf(-c / a, 0.0f, 1.0f); }
            else if (t > 1.0f) { t = 1.0f; s = clampf((b - c) / a, 0.0f, 1.0f); }
        }
    }
    c0 = add(p0, scl(d1, s));
    c1 = add(q0, scl(d2, t));
}

// Manifold::collide with flip=false.  collision.cpp:420-489
static int collide(const PoseRef& pa, V3 sizeA, const PoseRef& pb, V3 sizeB, Contact* out) {
    Box A = boxOf(pa.pos, pa.rot, sizeA), B = boxOf(pb.pos, pb.rot, sizeB);
    V3 d = sub(B.c, A.c);
    Sat face{}; face.sep = -FLT_MAX; face.valid = false;
    Sat edge{}; edge.sep = -FLT_MAX; edge.valid = false;
    for (int i = 0; i < 3; ++i) if (!satAxis(A, B, d, A.ax[i], 0, i, -1, face)) return 0;
    for (int i = 0; i < 3; ++i) if (!satAxis(A, B, d, B.ax[i], 1, -1, i, face)) return 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (!satAxis(A, B, d, cross(A.ax[i], B.ax[j]), 2, i, j, edge)) return 0;
    if (!face.valid) return 0;
    Sat best = face;
    if (edge.valid && 0.95f * edge.sep > face.sep + 0.01f) best = edge;     // :459-468
    if (best.type == 2) {                                                    // buildEdgeContact :396-416
        V3 a0, a1, b0, b1, xA, xB;
        supportEdge(A, best.ia, best.n, a0, a1);
        supportEdge(B, best.ib, neg(best.n), b0, b1);
        segClosest(a0, a1, b0, b1, xA, xB);
        int n = 0; V3 mids[kMaxContacts];
        int key = (2 << 24) | ((best.ia & 0xFF) << 8) | (best.ib & 0xFF);
        pushContact(pa, pb, out, n, mids, xA, xB, key, neg(best.n));
        return n;
    }
    if (best.type == 0) return faceManifold(pa, pb, A, B, true, best.ia, best.n, out);
    return faceManifold(pa, pb, A, B, false, best.ib, best.n, out);
}

// ----------------------------------------------------------------------------- manifold rows (manifold.cpp)
static V3 unitOr(V3 v, V3 fb) { float l2 = len2(v); if (l2 < kVecEps) return fb; return dvd(v, sqrtf(l2)); }   // :30-37
static void basis(V3 nin, V3& n, V3& t1, V3& t2) {                                                            // :39-50
    n = unitOr(nin, mk(0.0f, 1.0f, 0.0f));
    if (fabsf(n.x) >= fabsf(n.z)) t1 = mk(-n.y, n.x, 0.0f); else t1 = mk(0.0f, -n.z, n.y);
    t1 = unitOr(t1, mk(1.0f, 0.0f, 0.0f));
    t2 = unitOr(cross(n, t1), mk(0.0f, 0.0f, 1.0f));
}
static V3 toWorld(const Body& b, V3 local) { return add(b.pos, qrot(b.rot, local)); }                        // :25-28

// Manifold::initialize — re-collide, carry warm-start by feature id.  manifold.cpp:71-175
static bool manifoldInit(World& w, Force& f) {
    const Body& A = w.bodies[f.a]; const Body& B = w.bodies[f.b];
    f.mu = sqrtf(A.friction * B.friction);
    Contact old[4]; float oldLam[12], oldPen[12]; bool used[4] = {false, false, false, false};
    int nOld = f.nct;
    for (int i = 0; i < nOld; ++i) {
        old[i] = f.ct[i];
        for (int k = 0; k < 3; ++k) { oldLam[i * 3 + k] = f.lambda[i * 3 + k]; oldPen[i * 3 + k] = f.penalty[i * 3 + k]; }
    }
    PoseRef pa{A.pos, A.rot}, pb{B.pos, B.rot};
    f.nct = collide(pa, A.size, pb, B.size, f.ct);
    if (f.nct == 0) return false;
    for (int i = 0; i < f.nct; ++i) {
        int base = i * 3;
        for (int k = 0; k < 3; ++k) { f.stiffness[base + k] = FLT_MAX; f.lambda[base + k] = 0.0f; f.penalty[base + k] = kPenaltyMin; f.motor[base + k] = 0.0f; }
        f.ct[i].stick = false;
        int hit = -1;
        for (int j = 0; j < nOld; ++j) { if (used[j]) continue; if (f.ct[i].feature == old[j].feature) { hit = j; break; } }
        if (hit >= 0) {
            used[hit] = true;
            V3 nn = unitOr(f.ct[i].normal, mk(0.0f, 1.0f, 0.0f));
            V3 on = unitOr(old[hit].normal, nn);
            float nd = dot(nn, on);
            V3 oldMid = scl(add(toWorld(A, old[hit].rA), toWorld(B, old[hit].rB)), 0.5f);
            V3 newMid = scl(add(toWorld(A, f.ct[i].rA), toWorld(B, f.ct[i].rB)), 0.5f);
            float drift2 = len2(sub(newMid, oldMid));
            bool warm = (nd >= kWarmNormalMinDot) && (drift2 <= kWarmMaxDrift * kWarmMaxDrift);
            if (warm) for (int k = 0; k < 3; ++k) {
                f.lambda[base + k] = oldLam[hit * 3 + k];
                f.penalty[base + k] = clampf(oldPen[hit * 3 + k], kPenaltyMin, kManifoldPenaltyCap);
            }
            bool reuse = false;
            if (old[hit].stick && warm) reuse = (nd >= kStickNormalMinDot) && (drift2 <= kStickAnchorMaxDrift * kStickAnchorMaxDrift);
            f.ct[i].stick = old[hit].stick && reuse;
            if (reuse) { f.ct[i].rA = old[hit].rA; f.ct[i].rB = old[hit].rB; }
        }
        V3 n, t1, t2;
        basis(f.ct[i].normal, n, t1, t2);
        f.ct[i].normal = n;
        V3 dlt = sub(toWorld(A, f.ct[i].rA), toWorld(B, f.ct[i].rB));
        f.ct[i].C0n = dot(dlt, n) - kNormalContactMargin;
        f.ct[i].C0t.x = dot(dlt, t1); f.ct[i].C0t.y = dot(dlt, t2); f.ct[i].C0t.z = 0.0f;
        f.ct[i].penetration = fmax2(0.0f, -dot(dlt, n));
    }
    return true;
}

// Manifold::computeConstraint — stateful (clamps warm tangential lambda, flips stick).  manifold.cpp:177-245
static void manifoldConstraint(World& w, Force& f, float alpha) {
    const Body& A = w.bodies[f.a]; const Body& B = w.bodies[f.b];
    float bias = clampf(1.0f - alpha, 0.0f, 1.0f);
    for (int i = 0; i < f.nct; ++i) {
        int base = i * 3;
        Contact& c = f.ct[i];
        V3 n, t1, t2;
        basis(c.normal, n, t1, t2);
        c.normal = n;
        V3 dlt = sub(toWorld(A, c.rA), toWorld(B, c.rB));
        float sepn = dot(dlt, n) - kNormalContactMargin;
        float s1 = dot(dlt, t1), s2 = dot(dlt, t2);
        f.C[base] = sepn + bias * c.C0n;
        float ims = A.invMass + B.invMass;
        float mscale = (ims > 1.0e-6f) ? (1.0f / ims) : 1.0f;
        float cap = kNormalForceCap * mscale;
        f.fmin[base] = -cap; f.fmax[base] = 0.0f;
        f.C[base + 1] = s1 + bias * c.C0t.x;
        f.C[base + 2] = s2 + bias * c.C0t.y;
        float warmN = fabsf(fmin2(f.lambda[base], 0.0f));
        float trial = f.penalty[base] * f.C[base] + f.lambda[base];
        float trialN = fabsf(fmin2(trial, 0.0f));
        float nmag = fmax2(warmN, trialN);
        nmag = fmin2(nmag, cap);
        float mu = f.mu;
        if (!c.stick) mu *= 0.9f;
        float lim = mu * nmag;
        float l1 = f.lambda[base + 1], l2 = f.lambda[base + 2];
        float tm = sqrtf(l1 * l1 + l2 * l2);
        if (tm > lim && tm > 1.0e-8f) { float s = lim / tm; f.lambda[base + 1] *= s; f.lambda[base + 2] *= s; }
        f.fmin[base + 1] = -lim; f.fmax[base + 1] = lim; f.fmin[base + 2] = -lim; f.fmax[base + 2] = lim;
        float slip2 = f.C[base + 1] * f.C[base + 1] + f.C[base + 2] * f.C[base + 2];
        float tl2 = f.lambda[base + 1] * f.lambda[base + 1] + f.lambda[base + 2] * f.lambda[base + 2];
        c.stick = (slip2 <= kStickThresh * kStickThresh) && (tl2 <= lim * lim + 1.0e-8f);
        c.penetration = fmax2(0.0f, -dot(dlt, n));
    }
}

// manifold.cpp:247-271
static void manifoldJac(const World& w, const Force& f, int body, int row, V3& Jl, V3& Ja) {
    const Contact& c = f.ct[row / 3];
    int type = row % 3;
    V3 n, t1, t2;
    basis(c.normal, n, t1, t2);
    V3 bs = type == 0 ? n : (type == 1 ? t1 : t2);
    bool isA = body == f.a;
    float sg = isA ? 1.0f : -1.0f;
    V3 r = isA ? qrot(w.bodies[f.a].rot, c.rA) : qrot(w.bodies[f.b].rot, c.rB);
    Jl = scl(bs, sg);
    Ja = scl(cross(r, bs), sg);
}

// ----------------------------------------------------------------------------- joint (joint.cpp) / spring (spring.cpp)
static void jointConstraint(World& w, Force& f) {        // joint.cpp:68-106
    V3 pA, pB; Q4 qA;
    if (f.a >= 0) { qA = w.bodies[f.a].rot; pA = add(w.bodies[f.a].pos, qrot(qA, f.rA)); }
    else { qA = qid(); pA = f.rA; }
    const Body& B = w.bodies[f.b];
    pB = add(B.pos, qrot(B.rot, f.rB));
    V3 lc = sub(pA, pB);
    f.C[0] = lc.x; f.C[1] = lc.y; f.C[2] = lc.z;
    Q4 cur = qmul(qconj(qA), B.rot);
    Q4 dq = qmul(cur, qconj(f.rel0));
    V3 ac = scl(mk(dq.x, dq.y, dq.z), 2.0f);
    f.C[3] = ac.x; f.C[4] = ac.y; f.C[5] = ac.z;
    for (int i = 0; i < 6; ++i) { f.fmin[i] = -FLT_MAX; f.fmax[i] = FLT_MAX; }
}
static void jointJac(const World& w, const Force& f, int body, int row, V3& Jl, V3& Ja) {   // joint.cpp:108-139
    Jl = zero3(); Ja = zero3();
    bool isA = body == f.a;
    float sg = isA ? 1.0f : -1.0f;
    if (isA && f.a < 0) return;
    if (row < 3) {
        V3 ax = zero3(); (row == 0 ? ax.x : (row == 1 ? ax.y : ax.z)) = 1.0f;
        V3 r = isA ? qrot(w.bodies[body].rot, f.rA) : qrot(w.bodies[body].rot, f.rB);
        Jl = scl(ax, sg); Ja = scl(cross(r, ax), sg);
    } else {
        V3 ax = zero3(); (row == 3 ? ax.x : (row == 4 ? ax.y : ax.z)) = 1.0f;
        Jl = zero3(); Ja = scl(ax, sg);
    }
}
static void springEnds(const World& w, const Force& f, V3& pA, V3& pB) {
    Q4 qA = f.a >= 0 ? w.bodies[f.a].rot : qid();
    pA = f.a >= 0 ? add(w.bodies[f.a].pos, qrot(qA, f.rA)) : f.rA;
    pB = add(w.bodies[f.b].pos, qrot(w.bodies[f.b].rot, f.rB));
}
static void springConstraint(World& w, Force& f) {       // spring.cpp:33-56 (H_ll is dead state, omitted)
    V3 pA, pB; springEnds(w, f, pA, pB);
    f.C[0] = len(sub(pA, pB)) - f.rest;
}
static void springJac(const World& w, const Force& f, int body, V3& Jl, V3& Ja) {           // spring.cpp:59-90
    V3 pA, pB; springEnds(w, f, pA, pB);
    V3 d = sub(pA, pB);
    float L = len(d);
    if (L < kVecEps) { Jl = zero3(); Ja = zero3(); return; }
    V3 n = dvd(d, L);
    bool isA = body == f.a;
    float sg = isA ? 1.0f : -1.0f;
    Jl = scl(n, sg);
    if (body >= 0) { V3 r = isA ? qrot(w.bodies[body].rot, f.rA) : qrot(w.bodies[body].rot, f.rB); Ja = scl(cross(r, n), sg); }
    else Ja = zero3();
}

static int rowCount(const Force& f) { return f.kind == MANIFOLD ? f.nct * 3 : (f.kind == JOINT ? 6 : (f.kind == SPRING ? 1 : 0)); }
static void evalConstraint(World& w, Force& f, float alpha) {
    if (f.kind == MANIFOLD) manifoldConstraint(w, f, alpha);
    else if (f.kind == JOINT) jointConstraint(w, f);
    else if (f.kind == SPRING) springConstraint(w, f);
}
static void evalJac(const World& w, const Force& f, int body, int row, V3& Jl, V3& Ja) {
    if (f.kind == MANIFOLD) manifoldJac(w, f, body, row, Jl, Ja);
    else if (f.kind == JOINT) jointJac(w, f, body, row, Jl, Ja);
    else if (f.kind == SPRING) springJac(w, f, body, Jl, Ja);
    else { /* IgnoreCollision leaves its outputs untouched (ignorecollision.h:20); it has no rows */ }
}

// ----------------------------------------------------------------------------- step stages (solver.cpp)
static bool finite3(V3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }
static bool finite4(Q4 q) { return std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z) && std::isfinite(q.w); }
static void scrub3(V3& v, const char* what, int id) {     // solver.cpp:51-58
    if (!finite3(v)) {
        std::printf("[Physics] Warning: body %d produced non-finite %s (%.3f, %.3f, %.3f); resetting to zero.\n", id, what, v.x, v.y, v.z);
        v = zero3();
    }
}
static void scrub4(Q4& q, const char* what, int id) {     // solver.cpp:60-66
    if (!finite4(q)) {
        std::printf("[Physics] Warning: body %d produced non-finite %s; resetting to identity.\n", id, what);
        q = qid();
    }
}

// solver.cpp:262-270.  Reference order: A walks newest->oldest, B is every older body.
static void stageBroadphase(World& w) {
    int n = (int)w.bodies.size();
    for (int a = n - 1; a >= 0; --a) {
        const Body& A = w.bodies[a];
        for (int b = a - 1; b >= 0; --b) {
            const Body& B = w.bodies[b];
            V3 dp = sub(A.pos, B.pos);
            float r = A.radius + B.radius;
            if (dot(dp, dp) <= r * r && !w.linked.count(pairKey(a, b))) newForce(w, MANIFOLD, a, b);
        }
    }
}

// solver.cpp:273-296
static void stageInit(World& w) {
    for (int s = (int)w.forces.size() - 1; s >= 0; --s) {
        Force& f = w.forces[s];
        if (!f.alive) continue;
        bool ok = f.kind == MANIFOLD ? manifoldInit(w, f) : true;
        if (!ok) { killForce(w, s); continue; }
        int rows = rowCount(f);
        for (int i = 0; i < rows; ++i) {
            if (w.postStab) {
                f.penalty[i] = clampf(f.penalty[i] * w.gamma, kPenaltyMin, kPenaltyMax);
            } else {
                f.lambda[i] *= w.alpha * w.gamma;
                f.penalty[i] = clampf(f.penalty[i] * w.gamma, kPenaltyMin, kPenaltyMax);
            }
            if (f.stiffness[i] > 0.0f && f.stiffness[i] < FLT_MAX) f.penalty[i] = fmin2(f.penalty[i], f.stiffness[i]);
        }
    }
    compactForces(w);
}

// solver.cpp:299-337
static void stagePredict(World& w) {
    w.dynCount = 0;
    float dt = w.dt;
    for (int i = (int)w.bodies.size() - 1; i >= 0; --i) {
        Body& b = w.bodies[i];
        { float l = len(b.ang); if (l > 80.0f && l > kVecEps) b.ang = scl(b.ang, 80.0f / l); }   // :85-92
        b.pos0 = b.pos; b.rot0 = b.rot;
        b.posI = b.pos; b.rotI = b.rot;
        if (b.invMass > 0.0f) {
            ++w.dynCount;
            scrub3(b.lin, "linear velocity", b.id);
            scrub3(b.ang, "angular velocity", b.id);
            b.posI = add(add(b.pos, scl(b.lin, dt)), scl(w.g, dt * dt));
            Q4 om{b.ang.x, b.ang.y, b.ang.z, 0.0f};
            b.rotI = qunit(qadd(b.rot, qscl(qmul(om, b.rot), 0.5f * dt)));
            float gl = len(w.g);
            float aw = 0.0f;
            if (gl > 1e-5f) {
                V3 acc = dvd(sub(b.lin, b.prevLin), dt);
                float proj = dot(acc, dvd(w.g, gl));
                aw = clampf(proj / gl, 0.0f, 1.0f);
                if (!std::isfinite(aw)) aw = 0.0f;
            }
            b.pos = add(b.pos, add(scl(b.lin, dt), scl(w.g, aw * dt * dt)));
            b.rot = b.rotI;
            scrub3(b.pos, "predicted position", b.id);
            scrub4(b.rot, "predicted orientation", b.id);
        } else {
            b.pos = b.posI; b.rot = b.rotI;
        }
    }
}

// One body's block solve + pose update.  solver.cpp:345-408
static V6 primalBody(World& w, int bi, float alpha) {
    Body& b = w.bodies[bi];
    float dt = w.dt;
    M66 lhs{m3zero(), m3zero(), m3zero(), m3zero()};
    V6 rhs{zero3(), zero3()};
    M3 massM = m3diag(mk(b.mass, b.mass, b.mass));
    M3 Iw = worldInertia(b);
    float invDt2 = 1.0f / (dt * dt);
    lhs.ll = madd(lhs.ll, mscl(massM, invDt2));
    lhs.aa = madd(lhs.aa, mscl(Iw, invDt2));
    rhs.l = mv(massM, scl(sub(b.pos, b.posI), invDt2));
    Q4 qe = qmul(b.rot, qconj(b.rotI));
    V3 re = scl(mk(qe.x, qe.y, qe.z), 2.0f);
    if (qe.w < 0.0f) re = neg(re);
    rhs.a = mv(Iw, scl(re, invDt2));

    for (int k = (int)b.forces.size() - 1; k >= 0; --k) {       // newest force first == reference per-body list
        Force& f = w.forces[b.forces[k]];
        evalConstraint(w, f, alpha);
        int rows = rowCount(f);
        for (int r = 0; r < rows; ++r) {
            V3 Jl, Ja;
            evalJac(w, f, bi, r, Jl, Ja);
            float lamWarm = (f.stiffness[r] == FLT_MAX) ? f.lambda[r] : 0.0f;
            float want = f.penalty[r] * f.C[r] + lamWarm + f.motor[r];
            float fr = clampf(want, f.fmin[r], f.fmax[r]);
            rhs.l = add(rhs.l, scl(Jl, fr));
            rhs.a = add(rhs.a, scl(Ja, fr));
            float pen = f.penalty[r];
            if (pen > 0.0f && std::isfinite(pen)) {
                lhs.ll = madd(lhs.ll, mscl(outerp(Jl, Jl), pen));
                lhs.la = madd(lhs.la, mscl(outerp(Jl, Ja), pen));
                lhs.al = madd(lhs.al, mscl(outerp(Ja, Jl), pen));
                lhs.aa = madd(lhs.aa, mscl(outerp(Ja, Ja), pen));
                if (f.kind == MANIFOLD) {
                    M3 iIw = worldInvInertia(b);
                    V3 gy = scl(vabs(cross(Ja, mv(iIw, Ja))), fabsf(fr));
                    lhs.aa = madd(lhs.aa, m3diag(gy));
                }
            }
        }
    }
    V6 dx = schur6(lhs, rhs);
    b.pos = sub(b.pos, dx.l);
    Q4 dq{dx.a.x, dx.a.y, dx.a.z, 0.0f};
    b.rot = qunit(qsub(b.rot, qscl(qmul(dq, b.rot), 0.5f)));
    scrub3(b.pos, "position", b.id);
    scrub4(b.rot, "orientation", b.id);
    return dx;
}

static void stagePrimal(World& w, float alpha, const int* order, int n, float* dxOut) {
    if (order == nullptr) {
        for (int i = (int)w.bodies.size() - 1; i >= 0; --i) {
            if (w.bodies[i].invMass <= 0.0f) continue;
            V6 dx = primalBody(w, i, alpha);
            if (dxOut) { float* o = dxOut + 6 * i; o[0] = dx.l.x; o[1] = dx.l.y; o[2] = dx.l.z; o[3] = dx.a.x; o[4] = dx.a.y; o[5] = dx.a.z; }
        }
    } else {
        for (int k = 0; k < n; ++k) {
            int i = order[k];
            if (w.bodies[i].invMass <= 0.0f) continue;
            V6 dx = primalBody(w, i, alpha);
            if (dxOut) { float* o = dxOut + 6 * i; o[0] = dx.l.x; o[1] = dx.l.y; o[2] = dx.l.z; o[3] = dx.a.x; o[4] = dx.a.y; o[5] = dx.a.z; }
        }
    }
}

// solver.cpp:94-125
static float penaltyGain(const World& w, const Force& f, int row, float beta) {
    float lw = 0.0f, aw = 0.0f;
    for (int side = 0; side < 2; ++side) {
        int bi = side == 0 ? f.a : f.b;
        if (bi < 0) continue;
        V3 Jl, Ja;
        if (f.kind == IGNORE) continue;
        evalJac(w, f, bi, row, Jl, Ja);
        lw += len2(Jl); aw += len2(Ja);
    }
    float tot = lw + aw;
    if (tot < 1.0e-8f) return beta;
    float bl = beta, ba = beta * 0.01f;
    return (bl * lw + ba * aw) / tot;
}

// solver.cpp:411-430
static void stageDual(World& w, float alpha) {
    for (int s = (int)w.forces.size() - 1; s >= 0; --s) {
        Force& f = w.forces[s];
        evalConstraint(w, f, alpha);
        int rows = rowCount(f);
        for (int r = 0; r < rows; ++r) {
            if (f.stiffness[r] != FLT_MAX) continue;
            float lu = clampf(f.penalty[r] * f.C[r] + f.lambda[r], f.fmin[r], f.fmax[r]);
            bool active = lu > f.fmin[r] && lu < f.fmax[r];
            f.lambda[r] = lu;
            if (active) {
                float br = penaltyGain(w, f, r, w.beta);
                float cap = f.kind == MANIFOLD ? kManifoldPenaltyCap : kPenaltyMax;
                f.penalty[r] = fmin2(f.penalty[r] + br * fabsf(f.C[r]), cap);
            }
        }
    }
}

// solver.cpp:434-469
static void stageVelocity(World& w) {
    float dt = w.dt;
    w.statMaxLin = 0.0f; w.statMaxAng = 0.0f;
    for (int i = (int)w.bodies.size() - 1; i >= 0; --i) {
        Body& b = w.bodies[i];
        if (b.invMass <= 0.0f) continue;
        b.prevLin = b.lin; b.prevAng = b.ang;
        b.lin = dvd(sub(b.pos, b.pos0), dt);
        Q4 dq = qmul(b.rot, qconj(b.rot0));
        V3 av = scl(mk(dq.x, dq.y, dq.z), 2.0f / dt);
        if (dq.w < 0.0f) av = neg(av);
        b.ang = av;
        b.lin = scl(b.lin, 0.995f);
        b.ang = scl(b.ang, 0.97f);
        scrub3(b.lin, "linear velocity", b.id);
        scrub3(b.ang, "angular velocity", b.id);
        w.statMaxLin = fmax2(w.statMaxLin, len(b.lin));
        w.statMaxAng = fmax2(w.statMaxAng, len(b.ang));
    }
}

// solver.cpp:472-497
static void stageDiagnostics(World& w) {
    float maxPen = 0, maxViol = 0, maxLam = 0; int nc = 0, nm = 0;
    for (int s = (int)w.forces.size() - 1; s >= 0; --s) {
        const Force& f = w.forces[s];
        if (f.kind != MANIFOLD) continue;
        ++nm; nc += f.nct;
        const Body& A = w.bodies[f.a]; const Body& B = w.bodies[f.b];
        for (int i = 0; i < f.nct; ++i) {
            const Contact& c = f.ct[i];
            V3 pA = add(A.pos, qrot(A.rot, c.rA)), pB = add(B.pos, qrot(B.rot, c.rB));
            float sepn = dot(sub(pA, pB), c.normal);
            maxPen = fmax2(maxPen, fmax2(0.0f, -sepn));
            maxViol = fmax2(maxViol, fmax2(0.0f, kPenetrationSlop - sepn));
            maxLam = fmax2(maxLam, fabsf(f.lambda[i * 3]));
        }
    }
    w.diagF[0] = maxPen; w.diagF[1] = maxViol; w.diagF[2] = w.statMaxLin; w.diagF[3] = w.statMaxAng; w.diagF[4] = maxLam;
    w.diagI[0] = nc; w.diagI[1] = nm; w.diagI[2] = w.dynCount;
}

static void stepOnce(World& w, const int* order, int n) {    // solver.cpp:255-514
    ++w.stepIndex;
    stageBroadphase(w);
    stageInit(w);
    stagePredict(w);
    int total = w.iterations + (w.postStab ? 1 : 0);
    for (int it = 0; it < total; ++it) {
        float a = w.postStab ? (it < w.iterations ? 1.0f : 0.0f) : w.alpha;
        stagePrimal(w, a, order, n, nullptr);
        if (it < w.iterations) stageDual(w, a);
    }
    stageVelocity(w);
    stageDiagnostics(w);
    if (w.logOn) {                                           // solver.cpp:499-512
        int fq = w.logFreq > 0 ? w.logFreq : 1;
        if (w.stepIndex % fq == 0)
            std::printf("[Physics] step %d | manifolds: %d | contacts: %d | dyn bodies: %d | maxPen: %.6f | maxDrift: %.6f | maxLin: %.3f | maxAng: %.3f | maxLambda: %.3f\n",
                        w.stepIndex, w.diagI[1], w.diagI[0], w.diagI[2], w.diagF[0], w.diagF[1], w.diagF[2], w.diagF[3], w.diagF[4]);
    }
}

// ----------------------------------------------------------------------------- scenes (scenes.h)
static void ground(World& w, float sx, float sz) {           // scenes.h:27-31
    addBody(w, mk(sx, 1, sz), 0.0f, 0.5f, mk(0, -0.5f, 0), qid(), zero3(), zero3());
}
static float hash01(unsigned x) {                            // scenes.h:108-115
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return (x & 0x00FFFFFFU) / 16777215.0f;
}
static void stressGrid(World& w, int NX, int NY, int NZ, float spacingY, float startY, bool wide) {   // scenes.h:86-132
    clearWorld(w);
    float gx = 100.0f, gz = 100.0f;
    if (wide) { gx = fmax2(100.0f, NX * 1.15f + 20.0f); gz = fmax2(100.0f, NZ * 1.15f + 20.0f); }
    ground(w, gx, gz);
    w.iterations = 20; w.beta = 30000.0f; w.gamma = 0.995f;
    const V3 size = mk(1.0f, 1.0f, 1.0f);
    const float spacingXZ = 1.15f, jitterXZ = 0.04f, jitterY = 0.25f;
    for (int y = 0; y < NY; ++y)
        for (int z = 0; z < NZ; ++z)
            for (int x = 0; x < NX; ++x) {
                unsigned seed = (unsigned)(x + NX * (z + NZ * y) + 1);
                float jx = (hash01(seed * 9781U) * 2.0f - 1.0f) * jitterXZ;
                float jz = (hash01(seed * 6271U) * 2.0f - 1.0f) * jitterXZ;
                float jy = hash01(seed * 3343U) * jitterY;
                float px = (x - (NX - 1) * 0.5f) * spacingXZ + jx;
                float py = startY + y * spacingY + jy;
                float pz = (z - (NZ - 1) * 0.5f) * spacingXZ + jz;
                addBody(w, size, 1.0f, 0.5f, mk(px, py, pz), qid(), zero3(), zero3());
            }
}

static const char* kSceneNames[] = {"Empty", "Ground", "Stack", "Pyramid", "Wall", "TwoBlockDrop", "Stress1000",
                                    "Rod (WIP)", "Soft Body (WIP)"};                       // scenes.h:199-209
constexpr int kSceneCount = 9;

static void loadScene(World& w, int idx) {
    clearWorld(w);
    switch (idx) {
    case 0: break;                                                                           // :23-25
    case 1: ground(w, 100, 100); break;
    case 2:                                                                                  // :33-40
        ground(w, 100, 100);
        for (int i = 0; i < 10; ++i) addBody(w, mk(1, 1, 1), 1.0f, 0.5f, mk(0, i * 1.1f + 0.5f, 0), qid(), zero3(), zero3());
        break;
    case 3: {                                                                                // :42-53
        ground(w, 100, 100);
        const int P = 10;
        for (int y = 0; y < P; ++y)
            for (int x = 0; x < P - y; ++x) {
                float xp = (x - (P - y - 1) * 0.5f) * 1.1f;
                float yp = y * 1.05f + 0.5f;
                addBody(w, mk(1, 1, 1), 1.0f, 0.5f, mk(xp, yp, 0), qid(), zero3(), zero3());
            }
        break; }
    case 4: {                                                                                // :55-73
        ground(w, 100, 100);
        const int W = 8, H = 8;
        const V3 brick = mk(1.0f, 0.5f, 0.5f);
        const float sx = 1.03f, sy = 0.52f;
        const float baseY = brick.y * 0.5f;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                float xo = (i % 2 == 0) ? 0.0f : 0.5f * sx;
                float x = (j - (W - 1) * 0.5f) * sx + xo;
                float y = i * sy + baseY;
                addBody(w, brick, 1.0f, 0.4f, mk(x, y, -5), qid(), zero3(), zero3());
            }
        break; }
    case 5: {                                                                                // :75-84
        ground(w, 100, 100);
        addBody(w, mk(1.0f, 1.0f, 1.0f), 1.0f, 0.5f, mk(0.0f, 0.5f, 0.0f), qid(), zero3(), zero3());
        Q4 tilt = qaxis(mk(0.0f, 0.0f, 1.0f), 0.45f);
        addBody(w, mk(1.0f, 1.0f, 1.0f), 1.0f, 0.5f, mk(0.18f, 2.2f, 0.0f), tilt, zero3(), mk(0.0f, 0.0f, 1.0f));
        break; }
    case 6: stressGrid(w, 10, 10, 10, 2.0f, 20.0f, false); break;                            // :86-132
    case 7:                                                                                  // :138-151 (joints are placeholders upstream)
        for (int i = 0; i < 15; ++i)
            addBody(w, mk(0.25f, 1, 0.25f), (i == 0) ? 0.0f : 1.0f, 0.5f, mk(0, 10.0f - i * 1.0f, 0), qid(), zero3(), zero3());
        break;
    case 8: {                                                                                // :153-179
        ground(w, 100, 100);
        const int W = 10, H = 10;
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < H; ++j)
                addBody(w, mk(0.5f, 0.5f, 0.5f), 1.0f, 0.3f, mk(i * 0.6f - W * 0.3f, j * 0.6f + 2.0f, 0), qid(), zero3(), zero3());
        break; }
    default: break;
    }
}

} // namespace orc

// ============================================================================= C ABI
using namespace orc;
static World& W(void* h) { return *(World*)h; }
static V3 p3(const float* p) { return mk(p[0], p[1], p[2]); }
static Q4 p4(const float* p) { return Q4{p[0], p[1], p[2], p[3]}; }

extern "C" {

void* orc_create(void) { World* w = new World(); defaults(*w); return w; }
void orc_destroy(void* h) { delete (World*)h; }
void orc_clear(void* h) { clearWorld(W(h)); }
void orc_default_params(void* h) { defaults(W(h)); }
void orc_set_params(void* h, float dt, const float* g, int it, float alpha, float beta, float gamma, int ps) {
    World& w = W(h); w.dt = dt; w.g = p3(g); w.iterations = it; w.alpha = alpha; w.beta = beta; w.gamma = gamma; w.postStab = ps != 0;
}
void orc_get_params(void* h, float* o) {
    World& w = W(h); o[0] = w.dt; o[1] = w.g.x; o[2] = w.g.y; o[3] = w.g.z; o[4] = (float)w.iterations; o[5] = w.alpha; o[6] = w.beta; o[7] = w.gamma;
}
void orc_set_logging(void* h, int on, int freq) { W(h).logOn = on != 0; W(h).logFreq = freq; }

int orc_scene_count(void) { return kSceneCount; }
const char* orc_scene_name(int i) { return (i >= 0 && i < kSceneCount) ? kSceneNames[i] : ""; }
int orc_load_scene_index(void* h, int idx) { loadScene(W(h), idx); return (int)W(h).bodies.size(); }
int orc_load_scene(void* h, const char* name) {
    for (int i = 0; i < kSceneCount; ++i) if (std::strcmp(kSceneNames[i], name) == 0) return orc_load_scene_index(h, i);
    return -1;
}
int orc_load_stress_grid(void* h, int nx, int ny, int nz, float sy, float y0, int wide) {
    stressGrid(W(h), nx, ny, nz, sy, y0, wide != 0); return (int)W(h).bodies.size();
}

int orc_add_body(void* h, const float* size, float density, float friction, const float* pos, const float* q, const float* lin, const float* ang) {
    return addBody(W(h), p3(size), density, friction, p3(pos), p4(q), p3(lin), p3(ang));
}

void orc_add_joint(void* h, int a, int b, const float* anchorA, const float* anchorB, float linK, float angK) {   // joint.cpp:11-63
    World& w = W(h);
    int s = newForce(w, JOINT, a, b);
    Force& f = w.forces[s];
    const Body& B = w.bodies[b];
    if (a >= 0) {
        f.rA = p3(anchorA); f.rB = p3(anchorB);
        f.rel0 = qmul(qconj(w.bodies[a].rot), B.rot);
    } else {
        V3 wa = p3(anchorA);
        f.rA = wa;
        f.rB = mv(m3T(qmat(B.rot)), sub(wa, B.pos));
        f.rel0 = B.rot;
    }
    for (int i = 0; i < 3; ++i) { f.stiffness[i] = linK; f.lambda[i] = 0; f.penalty[i] = kPenaltyMin; }
    for (int i = 3; i < 6; ++i) { f.stiffness[i] = angK; f.lambda[i] = 0; f.penalty[i] = kPenaltyMin; }
}

void orc_add_spring(void* h, int a, int b, const float* anchorA, const float* anchorB, float k, float rest) {      // spring.cpp:10-30
    World& w = W(h);
    int s = newForce(w, SPRING, a, b);
    Force& f = w.forces[s];
    f.rA = p3(anchorA); f.rB = p3(anchorB); f.rest = rest;
    f.stiffness[0] = k;
    if (f.rest < 0) { V3 pA, pB; springEnds(w, f, pA, pB); f.rest = len(sub(pA, pB)); }
    f.lambda[0] = 0.0f; f.penalty[0] = kPenaltyMin; f.fmin[0] = -FLT_MAX; f.fmax[0] = FLT_MAX;
}

void orc_add_ignore(void* h, int a, int b) { newForce(W(h), IGNORE, a, b); }

void orc_step(void* h, int n) { for (int i = 0; i < n; ++i) stepOnce(W(h), nullptr, 0); }
void orc_step_ordered(void* h, const int* order, int n) { stepOnce(W(h), order, n); }

void orc_stage_broadphase(void* h) { ++W(h).stepIndex; stageBroadphase(W(h)); }
void orc_stage_init(void* h) { stageInit(W(h)); }
void orc_stage_predict(void* h) { stagePredict(W(h)); }
void orc_stage_primal(void* h, float alpha, const int* order, int n, float* dx) { stagePrimal(W(h), alpha, order, n, dx); }
void orc_stage_dual(void* h, float alpha) { stageDual(W(h), alpha); }
void orc_stage_velocity(void* h) { stageVelocity(W(h)); }
void orc_stage_diagnostics(void* h) { stageDiagnostics(W(h)); }

int orc_num_bodies(void* h) { return (int)W(h).bodies.size(); }
int orc_body_id(void* h, int i) { return W(h).bodies[i].id; }

void orc_get_state(void* h, float* o) {
    for (const Body& b : W(h).bodies) {
        *o++ = b.pos.x; *o++ = b.pos.y; *o++ = b.pos.z;
        *o++ = b.rot.x; *o++ = b.rot.y; *o++ = b.rot.z; *o++ = b.rot.w;
        *o++ = b.lin.x; *o++ = b.lin.y; *o++ = b.lin.z;
        *o++ = b.ang.x; *o++ = b.ang.y; *o++ = b.ang.z;
    }
}
void orc_set_state(void* h, const float* in) {
    for (Body& b : W(h).bodies) { b.pos = p3(in); in += 3; b.rot = p4(in); in += 4; b.lin = p3(in); in += 3; b.ang = p3(in); in += 3; }
}
void orc_get_prev_linvel(void* h, float* o) { for (const Body& b : W(h).bodies) { *o++ = b.prevLin.x; *o++ = b.prevLin.y; *o++ = b.prevLin.z; } }
void orc_set_prev_linvel(void* h, const float* in) { for (Body& b : W(h).bodies) { b.prevLin = p3(in); in += 3; } }
void orc_get_body_props(void* h, float* o) {
    for (const Body& b : W(h).bodies) {
        *o++ = b.size.x; *o++ = b.size.y; *o++ = b.size.z; *o++ = b.mass; *o++ = b.invMass;
        *o++ = b.inertia.c[0].x; *o++ = b.inertia.c[1].y; *o++ = b.inertia.c[2].z; *o++ = b.friction; *o++ = b.radius;
    }
}
void orc_get_diagnostics(void* h, float* f5, int* i3) {
    World& w = W(h); for (int i = 0; i < 5; ++i) f5[i] = w.diagF[i]; for (int i = 0; i < 3; ++i) i3[i] = w.diagI[i];
}

int orc_num_manifolds(void* h) { int n = 0; for (const Force& f : W(h).forces) n += (f.alive && f.kind == MANIFOLD) ? 1 : 0; return n; }

void orc_get_manifolds(void* h, int* ints, int* feats, int* stick, float* flts) {   // newest first, same layout as ref_get_manifolds
    World& w = W(h);
    for (int s = (int)w.forces.size() - 1; s >= 0; --s) {
        const Force& f = w.forces[s];
        if (!f.alive || f.kind != MANIFOLD) continue;
        *ints++ = f.a; *ints++ = f.b; *ints++ = f.nct;
        *flts++ = f.mu;
        for (int i = 0; i < 4; ++i) {
            const Contact& c = f.ct[i];
            bool live = i < f.nct;
            *feats++ = live ? c.feature : 0;
            *stick++ = live ? (c.stick ? 1 : 0) : 0;
            const float v[14] = {c.rA.x, c.rA.y, c.rA.z, c.rB.x, c.rB.y, c.rB.z, c.normal.x, c.normal.y, c.normal.z,
                                 c.penetration, c.C0n, c.C0t.x, c.C0t.y, c.C0t.z};
            for (int k = 0; k < 14; ++k) *flts++ = live ? v[k] : 0.0f;
        }
        for (int k = 0; k < 12; ++k) *flts++ = k < f.nct * 3 ? f.lambda[k] : 0.0f;
        for (int k = 0; k < 12; ++k) *flts++ = k < f.nct * 3 ? f.penalty[k] : 0.0f;
    }
}

// Replaces the world's manifold set (test seam: lets a parity test start BOTH sides of a per-stage comparison from the same
// warm-start history).  Same layout and order as orc_get_manifolds (newest first), so get -> set is the identity.
void orc_set_manifolds(void* h, int count, const int* ints, const int* feats, const int* stick, const float* flts) {
    World& w = W(h);
    for (int s = 0; s < (int)w.forces.size(); ++s) if (w.forces[s].alive && w.forces[s].kind == MANIFOLD) killForce(w, s);
    compactForces(w);
    for (int m = count - 1; m >= 0; --m) {               // oldest first = creation order
        const int* I = ints + 3 * m; const float* F = flts + 81 * m;
        int s = newForce(w, MANIFOLD, I[0], I[1]);
        Force& f = w.forces[s];
        f.nct = I[2]; f.mu = F[0];
        for (int i = 0; i < 4; ++i) {
            Contact& c = f.ct[i]; const float* v = F + 1 + 14 * i;
            c.feature = feats[4 * m + i]; c.stick = stick[4 * m + i] != 0;
            c.rA = mk(v[0], v[1], v[2]); c.rB = mk(v[3], v[4], v[5]); c.normal = mk(v[6], v[7], v[8]);
            c.penetration = v[9]; c.C0n = v[10]; c.C0t = mk(v[11], v[12], v[13]);
        }
        for (int k = 0; k < 12; ++k) {
            f.lambda[k] = F[57 + k]; f.penalty[k] = F[69 + k];
            if (k < f.nct * 3) { f.stiffness[k] = FLT_MAX; f.motor[k] = 0.0f; }
        }
    }
}

int orc_overlap_pairs(void* h, int* pairs, int cap) {       // the sphere test of solver.cpp:264-266, no exclusion
    World& w = W(h);
    int n = (int)w.bodies.size(), cnt = 0;
    for (int a = n - 1; a >= 0; --a)
        for (int b = a - 1; b >= 0; --b) {
            V3 dp = sub(w.bodies[a].pos, w.bodies[b].pos);
            float r = w.bodies[a].radius + w.bodies[b].radius;
            if (dot(dp, dp) <= r * r) { if (cnt < cap) { pairs[2 * cnt] = a; pairs[2 * cnt + 1] = b; } ++cnt; }
        }
    return cnt;
}

int orc_collide(const float* a, const float* b, int* feats, float* out) {
    PoseRef pa{p3(a + 3), p4(a + 6)}, pb{p3(b + 3), p4(b + 6)};
    Contact c[4];
    int n = collide(pa, p3(a), pb, p3(b), c);
    for (int i = 0; i < n; ++i) {
        feats[i] = c[i].feature;
        float* o = out + i * 10;
        o[0] = c[i].rA.x; o[1] = c[i].rA.y; o[2] = c[i].rA.z; o[3] = c[i].rB.x; o[4] = c[i].rB.y; o[5] = c[i].rB.z;
        o[6] = c[i].normal.x; o[7] = c[i].normal.y; o[8] = c[i].normal.z; o[9] = c[i].penetration;
    }
    return n;
}

void orc_solve6x6(const float* lhs, const float* rhs, float* out) {
    M66 A; M3* blk[4] = {&A.ll, &A.la, &A.al, &A.aa};
    for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) blk[b]->c[c] = p3(lhs + b * 9 + c * 3);
    V6 x = schur6(A, V6{p3(rhs), p3(rhs + 3)});
    out[0] = x.l.x; out[1] = x.l.y; out[2] = x.l.z; out[3] = x.a.x; out[4] = x.a.y; out[5] = x.a.z;
}
// Solver::pick, solver.cpp:145-228.  Returns the creation index of the closest dynamic body hit, or -1.
int orc_pick(void* h, const float* origin3, const float* dir3, float* local3) {
    World& w = W(h);
    const float eps = 1.0e-6f;
    float bestT = FLT_MAX; int best = -1; V3 bestLocal = zero3();
    V3 origin = p3(origin3), rd = p3(dir3);
    float dl2 = len2(rd);
    if (dl2 < eps) return -1;
    rd = dvd(rd, sqrtf(dl2));
    for (int i = (int)w.bodies.size() - 1; i >= 0; --i) {          // list order: newest first
        const Body& b = w.bodies[i];
        if (b.invMass <= 0.0f) continue;
        Q4 inv = qconj(b.rot);
        V3 lo = qrot(inv, sub(origin, b.pos)), ld = qrot(inv, rd), half = scl(b.size, 0.5f);
        float tIn = 0.0f, tOut = FLT_MAX; bool hit = true;
        for (int ax = 0; ax < 3; ++ax) {
            float o = comp(lo, ax), d = comp(ld, ax), mn = -comp(half, ax), mx = comp(half, ax);
            if (fabsf(d) < eps) { if (o < mn || o > mx) { hit = false; break; } continue; }
            float invD = 1.0f / d, t0 = (mn - o) * invD, t1 = (mx - o) * invD;
            if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
            tIn = fmax2(tIn, t0); tOut = fmin2(tOut, t1);
            if (tIn > tOut) { hit = false; break; }
        }
        if (!hit) continue;
        float tHit = (tIn >= 0.0f) ? tIn : tOut;
        if (tHit < 0.0f) continue;
        if (tHit < bestT) { bestT = tHit; best = i; bestLocal = add(lo, scl(ld, tHit)); }
    }
    if (best < 0) return -1;
    local3[0] = bestLocal.x; local3[1] = bestLocal.y; local3[2] = bestLocal.z;
    return best;
}

void orc_solve3(const float* A, const float* b, float* out) {
    V3 x = ldl3(m3(p3(A), p3(A + 3), p3(A + 6)), p3(b)); out[0] = x.x; out[1] = x.y; out[2] = x.z;
}

} // extern "C"
