// avbd_collide.cuh — OBB-vs-OBB SAT and contact generation, device side.
//
// Behavioural contract: alxspiker/avbd-demo3d source/collision.cpp:56-489
// (Manifold::collide with flip=false).  Same axis order, same strict ">" tie
// breaking, same epsilons, same quantised feature keys; compiled with
// -fmad=false so the outputs are bit-identical to the CPU reference.
//
// The SAT is split in two so the pipeline can cull before it builds:
//   sat_test()      15 axes -> separated? + winning axis packed in one int
//   build_contacts() recomputes the winning axis' normal (same expression =>
//                    same bits) and emits <=4 contacts.
#pragma once
#include "avbd_math.cuh"

namespace avbd {

struct Obb { V3 c, h, ax[3]; };

struct RawContact {      // collision.cpp:176-206 output, before Manifold::initialize
    int feature;
    V3 rA, rB, normal;
};

AVBD_HD Obb make_obb(V3 pos, Q4 rot, V3 size) {             // collision.cpp:56-66
    Obb b; b.c = pos; b.h = size * 0.5f;
    M3 R = qmat(rot); b.ax[0] = R.c[0]; b.ax[1] = R.c[1]; b.ax[2] = R.c[2];
    return b;
}
AVBD_HD float adot(V3 a, V3 b) { return fabsf(dot(a, b)); }
// Indexing by a run-time axis number with selects: a dynamically indexed ax[k] would push the whole Obb to local memory.
AVBD_HD V3 obb_axis(const Obb& b, int k) { return k == 0 ? b.ax[0] : (k == 1 ? b.ax[1] : b.ax[2]); }

// One axis.  Returns false when separated beyond the persistence margin.
// `sep`/`n` are only meaningful when `counted` comes back true (degenerate
// axes are skipped, collision.cpp:211-214).
AVBD_HD bool sat_axis(const Obb& A, const Obb& B, V3 d, V3 axis, bool& counted, float& sep, V3& n) {   // collision.cpp:208-247
    counted = false; sep = 0.0f; n = zero3();
    float l2 = len2(axis);
    if (l2 < kSatEps) return true;
    n = axis / sqrtf(l2);
    if (dot(n, d) < 0.0f) n = -n;
    float dist = fabsf(dot(d, n));
    float ra = A.h.x * adot(n, A.ax[0]) + A.h.y * adot(n, A.ax[1]) + A.h.z * adot(n, A.ax[2]);
    float rb = B.h.x * adot(n, B.ax[0]) + B.h.y * adot(n, B.ax[1]) + B.h.z * adot(n, B.ax[2]);
    sep = dist - (ra + rb);
    if (sep > kCollisionMargin) return false;
    counted = true;
    return true;
}

AVBD_HD V3 sat_axis_dir(const Obb& A, const Obb& B, int k) {
    // k: 0-2 face of A, 3-5 face of B, 6-14 edge i*3+j
    if (k < 3) return obb_axis(A, k);
    if (k < 6) return obb_axis(B, k - 3);
    int e = k - 6;
    return cross(obb_axis(A, e / 3), obb_axis(B, e % 3));
}

// The 15-axis test in two halves (the cull kernel compacts the pairs that survive the face axes before it runs the
// edge axes, so its lanes stay dense).  Same axes, same order, same expressions as one loop over k = 0..14.
struct SatFaces { bool valid; float sep; int k; };
// Axes 0-5 (faces of A, faces of B).  Returns false when one of them separates the boxes.
AVBD_HD bool sat_faces(const Obb& A, const Obb& B, SatFaces& f) {
    V3 d = B.c - A.c;
    f.valid = false; f.sep = -FLT_MAX; f.k = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        bool counted; float sep; V3 n;
        if (!sat_axis(A, B, d, k < 3 ? A.ax[k % 3] : B.ax[k % 3], counted, sep, n)) return false;
        if (!counted) continue;
        if (!f.valid || sep > f.sep) { f.valid = true; f.sep = sep; f.k = k; }
    }
    return true;
}
// Axes 6-14 (edge x edge) and the face-vs-edge preference rule (collision.cpp:459-468).  Returns 0 when separated / no valid
// face axis, else 1 | (axisIndex << 1) with axisIndex in [0,15) naming the chosen axis.
AVBD_HD int sat_edges(const Obb& A, const Obb& B, const SatFaces& f) {
    V3 d = B.c - A.c;
    bool edgeValid = false;
    float edgeSep = -FLT_MAX;
    int edgeK = 0;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        bool counted; float sep; V3 n;
        if (!sat_axis(A, B, d, cross(A.ax[e / 3], B.ax[e % 3]), counted, sep, n)) return 0;
        if (!counted) continue;
        if (!edgeValid || sep > edgeSep) { edgeValid = true; edgeSep = sep; edgeK = 6 + e; }
    }
    if (!f.valid) return 0;
    int best = f.k;
    if (edgeValid && kEdgeRelTol * edgeSep > f.sep + kEdgeAbsTol) best = edgeK;
    return 1 | (best << 1);
}
// Full 15-axis test (collision.cpp:420-468).
AVBD_HD int sat_test(const Obb& A, const Obb& B) {
    SatFaces f;
    if (!sat_faces(A, B, f)) return 0;
    return sat_edges(A, B, f);
}

// Per-body part of the SAT's face axes.  The cull tests ~9 sphere pairs per body of a dense pile, and for a face axis of box S
// everything but the cross terms depends on S alone: the normalised axis n = ax / |ax| and S's own projection radius on it.
// make_obb_frame evaluates those with the very expressions sat_axis uses (same operands, same order), once per body per step;
// sat_faces_frames then reproduces sat_faces bit for bit.  (The sign flip of n in sat_axis changes neither |dot(d, n)| nor the
// |dot(n, ax)| terms: negating every product negates the rounded sums exactly.)
struct ObbFrame { V3 c, h, ax[3], n[3]; float raSelf[3]; };       // raSelf[k] < 0: axis k is degenerate and skipped (collision.cpp:211-214)
AVBD_HD ObbFrame make_obb_frame(V3 pos, Q4 rot, V3 size) {
    Obb b = make_obb(pos, rot, size);
    ObbFrame f; f.c = b.c; f.h = b.h;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f.ax[k] = b.ax[k];
        float l2 = len2(b.ax[k]);
        if (l2 < kSatEps) { f.n[k] = zero3(); f.raSelf[k] = -1.0f; continue; }
        V3 n = b.ax[k] / sqrtf(l2);
        f.n[k] = n;
        f.raSelf[k] = b.h.x * adot(n, b.ax[0]) + b.h.y * adot(n, b.ax[1]) + b.h.z * adot(n, b.ax[2]);
    }
    return f;
}
AVBD_HD Obb frame_obb(const ObbFrame& f) { Obb b; b.c = f.c; b.h = f.h; b.ax[0] = f.ax[0]; b.ax[1] = f.ax[1]; b.ax[2] = f.ax[2]; return b; }
AVBD_HD bool sat_faces_frames(const ObbFrame& A, const ObbFrame& B, SatFaces& f) {
    V3 d = B.c - A.c;
    f.valid = false; f.sep = -FLT_MAX; f.k = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float self = k < 3 ? A.raSelf[k % 3] : B.raSelf[k % 3];
        if (self < 0.0f) continue;
        const V3 n = k < 3 ? A.n[k % 3] : B.n[k % 3];
        const float dist = fabsf(dot(d, n));
        const float ra = k < 3 ? self : A.h.x * adot(n, A.ax[0]) + A.h.y * adot(n, A.ax[1]) + A.h.z * adot(n, A.ax[2]);
        const float rb = k < 3 ? B.h.x * adot(n, B.ax[0]) + B.h.y * adot(n, B.ax[1]) + B.h.z * adot(n, B.ax[2]) : self;
        const float sep = dist - (ra + rb);
        if (sep > kCollisionMargin) return false;
        if (!f.valid || sep > f.sep) { f.valid = true; f.sep = sep; f.k = k; }
    }
    return true;
}

AVBD_HD void face_axes(const Obb& b, int k, V3& u, V3& v, float& eu, float& ev) {   // collision.cpp:73-92
    if (k == 0) { u = b.ax[1]; v = b.ax[2]; eu = b.h.y; ev = b.h.z; }
    else if (k == 1) { u = b.ax[0]; v = b.ax[2]; eu = b.h.x; ev = b.h.z; }
    else { u = b.ax[0]; v = b.ax[1]; eu = b.h.x; ev = b.h.y; }
}
AVBD_HD V3 sel3(bool c, V3 a, V3 b) { return mk3(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }
AVBD_HD Obb sel_obb(bool c, const Obb& a, const Obb& b) {
    Obb r; r.c = sel3(c, a.c, b.c); r.h = sel3(c, a.h, b.h);
    r.ax[0] = sel3(c, a.ax[0], b.ax[0]); r.ax[1] = sel3(c, a.ax[1], b.ax[1]); r.ax[2] = sel3(c, a.ax[2], b.ax[2]);
    return r;
}

constexpr int kMaxPoly = 16;

// Scratch of the Sutherland-Hodgman clipper: two buffers of kMaxPoly vertices.  Per-thread arrays of this size live in
// local memory; at a million manifolds per step that traffic does not fit L2 and spills to HBM, so the device kernels keep
// the polygons in shared memory (PolyShared, one column per thread) and only host code uses PolyLocal.
struct PolyLocal {
    V3 v[2][kMaxPoly];
    AVBD_HD V3 get(int b, int i) const { return v[b][i]; }
    AVBD_HD void set(int b, int i, V3 x) { v[b][i] = x; }
};
#ifdef __CUDACC__
struct PolyShared {          // element (buffer b, vertex i, component c) of this thread at col[((b * kMaxPoly + i) * 3 + c) * stride]
    float* col; int stride;
    __device__ __forceinline__ V3 get(int b, int i) const {
        const float* p = col + (size_t)((b * kMaxPoly + i) * 3) * stride;
        return mk3(p[0], p[stride], p[2 * stride]);
    }
    __device__ __forceinline__ void set(int b, int i, V3 x) {
        float* p = col + (size_t)((b * kMaxPoly + i) * 3) * stride;
        p[0] = x.x; p[stride] = x.y; p[2 * stride] = x.z;
    }
};
constexpr int kPolyFloatsPerThread = 2 * kMaxPoly * 3;
#endif

// Sutherland-Hodgman against dot(n,p) <= off, buffer `src` -> buffer `src ^ 1`.  collision.cpp:136-174
template <class Poly>
AVBD_HD int clip_poly(Poly& poly, int src, int nin, V3 n, float off) {
    if (nin <= 0) return 0;
    int no = 0, dst = src ^ 1;
    V3 a = poly.get(src, nin - 1);
    float da = dot(n, a) - off;
    for (int i = 0; i < nin; ++i) {
        V3 b = poly.get(src, i);
        float db = dot(n, b) - off;
        bool ain = da <= kPlaneEps, bin = db <= kPlaneEps;
        if (ain != bin) {
            float t = 0.0f, den = da - db;
            if (fabsf(den) > kSatEps) t = clampf(da / den, 0.0f, 1.0f);
            if (no < kMaxPoly) poly.set(dst, no++, a + (b - a) * t);
        }
        if (bin && no < kMaxPoly) poly.set(dst, no++, b);
        a = b; da = db;
    }
    return no;
}

// Contacts leave the builder one at a time through `emit(feature, rA, rB, normal)` (in manifold order) so a consumer can
// finish each one (warm start, C0, store) without holding the other three.
struct ContactDedupe {        // midpoints of the contacts emitted so far (collision.cpp:176-190)
    V3 mid[4]; int n;
};

// collision.cpp:176-206
template <class Emit>
AVBD_HD bool push_contact(V3 posA, Q4 rotA, V3 posB, Q4 rotB, ContactDedupe& dd, V3 xA, V3 xB, int key, V3 nBA, Emit& emit) {
    V3 mid = (xA + xB) * 0.5f;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < dd.n && len2(mid - dd.mid[i]) < kMergeDistSq) return false;
    if (dd.n >= 4) return false;
    emit(key, qrot(qconj(rotA), xA - posA), qrot(qconj(rotB), xB - posB), nBA);
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i == dd.n) dd.mid[i] = mid;
    ++dd.n;
    return true;
}

// collision.cpp:313-394 (+ helpers :94-134)
template <class Poly, class Emit>
AVBD_HD int face_manifold(V3 posA, Q4 rotA, V3 posB, Q4 rotB, const Obb& A, const Obb& B, bool refIsA, int refAxis, V3 nAB, Poly& poly, Emit& emit) {
    Obb R = sel_obb(refIsA, A, B);
    Obb I = sel_obb(refIsA, B, A);
    V3 outward = refIsA ? nAB : -nAB;
    V3 nBA = -nAB;
    V3 rax = obb_axis(R, refAxis);
    float sgn = dot(outward, rax) >= 0.0f ? 1.0f : -1.0f;
    V3 fn = rax * sgn;
    V3 fc = R.c + fn * comp(R.h, refAxis);
    V3 fu, fv; float eu, ev;
    face_axes(R, refAxis, fu, fv, eu, ev);
    int inc = 0; float bestd = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < 3; ++i) { float d = adot(I.ax[i], fn); if (d > bestd) { bestd = d; inc = i; } }
    V3 iax = obb_axis(I, inc);
    float isg = dot(iax, fn) > 0.0f ? -1.0f : 1.0f;
    V3 inrm = iax * isg;
    V3 icen = I.c + inrm * comp(I.h, inc);
    V3 iu, iv; float ieu, iev;
    face_axes(I, inc, iu, iv, ieu, iev);
    poly.set(0, 0, (icen + iu * ieu) + iv * iev);
    poly.set(0, 1, (icen - iu * ieu) + iv * iev);
    poly.set(0, 2, (icen - iu * ieu) - iv * iev);
    poly.set(0, 3, (icen + iu * ieu) - iv * iev);
    int cnt = 4;
    cnt = clip_poly(poly, 0, cnt, fu, dot(fu, fc) + eu); if (!cnt) return 0;
    V3 nu = -fu;
    cnt = clip_poly(poly, 1, cnt, nu, dot(nu, fc) + eu); if (!cnt) return 0;
    cnt = clip_poly(poly, 0, cnt, fv, dot(fv, fc) + ev); if (!cnt) return 0;
    V3 nv = -fv;
    cnt = clip_poly(poly, 1, cnt, nv, dot(nv, fc) + ev); if (!cnt) return 0;

    ContactDedupe dd; dd.n = 0;
    int prefix = ((refIsA ? 0 : 1) << 24) | ((refAxis & 0xFF) << 16) | ((inc & 0xFF) << 8);
    // Same walk as `for (i = 0; i < cnt && dd.n < 4; ++i) { if (dist > margin) continue; ... }`, in two passes: a cheap one marks the
    // vertices within the margin, the expensive one (projection, feature key, two inverse rotations, dedupe) then visits only those, in the
    // same ascending order.  On the device the lanes of a warp hold different polygons: with the single loop a lane did the expensive part
    // at ITS vertices' indices and idled at the others' (ncu: 9 of 32 lanes active there); now every lane's k-th contact runs together.
    unsigned hits = 0u;
    for (int i = 0; i < cnt; ++i) {
        V3 pi = poly.get(0, i);
        float dist = dot(pi - fc, fn);
        if (!(dist > kCollisionMargin)) hits |= 1u << i;
    }
    while (hits != 0u && dd.n < 4) {
#ifdef __CUDA_ARCH__
        const int i = __ffs((int)hits) - 1;
#else
        const int i = __builtin_ctz(hits);
#endif
        hits &= hits - 1u;
        V3 pi = poly.get(0, i);
        float dist = dot(pi - fc, fn);
        V3 pr = pi - fn * dist;
        V3 xA = refIsA ? pr : pi, xB = refIsA ? pi : pr;
        V3 rel = pr - fc;
        float uc = dot(rel, fu), vc = dot(rel, fv);
        float un = (eu > kSatEps) ? (uc / eu) : 0.0f;
        float vn = (ev > kSatEps) ? (vc / ev) : 0.0f;
        int qu = (int)floorf(clampf((un + 1.0f) * 7.5f, 0.0f, 15.0f));
        int qv = (int)floorf(clampf((vn + 1.0f) * 7.5f, 0.0f, 15.0f));
        int key = prefix | ((qu & 0x0F) << 4) | (qv & 0x0F);
        push_contact(posA, rotA, posB, rotB, dd, xA, xB, key, nBA, emit);
    }
    return dd.n;
}

AVBD_HD void support_edge(const Obb& b, int k, V3 dir, V3& e0, V3& e1) {      // collision.cpp:249-263
    int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    V3 a0 = obb_axis(b, k), a1 = obb_axis(b, k1), a2 = obb_axis(b, k2);
    float s1 = dot(dir, a1) >= 0.0f ? 1.0f : -1.0f;
    float s2 = dot(dir, a2) >= 0.0f ? 1.0f : -1.0f;
    V3 ec = (b.c + a1 * (comp(b.h, k1) * s1)) + a2 * (comp(b.h, k2) * s2);
    e0 = ec - a0 * comp(b.h, k);
    e1 = ec + a0 * comp(b.h, k);
}

AVBD_HD void seg_closest(V3 p0, V3 p1, V3 q0, V3 q1, V3& c0, V3& c1) {        // collision.cpp:265-311
    V3 d1 = p1 - p0, d2 = q1 - q0, r = p0 - q0;
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
    c0 = p0 + d1 * s;
    c1 = q0 + d2 * t;
}

// Emits the contacts for the axis `satCode` names (from sat_test on the same
// poses).  collision.cpp:470-476
template <class Poly, class Emit>
AVBD_HD int build_contacts_emit(V3 posA, Q4 rotA, V3 sizeA, V3 posB, Q4 rotB, V3 sizeB, int satCode, Poly& poly, Emit& emit) {
    Obb A = make_obb(posA, rotA, sizeA), B = make_obb(posB, rotB, sizeB);
    int k = satCode >> 1;
    V3 d = B.c - A.c;
    bool counted; float sep; V3 n;
    sat_axis(A, B, d, sat_axis_dir(A, B, k), counted, sep, n);   // same ops as the winning test => same normal bits
    if (k >= 6) {                                                // buildEdgeContact, collision.cpp:396-416
        int ia = (k - 6) / 3, ib = (k - 6) % 3;
        V3 a0, a1, b0, b1, xA, xB;
        support_edge(A, ia, n, a0, a1);
        support_edge(B, ib, -n, b0, b1);
        seg_closest(a0, a1, b0, b1, xA, xB);
        ContactDedupe dd; dd.n = 0;
        int key = (2 << 24) | ((ia & 0xFF) << 8) | (ib & 0xFF);
        push_contact(posA, rotA, posB, rotB, dd, xA, xB, key, -n, emit);
        return dd.n;
    }
    // one instance for both reference sides: two inlined copies would run back to back in every mixed warp
    bool refIsA = k < 3;
    return face_manifold(posA, rotA, posB, rotB, A, B, refIsA, refIsA ? k : k - 3, n, poly, emit);
}

// Array form (host mirror, parity harness): contacts collected into out[4].
struct RawContactCollector {
    RawContact* out; int n;
    AVBD_HD void operator()(int feature, V3 rA, V3 rB, V3 normal) {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i == n) { out[i].feature = feature; out[i].rA = rA; out[i].rB = rB; out[i].normal = normal; }
        ++n;
    }
};
template <class Poly>
AVBD_HD int build_contacts(V3 posA, Q4 rotA, V3 sizeA, V3 posB, Q4 rotB, V3 sizeB, int satCode, RawContact* out, Poly& poly) {
    RawContactCollector col{out, 0};
    return build_contacts_emit(posA, rotA, sizeA, posB, rotB, sizeB, satCode, poly, col);
}
inline int build_contacts(V3 posA, Q4 rotA, V3 sizeA, V3 posB, Q4 rotB, V3 sizeB, int satCode, RawContact* out) {   // host only
    PolyLocal poly;
    return build_contacts(posA, rotA, sizeA, posB, rotB, sizeB, satCode, out, poly);
}

} // namespace avbd
