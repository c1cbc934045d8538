// avbd_solve.cu — the per-iteration solver kernels (primal block solve per colour, dual / penalty ramp) and
// their launchers.  This translation unit is compiled WITH FMA contraction: its outputs are held to an FP32
// tolerance against the reference (BASELINE.json north_star), unlike the collision path in avbd_engine.cu.
//
// The reference walks bodies serially (Gauss-Seidel, solver.cpp:344); here bodies of one colour share no
// manifold, so a colour is one parallel phase, one contact visit (computeConstraint + 3 rows) per thread:
//   large worlds   per colour primal_sweep_warp: one contact visit per lane in warp-private pipelines, finished bodies solved in batches
//   small worlds   solve_loop_cluster: the whole iteration loop in one thread-block cluster, a tile of bodies per CTA and phase
//
// Deferred dual.  The dual / penalty-ramp pass of iteration k (solver.cpp:411-430) reads the poses left by sweep k, and
// nothing moves between it and sweep k+1.  When sweep k+1 reaches the FIRST visit of a contact (its other endpoint is
// static or has a higher colour) neither endpoint has moved yet, so that visit sees exactly the poses the dual pass
// would have seen: it applies the pending dual update in registers (same operations, same order), then evaluates the
// primal rows, and writes lambda / penalty back once.  A step therefore runs ONE stand-alone dual pass (after the last
// sweep, fused with the contact diagnostics) instead of `iterations` of them.  biasDual (= clamp(1 - alpha, 0, 1) of the pending pass; manifold.cpp:179) < 0 = nothing pending (first sweep
// of a step, stage API).
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include "avbd_launch.h"
#include "avbd_body.cuh"
#include "avbd_forces.cuh"

namespace avbd {
namespace cg = cooperative_groups;

__device__ __forceinline__ ContactState load_contact(const ManifoldSet& ms, int ci) {
    ContactLP q = ms.lp[ci];
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], q.l, q.p);
}

// ------------------------------------------------------------------ primal
// Rows of the user forces touching body i (lane 0 of the group, serial).
__device__ void accumulate_user_forces(BodySystem& s, const ForceView& fv, const BodyPose* pose, int i, V3 pos, Q4 rot, const M3& invIw) {
    for (int k = fv.adjStart[i]; k < fv.adjStart[i + 1]; ++k) {
        int e = fv.adj[k]; int idx = e >> 2; bool isA = e & 1;
        if (e & 2) {
            const SpringRec& sp = fv.springs[idx];
            int other = isA ? sp.b : sp.a;
            V3 po = zero3(); Q4 qo = qid();
            if (other >= 0) { BodyPose o = pose[other]; po = xyz(o.pos); qo = quat(o.rot); }
            bool hasA = sp.a >= 0;
            V3 pA = isA ? pos : po, pB = isA ? po : pos; Q4 qA = isA ? rot : qo, qB = isA ? qo : rot;
            float C = spring_constraint(sp, hasA, pA, qA, pB, qB);
            V3 Jl, Ja;
            spring_jacobian(sp, hasA, pA, qA, pB, qB, isA, Jl, Ja);
            float lamWarm = (sp.k == FLT_MAX) ? sp.lambda : 0.0f;
            float f = clampf(sp.penalty * C + lamWarm + sp.motor, -FLT_MAX, FLT_MAX);
            accumulate_row(s, Jl, Ja, f, sp.penalty, false, invIw);
        } else {
            const JointRec& j = fv.joints[idx];
            int other = isA ? j.b : j.a;
            V3 po = zero3(); Q4 qo = qid();
            if (other >= 0) { BodyPose o = pose[other]; po = xyz(o.pos); qo = quat(o.rot); }
            bool hasA = j.a >= 0;
            V3 pA = isA ? pos : po, pB = isA ? po : pos; Q4 qA = isA ? rot : qo, qB = isA ? qo : rot;
            ForceEval ev;
            joint_constraint(j, hasA, pA, qA, pB, qB, ev);
            for (int r = 0; r < 6; ++r) {
                V3 Jl, Ja;
                joint_jacobian(j, isA, rot, r, Jl, Ja);
                float lamWarm = (j.stiffness[r] == FLT_MAX) ? j.lambda[r] : 0.0f;
                float f = clampf(j.penalty[r] * ev.C[r] + lamWarm + j.motor[r], ev.fmin[r], ev.fmax[r]);
                accumulate_row(s, Jl, Ja, f, j.penalty[r], false, invIw);
            }
        }
    }
}

// L2 residency hints.  Poses (32 B per body) are gathered at random by every contact visit, several times per sweep, while
// contact data streams past once per visit; without a hint the stream evicts the poses and each pose gather goes back to
// HBM (dragging a 64-byte granule for 32 useful bytes).  Pose loads / stores carry an evict_last policy.
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld4_keep(const float4* ptr, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_keep(float4* ptr, float4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ BodyPose load_pose_keep(const BodyPose* p, unsigned long long pol) {
    BodyPose r; r.pos = ld4_keep(&p->pos, pol); r.rot = ld4_keep(&p->rot, pol); return r;
}
// The same through L2 only: inside a persistent loop another SM may have rewritten the pose since this SM's L1 last saw it.
__device__ __forceinline__ float4 ld4_keep_l2(const float4* ptr, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
    return v;
}
template <bool COH> __device__ __forceinline__ BodyPose load_pose_keep_c(const BodyPose* p, unsigned long long pol) {
    BodyPose r;
    if (COH) { r.pos = ld4_keep_l2(&p->pos, pol); r.rot = ld4_keep_l2(&p->rot, pol); }
    else { r.pos = ld4_keep(&p->pos, pol); r.rot = ld4_keep(&p->rot, pol); }
    return r;
}

// ------------------------------------------------------------------ row math of the solver kernels
// The large-world visit kernel is instruction-issue bound (ncu: ~45 % issue slots busy, DRAM at a third of peak), so the device
// copy of the row math is written for instruction count.  Same formulas as avbd_rows.cuh (which the host mirror and the host
// emulation build use, and which the parity tests compare this against), with:
//   * the contact seen from the VISITING body: geometry arrives as {r_self, r_other, n} (visit_geometry swaps once per step; the cluster loop swaps per visit) and
//     the A/B orientation is one sign, sg — (pA + wA) - (pB + wB) = sg ((pS + wS) - (pO + wO)) exactly, so no operand selects;
//   * clamps as fminf / fmaxf (one FMNMX each; a NaN operand yields the bound where the reference's ternary clamp passes the
//     NaN on — NaN states are scrubbed by the pose update either way);
//   * the dual's |w x b|^2 as |w|^2 - (w . b)^2 for the unit basis vectors b (|b|^2 taken as 1), one division per active row;
//   * approximate (2 ulp) reciprocal / square root / reciprocal square root instructions, IN THE ROWS ONLY: the block solve and the
//     pose update keep IEEE division and sqrt — building the whole translation unit with -prec-div=false -prec-sqrt=false
//     moved the Stack's rest heights by 2e-3 and toppled the Pyramid's apex box (tools/rest_probe.py).
#ifndef AVBD_SYSTEM_FORM
#define AVBD_SYSTEM_FORM 1
#endif
__device__ __forceinline__ float clampq(float x, float lo, float hi) { return fmaxf(lo, fminf(hi, x)); }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// contact_basis_unit of avbd_math.cuh (manifold.cpp:39-50 on an already-unit normal)
__device__ __forceinline__ void basis_fast(V3 n, V3& t1, V3& t2) {
    bool useX = fabsf(n.x) >= fabsf(n.z);
    float a = useX ? -n.y : -n.z, b = useX ? n.x : n.y;
    float l2 = a * a + b * b;
    float inv = rsqrtf(l2);
    t1 = useX ? mk3(a * inv, b * inv, 0.0f) : mk3(0.0f, a * inv, b * inv);
    if (l2 < kVecEps) t1 = mk3(1.0f, 0.0f, 0.0f);
    t2 = cross(n, t1);
}

// Manifold::computeConstraint, second half (manifold.cpp:198-241) — contact_limits of avbd_rows.cuh.
__device__ __forceinline__ void limits_fast(float cap, float mu0, float bias, const float (&sep)[3], ContactState& c, ContactEval& e) {
    e.C[0] = sep[0] + bias * c.C0n;
    e.C[1] = sep[1] + bias * c.C0t1;
    e.C[2] = sep[2] + bias * c.C0t2;
    e.fmin[0] = -cap; e.fmax[0] = 0.0f;
    float warmN = fabsf(fminf(c.lam[0], 0.0f));
    float trialN = fabsf(fminf(c.pen[0] * e.C[0] + c.lam[0], 0.0f));
    float nmag = fminf(fmaxf(warmN, trialN), cap);
    float lim = (c.stick ? mu0 : mu0 * kKineticFrictionScale) * nmag;
    float t2 = c.lam[1] * c.lam[1] + c.lam[2] * c.lam[2];
    float tm = sqrt_fast(t2);
    if (tm > lim && tm > 1.0e-8f) { float s = __fdividef(lim, tm); c.lam[1] *= s; c.lam[2] *= s; }
    e.fmin[1] = -lim; e.fmax[1] = lim; e.fmin[2] = -lim; e.fmax[2] = lim;
    float slip2 = e.C[1] * e.C[1] + e.C[2] * e.C[2];
    float tl2 = c.lam[1] * c.lam[1] + c.lam[2] * c.lam[2];
    c.stick = (slip2 <= kStickThresh * kStickThresh) && (tl2 <= lim * lim + 1.0e-8f);
}
// Dual + penalty ramp (solver.cpp:411-430, rowPenaltyGain :94-125) — dual_contact of avbd_rows.cuh.
__device__ __forceinline__ void dual_fast(ContactState& c, const ContactEval& e, float beta) {
    float w2 = len2(e.wrA) + len2(e.wrB);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float lu = clampq(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        bool active = lu > e.fmin[r] && lu < e.fmax[r];
        c.lam[r] = lu;
        float da = dot(e.wrA, e.basis[r]), db = dot(e.wrB, e.basis[r]);
        float aw = fmaxf(w2 - (da * da + db * db), 0.0f);
        float br = beta * __fdividef(2.0f + kAngularBetaScale * aw, 2.0f + aw);
        float grown = fminf(c.pen[r] + br * fabsf(e.C[r]), kManifoldPenaltyCap);
        c.pen[r] = active ? grown : c.pen[r];
    }
}
// contact_system_w of avbd_rows.cuh without the "does the row add stiffness" test (solver.cpp:381): a manifold row's penalty is
// always in [PENALTY_MIN, MANIFOLD_PENALTY_CAP] on the device (set at creation, clamped by the warm-start decay and by the ramp,
// whose fminf maps a NaN to the cap).
// The 27 products per row are issued as packed pairs (Blackwell's fma.rn.f32x2 / mul.rn.f32x2: two IEEE FP32 operations per
// issue slot, each lane rounded exactly like the scalar instruction) in the order of the shared-memory row the visit kernel
// stores: rl0 rl1 | rl2 ra0 | ra1 ra2 | ll0 ll1 | ll2 ll3 | ll4 ll5 | la0 la1 | la2 la3 | la4 la5 | la6 la7 | la8 aa0 | aa1 aa2 | aa3 aa4 | aa5 -.
__device__ __forceinline__ void system_pairs(float2 (&v)[14], const ContactState& c, const ContactEval& e, V3 w, float sg, bool gyro, const M3& invIw) {
#if AVBD_SYSTEM_FORM == 0
    // The contact's three rows share one lever arm w and an orthonormal basis b_r, J_r = [b_r, w x b_r], so their sum factors:
    //   S = sum_r pen_r b_r b_r^T (6 unique),  F = sum_r f_r b_r
    //   rhs = [F, w x F],   LL = S,   LA = S [w]x^T  (row i = w x S_i),   AA = [w]x LA  (column k = w x LA_k)
    // 72 multiply-adds against 3 x (6 + 27) for the row-by-row form, and no operand shuffling into register pairs.  Measured (build with
    // SOLVE_DEFS=-DAVBD_SYSTEM_FORM=0): Stress1000 1405 -> 1455 steps/s, 8192-world ensemble solve 2.34 -> 2.26 ms, 1M-box grid solve
    // 4.92 -> 4.84 ms.  NOT the default: AA comes out as differences of products of magnitude pen |w|^2 instead of a sum of squares, and
    // with penalties near the cap that rounding noise is of the order of the inertia term I / dt^2 — the Pyramid's apex box (balanced on
    // two supports) topples within 600 steps with this form and stays with the row-by-row one, as it does in the reference.
    float f0[3];
    float Sxx, Syx, Szx, Syy, Szy, Szz; V3 F;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const V3 b = e.basis[r];
        f0[r] = clampq(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        const float f = f0[r] * sg, pen = c.pen[r];
        const float px = pen * b.x, py = pen * b.y, pz = pen * b.z;
        if (r == 0) { Sxx = px * b.x; Syx = py * b.x; Szx = pz * b.x; Syy = py * b.y; Szy = pz * b.y; Szz = pz * b.z; F = b * f; }
        else {
            Sxx = fmaf(px, b.x, Sxx); Syx = fmaf(py, b.x, Syx); Szx = fmaf(pz, b.x, Szx); Syy = fmaf(py, b.y, Syy); Szy = fmaf(pz, b.y, Szy); Szz = fmaf(pz, b.z, Szz);
            F.x = fmaf(f, b.x, F.x); F.y = fmaf(f, b.y, F.y); F.z = fmaf(f, b.z, F.z);
        }
    }
    const V3 Ra = cross(w, F);
    // la[3 i + j] = (w x S_i)_j with S_i the i-th row of the symmetric S
    const V3 l0 = cross(w, mk3(Sxx, Syx, Szx)), l1 = cross(w, mk3(Syx, Syy, Szy)), l2 = cross(w, mk3(Szx, Szy, Szz));
    // aa = lower triangle of [w]x LA, column k of LA = (l0[k], l1[k], l2[k])
    float aa0 = w.y * l2.x - w.z * l1.x, aa1 = w.z * l0.x - w.x * l2.x, aa2 = w.x * l1.x - w.y * l0.x;
    float aa3 = w.z * l0.y - w.x * l2.y, aa4 = w.x * l1.y - w.y * l0.y;
    float aa5 = w.x * l1.z - w.y * l0.z;
    if (gyro) {                                                  // solver.cpp:393-397; exactly zero for isotropic inertia
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const V3 Ja = cross(w, e.basis[r]);
            const V3 g = vabs(cross(Ja, mv(invIw, Ja)));
            const float af = fabsf(f0[r]);
            aa0 += g.x * af; aa3 += g.y * af; aa5 += g.z * af;
        }
    }
    v[0] = make_float2(F.x, F.y); v[1] = make_float2(F.z, Ra.x); v[2] = make_float2(Ra.y, Ra.z);
    v[3] = make_float2(Sxx, Syx); v[4] = make_float2(Szx, Syy); v[5] = make_float2(Szy, Szz);
    v[6] = make_float2(l0.x, l0.y); v[7] = make_float2(l0.z, l1.x); v[8] = make_float2(l1.y, l1.z); v[9] = make_float2(l2.x, l2.y);
    v[10] = make_float2(l2.z, aa0); v[11] = make_float2(aa1, aa2); v[12] = make_float2(aa3, aa4); v[13] = make_float2(aa5, 0.0f);
#else
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V3 Jl = e.basis[r];
        V3 Ja = cross(w, e.basis[r]);
        float f0 = clampq(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        float f = f0 * sg;
        float pen = c.pen[r];
        float2 pp = make_float2(pen, pen), ff = make_float2(f, f);
        float2 lxy = __fmul2_rn(make_float2(Jl.x, Jl.y), pp), lzq = __fmul2_rn(make_float2(Jl.z, Ja.x), pp), qbc = __fmul2_rn(make_float2(Ja.y, Ja.z), pp);
        const float px = lxy.x, py = lxy.y, pz = lzq.x, qa = lzq.y, qb = qbc.x, qc = qbc.y;       // pen * Jl, pen * Ja
        const float2 A[14] = {make_float2(Jl.x, Jl.y), make_float2(Jl.z, Ja.x), make_float2(Ja.y, Ja.z),
                              make_float2(px, py), make_float2(pz, py), make_float2(pz, pz),
                              make_float2(px, px), make_float2(px, py), make_float2(py, py), make_float2(pz, pz),
                              make_float2(pz, qa), make_float2(qb, qc), make_float2(qb, qc), make_float2(qc, 0.0f)};
        const float2 B[14] = {ff, ff, ff,
                              make_float2(Jl.x, Jl.x), make_float2(Jl.x, Jl.y), make_float2(Jl.y, Jl.z),
                              make_float2(Ja.x, Ja.y), make_float2(Ja.z, Ja.x), make_float2(Ja.y, Ja.z), make_float2(Ja.x, Ja.y),
                              make_float2(Ja.z, Ja.x), make_float2(Ja.x, Ja.x), make_float2(Ja.y, Ja.y), make_float2(Ja.z, 0.0f)};
#pragma unroll
        for (int q = 0; q < 14; ++q) v[q] = r == 0 ? __fmul2_rn(A[q], B[q]) : __ffma2_rn(A[q], B[q], v[q]);
        if (gyro) {                                              // solver.cpp:393-397; exactly zero for isotropic inertia
            V3 g = vabs(cross(Ja, mv(invIw, Ja)));
            float af = fabsf(f0);
            v[10].y += g.x * af; v[12].x += g.y * af; v[13].x += g.z * af;
        }
    }
#endif
}
__device__ __forceinline__ void system_fast(BodySystem& s, const ContactState& c, const ContactEval& e, V3 w, float sg, bool gyro, const M3& invIw) {
    float2 v[14];
    system_pairs(v, c, e, w, sg, gyro, invIw);
    s.rl[0] = v[0].x; s.rl[1] = v[0].y; s.rl[2] = v[1].x; s.ra[0] = v[1].y; s.ra[1] = v[2].x; s.ra[2] = v[2].y;
    s.ll[0] = v[3].x; s.ll[1] = v[3].y; s.ll[2] = v[4].x; s.ll[3] = v[4].y; s.ll[4] = v[5].x; s.ll[5] = v[5].y;
    s.la[0] = v[6].x; s.la[1] = v[6].y; s.la[2] = v[7].x; s.la[3] = v[7].y; s.la[4] = v[8].x; s.la[5] = v[8].y; s.la[6] = v[9].x; s.la[7] = v[9].y;
    s.la[8] = v[10].x; s.aa[0] = v[10].y; s.aa[1] = v[11].x; s.aa[2] = v[11].y; s.aa[3] = v[12].x; s.aa[4] = v[12].y; s.aa[5] = v[13].x;
}
// One contact visit in the visiting body's frame: computeConstraint (with the pending dual update first), then the 3 rows'
// contribution to the body's 6x6 system.  `sp,sq` / `op,oq` = self / other pose, g0 g1 g2 = {r_self,C0n} {r_other,C0t.x} {n,C0t.y}.
// Every solver kernel (sweep kernel, cluster loop, dual pass) evaluates rows through these, so the deferred and the
// stand-alone dual agree to FMA-contraction rounding: `stick` is decided by comparing a just-clamped |lambda_t|^2 with lim^2, i.e.
// by rounding, and a flipped `stick` changes friction by 10 % — two row-math variants would drift apart within a few steps.
__device__ __forceinline__ float rows_geometry(float4 sp, float4 sq, float4 op, float4 oq, float sg, const ContactState& cs, ContactEval& ev, float (&sep)[3]) {
    ev.basis[0] = cs.n;
    basis_fast(cs.n, ev.basis[1], ev.basis[2]);
    ev.wrA = qrot(quat(sq), cs.rA);                 // self
    ev.wrB = qrot(quat(oq), cs.rB);                 // other
    V3 d = (xyz(sp) + ev.wrA) - (xyz(op) + ev.wrB);
    sep[0] = sg * dot(d, ev.basis[0]) - kNormalContactMargin; sep[1] = sg * dot(d, ev.basis[1]); sep[2] = sg * dot(d, ev.basis[2]);
    float ims = sp.w + op.w;
    return kNormalForceCap * ((ims > 1.0e-6f) ? __fdividef(1.0f, ims) : 1.0f);       // normal force cap (manifold.cpp:199-203)
}
__device__ __forceinline__ void visit_rows(float4 sp, float4 sq, float4 op, float4 oq, float sg, float mu, float alpha, bool pending, float biasDual,
                                           float beta, bool gyro, const M3& invIw, ContactState& cs, BodySystem& sys) {
    ContactEval ev; float sep[3];
    float cap = rows_geometry(sp, sq, op, oq, sg, cs, ev, sep);
    if (pending) {
        limits_fast(cap, mu, biasDual, sep, cs, ev);
        dual_fast(cs, ev, beta);
    }
    limits_fast(cap, mu, fminf(fmaxf(1.0f - alpha, 0.0f), 1.0f), sep, cs, ev);
    system_fast(sys, cs, ev, ev.wrA, sg, gyro, invIw);
}

// The same visit, leaving the 27 numbers as the 14 packed pairs system_pairs produces (order of FlatRow / the shared-memory rows).
__device__ __forceinline__ void visit_rows_pairs(float4 sp, float4 sq, float4 op, float4 oq, float sg, float mu, float alpha, bool pending, float biasDual,
                                                 float beta, bool gyro, const M3& invIw, ContactState& cs, float2 (&v)[14]) {
    ContactEval ev; float sep[3];
    float cap = rows_geometry(sp, sq, op, oq, sg, cs, ev, sep);
    if (pending) {
        limits_fast(cap, mu, biasDual, sep, cs, ev);
        dual_fast(cs, ev, beta);
    }
    limits_fast(cap, mu, fminf(fmaxf(1.0f - alpha, 0.0f), 1.0f), sep, cs, ev);
    system_pairs(v, cs, ev, ev.wrA, sg, gyro, invIw);
}

// Loads of data another CTA may have written earlier in the SAME launch (persistent loop): bypass L1.
template <bool COH> __device__ __forceinline__ float4 ld4(const float4* p) { return COH ? __ldcg(p) : *p; }
template <bool COH> __device__ __forceinline__ BodyPose load_pose(const BodyPose* p) {
    BodyPose r; r.pos = ld4<COH>(&p->pos); r.rot = ld4<COH>(&p->rot); return r;
}
template <bool COH> __device__ __forceinline__ ContactState load_contact_c(const ManifoldSet& ms, int ci) {
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], ld4<COH>(&ms.lp[ci].l), ld4<COH>(&ms.lp[ci].p));
}

// ------------------------------------------------------------------ primal, one tile of bodies (the cluster loop's phase)
// The visit list is laid out in colour order, so the visits of a tile of BPB consecutive bodies of one colour are ONE
// contiguous run.  The tile walks that run one visit per thread (every lane busy, every lane's gathers independent
// and in flight together) and solves its bodies in the same call — small worlds are latency bound, a phase must not
// be split over launches.  (Large worlds use the warp-private visit pipeline further down.)
//   phase 0  thread t < BPB stages body t of the tile in shared memory (pose, inertial target, mass, inverse inertia)
//   phase 1  thread t takes visit base+t: computeConstraint + 3 rows -> its 27 partial sums, parked transposed in
//            shared memory (row stride 257: conflict free)
//   phase 2  L = 256/BPB lanes per body add the body's run of partial sums in visit order (deterministic), each lane
//            owning ceil(27/L) of the 27 components
//   phase 3  thread t < BPB: inertial terms, Schur 3x3 LDL^T, pose update (solver.cpp:402-408)
// Runs longer than 256 visits loop over phases 1-2.
template <int BPB>
struct PrimalSmem {
    float c[27][kThreads + 1];
    float sys[BPB][27];
    float4 pos[BPB], rot[BPB], posI[BPB], rotI[BPB];
    float4 mass[BPB];              // mass, invMass, friction, radius
    float4 inert[BPB];             // Ixx Iyy Izz, w = 1 when the inertia is anisotropic (gyroscopic row term is non-zero)
    float inv[BPB][6];             // world inverse inertia (0,0) (1,0) (2,0) (1,1) (2,1) (2,2), only when anisotropic
    int vs[BPB + 1];
    int body[BPB];
};

// What a tile reads that does not change during a step's iteration loop: its bodies, their visit runs and inertial targets,
// the visit entries and the contact geometry.  The cluster loop keeps it in shared memory across the whole loop: a cluster
// barrier carries a gpu-scope fence and an L1 invalidation (CCTL.IVALL), so without the copy every one of the ~100 phases of a
// step would walk the chain visitStart -> visit entry -> geometry through L2 again (two dependent round trips per phase).
template <int BPB>
struct TileCache {
    int4 visit[kThreads];
    float4 a[kThreads], b[kThreads], n[kThreads];      // cA cB cN of the visit's contact (A / B frame)
    float4 posI[BPB], rotI[BPB], mass[BPB], inert[BPB];
    int vs[BPB + 1]; int body[BPB];
    int usable, pad[2];                                 // 0: the tile's run does not fit (more than kThreads visits) -> read from global
};
template <int BPB>
__device__ __forceinline__ void fill_tile_cache(TileCache<BPB>& tc, const BodyView& b, const int* __restrict__ vstart, const int4* __restrict__ visits,
                                                const ManifoldSet& ms, const int* __restrict__ order, int count, int tile) {
    const int t = threadIdx.x;
    const int k0 = tile * BPB;
    const int nb = (count - k0) < BPB ? (count - k0) : BPB;
    if (t <= nb) tc.vs[t] = vstart[k0 + t];
    __syncthreads();
    const int v0 = tc.vs[0], v1 = tc.vs[nb];
    const bool fits = v1 - v0 <= kThreads;
    if (t == 0) tc.usable = fits ? 1 : 0;
    if (fits) {
        if (t < nb) {
            int i = order[k0 + t];
            tc.body[t] = i;
            BodyAux aux = b.aux[i];
            tc.posI[t] = aux.posI; tc.rotI[t] = aux.rotI; tc.mass[t] = aux.mass; tc.inert[t] = aux.inert;
        }
        if (v0 + t < v1) {
            int4 e = visits[v0 + t];
            tc.visit[t] = e; tc.a[t] = ms.cA[e.x]; tc.b[t] = ms.cB[e.x]; tc.n[t] = ms.cN[e.x];
        }
    }
    __syncthreads();
}

template <int BPB, bool COH>
__device__ __forceinline__ void primal_tile_visits(const BodyView& b, const int* __restrict__ vstart, const int4* __restrict__ visits,
                                                   const ManifoldSet& ms, const ForceView& fv, const int* __restrict__ order, int count, int tile,
                                                   const SolveParams& prm, float alpha, float biasDual, float* dxOut, Diag* diag, PrimalSmem<BPB>& sm,
                                                   const TileCache<BPB>* tc = nullptr) {
    constexpr int L = kThreads / BPB;
    constexpr int CPL = (27 + L - 1) / L;
    const int t = threadIdx.x;
    const int k0 = tile * BPB;
    const int nb = (count - k0) < BPB ? (count - k0) : BPB;
    if (t <= nb) sm.vs[t] = tc ? tc->vs[t] : vstart[k0 + t];
    if (t < nb) {
        int i = tc ? tc->body[t] : order[k0 + t];
        sm.body[t] = i;
        BodyPose self = load_pose<COH>(b.pose + i);
        BodyAux aux;
        if (tc) { aux.posI = tc->posI[t]; aux.rotI = tc->rotI[t]; aux.mass = tc->mass[t]; aux.inert = tc->inert[t]; }
        else aux = b.aux[i];
        sm.pos[t] = self.pos; sm.rot[t] = self.rot; sm.posI[t] = aux.posI; sm.rotI[t] = aux.rotI; sm.mass[t] = aux.mass;
        V3 I = xyz(aux.inert);
        // isotropic inertia: R diag(c) R^T = c*Id, so Ja x (I^-1 Ja) of solver.cpp:393-397 is exactly zero; skip the term
        bool aniso = !(I.x == I.y && I.y == I.z);
        sm.inert[t] = make_float4(I.x, I.y, I.z, aniso ? 1.0f : 0.0f);
        if (aniso) {
            M3 inv = rot_diag(qmat(quat(self.rot)), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
            sm.inv[t][0] = inv.c[0].x; sm.inv[t][1] = inv.c[0].y; sm.inv[t][2] = inv.c[0].z;
            sm.inv[t][3] = inv.c[1].y; sm.inv[t][4] = inv.c[1].z; sm.inv[t][5] = inv.c[2].z;
        }
    }
    __syncthreads();
    const int v0 = sm.vs[0], v1 = sm.vs[nb];
    float acc[CPL];
#pragma unroll
    for (int m = 0; m < CPL; ++m) acc[m] = 0.0f;
    const int rs = t / L, rj = t % L;
    int rlo = 0, rhi = 0;
    if (rs < nb) { rlo = sm.vs[rs]; rhi = sm.vs[rs + 1]; }
    for (int base = v0; base < v1; base += kThreads) {
        int v = base + t;
        if (v < v1) {
            int4 e = tc ? tc->visit[v - v0] : visits[v];
            int ci = e.x; bool isA = (e.z & 1) != 0;
            BodyPose po = load_pose<COH>(b.pose + e.y);
            float4 l4 = ld4<COH>(&ms.lp[ci].l), p4 = ld4<COH>(&ms.lp[ci].p);
            float4 a4, b4, n4;
            if (tc) { a4 = tc->a[v - v0]; b4 = tc->b[v - v0]; n4 = tc->n[v - v0]; }
            else { a4 = ms.cA[ci]; b4 = ms.cB[ci]; n4 = ms.cN[ci]; }
            int lo = 0, hi = nb;                                  // slot: vs[lo] <= v < vs[lo + 1]
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sm.vs[mid] <= v) lo = mid; else hi = mid; }
            float4 sp = sm.pos[lo], sr = sm.rot[lo];
            bool pending = biasDual >= 0.0f && (e.z & 4) != 0;
            ContactState cs = unpack_contact(isA ? a4 : b4, isA ? b4 : a4, n4, l4, p4);     // rA = r_self, rB = r_other (visit_rows' frame)
            cs.C0n = a4.w; cs.C0t1 = b4.w;
            bool gyro = sm.inert[lo].w != 0.0f;
            M3 invIw;
            if (gyro) {
                const float* q = sm.inv[lo];
                invIw = m3(mk3(q[0], q[1], q[2]), mk3(q[1], q[3], q[4]), mk3(q[2], q[4], q[5]));
            } else {
                invIw = m3(zero3(), zero3(), zero3());
            }
            BodySystem sys;
            visit_rows(sp, sr, po.pos, po.rot, isA ? 1.0f : -1.0f, __int_as_float(e.w), alpha, pending, biasDual, prm.beta, gyro, invIw, cs, sys);
            // computeConstraint's side effects (manifold.cpp:224-241): written only when they changed something
            float4 nl = pack_lambda(cs);
            if (pending) { ContactLP q; q.l = nl; q.p = pack_penalty(cs); ms.lp[ci] = q; }
            else if (nl.y != l4.y || nl.z != l4.z || nl.w != l4.w) ms.lp[ci].l = nl;
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.c[k][t] = sys.rl[k]; sm.c[3 + k][t] = sys.ra[k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) { sm.c[6 + k][t] = sys.ll[k]; sm.c[21 + k][t] = sys.aa[k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) sm.c[12 + k][t] = sys.la[k];
        }
        __syncthreads();
        if (rs < nb) {
            int a = (rlo > base ? rlo : base) - base;
            int z = (rhi < base + kThreads ? rhi : base + kThreads) - base;
            for (int lv = a; lv < z; ++lv) {
#pragma unroll
                for (int m = 0; m < CPL; ++m) { int k = rj + m * L; if (k < 27) acc[m] += sm.c[k][lv]; }
            }
        }
        __syncthreads();
    }
    if (rs < nb) {
#pragma unroll
        for (int m = 0; m < CPL; ++m) { int k = rj + m * L; if (k < 27) sm.sys[rs][k] = acc[m]; }
    }
    __syncthreads();
    if (t < nb) {
        int i = sm.body[t];
        float4 sp = sm.pos[t];
        V3 pos = xyz(sp); Q4 rot = quat(sm.rot[t]);
        BodyAux aux; aux.posI = sm.posI[t]; aux.rotI = sm.rotI[t]; aux.mass = sm.mass[t]; aux.inert = sm.inert[t];
        BodySystem own; M3 invIw;
        body_self_system(pos, rot, aux, prm.dt, own, invIw);
        const float* o = sm.sys[t];
#pragma unroll
        for (int k = 0; k < 3; ++k) { own.rl[k] += o[k]; own.ra[k] += o[3 + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { own.ll[k] += o[6 + k]; own.aa[k] += o[21 + k]; }
#pragma unroll
        for (int k = 0; k < 9; ++k) own.la[k] += o[12 + k];
        if (fv.adjStart != nullptr && fv.adjStart[i + 1] > fv.adjStart[i]) accumulate_user_forces(own, fv, b.pose, i, pos, rot, invIw);
        V3 dl, da;
        solve_body_system(own, dl, da);
        int evn = apply_body_update(pos, rot, dl, da);
        BodyPose out; out.pos = f4(pos, sp.w); out.rot = f4(rot);
        b.pose[i] = out;
        if (dxOut) { float* d = dxOut + 6 * i; d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = da.x; d[4] = da.y; d[5] = da.z; }
        if (evn) atomicAdd(&diag[b.worldId[i]].nanEvents, evn);
    }
}

// ------------------------------------------------------------------ primal, warp-private visit pipeline (the large-world path)
// A colour's contact visits are one contiguous run of the colour-ordered visit list (graph stage).  The run is cut into one
// contiguous, BODY-ALIGNED range per warp (warp_ranges): a body's visits never straddle two warps, so nothing in this kernel is
// shared between warps — no block barrier, no atomics — and a body's 27 row sums are ONE sequence of additions in visit order
// (exactly the cluster loop's, whatever else is in the batch: ensembles are bit-identical however they are partitioned).
// Each warp runs its own software pipeline over chunks of 32 visits, one visit per lane.  The kernel is bound by the SM's load /
// store data pipe (ncu: l1tex data-pipe wavefronts at 68 % of peak with every operand staged through shared memory by cp.async —
// a gathering LDGSTS costs ~25-30 wavefronts, one per lane, however few sectors it touches), so operands take the cheapest road:
//   entries     two chunks ahead, in registers (streamed, 16 B per visit)
//   geometry    one chunk ahead, three coalesced 16-byte loads per lane straight into registers (visit-order copy, evict-first)
//   lambda /    one chunk ahead into registers; the two 16-byte halves of a contact's 32-byte record are fetched by a PAIR of lanes
//   penalty     in one instruction (16 whole sectors per instruction instead of 32 half-used ones) and swapped with one shuffle each
//   self pose   one chunk ahead, cp.async by the segment heads only (a body's visits are consecutive lanes: one fetch per body, not
//               per visit) into a per-segment slot every lane of the segment then reads (broadcast)
//   other pose  one chunk ahead, cp.async into the lane's own slot (L2 evict-last: poses are the data every sweep re-reads)
//   phase 1     computeConstraint (+ the pending dual update on a contact's first visit) + 3 rows -> 27 partial sums, one 112-byte
//               shared-memory row per lane (7 x STS.128, stride 28 words: conflict free)
//   phase 2     8 lanes per segment (= consecutive visits of one body, found with one ballot; 7 of the 8 carry a float4 column, so a
//               quarter warp reads one row: conflict free) add the segment's rows in visit order, four segments per pass, the pass's
//               trip count uniform over the warp; a segment that runs on into the next chunk leaves its partial sum in the carry slot
//   solve       finished bodies queue up in shared memory (sums + body index) and are solved in batches, one body per lane:
//               inertial terms, Schur 3x3 LDL^T, pose update (solver.cpp:351-369, :402-408) — no per-body sums round trip through
//               HBM, no second kernel, and the serial 6x6 solve runs with most lanes busy.
#ifndef AVBD_VG_REG
#define AVBD_VG_REG 0        // visit-order geometry: 1 = coalesced loads into registers, 0 = cp.async into the stage
#endif
#ifndef AVBD_VG_BULK
#define AVBD_VG_BULK 0       // visit-order geometry (when staged): 1 = three bulk copies per chunk by one lane (TMA + mbarrier; measured 4 % SLOWER:
                             // 4.39 vs 4.23 ms of sweeps on the 1M grid — the wait + proxy fence cost more than the wavefronts saved), 0 = cp.async per lane
#endif
#ifndef AVBD_LP_REG
#define AVBD_LP_REG 0        // lambda / penalty: 1 = lane-pair loads into registers + shuffle, 0 = cp.async into the stage
#endif
#ifndef AVBD_SELF_SEG
#define AVBD_SELF_SEG 1      // visiting-body pose: 1 = one cp.async per segment (broadcast read), 0 = one per lane
#endif
#ifndef AVBD_QUEUE_SLOTS
#define AVBD_QUEUE_SLOTS 26
#endif
constexpr int kQueueSlots = AVBD_QUEUE_SLOTS;
constexpr bool kVgReg = AVBD_VG_REG != 0, kLpReg = AVBD_LP_REG != 0, kSelfSeg = AVBD_SELF_SEG != 0, kVgBulk = AVBD_VG_BULK != 0 && !kVgReg;
struct WarpPipe {
#ifdef AVBD_TIMELINE
    int tlRow;
#endif
    float4 rows[32][7];              // this chunk's partial sums, one row of 28 floats per visit: rl(3) ra(3) ll(6) la(9) aa(6) pad
    float4 other[2][32];             // the NEXT chunk's other-body poses, per lane (cp.async)
    float4 selfp[2][32];             // the NEXT chunk's visiting-body poses, per segment (cp.async by the segment's first lane) or per lane
    float4 geom[kVgReg ? 1 : 3][32]; // geometry / lambda, penalty when they are staged rather than loaded into registers
    float4 lamp[kLpReg ? 1 : 2][32];
    float4 queue[kQueueSlots][7];    // row sums of finished bodies waiting for the batched solve
    float4 carry[7];                 // partial sum of the body whose run crosses into the next chunk
    int    qBody[kQueueSlots];
    unsigned char segStart[32];      // lane of each segment's first visit
    unsigned long long mbar;         // completion barrier of the chunk's bulk copies (kVgBulk)
};
constexpr int kSweepWarps = 4;                     // warps per block (they share nothing but the block's shared-memory allocation)

__device__ __forceinline__ void stage16(float4* dst, const float4* src, unsigned long long policy) {       // cp.async (LDGSTS), L1-allocating
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "l"(policy) : "memory");
}
template <bool COH> __device__ __forceinline__ void stage_pose(float4* dstPos, float4* dstRot, const BodyPose* src, unsigned long long policy);
__device__ __forceinline__ void stage16_nol1(float4* dst, const float4* src, unsigned long long policy) {  // bypasses L1 (single-use data)
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_stream_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Bulk asynchronous copy global -> shared (the TMA engine: no per-lane address math, no load / store pipe wavefronts) completing on
// an mbarrier.  `bytes` a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar), done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy) : "memory");
}
template <bool COH> __device__ __forceinline__ void stage_pose(float4* dstPos, float4* dstRot, const BodyPose* src, unsigned long long policy) {
    if (COH) { stage16_nol1(dstPos, &src->pos, policy); stage16_nol1(dstRot, &src->rot, policy); }
    else { stage16(dstPos, &src->pos, policy); stage16(dstRot, &src->rot, policy); }
}
__device__ __forceinline__ float4 ld4_keep_cg(const float4* ptr, unsigned long long pol) {                // single use in this launch: L2 only
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
    return v;
}

// A body no contact visits: inertial terms (plus user forces) -> 6x6 solve -> pose update (solver.cpp:351-369, :402-408).
template <bool COH>
__device__ __forceinline__ void solve_free_body(const BodyView& b, const ForceView& fv, int i, const SolveParams& prm, float* dxOut, Diag* diag,
                                                unsigned long long keep) {
    BodyPose self = load_pose_keep_c<COH>(b.pose + i, keep);
    BodyAux aux = b.aux[i];
    V3 pos = xyz(self.pos); Q4 rot = quat(self.rot);
    BodySystem own; M3 invIw;
    body_self_system(pos, rot, aux, prm.dt, own, invIw);
    if (fv.adjStart != nullptr && fv.adjStart[i + 1] > fv.adjStart[i]) accumulate_user_forces(own, fv, b.pose, i, pos, rot, invIw);
    V3 dl, da;
    solve_body_system(own, dl, da);
    int evn = apply_body_update(pos, rot, dl, da);
    st4_keep(&b.pose[i].pos, f4(pos, self.pos.w), keep);
    st4_keep(&b.pose[i].rot, f4(rot), keep);
    if (dxOut) { float* d = dxOut + 6 * i; d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = da.x; d[4] = da.y; d[5] = da.z; }
    if (evn) atomicAdd(&diag[b.worldId[i]].nanEvents, evn);
}

// One body per lane (lanes [0, qn)): queued row sums + inertial terms -> 6x6 solve -> pose update.
template <bool COH>
__device__ __forceinline__ void solve_queue(WarpPipe& w, int qn, int lane, const BodyView& b, const ForceView& fv, const SolveParams& prm,
                                            float* dxOut, Diag* diag, unsigned long long keep) {
    if (lane < qn) {
        const int i = w.qBody[lane];
        BodyPose self = load_pose_keep_c<COH>(b.pose + i, keep);
        BodyAux aux = b.aux[i];
        float o[28];
#pragma unroll
        for (int q = 0; q < 7; ++q) { float4 x = w.queue[lane][q]; o[4 * q] = x.x; o[4 * q + 1] = x.y; o[4 * q + 2] = x.z; o[4 * q + 3] = x.w; }
        V3 pos = xyz(self.pos); Q4 rot = quat(self.rot);
        BodySystem own; M3 invIw;
        body_self_system(pos, rot, aux, prm.dt, own, invIw);
#pragma unroll
        for (int q = 0; q < 3; ++q) { own.rl[q] += o[q]; own.ra[q] += o[3 + q]; }
#pragma unroll
        for (int q = 0; q < 6; ++q) { own.ll[q] += o[6 + q]; own.aa[q] += o[21 + q]; }
#pragma unroll
        for (int q = 0; q < 9; ++q) own.la[q] += o[12 + q];
        if (fv.adjStart != nullptr && fv.adjStart[i + 1] > fv.adjStart[i]) accumulate_user_forces(own, fv, b.pose, i, pos, rot, invIw);
        V3 dl, da;
        solve_body_system(own, dl, da);
        int evn = apply_body_update(pos, rot, dl, da);
        st4_keep(&b.pose[i].pos, f4(pos, self.pos.w), keep);
        st4_keep(&b.pose[i].rot, f4(rot), keep);
        if (dxOut) { float* d = dxOut + 6 * i; d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = da.x; d[4] = da.y; d[5] = da.z; }
        if (evn) atomicAdd(&diag[b.worldId[i]].nanEvents, evn);
    }
    __syncwarp();
}

// Debug build only (make variant SOLVE_DEFS=-DAVBD_TIMELINE, tools/sweep_timeline.py): %globaltimer stamps of the first warp of every
// sweep launch — entry, before / after the wait for the predecessor, operands of the first chunk landed, rows of the first chunk
// stored, range done.  Shows what a colour phase of a small world is made of.
#ifdef AVBD_TIMELINE
__device__ unsigned long long g_timeline[16384][6];
__device__ unsigned g_timelineCount;
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define AVBD_TL(row, slot) do { if ((row) >= 0 && (threadIdx.x & 31) == 0) g_timeline[(row)][(slot)] = global_ns(); } while (0)
extern "C" int avbd_debug_timeline(unsigned long long* out, int capRows) {
    unsigned n = 0;
    if (cudaMemcpyFromSymbol(&n, g_timelineCount, sizeof(n)) != cudaSuccess) return -1;
    int rows = (int)(n < 16384u ? n : 16384u); if (rows > capRows) rows = capRows;
    if (rows > 0 && cudaMemcpyFromSymbol(out, g_timeline, (size_t)rows * 6 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    unsigned zero = 0; cudaMemcpyToSymbol(g_timelineCount, &zero, sizeof(zero));
    return rows;
}
#else
#define AVBD_TL(row, slot) do { } while (0)
#endif

// One warp's pipeline over the visits [vBegin, vEnd) (a body-aligned range).  COH: poses are read through L2 only (persistent loop:
// other SMs rewrote them since this SM's L1 last saw them); the per-colour launches let L1 keep them (L1 is flushed between launches).
// DEP: the launch waits for its predecessor (cudaGridDependencySynchronize) only after it has issued everything that does not depend
// on it — the first chunk's entries and its streamed geometry are static for the whole step.
template <bool COH, bool DEP>
__device__ __forceinline__ void sweep_range(WarpPipe& w, const int lane, const int vBegin, const int vEnd, const BodyView& b, const int4* __restrict__ visits,
                                            const VisitGeom& vg, const ManifoldSet& ms, const ForceView& fv, const SolveParams& prm,
                                            const float alpha, const float biasDual, float* __restrict__ dxOut, Diag* __restrict__ diag) {
    const unsigned long long keep = l2_keep_policy(), stream = l2_stream_policy();
    const int4 none = make_int4(0, 0, -8, 0);                                // body -1
    const bool odd = (lane & 1) != 0;
    unsigned bulkPhase = 0;
    if (kVgBulk) {
        if (lane == 0) { mbar_init(&w.mbar, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
        __syncwarp();
    }
    auto load_entry = [&](int v) { return v < vEnd ? __ldcs(visits + v) : none; };
    // Everything chunk [vb, vb + 32) needs, issued a whole chunk ahead.  `e` = the lane's entry of that chunk, `prevTail` = body of the
    // visit before the chunk (-1: none).  Returns the chunk's segment heads; nvA / nvB / nvN / nL0 / nL1 receive the register operands.
    float4 nvA, nvB, nvN, nL0, nL1;
    nvA = nvB = nvN = nL0 = nL1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    auto prefetch = [&](int vb, const int4& e, int prevTail, bool first) -> unsigned {
        const int v = vb + lane;
        const bool lv = v < vEnd;
        const int self = e.z >> 3;
        int prevSelf = __shfl_up_sync(0xffffffffu, self, 1);
        if (lane == 0) prevSelf = prevTail;
        const bool head = lv && (lane == 0 || prevSelf != self);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (kVgBulk) {
            int cnt = vEnd - vb; if (cnt > 32) cnt = 32;
            if (lane == 0 && cnt > 0) {                                      // every lane has read the previous contents (callers sync the warp first)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect(&w.mbar, 3u * 16u * (unsigned)cnt);
                bulk_copy(&w.geom[0][0], vg.a + vb, 16u * (unsigned)cnt, &w.mbar, stream);
                bulk_copy(&w.geom[kVgReg ? 0 : 1][0], vg.b + vb, 16u * (unsigned)cnt, &w.mbar, stream);
                bulk_copy(&w.geom[kVgReg ? 0 : 2][0], vg.n + vb, 16u * (unsigned)cnt, &w.mbar, stream);
            }
        } else if (!kVgReg && lv) { stage16_nol1(&w.geom[0][lane], vg.a + v, stream); stage16_nol1(&w.geom[kVgReg ? 0 : 1][lane], vg.b + v, stream); stage16_nol1(&w.geom[kVgReg ? 0 : 2][lane], vg.n + v, stream); }
        // what follows was written by the previous launch (poses, lambda / penalty): wait for it now, not before
#ifdef AVBD_TIMELINE
        if (first) AVBD_TL(w.tlRow, 1);
#endif
        if (DEP && first) cudaGridDependencySynchronize();
#ifdef AVBD_TIMELINE
        if (first) AVBD_TL(w.tlRow, 2);
#endif
        if (kSelfSeg) {
            if (head) {                                                      // one fetch of the visiting body's pose per segment
                const int seg = __popc(heads & ((1u << lane) - 1u));
                stage_pose<COH>(&w.selfp[0][seg], &w.selfp[1][seg], b.pose + self, keep);
            }
        } else if (lv) { stage_pose<COH>(&w.selfp[0][lane], &w.selfp[1][lane], b.pose + self, keep); }
        if (lv) { stage_pose<COH>(&w.other[0][lane], &w.other[1][lane], b.pose + e.y, keep); }
        if (!kLpReg && lv) { stage16_nol1(&w.lamp[0][lane], &ms.lp[e.x].l, keep); stage16_nol1(&w.lamp[kLpReg ? 0 : 1][lane], &ms.lp[e.x].p, keep); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (kVgReg && lv) { nvA = __ldcs(vg.a + v); nvB = __ldcs(vg.b + v); nvN = __ldcs(vg.n + v); }
        if (kLpReg) {
            // lambda / penalty: lanes 2k and 2k + 1 fetch the two halves of ONE contact's record per instruction
            const int cPar = __shfl_xor_sync(0xffffffffu, e.x, 1);
            const bool lvPar = __shfl_xor_sync(0xffffffffu, (int)lv, 1) != 0;
            const int c0 = odd ? cPar : e.x, c1 = odd ? e.x : cPar;          // instruction 0: the even lane's contact, 1: the odd lane's
            const bool ok0 = odd ? lvPar : lv, ok1 = odd ? lv : lvPar;
            if (ok0) nL0 = ld4_keep_cg(odd ? &ms.lp[c0].p : &ms.lp[c0].l, keep);
            if (ok1) nL1 = ld4_keep_cg(odd ? &ms.lp[c1].p : &ms.lp[c1].l, keep);
        }
        return heads;
    };
    int4 eCur = load_entry(vBegin + lane), eNext = load_entry(vBegin + 32 + lane);
    unsigned headsNext = prefetch(vBegin, eCur, -1, true);
    int qn = 0;                      // bodies waiting in the solve queue (warp-uniform)
    int tailSelf = -1;               // body of the previous chunk's last visit
    const int sub = lane >> 3, j = lane & 7;                                 // phase 2: segment of the pass / float4 column (7 = idle)
    for (int base = vBegin; base < vEnd; base += 32) {
        const int v = base + lane;
        const int4 e = eCur;
        const bool live = v < vEnd;
        const int self = e.z >> 3;
        const unsigned heads = headsNext;
        int liveCount = vEnd - base; if (liveCount > 32) liveCount = 32;
        const int lastSelf = __shfl_sync(0xffffffffu, self, liveCount - 1);
        const bool cont = tailSelf == __shfl_sync(0xffffffffu, self, 0);                  // the first segment continues the previous chunk's last
        // ---- this chunk's operands: registers filled a chunk ago (+ the pair swap), staged poses; then refill for the next chunk at once
        float4 a4 = nvA, b4 = nvB, n4 = nvN;
        float4 l4 = nL0, p4 = nL1;
        if (kLpReg) {
            const float4 give = odd ? nL0 : nL1;                             // the half fetched for the partner
            float4 got;
            got.x = __shfl_xor_sync(0xffffffffu, give.x, 1); got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
            got.z = __shfl_xor_sync(0xffffffffu, give.z, 1); got.w = __shfl_xor_sync(0xffffffffu, give.w, 1);
            l4 = odd ? got : nL0; p4 = odd ? nL1 : got;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (kSelfSeg) __syncwarp();                                          // the self poses were fetched by the segment heads
#ifdef AVBD_TIMELINE
        if (base == vBegin) AVBD_TL(w.tlRow, 3);
#endif
        const int seg = __popc(heads & ((2u << lane) - 1u)) - 1;
        const int sslot = kSelfSeg ? (seg < 0 ? 0 : seg) : lane;
        BodyPose ps, po;
        ps.pos = w.selfp[0][sslot]; ps.rot = w.selfp[1][sslot]; po.pos = w.other[0][lane]; po.rot = w.other[1][lane];
        if (kVgBulk) { mbar_wait(&w.mbar, bulkPhase); bulkPhase ^= 1u; }      // a chunk inside the loop always has visits: its copies were issued
        if (!kVgReg) { a4 = w.geom[0][lane]; b4 = w.geom[kVgReg ? 0 : 1][lane]; n4 = w.geom[kVgReg ? 0 : 2][lane]; }
        if (!kLpReg) { l4 = w.lamp[0][lane]; p4 = w.lamp[kLpReg ? 0 : 1][lane]; }
        if (live && (heads >> lane & 1u)) w.segStart[seg] = (unsigned char)lane;
        __syncwarp();                                                        // everyone holds its poses: the stage may be refilled
        eCur = eNext;
        headsNext = prefetch(base + 32, eCur, lastSelf, false);                     // next chunk (its entry was loaded a whole chunk ago)
        eNext = load_entry(v + 64);                                          // the entry after that
        // ---- phase 1
        if (live) {
            const bool gyro = (e.z & 2) != 0, pending = biasDual >= 0.0f && (e.z & 4) != 0;
            const float sg = (e.z & 1) ? 1.0f : -1.0f;                         // visiting body is A / B of the manifold
            ContactState cs = unpack_contact(a4, b4, n4, l4, p4);               // rA = r_self, rB = r_other here
            M3 invIw = m3(zero3(), zero3(), zero3());
            if (gyro) {   // anisotropic inertia only: for R diag(c) R^T = c Id the term Ja x (I^-1 Ja) of solver.cpp:393-397 is exactly zero
                V3 I = xyz(b.aux[self].inert);
                invIw = rot_diag(qmat(quat(ps.rot)), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
            }
            float2 vv[14];
            visit_rows_pairs(ps.pos, ps.rot, po.pos, po.rot, sg, __int_as_float(e.w), alpha, pending, biasDual, prm.beta, gyro, invIw, cs, vv);
            // computeConstraint's side effects (manifold.cpp:224-241): written only when they changed something
            ContactLP* lp = ms.lp + e.x;
            float4 nl = pack_lambda(cs);
            if (pending) { ContactLP q; q.l = nl; q.p = pack_penalty(cs); *lp = q; }
            else if (nl.y != l4.y || nl.z != l4.z || nl.w != l4.w) lp->l = nl;
            float4* row = w.rows[lane];
#pragma unroll
            for (int q = 0; q < 7; ++q) row[q] = make_float4(vv[2 * q].x, vv[2 * q].y, vv[2 * q + 1].x, vv[2 * q + 1].y);      // the pairs ARE the row's order
        }
        __syncwarp();
#ifdef AVBD_TIMELINE
        if (base == vBegin) AVBD_TL(w.tlRow, 4);
#endif
        // ---- phase 2: segment sums
        const int nSeg = __popc(heads);
        const int nextSelf0 = __shfl_sync(0xffffffffu, eCur.z >> 3, 0);           // body the next chunk opens with (-1 past the warp's range)
        const bool lastContinues = nextSelf0 == lastSelf;
        const int nDone = nSeg - (lastContinues ? 1 : 0);
        for (int s0 = 0; s0 < nSeg; s0 += 4) {
            // queue slots in use: qn + s0 (every segment before this pass is finished: only a chunk's LAST segment can run on).
            // No room for four more: solve what is queued first.
            if (qn + s0 + 4 > kQueueSlots) { __syncwarp(); solve_queue<COH>(w, qn + s0, lane, b, fv, prm, dxOut, diag, keep); qn = -s0; }
            const int sgi = s0 + sub;
            const bool active = j < 7 && sgi < nSeg;
            const int start = active ? (int)w.segStart[sgi] : 0;
            const int end = (active && sgi + 1 < nSeg) ? (int)w.segStart[sgi + 1] : liveCount;
            const int body = __shfl_sync(0xffffffffu, self, start);
            const int len = active ? end - start : 0;
            const int maxLen = __reduce_max_sync(0xffffffffu, len);
            float4 acc = (active && sgi == 0 && cont) ? w.carry[j] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float4* src = &w.rows[start][j < 7 ? j : 0];
            for (int k = 0; k < maxLen; ++k) {                                    // warp-uniform trip count, predicated adds: no divergence
                if (k < len) {
                    float4 x = src[k * 7];
                    float2 lo = __fadd2_rn(make_float2(acc.x, acc.y), make_float2(x.x, x.y)), hi = __fadd2_rn(make_float2(acc.z, acc.w), make_float2(x.z, x.w));
                    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
            }
            if (active) {
                if (sgi == nSeg - 1 && lastContinues) w.carry[j] = acc;          // the run goes on in the warp's next chunk
                else { w.queue[qn + sgi][j] = acc; if (j == 0) w.qBody[qn + sgi] = body; }
            }
        }
        qn += nDone;
        tailSelf = lastSelf;
        __syncwarp();               // rows consumed, carry / queue visible, before the next chunk overwrites the rows
    }
    if (qn > 0) solve_queue<COH>(w, qn, lane, b, fv, prm, dxOut, diag, keep);
#ifdef AVBD_TIMELINE
    AVBD_TL(w.tlRow, 5);
#endif
    if (kVgBulk) { __syncwarp(); if (lane == 0) mbar_inval(&w.mbar); __syncwarp(); }
}

template <int MINB>
__global__ void __launch_bounds__(32 * kSweepWarps, MINB) primal_sweep_warp(BodyView b, const int4* __restrict__ visits, VisitGeom vg, ManifoldSet ms, ForceView fv,
                                                                           const int* __restrict__ range, int nWarps, SolveParams prm,
                                                                           float alpha, float biasDual, float* __restrict__ dxOut, Diag* __restrict__ diag,
                                                                           const int* __restrict__ freeList, int nFree, int flags) {
    __shared__ WarpPipe pipes[kSweepWarps];
    // Let the NEXT launch become resident now and run its own static prologue (every kernel of this library waits at
    // cudaGridDependencySynchronize before it touches anything a predecessor writes): a small colour is a chain of latencies and its
    // grid leaves the machine empty.  Measured: Stress1000 0.746 -> 0.722 ms per step, 8192-world ensemble 3.50 -> 3.46, 1M grid
    // 7.61 -> 7.59 — never a loss with this kernel, so the host sets it for every grid (AVBD_EARLY_TRIGGER_BLOCKS caps the grid size).
    if (flags & 1) cudaTriggerProgrammaticLaunchCompletion();
    // bit 1: the immediate predecessor wrote the static inputs (visit-order geometry refreshed just before this launch): wait first
    if (flags & 2) cudaGridDependencySynchronize();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kSweepWarps + warp;
#ifdef AVBD_TIMELINE
    {
        int row = -1;
        if (gw == 0 && lane == 0) { unsigned r = atomicAdd(&g_timelineCount, 1u); row = r < 16384u ? (int)r : -1; }
        row = __shfl_sync(0xffffffffu, row, 0);
        if (lane == 0) pipes[warp].tlRow = gw == 0 ? row : -1;
        __syncwarp();
        AVBD_TL(pipes[warp].tlRow, 0);
    }
#endif
    if (gw >= nWarps) {                                        // no block-wide barrier anywhere in this kernel: a warp may leave on its own
        // the warps past the colour's ranges (first colour of a sweep only) take the bodies no contact visits and no user force
        // touches, one per lane: nothing they read is written by anyone else, so their colour does not matter
        const int t = (gw - nWarps) * 32 + lane;
        if (t < nFree) {
            const int i = __ldg(freeList + t);
            cudaGridDependencySynchronize();
            solve_free_body<false>(b, fv, i, prm, dxOut, diag, l2_keep_policy());
        }
        return;
    }
    const int vBegin = __ldg(range + gw), vEnd = __ldg(range + gw + 1);      // written by the graph stage, many launches ago
    if (vBegin >= vEnd) return;
    // launched with programmatic stream serialization: poses and lambda / penalty may still be in flight from the previous colour —
    // sweep_range waits for them after it has issued its static loads
    sweep_range<false, true>(pipes[warp], lane, vBegin, vEnd, b, visits, vg, ms, fv, prm, alpha, biasDual, dxOut, diag);
}

// The whole iteration loop of solver.cpp:340-431 (manifold rows) in ONE cooperative launch: the same warp pipelines, a grid barrier
// where the per-colour path has a kernel boundary.  A step of a mid-size batch (tens of thousands of bodies) is ~40-100 dependent
// colour phases of a few microseconds each; a grid barrier costs less than a launch + drain + ramp-up, and nothing but the barrier
// sits between two phases.  Poses / lambda / penalty are read through L2 only (COH).  The step's last dual pass stays a launch of its own.
struct SweepPlan { int nColours; int nWarps[64]; int off[64]; };
// CLUSTER: the grid is ONE thread-block cluster (<= 16 CTAs = 64 warps) and phases are separated by the hardware cluster barrier
// (release / acquire at cluster scope, a fraction of a microsecond) instead of a grid-wide barrier: the small-world form.
__device__ __forceinline__ void phase_barrier_cluster() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int MINB, bool CLUSTER>
__global__ void __launch_bounds__(32 * kSweepWarps, MINB) solve_loop_grid(BodyView b, const int4* __restrict__ visits, VisitGeom vg, ManifoldSet ms, ForceView fv,
                                                                         const int* __restrict__ ranges, SweepPlan plan, SolveParams prm, Diag* __restrict__ diag,
                                                                         const int* __restrict__ freeList, int nFree) {
    __shared__ WarpPipe pipes[kSweepWarps];
    cg::grid_group grid = cg::this_grid();
    if (CLUSTER) cudaGridDependencySynchronize();                           // launched with programmatic stream serialization
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw0 = blockIdx.x * kSweepWarps + warp, nGridWarps = gridDim.x * kSweepWarps;
    const int total = prm.iterations + (prm.postStabilize ? 1 : 0);
    float biasDual = -1.0f;                                                  // dual pass of the previous iteration still to apply (deferred dual)
    for (int it = 0; it < total; ++it) {
        const float alpha = prm.postStabilize ? (it < prm.iterations ? 1.0f : 0.0f) : prm.alpha;      // solver.cpp:340-342
        bool freeDone = nFree <= 0;
        for (int c = 0; c < plan.nColours; ++c) {
            const int nW = plan.nWarps[c];
            const int* range = ranges + plan.off[c];
            const int extra = freeDone ? 0 : (nFree + 31) / 32;             // contact-free bodies ride on the first phase
            for (int gw = gw0; gw < nW + extra; gw += nGridWarps) {
                if (gw < nW) {
                    const int vBegin = range[gw], vEnd = range[gw + 1];
                    if (vBegin < vEnd) sweep_range<true, false>(pipes[warp], lane, vBegin, vEnd, b, visits, vg, ms, fv, prm, alpha, biasDual, nullptr, diag);
                } else {
                    const int t = (gw - nW) * 32 + lane;
                    if (t < nFree) solve_free_body<true>(b, fv, freeList[t], prm, nullptr, diag, l2_keep_policy());
                }
            }
            freeDone = true;
            if (CLUSTER) phase_barrier_cluster(); else grid.sync();
        }
        biasDual = it < prm.iterations ? fminf(fmaxf(1.0f - alpha, 0.0f), 1.0f) : -1.0f;
    }
}


// Dynamic bodies no contact visits that a joint / spring links to another body: inertial terms + user forces -> 6x6 solve -> pose
// update, one colour per launch like every other body.  (The ones no user force touches ride on the first colour's sweep launch.)
__global__ void __launch_bounds__(kThreads) primal_free_bodies(BodyView b, ForceView fv, const int* __restrict__ freeList, int nFree, const int* __restrict__ colour,
                                                               int onlyColour, SolveParams prm, float* __restrict__ dxOut, Diag* __restrict__ diag) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nFree) return;
    const int i = __ldg(freeList + t);
    if (onlyColour >= 0 && __ldg(colour + i) != onlyColour) return;
    cudaGridDependencySynchronize();
    solve_free_body<false>(b, fv, i, prm, dxOut, diag, l2_keep_policy());
}

// Body-aligned warp ranges of every colour's visits, one launch: range[c][r] (r = 0 .. nWarps[c]) starts at the first body whose run
// starts at or after the r-th equal share of the colour's visits — every body's visits then belong to exactly one warp.
struct SweepGrids { int nWarps[64]; int off[64]; };
__global__ void warp_ranges(const int* __restrict__ vstart, const int2* __restrict__ colRange, SweepGrids g, int* __restrict__ range) {
    cudaGridDependencySynchronize();
    const int c = blockIdx.y;
    const int nW = g.nWarps[c];
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nW) return;
    int2 cr = colRange[c];
    const int* vs = vstart + cr.x; int count = cr.y - cr.x;
    int vBegin = vs[0], vEnd = vs[count];
    long long target = (long long)vBegin + ((long long)(vEnd - vBegin) * r) / nW;
    int lo = 0, hi = count;                        // first k in [0, count] with vs[k] >= target (vs[count] == vEnd)
    while (lo < hi) { int mid = (lo + hi) >> 1; if (vs[mid] < target) lo = mid + 1; else hi = mid; }
    range[g.off[c] + r] = vs[lo];
}

// ------------------------------------------------------------------ dual
// One LIVE contact (solver.cpp:411-430 for manifold rows).  Returns what the diagnostics need.
struct DualOut { float sepn, lamN; int visits, world, first; };
// `unvisitedReps` > 0 is the deferred-dual form (the primal sweeps already applied every pass but the last to each
// contact they visit): a contact NO dynamic body visits (both endpoints static) takes all `unvisitedReps` passes here —
// its poses never move, so repeating the update in place equals the reference's once-per-iteration pass — and with
// `onlyUnvisited` every other contact is left alone (postStabilize: the extra sweep already applied the last pass).
template <bool COH, bool DIAG>
__device__ __forceinline__ DualOut dual_one(const BodyView& b, const ManifoldSet& ms, int ci, const SolveParams& prm, float alpha,
                                            int unvisitedReps = 0, bool onlyUnvisited = false) {
    int m = ms.cM[ci];
    int4 h = ms.hdr[m];
    BodyPose pa = load_pose<COH>(b.pose + h.x), pb = load_pose<COH>(b.pose + h.y);
    ContactState cs = load_contact_c<COH>(ms, ci);
    ContactEval ev;
    DualOut o;
    o.visits = (pa.pos.w > 0.0f ? 1 : 0) + (pb.pos.w > 0.0f ? 1 : 0);
    int reps = 1;
    if (unvisitedReps > 0) reps = o.visits == 0 ? unvisitedReps : (onlyUnvisited ? 0 : 1);
    float sep[3];
    float cap = rows_geometry(pa.pos, pa.rot, pb.pos, pb.rot, 1.0f, cs, ev, sep);
    float bias = fminf(fmaxf(1.0f - alpha, 0.0f), 1.0f);
    for (int r = 0; r < reps; ++r) {
        limits_fast(cap, __int_as_float(h.w), bias, sep, cs, ev);
        dual_fast(cs, ev, prm.beta);
    }
    if (reps > 0) { ContactLP q; q.l = pack_lambda(cs); q.p = pack_penalty(cs); ms.lp[ci] = q; }
    o.sepn = dot((xyz(pa.pos) + ev.wrA) - (xyz(pb.pos) + ev.wrB), cs.n);
    o.lamN = cs.lam[0];
    o.world = -1; o.first = 0;
    if (DIAG) { o.world = b.worldId[h.x]; o.first = (ci == 0 || ms.cM[ci - 1] != m) ? 1 : 0; }
    return o;
}

template <bool DIAG>
__global__ void __launch_bounds__(kThreads) dual_contacts(BodyView b, ManifoldSet ms, int nContacts, SolveParams prm, float alpha, int unvisitedReps,
                                                          bool onlyUnvisited, Diag* diag) {
    cudaGridDependencySynchronize();
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    DualOut o{0.0f, 0.0f, 0, -1, 0};
    if (ci < nContacts) o = dual_one<false, DIAG>(b, ms, ci, prm, alpha, unvisitedReps, onlyUnvisited);
    if (DIAG) reduce_contact_diag_block(o.world, o.sepn, o.lamN, o.world >= 0 ? 1 : 0, o.first, o.visits, diag);
}

// ------------------------------------------------------------------ persistent iteration loop (small worlds)
// A Stress1000-sized world is latency bound: iterations x (colours + dual) dependent phases of a few hundred threads
// each, so per-colour launches pay two launch latencies per phase and a grid-wide atomic barrier is no cheaper.  This
// kernel runs the WHOLE loop of solver.cpp:340-431 in one launch of ONE thread-block cluster (up to 16 CTAs on the SMs
// of one GPC): phases are separated by the hardware cluster barrier (barrier.cluster, release / acquire at cluster
// scope) instead of a kernel boundary, and each phase is the tile routine above (one contact visit per thread, in-order
// shared-memory sums, block solve) on tiles whose static inputs the CTA caches in shared memory (TileCache).  Data other CTAs write between barriers (poses, lambda,
// penalty) is read with ld.global.cg (L2); stores are write-through.
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BPB>
__global__ void __launch_bounds__(kThreads, 1) solve_loop_cluster(BodyView b, const int* __restrict__ visitStart, const int4* __restrict__ visits,
                                                                  ManifoldSet ms, ForceView fv, const int* __restrict__ order,
                                                                  const int2* __restrict__ colRange, int nColours, int nContacts, SolveParams prm,
                                                                  Diag* diag, bool contactDiag, bool anyUnvisited, int nSlots) {
    cudaGridDependencySynchronize();
    __shared__ PrimalSmem<BPB> sm;
    extern __shared__ __align__(16) unsigned char dynSmem[];
    TileCache<BPB>* slots = reinterpret_cast<TileCache<BPB>*>(dynSmem);      // nSlots of them: this CTA's first tiles, in processing order
    const int rank = (int)cluster_rank(), nCta = (int)cluster_size();       // the grid is one cluster
    int total = prm.iterations + (prm.postStabilize ? 1 : 0);
    {
        int slot = 0;
        for (int c = 0; c < nColours && slot < nSlots; ++c) {
            int2 r = colRange[c];
            int count = r.y - r.x;
            for (int tile = rank; tile * BPB < count && slot < nSlots; tile += nCta, ++slot)
                fill_tile_cache<BPB>(slots[slot], b, visitStart + r.x, visits, ms, order + r.x, count, tile);
        }
    }
    float biasDual = -1.0f;                                    // dual pass of the previous iteration still to apply (deferred dual)
    for (int it = 0; it < total; ++it) {
        float alpha = prm.postStabilize ? (it < prm.iterations ? 1.0f : 0.0f) : prm.alpha;      // solver.cpp:340-342
        int slot = 0;
        for (int c = 0; c < nColours; ++c) {
            int2 r = colRange[c];
            int count = r.y - r.x;
            for (int tile = rank; tile * BPB < count; tile += nCta, ++slot) {
                const TileCache<BPB>* tc = (slot < nSlots && slots[slot].usable) ? &slots[slot] : nullptr;
                primal_tile_visits<BPB, true>(b, visitStart + r.x, visits, ms, fv, order + r.x, count, tile, prm, alpha, biasDual, nullptr, diag, sm, tc);
                __syncthreads();
            }
            cluster_barrier();
        }
        biasDual = it < prm.iterations ? fminf(fmaxf(1.0f - alpha, 0.0f), 1.0f) : -1.0f;
    }
    // what the sweeps could not apply: the last iteration's dual pass (nothing moves after it, so the contact diagnostics
    // are reduced from the same registers), or — when postStabilize's extra sweep already applied it — only the contacts
    // no dynamic body visits
    bool lastPending = biasDual >= 0.0f;
    if (prm.iterations > 0 && (lastPending || anyUnvisited)) {
        float alpha = prm.postStabilize ? 1.0f : prm.alpha;
        bool reduce = contactDiag && lastPending;
        int rounded = (nContacts + 31) & ~31;                   // whole warps stay in the loop (warp-level reductions below)
        for (int t = rank * kThreads + (int)threadIdx.x; t < rounded; t += nCta * kThreads) {
            DualOut o{0.0f, 0.0f, 0, -1, 0};
            if (t < nContacts) o = dual_one<true, true>(b, ms, t, prm, alpha, prm.iterations, !lastPending);
            if (reduce) reduce_contact_diag(o.world, o.sepn, o.lamN, o.world >= 0 ? 1 : 0, o.first, o.visits, diag);
        }
    }
}

__global__ void dual_user_forces(BodyView b, ForceView fv, SolveParams prm) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < fv.nJoints) {
        JointRec& j = fv.joints[t];
        bool hasA = j.a >= 0;
        V3 pA = zero3(); Q4 qA = qid();
        if (hasA) { BodyPose a = b.pose[j.a]; pA = xyz(a.pos); qA = quat(a.rot); }
        BodyPose pbp = b.pose[j.b]; V3 pB = xyz(pbp.pos); Q4 qB = quat(pbp.rot);
        ForceEval ev;
        joint_constraint(j, hasA, pA, qA, pB, qB, ev);
        for (int r = 0; r < 6; ++r) {
            if (j.stiffness[r] != FLT_MAX) continue;
            float lu = clampf(j.penalty[r] * ev.C[r] + j.lambda[r], ev.fmin[r], ev.fmax[r]);
            bool active = lu > ev.fmin[r] && lu < ev.fmax[r];
            j.lambda[r] = lu;
            if (active) {
                float lw = 0.0f, aw = 0.0f; V3 Jl, Ja;
                if (hasA) { joint_jacobian(j, true, qA, r, Jl, Ja); lw += len2(Jl); aw += len2(Ja); }
                joint_jacobian(j, false, qB, r, Jl, Ja); lw += len2(Jl); aw += len2(Ja);
                j.penalty[r] = fmin2(j.penalty[r] + penalty_gain(lw, aw, prm.beta) * fabsf(ev.C[r]), kPenaltyMax);
            }
        }
    } else if (t - fv.nJoints < fv.nSprings) {
        SpringRec& s = fv.springs[t - fv.nJoints];
        if (s.k != FLT_MAX) return;                 // soft rows skip the dual (solver.cpp:416-418)
        bool hasA = s.a >= 0;
        V3 pA = zero3(); Q4 qA = qid();
        if (hasA) { BodyPose a = b.pose[s.a]; pA = xyz(a.pos); qA = quat(a.rot); }
        BodyPose pbp = b.pose[s.b]; V3 pB = xyz(pbp.pos); Q4 qB = quat(pbp.rot);
        float C = spring_constraint(s, hasA, pA, qA, pB, qB);
        float lu = clampf(s.penalty * C + s.lambda, -FLT_MAX, FLT_MAX);
        bool active = lu > -FLT_MAX && lu < FLT_MAX;
        s.lambda = lu;
        if (active) {
            float lw = 0.0f, aw = 0.0f; V3 Jl, Ja;
            if (hasA) { spring_jacobian(s, hasA, pA, qA, pB, qB, true, Jl, Ja); lw += len2(Jl); aw += len2(Ja); }
            spring_jacobian(s, hasA, pA, qA, pB, qB, false, Jl, Ja); lw += len2(Jl); aw += len2(Ja);
            s.penalty = fmin2(s.penalty + penalty_gain(lw, aw, prm.beta) * fabsf(C), kPenaltyMax);
        }
    }
}

// Batched 6x6 solves on caller data (parity harness for solve6x6, solver.cpp:68-83).
// lhs: ll la al aa blocks, each 9 floats column-major; only what the solve reads is used.
__global__ void solve6_batch(const float* lhs36, const float* rhs6, int n, float* out6) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* L = lhs36 + 36 * i; const float* r = rhs6 + 6 * i;
    BodySystem s;
    // column-major blocks: element (row r, col c) at c*3+r
    s.ll[0] = L[0]; s.ll[1] = L[1]; s.ll[2] = L[2]; s.ll[3] = L[4]; s.ll[4] = L[5]; s.ll[5] = L[8];
    for (int rr = 0; rr < 3; ++rr) for (int c = 0; c < 3; ++c) s.la[rr * 3 + c] = L[9 + c * 3 + rr];
    const float* A = L + 27;
    s.aa[0] = A[0]; s.aa[1] = A[1]; s.aa[2] = A[2]; s.aa[3] = A[4]; s.aa[4] = A[5]; s.aa[5] = A[8];
    for (int k = 0; k < 3; ++k) { s.rl[k] = r[k]; s.ra[k] = r[3 + k]; }
    V3 dl, da;
    solve_body_system(s, dl, da);
    float* o = out6 + 6 * i;
    o[0] = dl.x; o[1] = dl.y; o[2] = dl.z; o[3] = da.x; o[4] = da.y; o[5] = da.z;
}


// ------------------------------------------------------------------ launchers (declared in avbd_launch.h)
static inline int blocks_of(long long n, int per) { long long b = (n + per - 1) / per; return (int)(b < 1 ? 1 : b); }

// Function attributes and occupancy are PER DEVICE (a process may hold worlds on several GPUs): cached by device index.
constexpr int kMaxDevices = 64;
static int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < kMaxDevices) ? dev : 0; }

// Resident warps of the sweep kernel on the current device (persistent-style sizing: a colour never gets more warps than fit at once).
static int sweep_cfg() { static const int cfg = [] { const char* e = getenv("AVBD_SWEEP"); return e ? atoi(e) : 5; }(); return cfg; }
template <int MINB> static int sweep_blocks_per_sm() {
    int per = 0;
    cudaFuncSetAttribute(primal_sweep_warp<MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, primal_sweep_warp<MINB>, 32 * kSweepWarps, 0) != cudaSuccess || per < 1) { cudaGetLastError(); per = 1; }
    return per;
}
int primal_sweep_max_warps() {
    static int cache[kMaxDevices] = {0};
    const int dev = current_device();
    if (!cache[dev]) {
        int sms = 148, per = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        switch (sweep_cfg()) {
            case 3:  per = sweep_blocks_per_sm<3>(); break;
            case 4:  per = sweep_blocks_per_sm<4>(); break;
            default: per = sweep_blocks_per_sm<5>(); break;
        }
        if (getenv("AVBD_DEBUG")) fprintf(stderr, "primal_sweep_warp<%d>: %d blocks per SM resident, %d B shared memory per block (device %d)\n", sweep_cfg(), per,
                                          (int)(kSweepWarps * sizeof(WarpPipe)), dev);
        cache[dev] = sms * per * kSweepWarps;
    }
    return cache[dev];
}
// Warps a colour with nVisits contact visits is cut into: one per chunk of 32 visits until every resident slot is taken (a small
// colour is latency bound: short ranges finish sooner), then longer ranges on the resident warps.
int primal_sweep_warps(int nVisits) {
    int want = (nVisits + 31) / 32;
    int mx = primal_sweep_max_warps();
    return want < 1 ? 1 : (want < mx ? want : mx);
}
void launch_warp_ranges(cudaStream_t s, const int* vstart, const int2* colRange, int nColours, const int* nWarps, const int* off, int* range) {
    SweepGrids g{};
    int mx = 0;
    for (int c = 0; c < nColours && c < 64; ++c) { g.nWarps[c] = nWarps[c]; g.off[c] = off[c]; mx = nWarps[c] > mx ? nWarps[c] : mx; }
    if (nColours > 0) launch_dep(warp_ranges, dim3(blocks_of(mx + 1, kThreads), nColours), dim3(kThreads), 0, s, vstart, colRange, g, range);
}
void launch_primal_free(cudaStream_t s, BodyView b, ForceView fv, const int* freeList, int nFree, const int* colour, int onlyColour, SolveParams prm,
                        float* dxOut, Diag* diag) {
    if (nFree > 0) launch_dep(primal_free_bodies, dim3(blocks_of(nFree, kThreads)), dim3(kThreads), 0, s, b, fv, freeList, nFree, colour, onlyColour, prm, dxOut, diag);
}
// One colour of the large-world sweep: its visits, cut into nWarps body-aligned warp ranges (range[0 .. nWarps]).
void launch_primal_sweep(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* range, int nWarps, SolveParams prm,
                         float alpha, float biasDual, float* dxOut, Diag* diag, const int* freeList, int nFree, bool staticsJustWritten) {
    if (nWarps <= 0 && nFree <= 0) return;
    if (nWarps < 0) nWarps = 0;
    if (nFree < 0) nFree = 0;
    dim3 grid(blocks_of(nWarps + (nFree + 31) / 32, kSweepWarps)), block(32 * kSweepWarps);
    static const int triggerBlocks = [] { const char* e = getenv("AVBD_EARLY_TRIGGER_BLOCKS"); return e ? atoi(e) : (1 << 30); }();
    const int early = ((int)grid.x <= triggerBlocks ? 1 : 0) | (staticsJustWritten ? 2 : 0);
    switch (sweep_cfg()) {
        case 3:  launch_dep(primal_sweep_warp<3>, grid, block, 0, s, b, visits, vg, ms, fv, range, nWarps, prm, alpha, biasDual, dxOut, diag, freeList, nFree, early); break;
        case 4:  launch_dep(primal_sweep_warp<4>, grid, block, 0, s, b, visits, vg, ms, fv, range, nWarps, prm, alpha, biasDual, dxOut, diag, freeList, nFree, early); break;
        default: launch_dep(primal_sweep_warp<5>, grid, block, 0, s, b, visits, vg, ms, fv, range, nWarps, prm, alpha, biasDual, dxOut, diag, freeList, nFree, early); break;
    }
}

// The whole iteration loop in one cooperative launch (solve_loop_grid).  nWarps / off: per colour, as for launch_warp_ranges.
// Returns false if the launch was refused (caller falls back to per-colour launches).
template <int MINB> static int loop_blocks_per_sm() {
    int per = 0;
    cudaFuncSetAttribute(solve_loop_grid<MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, solve_loop_grid<MINB, false>, 32 * kSweepWarps, 0) != cudaSuccess || per < 1) { cudaGetLastError(); per = 0; }
    return per;
}
bool launch_solve_loop_grid(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* ranges, int nColours,
                            const int* nWarps, const int* off, SolveParams prm, Diag* diag, const int* freeList, int nFree) {
    static int residentDev[kMaxDevices] = {0};
    const int dev = current_device();
    if (!residentDev[dev]) {
        int sms = 148, coop = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        int per = sweep_cfg() == 4 ? loop_blocks_per_sm<4>() : loop_blocks_per_sm<5>();
        residentDev[dev] = (coop && per > 0) ? sms * per : -1;
        if (getenv("AVBD_DEBUG")) fprintf(stderr, "solve_loop_grid: %d blocks per SM resident (device %d)\n", per, dev);
    }
    if (residentDev[dev] <= 0 || nColours <= 0 || nColours > 64) return false;
    SweepPlan plan{};
    plan.nColours = nColours;
    int mx = 1;
    for (int c = 0; c < nColours; ++c) {
        plan.nWarps[c] = nWarps[c]; plan.off[c] = off[c];
        int need = nWarps[c] + (c == 0 ? (nFree + 31) / 32 : 0);
        mx = need > mx ? need : mx;
    }
    int grid = blocks_of(mx, kSweepWarps);
    if (grid > residentDev[dev]) grid = residentDev[dev];
    void* args[] = {&b, &visits, &vg, &ms, &fv, &ranges, &plan, &prm, &diag, &freeList, &nFree};
    cudaError_t e = sweep_cfg() == 4
        ? cudaLaunchCooperativeKernel((void*)solve_loop_grid<4, false>, dim3(grid), dim3(32 * kSweepWarps), args, 0, s)
        : cudaLaunchCooperativeKernel((void*)solve_loop_grid<5, false>, dim3(grid), dim3(32 * kSweepWarps), args, 0, s);
    if (e != cudaSuccess) { cudaGetLastError(); residentDev[dev] = -1; return false; }
    return true;
}

// The same loop as ONE thread-block cluster (small worlds: every colour fits the cluster's warps a few times over).
bool launch_solve_loop_warps(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* ranges, int nColours,
                             const int* nWarps, const int* off, SolveParams prm, Diag* diag, const int* freeList, int nFree) {
    static int maxClusterDev[kMaxDevices] = {0};
    const int dev = current_device();
    if (!maxClusterDev[dev]) {
        int mc = 16;
        if (cudaFuncSetAttribute(solve_loop_grid<4, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); mc = 8; }
        else if (const char* e = getenv("AVBD_CLUSTER_MAX")) { int v = atoi(e); if (v >= 1 && v <= 16) mc = v; }
        maxClusterDev[dev] = mc;
    }
    if (nColours <= 0 || nColours > 64) return false;
    SweepPlan plan{};
    plan.nColours = nColours;
    int mx = 1;
    for (int c = 0; c < nColours; ++c) {
        plan.nWarps[c] = nWarps[c]; plan.off[c] = off[c];
        int need = nWarps[c] + (c == 0 ? (nFree + 31) / 32 : 0);
        mx = need > mx ? need : mx;
    }
    int nCta = 1;
    while (nCta * kSweepWarps < mx && nCta < maxClusterDev[dev]) nCta <<= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nCta); cfg.blockDim = dim3(32 * kSweepWarps); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nCta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, solve_loop_grid<4, true>, b, visits, vg, ms, fv, ranges, plan, prm, diag, freeList, nFree);
    if (e != cudaSuccess && nCta > 8) {          // a 16-CTA cluster may not be placeable: retry with the portable size
        cudaGetLastError();
        maxClusterDev[dev] = 8;
        attr[0].val.clusterDim.x = 8; cfg.gridDim = dim3(8);
        e = cudaLaunchKernelEx(&cfg, solve_loop_grid<4, true>, b, visits, vg, ms, fv, ranges, plan, prm, diag, freeList, nFree);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}

bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, int nContacts, SolveParams prm,
                       Diag* diag, bool contactDiag, bool anyUnvisited) {
    constexpr int BPB = kClusterBodiesPerTile;
    static int maxClusterDev[kMaxDevices] = {0}, maxSlotsDev[kMaxDevices] = {0};
    const int dev = current_device();
    if (!maxClusterDev[dev]) {
        // 16 CTAs need the non-portable opt-in; fall back to the portable 8 if the device refuses it
        int mc = 16;
        if (cudaFuncSetAttribute(solve_loop_cluster<BPB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); mc = 8; }
        else if (const char* e = getenv("AVBD_CLUSTER_MAX")) { int v = atoi(e); if (v >= 1 && v <= 16) mc = v; }
        // shared-memory slots for the tile cache: whatever the SM has left next to the kernel's static shared memory
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa{}; cudaFuncGetAttributes(&fa, solve_loop_cluster<BPB>);
        long long room = (long long)optin - (long long)fa.sharedSizeBytes - 1024;
        int n = room > 0 ? (int)(room / (long long)sizeof(TileCache<BPB>)) : 0;
        if (const char* e = getenv("AVBD_TILE_CACHE_SLOTS")) n = atoi(e) < n ? atoi(e) : n;
        if (n > 0 && cudaFuncSetAttribute(solve_loop_cluster<BPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * (int)sizeof(TileCache<BPB>)) != cudaSuccess) { cudaGetLastError(); n = 0; }
        maxClusterDev[dev] = mc; maxSlotsDev[dev] = n + 1;          // + 1: 0 means "not initialised"
    }
    int& maxCluster = maxClusterDev[dev];
    const int maxSlots = maxSlotsDev[dev] - 1;
    int want = blocks_of(maxColourCount, BPB);
    int wantDual = blocks_of(nContacts, kThreads);
    int need = want > wantDual ? want : wantDual;
    int nCta = 1;
    while (nCta < need && nCta < maxCluster) nCta <<= 1;
    cudaLaunchConfig_t cfg = {};
    // a CTA runs ceil(tiles of a colour / nCta) tiles per colour: no point in more slots than it has tiles
    int tilesPerCta = 0;
    for (int c = 0; c < nColours; ++c) tilesPerCta += (blocks_of(maxColourCount, BPB) + nCta - 1) / nCta;      // upper bound (largest colour for all)
    int nSlots = tilesPerCta < maxSlots ? tilesPerCta : maxSlots;
    cfg.gridDim = dim3(nCta); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = (size_t)nSlots * sizeof(TileCache<BPB>); cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nCta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;     // see launch_dep
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, solve_loop_cluster<BPB>, b, visitStart, visits, ms, fv, order, colRange, nColours, nContacts, prm, diag, contactDiag, anyUnvisited, nSlots);
    if (e != cudaSuccess && nCta > 8) {          // a 16-CTA cluster may not be placeable (MIG slices, busy GPCs): retry with the portable size
        cudaGetLastError();
        maxCluster = 8;
        attr[0].val.clusterDim.x = 8; cfg.gridDim = dim3(8);
        e = cudaLaunchKernelEx(&cfg, solve_loop_cluster<BPB>, b, visitStart, visits, ms, fv, order, colRange, nColours, nContacts, prm, diag, contactDiag, anyUnvisited, nSlots);
    }
    return e == cudaSuccess;
}

void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, int nContacts, SolveParams prm, float alpha, int unvisitedReps, bool onlyUnvisited, Diag* diag) {
    Diag* none = nullptr;
    if (diag) launch_dep(dual_contacts<true>, dim3(blocks_of(nContacts, kThreads)), dim3(kThreads), 0, s, b, ms, nContacts, prm, alpha, unvisitedReps, onlyUnvisited, diag);
    else      launch_dep(dual_contacts<false>, dim3(blocks_of(nContacts, kThreads)), dim3(kThreads), 0, s, b, ms, nContacts, prm, alpha, unvisitedReps, onlyUnvisited, none);
}
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm) {
    launch_dep(dual_user_forces, dim3(blocks_of(fv.nJoints + fv.nSprings, kThreads)), dim3(kThreads), 0, s, b, fv, prm);
}
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6) {
    solve6_batch<<<blocks_of(n, 128), 128, 0, s>>>(lhs36, rhs6, n, out6);
}

} // namespace avbd
