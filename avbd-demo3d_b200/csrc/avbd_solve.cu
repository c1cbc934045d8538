// avbd_solve.cu — the per-iteration solver kernels (primal block solve per colour, dual / penalty ramp) and
// their launchers.  This translation unit is compiled WITH FMA contraction: its outputs are held to an FP32
// tolerance against the reference (BASELINE.json north_star), unlike the collision path in avbd_engine.cu.
//
// The reference walks bodies serially (Gauss-Seidel, solver.cpp:344); here bodies of one colour share no
// manifold, so a colour is one parallel phase, one contact visit (computeConstraint + 3 rows) per thread:
//   large worlds   per colour primal_visit_flat (flat visit partition -> per-body sums) + primal_solve_flat (block solve)
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
#include "avbd_launch.h"
#include "avbd_body.cuh"
#include "avbd_forces.cuh"

namespace avbd {

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

// ------------------------------------------------------------------ row math of the solver kernels
// The large-world visit kernel is instruction-issue bound (ncu: ~45 % issue slots busy, DRAM at a third of peak), so the device
// copy of the row math is written for instruction count.  Same formulas as avbd_rows.cuh (which the host mirror and the host
// emulation build use, and which the parity tests compare this against), with:
//   * the contact seen from the VISITING body: geometry arrives as {r_self, r_other, n} (visit_geometry swaps once per step) and
//     the A/B orientation is one sign, sg — (pA + wA) - (pB + wB) = sg ((pS + wS) - (pO + wO)) exactly, so no operand selects;
//   * clamps as fminf / fmaxf (one FMNMX each; a NaN operand yields the bound where the reference's ternary clamp passes the
//     NaN on — NaN states are scrubbed by the pose update either way);
//   * the dual's |w x b|^2 as |w|^2 - (w . b)^2 for the unit basis vectors b (|b|^2 taken as 1), one division per active row;
//   * approximate (2 ulp) reciprocal / square root / reciprocal square root instructions, IN THE ROWS ONLY: the block solve and the
//     pose update keep IEEE division and sqrt — building the whole translation unit with -prec-div=false -prec-sqrt=false
//     moved the Stack's rest heights by 2e-3 and toppled the Pyramid's apex box (tools/rest_probe.py).
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
__device__ __forceinline__ void system_fast(BodySystem& s, const ContactState& c, const ContactEval& e, V3 w, float sg, bool gyro, const M3& invIw) {
    float2 v[14];
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
    s.rl[0] = v[0].x; s.rl[1] = v[0].y; s.rl[2] = v[1].x; s.ra[0] = v[1].y; s.ra[1] = v[2].x; s.ra[2] = v[2].y;
    s.ll[0] = v[3].x; s.ll[1] = v[3].y; s.ll[2] = v[4].x; s.ll[3] = v[4].y; s.ll[4] = v[5].x; s.ll[5] = v[5].y;
    s.la[0] = v[6].x; s.la[1] = v[6].y; s.la[2] = v[7].x; s.la[3] = v[7].y; s.la[4] = v[8].x; s.la[5] = v[8].y; s.la[6] = v[9].x; s.la[7] = v[9].y;
    s.la[8] = v[10].x; s.aa[0] = v[10].y; s.aa[1] = v[11].x; s.aa[2] = v[11].y; s.aa[3] = v[12].x; s.aa[4] = v[12].y; s.aa[5] = v[13].x;
}
// One contact visit in the visiting body's frame: computeConstraint (with the pending dual update first), then the 3 rows'
// contribution to the body's 6x6 system.  `sp,sq` / `op,oq` = self / other pose, g0 g1 g2 = {r_self,C0n} {r_other,C0t.x} {n,C0t.y}.
// Every solver kernel (flat visit kernel, cluster loop, dual pass) evaluates rows through these, so the deferred and the
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
// be split over launches.  (Large worlds use the flat visit partition further down.)
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

constexpr int kSumStride = 28;

// ------------------------------------------------------------------ primal, flat visit partition (the default large-world path)
// Giving each block a tile of BODIES makes it load where its visits start (one DRAM round trip), then the visit entries (a
// second), then what they point at (a third), and leaves ~4 % of the lanes idle because a tile's visits rarely fill the block.
// This kernel partitions the colour's VISITS instead: chunk c is visits [vBegin + c*T, vBegin + (c+1)*T), every lane busy,
// start known from the block index.  Blocks are persistent (grid = a few per SM, chunks round-robin) and load the NEXT chunk's
// visit entry before working on the current one, so only the gathers (poses, lambda / penalty) remain on the critical path.
//   phase 0  segment heads: lane v starts a segment when visit v-1 belongs to another body (entries carry the visiting body);
//            ballot -> each warp's compact list of the segments that start in its 32 visits (no barrier)
//   phase 1  thread t takes visit base+t: computeConstraint (+ pending dual) + 3 rows -> 27 partial sums, one 112-byte row of
//            shared memory per visit (7 x STS.128)
//   ---- the chunk's only block barrier (rows and segment lists are double buffered) ----
//   phase 2  7 lanes per segment add the segment's run of rows in visit order, one float4 column each (LDS.128, 4 in flight);
//            a warp sums the segments that start in its visits
//   output   a segment that STARTS in the chunk writes sums[k], k = the body's position in the colour order; a segment continuing from the previous chunk (a body whose
//            visits straddle a chunk boundary) writes carry[chunk] instead, and primal_solve_flat adds main + carries in
//            chunk order — deterministic, no atomics, nothing to zero.
constexpr int kFlatLanes = 7;                 // lanes per segment in phase 2 (one float4 = 4 of the 27(+1) components each)
template <int T>
struct FlatSmem {
    float4 c[2][T][7];   // one row of 28 floats per visit: rl(3) ra(3) ll(6) la(9) aa(6) pad — 128-bit stores / loads, row stride 28 words: conflict free
    float4 stage[9][T];  // the NEXT chunk's gathered operands, filled by cp.async: self pose (2), other pose (2), geometry (3), lambda, penalty
    unsigned char segPos[2][T / 32][32];     // per warp: chunk-local positions (< T <= 256) of the segment heads among its 32 visits
    int segK[2][T / 32][32];       // ... and each segment's row of `sums` (complemented: it continues the previous chunk's last segment)
    unsigned headMask[2][T / 32];
    float4 carry[2][8];            // partial sum of the body whose run crosses into the next chunk (by chunk parity)
    int nextK[2];                  // row of the body the next chunk starts with (-1: this is the block's last chunk)
};

// cp.async (LDGSTS) of one 16-byte item into this thread's slot of the gather stage, with an L2 eviction policy.
__device__ __forceinline__ void stage16(float4* dst, const float4* src, unsigned long long policy) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void stage16_nol1(float4* dst, const float4* src, unsigned long long policy) {      // bypasses L1 (streamed / single-use data)
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_stream_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int T, int MINB, bool ALIGNED>
__global__ void __launch_bounds__(T, MINB) primal_visit_flat(BodyView b, const int4* __restrict__ visits, VisitGeom vg, ManifoldSet ms,
                                                             int vFirst, int vLast, const int* __restrict__ range, const int* __restrict__ kOf,
                                                             float alpha, float biasDual, float beta, float* __restrict__ sums, float* __restrict__ carry) {
    __shared__ FlatSmem<T> sm;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long keep = l2_keep_policy(), stream = l2_stream_policy();
    // launched with programmatic stream serialization: the grid may already be resident while the previous colour's block solve
    // drains; nothing it wrote (poses) is read before this point
    cudaGridDependencySynchronize();
    // Two ways to hand out the colour's visits [vFirst, vLast):
    //   range == nullptr  chunks round-robin over the blocks (chunk c = visits vFirst + c*T ...).  Fastest; a body whose run crosses a
    //                     chunk boundary gets its sum in two pieces (sums[k] + carry[chunk], added by primal_solve_flat), so its
    //                     rounding depends on where the chunk grid falls — harmless for ONE world, but it makes a world's
    //                     trajectory depend on what else is in the batch;
    //   range != nullptr  (batches of several worlds) block r owns the contiguous visits [range[r], range[r + 1]), cut on BODY
    //                     boundaries (flat_ranges): a run that crosses a chunk boundary stays inside the block, its partial sum
    //                     waits in shared memory and the sum is one sequence in visit order — exactly the cluster loop's, and
    //                     independent of the batch (bit-identical ensembles however they are partitioned over GPUs).
    constexpr bool aligned = ALIGNED;           // compile-time: the round-robin instantiation carries none of the other mode's code
    const int vBegin = aligned ? range[blockIdx.x] : vFirst, vEnd = aligned ? range[blockIdx.x + 1] : vLast;
    const int stride = aligned ? T : (int)gridDim.x * T;
    const int start = aligned ? vBegin : vBegin + (int)blockIdx.x * T;
    if (start >= vEnd) return;
    const int4 none = make_int4(0, 0, -8, 0);                        // body -1
    // Software pipeline over the block's chunks, two deep: the visit ENTRIES are loaded two chunks ahead (registers), and as soon
    // as an entry is there the data it points at — two poses, the lambda / penalty record, the streamed geometry: 9 x 16 B per
    // visit — is fetched one chunk ahead with cp.async into this thread's slots of a shared-memory stage.  A chunk therefore
    // starts with its operands already on chip; both DRAM round trips hide behind the previous chunk's row math and reduction.
    auto load_entry = [&](int v, int4& e, int& prevZ) {
        e = none; prevZ = -8;
        if (v < vEnd) { e = __ldcs(visits + v); if (lane == 0 && v > vBegin) prevZ = __ldg(&visits[v - 1].z); }
    };
    auto issue_gathers = [&](int v, const int4& e, int& kSelf) {         // into the stage; the caller commits the group
        kSelf = 0;
        if (v < vEnd) {
            int self = e.z >> 3;
            kSelf = __ldg(kOf + self);                                    // the body's position in the colour order = its row of `sums`
            stage16(&sm.stage[0][t], &b.pose[self].pos, keep); stage16(&sm.stage[1][t], &b.pose[self].rot, keep);
            stage16(&sm.stage[2][t], &b.pose[e.y].pos, keep);  stage16(&sm.stage[3][t], &b.pose[e.y].rot, keep);
            stage16_nol1(&sm.stage[4][t], vg.a + v, stream); stage16_nol1(&sm.stage[5][t], vg.b + v, stream); stage16_nol1(&sm.stage[6][t], vg.n + v, stream);
            stage16_nol1(&sm.stage[7][t], &ms.lp[e.x].l, keep); stage16_nol1(&sm.stage[8][t], &ms.lp[e.x].p, keep);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int4 eCur, eNext; int prevCur, prevNext, kCur;
    load_entry(start + t, eCur, prevCur);
    load_entry(start + t + stride, eNext, prevNext);
    issue_gathers(start + t, eCur, kCur);
    for (int it = 0, base = start; base < vEnd; base += stride, ++it) {
        const int v = base + t;
        const int4 e = eCur; const int prevZ = prevCur; const int kSelf = kCur;
        const bool live = v < vEnd;
        const int self = e.z >> 3;
        ContactLP* lp = ms.lp + e.x;
        // ---- this chunk's operands: wait for the thread's own copies, move them to registers, and refill the stage at once
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        BodyPose ps, po; float4 a4, b4, n4, l4, p4;
        ps.pos = sm.stage[0][t]; ps.rot = sm.stage[1][t]; po.pos = sm.stage[2][t]; po.rot = sm.stage[3][t];
        a4 = sm.stage[4][t]; b4 = sm.stage[5][t]; n4 = sm.stage[6][t]; l4 = sm.stage[7][t]; p4 = sm.stage[8][t];
        eCur = eNext; prevCur = prevNext;
        issue_gathers(v + stride, eCur, kCur);                        // next chunk (its entry was loaded a whole chunk ago)
        load_entry(v + 2 * stride, eNext, prevNext);                  // the entry after that
        // ---- phase 0: this warp's segment heads (no barrier: each warp lists its own, the block barrier after phase 1 publishes them)
        const int buf = it & 1;                                       // rows and segment lists are double buffered: ONE barrier per chunk
        int prevSelf = __shfl_up_sync(0xffffffffu, self, 1);
        if (lane == 0) prevSelf = prevZ >> 3;                         // -1 at the colour's first visit
        const bool head = live && (prevSelf != self || t == 0);       // lane 0 of the chunk always opens a segment (maybe a continuation)
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (head) {
            int r = __popc(heads & ((1u << lane) - 1u));
            sm.segPos[buf][warp][r] = (unsigned char)t;
            sm.segK[buf][warp][r] = (t == 0 && prevSelf == self) ? ~kSelf : kSelf;    // complemented: continues the previous chunk's last segment
        }
        if (lane == 0) sm.headMask[buf][warp] = heads;
        if (aligned && t == 0) sm.nextK[buf] = base + T < vEnd ? kCur : -1;     // row of the body the block's NEXT chunk opens with (kCur was loaded for it above)
        // ---- phase 1
        if (live) {
            bool gyro = (e.z & 2) != 0, pending = biasDual >= 0.0f && (e.z & 4) != 0;
            float sg = (e.z & 1) ? 1.0f : -1.0f;                       // visiting body is A / B of the manifold
            ContactState cs = unpack_contact(a4, b4, n4, l4, p4);       // rA = r_self, rB = r_other here
            M3 invIw = m3(zero3(), zero3(), zero3());
            if (gyro) {   // anisotropic inertia only: for R diag(c) R^T = c Id the term Ja x (I^-1 Ja) of solver.cpp:393-397 is exactly zero
                V3 I = xyz(b.aux[self].inert);
                invIw = rot_diag(qmat(quat(ps.rot)), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
            }
            BodySystem sys;
            visit_rows(ps.pos, ps.rot, po.pos, po.rot, sg, __int_as_float(e.w), alpha, pending, biasDual, beta, gyro, invIw, cs, sys);
            // computeConstraint's side effects (manifold.cpp:224-241): written only when they changed something
            float4 nl = pack_lambda(cs);
            if (pending) { ContactLP q; q.l = nl; q.p = pack_penalty(cs); *lp = q; }
            else if (nl.y != l4.y || nl.z != l4.z || nl.w != l4.w) lp->l = nl;
            float4* row = sm.c[buf][t];
            row[0] = make_float4(sys.rl[0], sys.rl[1], sys.rl[2], sys.ra[0]);
            row[1] = make_float4(sys.ra[1], sys.ra[2], sys.ll[0], sys.ll[1]);
            row[2] = make_float4(sys.ll[2], sys.ll[3], sys.ll[4], sys.ll[5]);
            row[3] = make_float4(sys.la[0], sys.la[1], sys.la[2], sys.la[3]);
            row[4] = make_float4(sys.la[4], sys.la[5], sys.la[6], sys.la[7]);
            row[5] = make_float4(sys.la[8], sys.aa[0], sys.aa[1], sys.aa[2]);
            row[6] = make_float4(sys.aa[3], sys.aa[4], sys.aa[5], 0.0f);
        }
        __syncthreads();          // the only barrier of the chunk: buffer `buf` is next written two chunks on, i.e. after the NEXT barrier
        // ---- phase 2: each warp sums the segments that START in its 32 visits (their rows may run on into later warps')
        {
            const int nSegW = __popc(heads);
            int liveCount = vEnd - base; if (liveCount > T) liveCount = T;
            int tailEnd = liveCount;                                  // where this warp's last segment ends: the next head of a later warp
#pragma unroll
            for (int w2 = T / 32 - 1; w2 >= 1; --w2) { unsigned m = sm.headMask[buf][w2]; if (w2 > warp && m) tailEnd = w2 * 32 + __ffs(m) - 1; }
            for (int wi = lane; wi < nSegW * kFlatLanes; wi += 32) {
                int sgi = wi / kFlatLanes, j = wi - sgi * kFlatLanes;
                int lv = sm.segPos[buf][warp][sgi], z = sgi + 1 < nSegW ? sm.segPos[buf][warp][sgi + 1] : tailEnd;
                int idx = sm.segK[buf][warp][sgi];          // the body's row of `sums`, complemented when the segment continues the previous chunk's last
                // a body whose run crosses a chunk boundary is summed in one sequence all the same: the partial sum waits in shared memory
                const bool cont = idx < 0;
                float4 acc = (cont && aligned) ? sm.carry[buf ^ 1][j] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (cont) idx = ~idx;
                auto add = [&](float4 x) {                   // two packed adds (add.rn.f32x2) instead of four scalar ones
                    float2 lo = __fadd2_rn(make_float2(acc.x, acc.y), make_float2(x.x, x.y)), hi = __fadd2_rn(make_float2(acc.z, acc.w), make_float2(x.z, x.w));
                    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
                };
                for (; lv + 4 <= z; lv += 4) {
                    float4 x0 = sm.c[buf][lv][j], x1 = sm.c[buf][lv + 1][j], x2 = sm.c[buf][lv + 2][j], x3 = sm.c[buf][lv + 3][j];
                    add(x0); add(x1); add(x2); add(x3);
                }
                for (; lv < z; ++lv) add(sm.c[buf][lv][j]);
                if (aligned && z == liveCount && sm.nextK[buf] == idx) sm.carry[buf][j] = acc;         // aligned: the run goes on in the block's next chunk
                else if (cont && !aligned) reinterpret_cast<float4*>(carry + (size_t)((base - vBegin) / T) * kSumStride)[j] = acc;   // round-robin: a piece
                else reinterpret_cast<float4*>(sums + (size_t)idx * kSumStride)[j] = acc;
            }
        }
    }
}

// One body per thread: the body's row sums (+ with round-robin chunks the pieces of every later chunk its run reaches, in chunk
// order), inertial terms, Schur solve, pose update.  `order`, `vstart`, `sums` point at the colour's first body.
__global__ void __launch_bounds__(kThreads) primal_solve_flat(BodyView b, ForceView fv, const int* __restrict__ order, const int* __restrict__ vstart, int count,
                                                              int vBegin, int chunkT, const float* __restrict__ sums, const float* __restrict__ carry,
                                                              SolveParams prm, float* dxOut, Diag* diag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = order[k];
    int vs = vstart[k], ve = vstart[k + 1];
    const unsigned long long keep = l2_keep_policy();
    cudaGridDependencySynchronize();          // the visit kernel's sums (programmatic stream serialization, see launch_flat)
    BodyPose self = load_pose_keep(b.pose + i, keep);
    BodyAux aux = b.aux[i];
    float o[kSumStride];
#pragma unroll
    for (int q = 0; q < kSumStride; ++q) o[q] = 0.0f;
    if (ve > vs) {                            // a body no contact visits has no row: nothing was written for it
        const float4* s4 = reinterpret_cast<const float4*>(sums + (size_t)k * kSumStride);
#pragma unroll
        for (int q = 0; q < kSumStride / 4; ++q) { float4 x = __ldcs(s4 + q); o[4 * q] = x.x; o[4 * q + 1] = x.y; o[4 * q + 2] = x.z; o[4 * q + 3] = x.w; }
        if (carry) {
            int c0 = (vs - vBegin) / chunkT, c1 = (ve - 1 - vBegin) / chunkT;
            for (int c = c0 + 1; c <= c1; ++c) {
                const float4* c4 = reinterpret_cast<const float4*>(carry + (size_t)c * kSumStride);
#pragma unroll
                for (int q = 0; q < kSumStride / 4; ++q) { float4 x = c4[q]; o[4 * q] += x.x; o[4 * q + 1] += x.y; o[4 * q + 2] += x.z; o[4 * q + 3] += x.w; }
            }
        }
    }
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

// Body-aligned ranges of one colour's visits: range r of `grid` starts at the first body whose run starts at or after
// the r-th equal share of [vBegin, vEnd) — every body's visits then belong to exactly one block, which sums them in one
// sequence (no partial sums to merge across blocks, the result does not depend on where the colour's visit list is cut).
__global__ void flat_ranges(const int* __restrict__ vstart, int count, int vBegin, int vEnd, int grid, int* __restrict__ range) {
    cudaGridDependencySynchronize();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > grid) return;
    long long target = (long long)vBegin + ((long long)(vEnd - vBegin) * r) / grid;
    int lo = 0, hi = count;                        // first k in [0, count] with vstart[k] >= target (vstart[count] == vEnd)
    while (lo < hi) { int mid = (lo + hi) >> 1; if (vstart[mid] < target) lo = mid + 1; else hi = mid; }
    range[r] = vstart[lo];
}

// Function attributes and occupancy are PER DEVICE (a process may hold worlds on several GPUs): cached by device index.
constexpr int kMaxDevices = 64;
static int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < kMaxDevices) ? dev : 0; }

template <int T, int MINB>
static int flat_resident_blocks() {
    static int cache[kMaxDevices] = {0};
    const int dev = current_device();
    if (!cache[dev]) {
        cudaFuncSetAttribute(primal_visit_flat<T, MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(primal_visit_flat<T, MINB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int sms = 148, per = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, primal_visit_flat<T, MINB, true>, T, 0) != cudaSuccess || per < 1) { cudaGetLastError(); per = 1; }
        if (getenv("AVBD_DEBUG")) fprintf(stderr, "primal_visit_flat<%d,%d>: %d blocks per SM resident (device %d)\n", T, MINB, per, dev);
        cache[dev] = sms * per;
    }
    return cache[dev];
}
static int flat_config() { static int cfg = [] { const char* e = getenv("AVBD_FLAT"); return e ? atoi(e) : 1284; }(); return cfg; }

// The default large-world sweep of one colour: flat visit partition + block solve.  `order` / `vstart` are the WHOLE colour-ordered
// arrays, the colour is their bodies [first, first + count) with visits [vBegin, vEnd); kOf[body] = its position in `order`.
// `sums`: 28 floats per dynamic body.
// AVBD_FLAT (tuning aid) = "<threads per block><blocks per SM>": 1284 (default) 1285 1286 — within 2 % of each other on the 1M-box grid.
int primal_flat_chunk_threads() { return 128; }
// Persistent grid of a colour with nVisits visits = what is actually resident (a block that has to wait for a slot would start
// its share late), at most one block per chunk.
int primal_flat_grid(int nVisits) {
    int nChunks = (nVisits + 127) / 128;
    int resident = flat_config() == 1286 ? flat_resident_blocks<128, 6>() : (flat_config() == 1285 ? flat_resident_blocks<128, 5>() : flat_resident_blocks<128, 4>());
    return nChunks < resident ? nChunks : resident;
}
void launch_flat_ranges(cudaStream_t s, const int* vstart, int first, int count, int vBegin, int vEnd, int grid, int* range) {
    if (grid > 0) launch_dep(flat_ranges, dim3(blocks_of(grid + 1, kThreads)), dim3(kThreads), 0, s, vstart + first, count, vBegin, vEnd, grid, range);
}

template <int T, int MINB>
static void launch_flat(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* order, const int* vstart,
                        const int* kOf, int first, int count, int vBegin, int vEnd, int grid, const int* range, SolveParams prm, float alpha, float biasDual,
                        float* sums, float* carry, float* dxOut, Diag* diag) {
    // Programmatic dependent launch: each kernel of a sweep is launched while its predecessor still runs and blocks at
    // cudaGridDependencySynchronize() until that one has completed — a step has ~160 of these dependent launches, and the
    // launch latency of each would otherwise sit on the critical path.
    if (grid > 0 && range) launch_dep(primal_visit_flat<T, MINB, true>, dim3(grid), dim3(T), 0, s, b, visits, vg, ms, vBegin, vEnd, range, kOf, alpha, biasDual, prm.beta, sums, carry);
    else if (grid > 0) launch_dep(primal_visit_flat<T, MINB, false>, dim3(grid), dim3(T), 0, s, b, visits, vg, ms, vBegin, vEnd, range, kOf, alpha, biasDual, prm.beta, sums, carry);
    const int* orderC = order + first; const int* vstartC = vstart + first; const float* sumsC = sums + (size_t)first * kSumStride;
    const float* carryC = range ? nullptr : carry;                       // body-aligned ranges leave no pieces to add
    launch_dep(primal_solve_flat, dim3(blocks_of(count, kThreads)), dim3(kThreads), 0, s, b, fv, orderC, vstartC, count, vBegin, (int)T, sumsC, carryC, prm, dxOut, diag);
}
// `range` == nullptr: round-robin chunks (+ `carry`: 28 floats per chunk of the colour); else the colour's grid + 1 body-aligned
// block boundaries (launch_flat_ranges, once per graph build).
int launch_primal_flat(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* order, const int* vstart,
                       const int* kOf, int first, int count, int vBegin, int vEnd, int grid, const int* range, SolveParams prm, float alpha, float biasDual,
                       float* sums, float* carry, float* dxOut, Diag* diag) {
#define AVBD_FL(T, M) launch_flat<T, M>(s, b, visits, vg, ms, fv, order, vstart, kOf, first, count, vBegin, vEnd, grid, range, prm, alpha, biasDual, sums, carry, dxOut, diag)
    switch (flat_config()) {
        case 1286: AVBD_FL(128, 6); break;
        case 1285: AVBD_FL(128, 5); break;
        default:   AVBD_FL(128, 4); break;
    }
#undef AVBD_FL
    return grid > 0 ? 2 : 1;
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
