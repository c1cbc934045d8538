// avbd_solve.cu — the per-iteration solver kernels (primal block solve per colour, dual / penalty ramp) and
// their launchers.  This translation unit is compiled WITH FMA contraction: its outputs are held to an FP32
// tolerance against the reference (BASELINE.json north_star), unlike the collision path in avbd_engine.cu.
//
// The reference walks bodies serially (Gauss-Seidel, solver.cpp:344); here bodies of one colour share no
// manifold, so a colour is solved in one launch, one contact visit (computeConstraint + 3 rows) per thread.
//
// Deferred dual.  The dual / penalty-ramp pass of iteration k (solver.cpp:411-430) reads the poses left by sweep k, and
// nothing moves between it and sweep k+1.  When sweep k+1 reaches the FIRST visit of a contact (its other endpoint is
// static or has a higher colour) neither endpoint has moved yet, so that visit sees exactly the poses the dual pass
// would have seen: it applies the pending dual update in registers (same operations, same order), then evaluates the
// primal rows, and writes lambda / penalty back once.  A step therefore runs ONE stand-alone dual pass (after the last
// sweep, fused with the contact diagnostics) instead of `iterations` of them.  alphaDual < 0 = nothing pending (first sweep
// of a step, stage API).
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
            float f = clampf(sp.penalty * C + lamWarm + 0.0f, -FLT_MAX, FLT_MAX);
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
                float k_ = r < 3 ? j.kLin : j.kAng;
                float lamWarm = (k_ == FLT_MAX) ? j.lambda[r] : 0.0f;
                float f = clampf(j.penalty[r] * ev.C[r] + lamWarm + 0.0f, ev.fmin[r], ev.fmax[r]);
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

// One contact visit: computeConstraint for the visiting body, with the pending dual update first when `pending`.
__device__ __forceinline__ void visit_constraint(float4 pa4, float4 qa4, float4 pb4, float4 qb4, float mu, float alpha, bool pending, float alphaDual,
                                                 float beta, ContactState& cs, ContactEval& ev) {
    float sep[3];
    contact_geometry(xyz(pa4), quat(qa4), xyz(pb4), quat(qb4), cs, ev, sep);
    if (pending) {
        contact_limits(pa4.w, pb4.w, mu, alphaDual, sep, cs, ev);
        dual_contact(cs, ev, beta);
    }
    contact_limits(pa4.w, pb4.w, mu, alpha, sep, cs, ev);
}

// Loads of data another CTA may have written earlier in the SAME launch (persistent loop): bypass L1.
template <bool COH> __device__ __forceinline__ float4 ld4(const float4* p) { return COH ? __ldcg(p) : *p; }
template <bool COH> __device__ __forceinline__ BodyPose load_pose(const BodyPose* p) {
    BodyPose r; r.pos = ld4<COH>(&p->pos); r.rot = ld4<COH>(&p->rot); return r;
}
template <bool COH> __device__ __forceinline__ ContactState load_contact_c(const ManifoldSet& ms, int ci) {
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], ld4<COH>(&ms.lp[ci].l), ld4<COH>(&ms.lp[ci].p));
}

// ------------------------------------------------------------------ primal, visit-parallel (the large-world path)
// The visit list is laid out in colour order, so the visits of a tile of BPB consecutive bodies of one colour are ONE
// contiguous run.  The tile walks that run one visit per thread (every lane busy, every lane's gathers independent
// and in flight together — the kernel is latency bound otherwise):
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

template <int BPB, bool COH>
__device__ __forceinline__ void primal_tile_visits(const BodyView& b, const int* __restrict__ vstart, const int4* __restrict__ visits,
                                                   const ManifoldSet& ms, const ForceView& fv, const int* __restrict__ order, int count, int tile,
                                                   const SolveParams& prm, float alpha, float alphaDual, float* dxOut, Diag* diag, PrimalSmem<BPB>& sm) {
    constexpr int L = kThreads / BPB;
    constexpr int CPL = (27 + L - 1) / L;
    const int t = threadIdx.x;
    const int k0 = tile * BPB;
    const int nb = (count - k0) < BPB ? (count - k0) : BPB;
    if (t <= nb) sm.vs[t] = vstart[k0 + t];
    if (t < nb) {
        int i = order[k0 + t];
        sm.body[t] = i;
        BodyPose self = load_pose<COH>(b.pose + i);
        BodyAux aux = b.aux[i];
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
            int4 e = visits[v];
            int ci = e.x; bool isA = (e.z & 1) != 0;
            BodyPose po = load_pose<COH>(b.pose + e.y);
            float4 a4 = ms.cA[ci], b4 = ms.cB[ci], n4 = ms.cN[ci];
            float4 l4 = ld4<COH>(&ms.lp[ci].l), p4 = ld4<COH>(&ms.lp[ci].p);
            int lo = 0, hi = nb;                                  // slot: vs[lo] <= v < vs[lo + 1]
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sm.vs[mid] <= v) lo = mid; else hi = mid; }
            float4 sp = sm.pos[lo], sr = sm.rot[lo];
            ContactState cs = unpack_contact(a4, b4, n4, l4, p4);
            ContactEval ev;
            float mu = __int_as_float(e.w);
            bool pending = alphaDual >= 0.0f && (e.z & 4) != 0;
            {
                float4 pa4 = isA ? sp : po.pos, qa4 = isA ? sr : po.rot, pb4 = isA ? po.pos : sp, qb4 = isA ? po.rot : sr;
                visit_constraint(pa4, qa4, pb4, qb4, mu, alpha, pending, alphaDual, prm.beta, cs, ev);
            }
            bool gyro = sm.inert[lo].w != 0.0f;
            M3 invIw;
            if (gyro) {
                const float* q = sm.inv[lo];
                invIw = m3(mk3(q[0], q[1], q[2]), mk3(q[1], q[3], q[4]), mk3(q[2], q[4], q[5]));
            } else {
                invIw = m3(zero3(), zero3(), zero3());
            }
            BodySystem sys;
            contact_system(sys, cs, ev, isA, gyro, invIw);
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

template <int BPB, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) primal_visits(BodyView b, const int* __restrict__ vstart, const int4* __restrict__ visits,
                                                                ManifoldSet ms, ForceView fv, const int* __restrict__ order, int count,
                                                                SolveParams prm, float alpha, float alphaDual, float* dxOut, Diag* diag) {
    __shared__ PrimalSmem<BPB> sm;
    primal_tile_visits<BPB, false>(b, vstart, visits, ms, fv, order, count, blockIdx.x, prm, alpha, alphaDual, dxOut, diag, sm);
}

// ------------------------------------------------------------------ primal, split (visit sums -> block solve)
// Keeping the 6x6 solve inside the visit kernel leaves 1/4 .. 1/8 of a block's threads running a long serial chain while
// the block pins its registers and shared memory.  The large-world path therefore runs two lane-dense kernels per colour:
//   primal_visit_sums  one visit per thread -> per-body sums of the 27 row contributions (shared-memory transpose +
//                      in-order reduction), written to `sums` in colour order (28 floats per body)
//   primal_solve       one body per thread: inertial terms + sums -> Schur solve -> pose update
// Visit entry (built with the graph): {contact id, other body, (self body << 3) | first-visit << 2 | anisotropic-inertia << 1 | body-is-A, mu}.
template <int BPB>
struct VisitSmem {
    float c[27][kThreads + 1];
    int vs[BPB + 1];
};
constexpr int kSumStride = 28;

// MODE 0 is the product.  MODE 1 / 2 are measurement aids behind avbd_debug_time_primal (never on the step path): 1 keeps the
// memory accesses and drops the row math (the kernel's memory-system floor), 2 keeps the math and makes every index
// sequential (its instruction-issue floor); neither writes solver state.
template <int BPB, int MINB, int MODE = 0>
__global__ void __launch_bounds__(kThreads, MINB) primal_visit_sums(BodyView b, const int* __restrict__ vstart, const int4* __restrict__ visits, VisitGeom vg,
                                                                    ManifoldSet ms, int count, float alpha, float alphaDual, float beta, float* __restrict__ sums,
                                                                    int nContacts = 0) {
    constexpr int L = kThreads / BPB;
    constexpr int CPL = (27 + L - 1) / L;
    __shared__ VisitSmem<BPB> sm;
    const int t = threadIdx.x;
    const int k0 = blockIdx.x * BPB;
    const int nb = (count - k0) < BPB ? (count - k0) : BPB;
    if (t <= nb) sm.vs[t] = vstart[k0 + t];
    const int v0 = vstart[k0], v1 = vstart[k0 + nb];
    float acc[CPL];
#pragma unroll
    for (int m = 0; m < CPL; ++m) acc[m] = 0.0f;
    const int rs = t / L, rj = t % L;
    const unsigned long long keep = l2_keep_policy();
    for (int base = v0; base < v1; base += kThreads) {
        int v = base + t;
        if (v < v1) {
            int4 e = __ldcs(visits + v);
            int ci = e.x, self = e.z >> 3; bool isA = (e.z & 1) != 0, gyro = (e.z & 2) != 0, pending = alphaDual >= 0.0f && (e.z & 4) != 0;
            if (MODE == 2) { ci = v % nContacts; self = v % b.n; e.y = (v + 1) % b.n; }
            BodyPose ps = load_pose_keep(b.pose + self, keep);
            BodyPose po = load_pose_keep(b.pose + e.y, keep);
            // the visit entry and its copy of the contact geometry stream past once per sweep, fully coalesced: evict-first, so
            // they do not push the poses and the lambda / penalty records (the gathered, re-used data) out of L2
            float4 a4 = __ldcs(vg.a + v), b4 = __ldcs(vg.b + v), n4 = __ldcs(vg.n + v);
            ContactLP* lp = ms.lp + ci;
            float4 l4 = lp->l, p4 = lp->p;
            V3 pos = xyz(ps.pos); Q4 rot = quat(ps.rot);
            ContactState cs = unpack_contact(a4, b4, n4, l4, p4);
            ContactEval ev;
            float mu = __int_as_float(e.w);
            if (MODE == 1) {
                float q = ps.pos.x + ps.rot.y + po.pos.z + po.rot.w + a4.x + b4.y + n4.z + l4.w + p4.x + mu;
#pragma unroll
                for (int k = 0; k < 27; ++k) sm.c[k][t] = q;
            } else {
            {   // one call with the operands swapped by selects: an if/else duplicates the code and mixed warps run both arms
                float4 pa4 = isA ? ps.pos : po.pos, qa4 = isA ? ps.rot : po.rot, pb4 = isA ? po.pos : ps.pos, qb4 = isA ? po.rot : ps.rot;
                visit_constraint(pa4, qa4, pb4, qb4, mu, alpha, pending, alphaDual, beta, cs, ev);
            }
            M3 invIw = m3(zero3(), zero3(), zero3());
            if (gyro) {   // anisotropic inertia only: for R diag(c) R^T = c Id the term Ja x (I^-1 Ja) of solver.cpp:393-397 is exactly zero
                V3 I = xyz(b.aux[self].inert);
                invIw = rot_diag(qmat(rot), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
            }
            BodySystem sys;
            contact_system(sys, cs, ev, isA, gyro, invIw);
            // computeConstraint's side effects (manifold.cpp:224-241): written only when they changed something
            float4 nl = pack_lambda(cs);
            if (MODE == 0) {
                if (pending) { ContactLP q; q.l = nl; q.p = pack_penalty(cs); *lp = q; }
                else if (nl.y != l4.y || nl.z != l4.z || nl.w != l4.w) lp->l = nl;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.c[k][t] = sys.rl[k]; sm.c[3 + k][t] = sys.ra[k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) { sm.c[6 + k][t] = sys.ll[k]; sm.c[21 + k][t] = sys.aa[k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) sm.c[12 + k][t] = sys.la[k];
            }
        }
        __syncthreads();
        if (rs < nb) {
            int rlo = sm.vs[rs], rhi = sm.vs[rs + 1];
            int a = (rlo > base ? rlo : base) - base;
            int z = (rhi < base + kThreads ? rhi : base + kThreads) - base;
            for (int lv = a; lv < z; ++lv) {
#pragma unroll
                for (int m = 0; m < CPL; ++m) { int k = rj + m * L; if (k < 27) acc[m] += sm.c[k][lv]; }
            }
        }
        if (base + kThreads < v1) __syncthreads();
    }
    if (rs < nb) {
        float* o = sums + (size_t)(k0 + rs) * kSumStride;
#pragma unroll
        for (int m = 0; m < CPL; ++m) { int k = rj + m * L; if (k < 27) o[k] = acc[m]; }
    }
}

__global__ void __launch_bounds__(kThreads) primal_solve(BodyView b, ForceView fv, const int* __restrict__ order, int count,
                                                         const float* __restrict__ sums, SolveParams prm, float* dxOut, Diag* diag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = order[k];
    const unsigned long long keep = l2_keep_policy();
    BodyPose self = load_pose_keep(b.pose + i, keep);
    BodyAux aux = b.aux[i];
    const float4* s4 = reinterpret_cast<const float4*>(sums + (size_t)k * kSumStride);
    float o[kSumStride];
#pragma unroll
    for (int q = 0; q < kSumStride / 4; ++q) { float4 x = s4[q]; o[4 * q] = x.x; o[4 * q + 1] = x.y; o[4 * q + 2] = x.z; o[4 * q + 3] = x.w; }
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
    contact_geometry(xyz(pa.pos), quat(pa.rot), xyz(pb.pos), quat(pb.rot), cs, ev, sep);
    for (int r = 0; r < reps; ++r) {
        contact_limits(pa.pos.w, pb.pos.w, __int_as_float(h.w), alpha, sep, cs, ev);
        dual_contact(cs, ev, prm.beta);
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
// scope) instead of a kernel boundary, and each phase is the visit-parallel tile of the large-world path (one contact
// visit per thread, in-order shared-memory sums, block solve).  Data other CTAs write between barriers (poses, lambda,
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
                                                                  Diag* diag, bool contactDiag, bool anyUnvisited) {
    __shared__ PrimalSmem<BPB> sm;
    const int rank = (int)cluster_rank(), nCta = (int)cluster_size();       // the grid is one cluster
    int total = prm.iterations + (prm.postStabilize ? 1 : 0);
    float alphaDual = -1.0f;                                    // dual pass of the previous iteration still to apply (deferred dual)
    for (int it = 0; it < total; ++it) {
        float alpha = prm.postStabilize ? (it < prm.iterations ? 1.0f : 0.0f) : prm.alpha;      // solver.cpp:340-342
        for (int c = 0; c < nColours; ++c) {
            int2 r = colRange[c];
            int count = r.y - r.x;
            for (int tile = rank; tile * BPB < count; tile += nCta) {
                primal_tile_visits<BPB, true>(b, visitStart + r.x, visits, ms, fv, order + r.x, count, tile, prm, alpha, alphaDual, nullptr, diag, sm);
                __syncthreads();
            }
            cluster_barrier();
        }
        alphaDual = it < prm.iterations ? alpha : -1.0f;
    }
    // what the sweeps could not apply: the last iteration's dual pass (nothing moves after it, so the contact diagnostics
    // are reduced from the same registers), or — when postStabilize's extra sweep already applied it — only the contacts
    // no dynamic body visits
    bool lastPending = alphaDual >= 0.0f;
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
            float k_ = r < 3 ? j.kLin : j.kAng;
            if (k_ != FLT_MAX) continue;
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

template <int BPB, int MINB>
static void launch_primal_visits(cudaStream_t s, BodyView b, const int* vstart, const int4* visits, ManifoldSet ms, ForceView fv,
                                 const int* order, int count, SolveParams prm, float alpha, float alphaDual, float* dxOut, Diag* diag) {
    primal_visits<BPB, MINB><<<blocks_of(count, BPB), kThreads, 0, s>>>(b, vstart, visits, ms, fv, order, count, prm, alpha, alphaDual, dxOut, diag);
}

template <int BPB, int MINB>
static void launch_split(cudaStream_t s, BodyView b, const int* vstart, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv,
                         const int* order, int count, SolveParams prm, float alpha, float alphaDual, float* sums, float* dxOut, Diag* diag) {
    primal_visit_sums<BPB, MINB><<<blocks_of(count, BPB), kThreads, 0, s>>>(b, vstart, visits, vg, ms, count, alpha, alphaDual, prm.beta, sums);
    primal_solve<<<blocks_of(count, kThreads), kThreads, 0, s>>>(b, fv, order, count, sums, prm, dxOut, diag);
}

// Default: the split path (visit sums + block solve), bodies per tile chosen so a tile's visits fill the block once.
// AVBD_PRIMAL_VARIANT (tuning aid): "s<BPB>[m<MINB>]" split path with a forced tile size (s16 s28 s64; m3 m4);
// "v<BPB>[m<MINB>]" the single-kernel visit-parallel tile.  Returns the number of kernels launched.
int launch_primal(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv,
                  const int* order, int count, float avgVisits, SolveParams prm, float alpha, float alphaDual, float* sums, float* dxOut, Diag* diag) {
    static int variant = 0, minb = 3; static char kind = 's';
    static bool init = [] {
        const char* e = getenv("AVBD_PRIMAL_VARIANT");
        if (!e) return true;
        if (e[0] == 'v' || e[0] == 's') {
            kind = e[0];
            variant = atoi(e + 1);
            for (const char* p = e; *p; ++p) if (*p == 'm') minb = atoi(p + 1);
        }
        return true;
    }();
    (void)init;
#define AVBD_VV(B, M) launch_primal_visits<B, M>(s, b, visitStart, visits, ms, fv, order, count, prm, alpha, alphaDual, dxOut, diag)
#define AVBD_SV(B, M) launch_split<B, M>(s, b, visitStart, visits, vg, ms, fv, order, count, prm, alpha, alphaDual, sums, dxOut, diag)
    int bpb = variant > 0 ? variant : (avgVisits <= 3.7f ? 64 : (avgVisits <= 9.0f ? 28 : (avgVisits <= 16.0f ? 16 : 8)));
    if (kind == 'v') {
        if (bpb >= 64) { if (minb == 4) AVBD_VV(64, 4); else AVBD_VV(64, 3); }
        else if (bpb >= 28) { if (minb == 4) AVBD_VV(28, 4); else AVBD_VV(28, 3); }
        else if (bpb >= 16) AVBD_VV(16, 3); else AVBD_VV(8, 3);
        return 1;
    }
    if (bpb >= 64) { if (minb == 4) AVBD_SV(64, 4); else AVBD_SV(64, 3); }
    else if (bpb >= 28) { if (minb == 4) AVBD_SV(28, 4); else AVBD_SV(28, 3); }
    else if (bpb >= 16) { if (minb == 4) AVBD_SV(16, 4); else AVBD_SV(16, 3); }
    else AVBD_SV(8, 3);
    return 2;
#undef AVBD_VV
#undef AVBD_SV
}
void launch_primal_experiment(cudaStream_t s, int mode, BodyView b, const int* vstart, const int4* visits, VisitGeom vg, ManifoldSet ms, int count, float alpha,
                              float* sums, int nContacts) {
    if (mode == 1) primal_visit_sums<28, 3, 1><<<blocks_of(count, 28), kThreads, 0, s>>>(b, vstart, visits, vg, ms, count, alpha, -1.0f, 0.0f, sums, nContacts);
    else if (mode == 2) primal_visit_sums<28, 3, 2><<<blocks_of(count, 28), kThreads, 0, s>>>(b, vstart, visits, vg, ms, count, alpha, -1.0f, 0.0f, sums, nContacts);
    else primal_visit_sums<28, 3, 0><<<blocks_of(count, 28), kThreads, 0, s>>>(b, vstart, visits, vg, ms, count, alpha, -1.0f, 0.0f, sums, nContacts);
}
bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, int nContacts, SolveParams prm,
                       Diag* diag, bool contactDiag, bool anyUnvisited) {
    constexpr int BPB = kClusterBodiesPerTile;
    static int maxCluster = [] {
        // 16 CTAs need the non-portable opt-in; fall back to the portable 8 if the device refuses it
        if (cudaFuncSetAttribute(solve_loop_cluster<BPB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 8; }
        if (const char* e = getenv("AVBD_CLUSTER_MAX")) { int v = atoi(e); if (v >= 1 && v <= 16) return v; }
        return 16;
    }();
    int want = blocks_of(maxColourCount, BPB);
    int wantDual = blocks_of(nContacts, kThreads);
    int need = want > wantDual ? want : wantDual;
    int nCta = 1;
    while (nCta < need && nCta < maxCluster) nCta <<= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nCta); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nCta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, solve_loop_cluster<BPB>, b, visitStart, visits, ms, fv, order, colRange, nColours, nContacts, prm, diag, contactDiag, anyUnvisited);
    if (e != cudaSuccess && nCta > 8) {          // a 16-CTA cluster may not be placeable (MIG slices, busy GPCs): retry with the portable size
        cudaGetLastError();
        maxCluster = 8;
        attr[0].val.clusterDim.x = 8; cfg.gridDim = dim3(8);
        e = cudaLaunchKernelEx(&cfg, solve_loop_cluster<BPB>, b, visitStart, visits, ms, fv, order, colRange, nColours, nContacts, prm, diag, contactDiag, anyUnvisited);
    }
    return e == cudaSuccess;
}

void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, int nContacts, SolveParams prm, float alpha, int unvisitedReps, bool onlyUnvisited, Diag* diag) {
    if (diag) dual_contacts<true><<<blocks_of(nContacts, kThreads), kThreads, 0, s>>>(b, ms, nContacts, prm, alpha, unvisitedReps, onlyUnvisited, diag);
    else      dual_contacts<false><<<blocks_of(nContacts, kThreads), kThreads, 0, s>>>(b, ms, nContacts, prm, alpha, unvisitedReps, onlyUnvisited, nullptr);
}
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm) {
    dual_user_forces<<<blocks_of(fv.nJoints + fv.nSprings, kThreads), kThreads, 0, s>>>(b, fv, prm);
}
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6) {
    solve6_batch<<<blocks_of(n, 128), 128, 0, s>>>(lhs36, rhs6, n, out6);
}

} // namespace avbd
