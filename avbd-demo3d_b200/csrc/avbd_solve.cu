// avbd_solve.cu — the per-iteration solver kernels (primal block solve per colour, dual / penalty ramp) and
// their launchers.  This translation unit is compiled WITH FMA contraction: its outputs are held to an FP32
// tolerance against the reference (BASELINE.json north_star), unlike the collision path in avbd_engine.cu.
//
// The reference walks bodies serially (Gauss-Seidel, solver.cpp:344); here bodies of one colour share no
// manifold, so a colour is solved in one launch with LPB lanes cooperating on each body (one contact visit =
// computeConstraint + 3 rows per lane).
#include <cstdlib>
#include "avbd_launch.h"
#include "avbd_body.cuh"
#include "avbd_forces.cuh"

namespace avbd {

__device__ __forceinline__ ContactState load_contact(const ManifoldSet& ms, int ci) {
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], ms.cL[ci], ms.cP[ci]);
}

// ------------------------------------------------------------------ primal
__device__ __forceinline__ float group_sum(float x, int width) {
    for (int off = width >> 1; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off, width);
    return x;
}
__device__ __forceinline__ void reduce_system(BodySystem& s, int width) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { s.rl[k] = group_sum(s.rl[k], width); s.ra[k] = group_sum(s.ra[k], width); }
#pragma unroll
    for (int k = 0; k < 6; ++k) { s.ll[k] = group_sum(s.ll[k], width); s.aa[k] = group_sum(s.aa[k], width); }
#pragma unroll
    for (int k = 0; k < 9; ++k) s.la[k] = group_sum(s.la[k], width);
}

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

// Loads of data another CTA may have written earlier in the SAME launch (persistent loop): bypass L1.
template <bool COH> __device__ __forceinline__ float4 ld4(const float4* p) { return COH ? __ldcg(p) : *p; }
template <bool COH> __device__ __forceinline__ BodyPose load_pose(const BodyPose* p) {
    BodyPose r; r.pos = ld4<COH>(&p->pos); r.rot = ld4<COH>(&p->rot); return r;
}
template <bool COH> __device__ __forceinline__ ContactState load_contact_c(const ManifoldSet& ms, int ci) {
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], ld4<COH>(ms.cL + ci), ld4<COH>(ms.cP + ci));
}

// One tile (kThreads/LPB bodies of one colour) of the primal sweep (solver.cpp:344-409), in two phases so both are
// lane-dense:
//   phase 1  LPB lanes per body walk the body's run of contact visits (lane l takes visits l, l+LPB, ...):
//            computeConstraint + 3 rows each, then a shuffle reduction leaves the 27 sums in lane 0, which
//            parks them in shared memory;
//   phase 2  one lane per body (the first kThreads/LPB threads = full warps): inertial terms, Schur solve,
//            pose update.
template <int LPB, bool COH>
__device__ __forceinline__ void primal_tile(const BodyView& b, const int* __restrict__ visitStart, const int4* __restrict__ visits,
                                            const ManifoldSet& ms, const ForceView& fv, const int* __restrict__ order, int count, int tile,
                                            const SolveParams& prm, float alpha, float* dxOut, Diag* diag, float* sSys) {
    constexpr int BPB = kThreads / LPB;
    int g = threadIdx.x / LPB, lane = threadIdx.x % LPB;
    int gid = tile * BPB + g;
    bool live = gid < count;
    BodySystem sys; sys.clear();
    if (live) {
        int i = order[gid];
        int v0 = visitStart[i], v1 = visitStart[i + 1];
        bool userForces = fv.adjStart != nullptr && lane == 0 && fv.adjStart[i + 1] > fv.adjStart[i];
        if (v0 + lane < v1 || userForces) {
            BodyPose self = load_pose<COH>(b.pose + i);
            V3 pos = xyz(self.pos); Q4 rot = quat(self.rot);
            float invMassSelf = self.pos.w;
            V3 I = xyz(b.aux[i].inert);
            M3 invIw = rot_diag(qmat(rot), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
            for (int v = v0 + lane; v < v1; v += LPB) {
                int4 e = visits[v];
                int ci = e.x; bool isA = e.z != 0;
                BodyPose po = load_pose<COH>(b.pose + e.y);
                ContactState cs = load_contact_c<COH>(ms, ci);
                ContactEval ev;
                float mu = __int_as_float(e.w);
                if (isA) contact_constraint(pos, rot, invMassSelf, xyz(po.pos), quat(po.rot), po.pos.w, mu, alpha, cs, ev);
                else     contact_constraint(xyz(po.pos), quat(po.rot), po.pos.w, pos, rot, invMassSelf, mu, alpha, cs, ev);
                accumulate_contact(sys, cs, ev, isA, invIw);
                ms.cL[ci] = pack_lambda(cs);      // computeConstraint's side effects (manifold.cpp:224-241)
            }
            if (userForces) accumulate_user_forces(sys, fv, b.pose, i, pos, rot, invIw);
        }
    }
    if (LPB > 1) reduce_system(sys, LPB);
    if (lane == 0) {
        float* o = sSys + g * 27;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = sys.rl[k]; o[3 + k] = sys.ra[k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { o[6 + k] = sys.ll[k]; o[21 + k] = sys.aa[k]; }
#pragma unroll
        for (int k = 0; k < 9; ++k) o[12 + k] = sys.la[k];
    }
    __syncthreads();
    gid = tile * BPB + threadIdx.x;
    if (threadIdx.x < BPB && gid < count) {
        int i = order[gid];
        BodyPose self = load_pose<COH>(b.pose + i);
        BodyAux aux = b.aux[i];
        V3 pos = xyz(self.pos); Q4 rot = quat(self.rot);
        BodySystem own; M3 invIw;
        body_self_system(pos, rot, aux, prm.dt, own, invIw);
        const float* o = sSys + threadIdx.x * 27;
#pragma unroll
        for (int k = 0; k < 3; ++k) { own.rl[k] += o[k]; own.ra[k] += o[3 + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { own.ll[k] += o[6 + k]; own.aa[k] += o[21 + k]; }
#pragma unroll
        for (int k = 0; k < 9; ++k) own.la[k] += o[12 + k];
        V3 dl, da;
        solve_body_system(own, dl, da);
        int ev = apply_body_update(pos, rot, dl, da);
        BodyPose out; out.pos = f4(pos, self.pos.w); out.rot = f4(rot);
        b.pose[i] = out;
        if (dxOut) { float* d = dxOut + 6 * i; d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = da.x; d[4] = da.y; d[5] = da.z; }
        if (ev) atomicAdd(&diag[b.worldId[i]].nanEvents, ev);
    }
}

template <int LPB, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) primal_colour(BodyView b, const int* __restrict__ visitStart, const int4* __restrict__ visits,
                                                                ManifoldSet ms, ForceView fv, const int* __restrict__ order, int count,
                                                                SolveParams prm, float alpha, float* dxOut, Diag* diag) {
    __shared__ float sSys[(kThreads / LPB) * 27];    // stride 27 is odd: conflict-free in phase 2
    primal_tile<LPB, false>(b, visitStart, visits, ms, fv, order, count, blockIdx.x, prm, alpha, dxOut, diag, sSys);
}

// ------------------------------------------------------------------ dual
// One LIVE contact (solver.cpp:411-430 for manifold rows).
template <bool COH>
__device__ __forceinline__ void dual_one(const BodyView& b, const ManifoldSet& ms, int ci, const SolveParams& prm, float alpha) {
    int4 h = ms.hdr[ci >> 2];
    BodyPose pa = load_pose<COH>(b.pose + h.x), pb = load_pose<COH>(b.pose + h.y);
    ContactState cs = load_contact_c<COH>(ms, ci);
    ContactEval ev;
    contact_constraint(xyz(pa.pos), quat(pa.rot), pa.pos.w, xyz(pb.pos), quat(pb.rot), pb.pos.w, __int_as_float(h.w), alpha, cs, ev);
    dual_contact(cs, ev, prm.beta);
    ms.cL[ci] = pack_lambda(cs);
    ms.cP[ci] = pack_penalty(cs);
}

__global__ void __launch_bounds__(kThreads) dual_contacts(BodyView b, ManifoldSet ms, const int* __restrict__ contactList, int nContacts,
                                                          SolveParams prm, float alpha) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nContacts) dual_one<false>(b, ms, contactList[t], prm, alpha);
}

// ------------------------------------------------------------------ persistent iteration loop (small worlds)
// A Stress1000-sized world is launch-latency bound: iterations x (colours + dual) dependent launches of a few
// hundred threads each.  This kernel runs the WHOLE loop of solver.cpp:340-431 in one cooperative launch; colours
// and the dual pass are separated by a grid-wide barrier instead of a kernel boundary.  Data other CTAs write
// between barriers (poses, lambda, penalty) is read with ld.global.cg.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <int LPB>
__global__ void __launch_bounds__(kThreads, 2) solve_loop_persistent(BodyView b, const int* __restrict__ visitStart, const int4* __restrict__ visits,
                                                                     ManifoldSet ms, ForceView fv, const int* __restrict__ order,
                                                                     const int2* __restrict__ colRange, int nColours,
                                                                     const int* __restrict__ contactList, int nContacts, SolveParams prm,
                                                                     Diag* diag, unsigned* barrier) {
    constexpr int BPB = kThreads / LPB;
    __shared__ float sSys[BPB * 27];
    unsigned target = 0;
    int total = prm.iterations + (prm.postStabilize ? 1 : 0);
    for (int it = 0; it < total; ++it) {
        float alpha = prm.postStabilize ? (it < prm.iterations ? 1.0f : 0.0f) : prm.alpha;      // solver.cpp:340-342
        for (int c = 0; c < nColours; ++c) {
            int2 r = colRange[c];
            int count = r.y - r.x;
            for (int tile = blockIdx.x; tile * BPB < count; tile += gridDim.x) {
                primal_tile<LPB, true>(b, visitStart, visits, ms, fv, order + r.x, count, tile, prm, alpha, nullptr, diag, sSys);
                __syncthreads();
            }
            grid_barrier(barrier, target);
        }
        if (it < prm.iterations) {
            for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nContacts; t += gridDim.x * blockDim.x)
                dual_one<true>(b, ms, contactList[t], prm, alpha);
            grid_barrier(barrier, target);
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

template <int LPB, int MINB>
static void launch_primal_variant(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv,
                                  const int* order, int count, SolveParams prm, float alpha, float* dxOut, Diag* diag) {
    primal_colour<LPB, MINB><<<blocks_of(count, kThreads / LPB), kThreads, 0, s>>>(b, visitStart, visits, ms, fv, order, count, prm, alpha, dxOut, diag);
}

// AVBD_PRIMAL_VARIANT (tuning aid): "<lanes per body><min blocks per SM>": 43 (default; measured best on the 1M grid,
// profiles/README.md), 82, 83, 42, 44, 23.
void launch_primal(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv,
                   const int* order, int count, SolveParams prm, float alpha, float* dxOut, Diag* diag) {
    static int variant = [] { const char* e = getenv("AVBD_PRIMAL_VARIANT"); return e ? atoi(e) : kLanesPerBody * 10 + 3; }();
#define AVBD_PV(L, M) case L * 10 + M: launch_primal_variant<L, M>(s, b, visitStart, visits, ms, fv, order, count, prm, alpha, dxOut, diag); break;
    switch (variant) {
        AVBD_PV(8, 2) AVBD_PV(8, 3) AVBD_PV(4, 2) AVBD_PV(4, 4) AVBD_PV(2, 3)
        default: launch_primal_variant<4, 3>(s, b, visitStart, visits, ms, fv, order, count, prm, alpha, dxOut, diag); break;
    }
#undef AVBD_PV
}
bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, const int* contactList, int nContacts, SolveParams prm,
                       Diag* diag, unsigned* barrier) {
    constexpr int LPB = kLanesPerBody;
    static int maxBlocks = [] {
        int dev = 0, sms = 0, perSm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, solve_loop_persistent<LPB>, kThreads, 0);
        return sms * (perSm < 1 ? 1 : perSm);
    }();
    int want = blocks_of(maxColourCount, kThreads / LPB);
    int wantDual = blocks_of(nContacts, kThreads);
    int grid = want > wantDual ? want : wantDual;
    if (grid > maxBlocks) grid = maxBlocks;
    if (grid < 1) grid = 1;
    cudaMemsetAsync(barrier, 0, sizeof(unsigned), s);
    void* args[] = {&b, &visitStart, &visits, &ms, &fv, &order, &colRange, &nColours, &contactList, &nContacts, &prm, &diag, &barrier};
    return cudaLaunchCooperativeKernel((void*)solve_loop_persistent<LPB>, dim3(grid), dim3(kThreads), args, 0, s) == cudaSuccess;
}

void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, const int* contactList, int nContacts, SolveParams prm, float alpha) {
    dual_contacts<<<blocks_of(nContacts, kThreads), kThreads, 0, s>>>(b, ms, contactList, nContacts, prm, alpha);
}
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm) {
    dual_user_forces<<<blocks_of(fv.nJoints + fv.nSprings, kThreads), kThreads, 0, s>>>(b, fv, prm);
}
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6) {
    solve6_batch<<<blocks_of(n, 128), 128, 0, s>>>(lhs36, rhs6, n, out6);
}

} // namespace avbd
