// avbd_kernels_solve.cuh — body/manifold graph (adjacency ranges, greedy
// colouring) and the per-iteration solver kernels: per-colour primal block
// solve, dual / penalty ramp, predict, velocity recovery, diagnostics.
//
// The reference walks bodies serially (Gauss-Seidel, solver.cpp:344); here
// bodies of one colour share no manifold, so a colour is solved in one launch
// with LPB lanes cooperating on each body (one contact = 3 rows per lane).
#pragma once
#include "avbd_kernels_collide.cuh"

namespace avbd {

// ------------------------------------------------------------------ adjacency
// Manifolds are sorted by (A,B): a body's "I am A" manifolds are one contiguous
// run [x,y).  Its "I am B" manifolds are a run [z,w) of bList (manifold ids
// stably sorted by B).  adjRange must be zeroed before these two kernels.
__global__ void adj_a_ranges(const int4* hdr, int nM, const int* flags, int nBodies, int4* adjRange, unsigned* bKey, int* bVal) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nM) return;
    int4 h = hdr[m];
    if (m == 0 || hdr[m - 1].x != h.x) adjRange[h.x].x = m;
    if (m == nM - 1 || hdr[m + 1].x != h.x) adjRange[h.x].y = m + 1;
    bKey[m] = (flags[h.y] & kDynamic) ? (unsigned)h.y : (unsigned)nBodies;   // static B never solves: park at the end
    bVal[m] = m;
}
__global__ void adj_b_ranges(const unsigned* bKeySorted, int nM, int nBodies, int4* adjRange) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nM) return;
    unsigned k = bKeySorted[t];
    if (k >= (unsigned)nBodies) return;
    if (t == 0 || bKeySorted[t - 1] != k) adjRange[k].z = t;
    if (t == nM - 1 || bKeySorted[t + 1] != k) adjRange[k].w = t + 1;
}

// User forces (joints / springs) per body: static CSR built on the host at upload.
struct ForceView {
    JointRec* joints; int nJoints;
    SpringRec* springs; int nSprings;
    const int* adjStart; const int* adj;    // entry = index*4 + type*2 + isA ; type 0 joint, 1 spring
};

// ------------------------------------------------------------------ colouring
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
// Priority is a function of the WORLD-LOCAL index only, so a world colours the
// same way wherever it sits in a batch (ensemble runs are partition-invariant).
__device__ __forceinline__ bool outranks(int localA, int localB) {
    unsigned ha = mix32((unsigned)localA + 0x9e3779b9u), hb = mix32((unsigned)localB + 0x9e3779b9u);
    return ha != hb ? ha > hb : localA > localB;
}

__global__ void colour_init(const int* flags, int n, int* colour) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) colour[i] = (flags[i] & kDynamic) ? -1 : -2;
}

// One Jones-Plassmann round: a body takes the smallest colour unused by its
// neighbours once every higher-priority neighbour is coloured.  The result is
// the sequential greedy colouring in priority order, independent of timing.
__global__ void colour_round(const int* dynList, int nDyn, const int4* adjRange, const int* bList, const int4* hdr,
                             ForceView fv, const int* localIdx, volatile int* colour, Counters* cnt) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    int i = dynList[t];
    if (colour[i] >= 0) return;
    int li = localIdx[i];
    unsigned long long used = 0ull;
    bool ready = true;
    int4 rg = adjRange[i];
    auto visit = [&](int other) {
        if (other < 0) return;
        int co = colour[other];
        if (co >= 0) used |= 1ull << co;
        else if (co == -1 && outranks(localIdx[other], li)) ready = false;
    };
    for (int m = rg.x; m < rg.y && ready; ++m) visit(hdr[m].y);
    for (int k = rg.z; k < rg.w && ready; ++k) visit(hdr[bList[k]].x);
    if (fv.adjStart) {
        for (int k = fv.adjStart[i]; k < fv.adjStart[i + 1] && ready; ++k) {
            int e = fv.adj[k]; int idx = e >> 2; bool isA = e & 1;
            int other = (e & 2) ? (isA ? fv.springs[idx].b : fv.springs[idx].a) : (isA ? fv.joints[idx].b : fv.joints[idx].a);
            visit(other);
        }
    }
    if (ready) {
        int c = __ffsll((long long)~used) - 1;
        if (c < 0) { c = 63; atomicOr(&cnt->overflow, 4); }
        colour[i] = c;
    } else {
        cg::coalesced_group grp = cg::coalesced_threads();
        if (grp.thread_rank() == 0) atomicAdd(&cnt->nUncoloured, (int)grp.size());
    }
}

__global__ void colour_keys(const int* dynList, int nDyn, const int* colour, unsigned* key, int* val) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    int i = dynList[t];
    key[t] = (unsigned)colour[i];
    val[t] = i;
}
// colourRange[c] = {first, last+1} in the colour-sorted body order; must be zeroed first.
__global__ void colour_bounds(const unsigned* keySorted, int nDyn, int2* colourRange, Counters* cnt) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    unsigned c = keySorted[t];
    if (t == 0 || keySorted[t - 1] != c) colourRange[c].x = t;
    if (t == nDyn - 1 || keySorted[t + 1] != c) colourRange[c].y = t + 1;
    if (t == nDyn - 1) cnt->nColours = (int)c + 1;
}

// ------------------------------------------------------------------ predict / warm-start decay of user forces
// Diagnostics are kept per world (an ensemble batch reports each world separately).  Lanes of a warp that
// belong to the same world combine first (match_any + masked reduce), then one atomic per (warp, world).
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));       // non-negative floats order like their bit patterns
}
struct WorldGroup {
    unsigned peers; bool leader;
    __device__ __forceinline__ WorldGroup(int world) {
        peers = __match_any_sync(0xffffffffu, world);
        leader = (__ffs(peers) - 1) == (int)(threadIdx.x & 31);
    }
    __device__ __forceinline__ float max_nonneg(float v) const { return __uint_as_float(__reduce_max_sync(peers, __float_as_uint(v))); }
    __device__ __forceinline__ int sum(int v) const { return __reduce_add_sync(peers, v); }
};

__global__ void predict_bodies(BodyView b, SolveParams prm, Diag* diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int dyn = 0, ev = 0, world = -1;
    if (i < b.n) {
        world = b.worldId[i];
        BodyPose pose = b.pose[i]; BodyVel vel = b.vel[i]; BodyAux aux = b.aux[i]; BodyInit init;
        float4 pl = b.prevLin[i];
        ev = predict_body(pose, vel, pl, aux, init, prm);
        b.pose[i] = pose; b.vel[i] = vel; b.init[i] = init;
        b.aux[i].posI = aux.posI; b.aux[i].rotI = aux.rotI;
        dyn = aux.mass.y > 0.0f ? 1 : 0;
    }
    WorldGroup wg(world);
    dyn = wg.sum(dyn); ev = wg.sum(ev);
    if (wg.leader && world >= 0) {
        if (dyn) atomicAdd(&diag[world].dynamicBodies, dyn);
        if (ev) atomicAdd(&diag[world].nanEvents, ev);
    }
}

__global__ void decay_user_forces(ForceView fv, SolveParams prm) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < fv.nJoints) {
        JointRec& j = fv.joints[t];
        for (int r = 0; r < 6; ++r) decay_row(j.lambda[r], j.penalty[r], r < 3 ? j.kLin : j.kAng, prm);
    } else if (t - fv.nJoints < fv.nSprings) {
        SpringRec& s = fv.springs[t - fv.nJoints];
        decay_row(s.lambda, s.penalty, s.k, prm);
    }
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

// One colour of the primal sweep (solver.cpp:344-409), in two phases so both are lane-dense:
//   phase 1  LPB lanes per body: lane l takes contact slots l, l+LPB, ... of the body's manifolds
//            (computeConstraint + 3 rows each), then a shuffle reduction leaves the 27 sums in lane 0,
//            which parks them in shared memory;
//   phase 2  one lane per body (the first kThreads/LPB threads = full warps): inertial terms, Schur
//            solve, pose update.  Running the serial solve on lane 0 of every group instead would
//            issue it at 32/LPB-fold cost (measured: 0.8 ms per launch at 1M bodies, issue-bound).
template <int LPB>
__global__ void __launch_bounds__(kThreads) primal_colour(BodyView b, const int4* adjRange, const int* bList, ManifoldSet ms,
                                                          ForceView fv, const int* order, int count, SolveParams prm, float alpha,
                                                          float* dxOut, Diag* diag) {
    constexpr int BPB = kThreads / LPB;
    __shared__ float sSys[BPB * 27];                 // stride 27 is odd: conflict-free in phase 2
    int g = threadIdx.x / LPB, lane = threadIdx.x % LPB;
    int gid = blockIdx.x * BPB + g;
    bool live = gid < count;
    BodySystem sys; sys.clear();
    if (live) {
        int i = order[gid];
        BodyPose self = b.pose[i];
        V3 pos = xyz(self.pos); Q4 rot = quat(self.rot);
        float invMassSelf = self.pos.w;
        V3 I = xyz(b.aux[i].inert);
        M3 invIw = rot_diag(qmat(rot), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
        int4 rg = adjRange[i];
        int nA = rg.y - rg.x, nB = rg.w - rg.z;
        int slots = (nA + nB) * 4;
        for (int slot = lane; slot < slots; slot += LPB) {
            int e = slot >> 2, c = slot & 3;
            bool isA = e < nA;
            int m = isA ? rg.x + e : bList[rg.z + (e - nA)];
            int4 h = ms.hdr[m];
            if (c >= h.z) continue;
            int other = isA ? h.y : h.x;
            BodyPose po = b.pose[other];
            int ci = m * 4 + c;
            ContactState cs = load_contact(ms, ci);
            ContactEval ev;
            float mu = __int_as_float(h.w);
            if (isA) contact_constraint(pos, rot, invMassSelf, xyz(po.pos), quat(po.rot), po.pos.w, mu, alpha, cs, ev);
            else     contact_constraint(xyz(po.pos), quat(po.rot), po.pos.w, pos, rot, invMassSelf, mu, alpha, cs, ev);
            accumulate_contact(sys, cs, ev, isA, invIw);
            ms.cL[ci] = pack_lambda(cs);          // computeConstraint's side effects (manifold.cpp:224-241)
        }
        if (lane == 0 && fv.adjStart) accumulate_user_forces(sys, fv, b.pose, i, pos, rot, invIw);
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
    if (threadIdx.x >= BPB) return;
    gid = blockIdx.x * BPB + threadIdx.x;
    if (gid >= count) return;
    int i = order[gid];
    BodyPose self = b.pose[i];
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

// ------------------------------------------------------------------ dual
__global__ void dual_contacts(BodyView b, ManifoldSet ms, int nM, SolveParams prm, float alpha) {
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nM * 4) return;
    int m = ci >> 2, c = ci & 3;
    int4 h = ms.hdr[m];
    if (c >= h.z) return;
    BodyPose pa = b.pose[h.x], pb = b.pose[h.y];
    ContactState cs = load_contact(ms, ci);
    ContactEval ev;
    contact_constraint(xyz(pa.pos), quat(pa.rot), pa.pos.w, xyz(pb.pos), quat(pb.rot), pb.pos.w, __int_as_float(h.w), alpha, cs, ev);
    dual_contact(cs, ev, prm.beta);
    ms.cL[ci] = pack_lambda(cs);
    ms.cP[ci] = pack_penalty(cs);
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

// ------------------------------------------------------------------ velocity + diagnostics
__global__ void velocity_bodies(BodyView b, SolveParams prm, Diag* diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float ls = 0.0f, as = 0.0f; int ev = 0; int world = -1;
    if (i < b.n) {
        world = b.worldId[i];
        if (b.pose[i].pos.w > 0.0f) {
            BodyVel vel = b.vel[i]; float4 pl;
            ev = velocity_body(b.pose[i], b.init[i], vel, pl, prm.dt, ls, as);
            b.vel[i] = vel; b.prevLin[i] = pl;
        }
    }
    WorldGroup wg(world);
    ls = wg.max_nonneg(ls); as = wg.max_nonneg(as); ev = wg.sum(ev);
    if (wg.leader && world >= 0) {
        Diag* d = diag + world;
        if (ls > 0.0f) atomic_max_nonneg(&d->maxLinearSpeed, ls);
        if (as > 0.0f) atomic_max_nonneg(&d->maxAngularSpeed, as);
        if (ev) atomicAdd(&d->nanEvents, ev);
    }
}

// solver.cpp:472-497
__global__ void diagnostics_contacts(BodyView b, ManifoldSet ms, int nM, Diag* diag) {
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    float pen = 0.0f, viol = 0.0f, lam = 0.0f; int nc = 0, nm = 0, nv = 0; int world = -1;
    if (ci < nM * 4) {
        int m = ci >> 2, c = ci & 3;
        int4 h = ms.hdr[m];
        world = b.worldId[h.x];
        if (c == 0 && h.z > 0) { nm = 1; nc = h.z; }
        if (c < h.z) {
            BodyPose pa = b.pose[h.x], pb = b.pose[h.y];
            nv = (pa.pos.w > 0.0f ? 1 : 0) + (pb.pos.w > 0.0f ? 1 : 0);
            float4 a4 = ms.cA[ci], b4 = ms.cB[ci], n4 = ms.cN[ci];
            V3 pA = xyz(pa.pos) + qrot(quat(pa.rot), xyz(a4));
            V3 pB = xyz(pb.pos) + qrot(quat(pb.rot), xyz(b4));
            float sepn = dot(pA - pB, xyz(n4));
            pen = fmax2(0.0f, -sepn);
            viol = fmax2(0.0f, kPenetrationSlop - sepn);
            lam = fabsf(ms.cL[ci].x);
        }
    }
    WorldGroup wg(world);
    pen = wg.max_nonneg(pen); viol = wg.max_nonneg(viol); lam = wg.max_nonneg(lam);
    nc = wg.sum(nc); nm = wg.sum(nm); nv = wg.sum(nv);
    if (wg.leader && world >= 0) {
        Diag* d = diag + world;
        if (pen > 0.0f) atomic_max_nonneg(&d->maxPenetration, pen);
        if (viol > 0.0f) atomic_max_nonneg(&d->maxViolation, viol);
        if (lam > 0.0f) atomic_max_nonneg(&d->maxNormalImpulse, lam);
        if (nc) atomicAdd(&d->activeContacts, nc);
        if (nm) atomicAdd(&d->activeManifolds, nm);
        if (nv) atomicAdd(&d->contactVisits, nv);
    }
}

// Rigid public state <-> the 13-float-per-body host layout (pos3 quat4 lin3 ang3), on the device so the
// host side of Solver::step() is one DMA each way.
__global__ void pack_state(BodyView b, float* out13) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    BodyPose p = b.pose[i]; BodyVel v = b.vel[i];
    float* o = out13 + 13 * (size_t)i;
    o[0] = p.pos.x; o[1] = p.pos.y; o[2] = p.pos.z; o[3] = p.rot.x; o[4] = p.rot.y; o[5] = p.rot.z; o[6] = p.rot.w;
    o[7] = v.lin.x; o[8] = v.lin.y; o[9] = v.lin.z; o[10] = v.ang.x; o[11] = v.ang.y; o[12] = v.ang.z;
}
__global__ void unpack_state(BodyView b, const float* in13) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const float* o = in13 + 13 * (size_t)i;
    float invMass = b.aux[i].mass.y;
    BodyPose p; p.pos = make_float4(o[0], o[1], o[2], invMass); p.rot = make_float4(o[3], o[4], o[5], o[6]);
    BodyVel v; v.lin = make_float4(o[7], o[8], o[9], 0.f); v.ang = make_float4(o[10], o[11], o[12], 0.f);
    b.pose[i] = p; b.vel[i] = v;
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

} // namespace avbd
