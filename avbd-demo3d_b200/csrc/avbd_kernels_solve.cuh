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
__global__ void predict_bodies(BodyView b, SolveParams prm, Diag* diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int dyn = 0, ev = 0;
    if (i < b.n) {
        BodyPose pose = b.pose[i]; BodyVel vel = b.vel[i]; BodyAux aux = b.aux[i]; BodyInit init;
        float4 pl = b.prevLin[i];
        // pose.pos.w carries invMass for neighbours; predict_body reads it from aux
        ev = predict_body(pose, vel, pl, aux, init, prm);
        b.pose[i] = pose; b.vel[i] = vel; b.init[i] = init;
        b.aux[i].posI = aux.posI; b.aux[i].rotI = aux.rotI;
        dyn = aux.mass.y > 0.0f ? 1 : 0;
    }
    dyn = __reduce_add_sync(0xffffffffu, dyn);
    ev = __reduce_add_sync(0xffffffffu, ev);
    if ((threadIdx.x & 31) == 0) {
        if (dyn) atomicAdd(&diag->dynamicBodies, dyn);
        if (ev) atomicAdd(&diag->nanEvents, ev);
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

// One colour of the primal sweep (solver.cpp:344-409).  LPB lanes per body;
// lane l takes contact slots l, l+LPB, ... of the body's manifolds.
template <int LPB>
__global__ void __launch_bounds__(kThreads) primal_colour(BodyView b, const int4* adjRange, const int* bList, ManifoldSet ms,
                                                          ForceView fv, const int* order, int count, SolveParams prm, float alpha,
                                                          float* dxOut, Diag* diag) {
    int gid = (blockIdx.x * blockDim.x + threadIdx.x) / LPB;
    int lane = threadIdx.x % LPB;
    bool live = gid < count;
    BodySystem sys; sys.clear();
    int i = -1; V3 pos = zero3(); Q4 rot = qid(); float invMassSelf = 0.0f;
    if (live) {
        i = order[gid];
        BodyPose self = b.pose[i];
        BodyAux aux = b.aux[i];
        pos = xyz(self.pos); rot = quat(self.rot);
        invMassSelf = aux.mass.y;
        M3 invIw;
        BodySystem own;
        body_self_system(pos, rot, aux, prm.dt, own, invIw);
        if (lane == 0) sys = own;
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
    if (live && lane == 0) {
        V3 dl, da;
        solve_body_system(sys, dl, da);
        int ev = apply_body_update(pos, rot, dl, da);
        BodyPose out; out.pos = f4(pos, invMassSelf); out.rot = f4(rot);
        b.pose[i] = out;
        if (dxOut) { float* o = dxOut + 6 * i; o[0] = dl.x; o[1] = dl.y; o[2] = dl.z; o[3] = da.x; o[4] = da.y; o[5] = da.z; }
        if (ev) atomicAdd(&diag->nanEvents, ev);
    }
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
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__global__ void velocity_bodies(BodyView b, SolveParams prm, Diag* diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float ls = 0.0f, as = 0.0f; int ev = 0;
    if (i < b.n && b.aux[i].mass.y > 0.0f) {
        BodyVel vel = b.vel[i]; float4 pl;
        ev = velocity_body(b.pose[i], b.init[i], vel, pl, prm.dt, ls, as);
        b.vel[i] = vel; b.prevLin[i] = pl;
    }
    for (int off = 16; off > 0; off >>= 1) {
        ls = fmax2(ls, __shfl_xor_sync(0xffffffffu, ls, off));
        as = fmax2(as, __shfl_xor_sync(0xffffffffu, as, off));
    }
    ev = __reduce_add_sync(0xffffffffu, ev);
    if ((threadIdx.x & 31) == 0) {
        if (ls > 0.0f) atomic_max_nonneg(&diag->maxLinearSpeed, ls);
        if (as > 0.0f) atomic_max_nonneg(&diag->maxAngularSpeed, as);
        if (ev) atomicAdd(&diag->nanEvents, ev);
    }
}

// solver.cpp:472-497
__global__ void diagnostics_contacts(BodyView b, ManifoldSet ms, int nM, Diag* diag) {
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    float pen = 0.0f, viol = 0.0f, lam = 0.0f; int nc = 0, nm = 0;
    if (ci < nM * 4) {
        int m = ci >> 2, c = ci & 3;
        int4 h = ms.hdr[m];
        if (c == 0 && h.z > 0) { nm = 1; nc = h.z; }
        if (c < h.z) {
            BodyPose pa = b.pose[h.x], pb = b.pose[h.y];
            float4 a4 = ms.cA[ci], b4 = ms.cB[ci], n4 = ms.cN[ci];
            V3 pA = xyz(pa.pos) + qrot(quat(pa.rot), xyz(a4));
            V3 pB = xyz(pb.pos) + qrot(quat(pb.rot), xyz(b4));
            float sepn = dot(pA - pB, xyz(n4));
            pen = fmax2(0.0f, -sepn);
            viol = fmax2(0.0f, kPenetrationSlop - sepn);
            lam = fabsf(ms.cL[ci].x);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        pen = fmax2(pen, __shfl_xor_sync(0xffffffffu, pen, off));
        viol = fmax2(viol, __shfl_xor_sync(0xffffffffu, viol, off));
        lam = fmax2(lam, __shfl_xor_sync(0xffffffffu, lam, off));
    }
    nc = __reduce_add_sync(0xffffffffu, nc);
    nm = __reduce_add_sync(0xffffffffu, nm);
    if ((threadIdx.x & 31) == 0) {
        if (pen > 0.0f) atomic_max_nonneg(&diag->maxPenetration, pen);
        if (viol > 0.0f) atomic_max_nonneg(&diag->maxViolation, viol);
        if (lam > 0.0f) atomic_max_nonneg(&diag->maxNormalImpulse, lam);
        if (nc) atomicAdd(&diag->activeContacts, nc);
        if (nm) atomicAdd(&diag->activeManifolds, nm);
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

} // namespace avbd
