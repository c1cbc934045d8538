// avbd_launch.h — the seam between the two translation units of libavbd_b200.so:
//   avbd_engine.cu  (-fmad=false)  host orchestration, C ABI, broadphase / narrowphase / graph kernels: everything
//                                  whose outputs are compared bit for bit with the reference
//   avbd_solve.cu   (FMA allowed)  the per-iteration primal / dual kernels, compared within an FP32 tolerance
#pragma once
#include "avbd_world.cuh"

namespace avbd {

struct BodyView {
    BodyPose* pose; BodyAux* aux; BodyVel* vel; BodyInit* init;
    float4* prevLin; float4* size;      // size: sx sy sz friction
    int* flags; int* worldId; int* localIdx;
    int n;
};

// User forces (joints / springs) per body: static CSR built on the host at upload.
struct ForceView {
    JointRec* joints; int nJoints;
    SpringRec* springs; int nSprings;
    const int* adjStart; const int* adj;    // entry = index*4 + type*2 + isA ; type 0 joint, 1 spring
};

// Contact geometry in VISIT order (one float4 per field per visit, refreshed once per step by visit_geometry): the primal's
// visit kernel streams it fully coalesced instead of gathering 3 x 16 B per visit by contact id.  {rA,C0n} {rB,C0t.x} {n,C0t.y}.
struct VisitGeom { float4* a; float4* b; float4* n; };

#ifdef __CUDACC__
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


// Contact part of Solver::Diagnostics (solver.cpp:472-497) for one warp of contacts: sepn = (pA - pB) . n with the
// final poses, |lambda_n| after the last dual update; one atomic per (warp, world) and field.
__device__ __forceinline__ void reduce_contact_diag(int world, float sepn, float lamN, int contacts, int manifolds, int visits, Diag* diag) {
    float pen = 0.0f, viol = 0.0f, lam = 0.0f;
    if (world >= 0 && contacts) {
        pen = (-sepn > 0.0f) ? -sepn : 0.0f;
        float v = 0.005f - sepn;                       // PENETRATION_SLOP, solver.h:36
        viol = v > 0.0f ? v : 0.0f;
        lam = fabsf(lamN);
    }
    WorldGroup wg(world);
    pen = wg.max_nonneg(pen); viol = wg.max_nonneg(viol); lam = wg.max_nonneg(lam);
    contacts = wg.sum(contacts); manifolds = wg.sum(manifolds); visits = wg.sum(visits);
    if (wg.leader && world >= 0) {
        Diag* d = diag + world;
        if (pen > 0.0f) atomic_max_nonneg(&d->maxPenetration, pen);
        if (viol > 0.0f) atomic_max_nonneg(&d->maxViolation, viol);
        if (lam > 0.0f) atomic_max_nonneg(&d->maxNormalImpulse, lam);
        if (contacts) atomicAdd(&d->activeContacts, contacts);
        if (manifolds) atomicAdd(&d->activeManifolds, manifolds);
        if (visits) atomicAdd(&d->contactVisits, visits);
    }
}
#endif

constexpr int kThreads = 256;
constexpr int kLanesPerBody = 4;
constexpr int kClusterBodiesPerTile = 28;     // small-world cluster loop: bodies per CTA tile (9 lanes per body in the sum phase)

// One colour of the primal sweep: `count` bodies listed in `order`; visitStart[k] .. visitStart[k+1] is the run of
// `visits` of body order[k]; avgVisits (visits per body, whole world) picks the tile shape.
// `sums` is scratch for the split path: 28 floats per body of the colour.  Returns the number of kernels launched.
int launch_primal(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv,
                  const int* order, int count, float avgVisits, SolveParams prm, float alpha, float* sums, float* dxOut, Diag* diag);
// Dual + penalty ramp over the nContacts live (densely stored) contacts.
// `diag` != nullptr: this is the step's last dual pass and no body moves after it, so the contact diagnostics
// (solver.cpp:472-497) are reduced here from the values already in registers instead of by a separate sweep.
void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, int nContacts, SolveParams prm, float alpha, Diag* diag);
// Small worlds: the whole iteration loop (solver.cpp:340-431, manifold rows only) in ONE launch of one thread-block
// cluster (<= 16 CTAs, hardware cluster barrier between phases).  Returns false if the launch was refused (caller
// falls back to per-colour launches).
bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, int nContacts, SolveParams prm,
                       Diag* diag, bool contactDiag);
// Measurement aid (avbd_debug_time_primal): one colour's visit-sum kernel in mode 0 (product), 1 (memory only), 2 (math only).
void launch_primal_experiment(cudaStream_t s, int mode, BodyView b, const int* vstart, const int4* visits, VisitGeom vg, ManifoldSet ms, int count, float alpha,
                              float* sums, int nContacts);
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm);
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6);

} // namespace avbd
