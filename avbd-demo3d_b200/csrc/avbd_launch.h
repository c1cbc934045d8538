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

// Contact geometry in VISIT order (one float4 per field per visit, refreshed once per step by visit_geometry): the sweeps stream it
// fully coalesced instead of gathering 3 x 16 B per visit by contact id, in the visiting body's frame:
// {r_self,C0n} {r_other,C0t.x} {n,C0t.y}.
struct VisitGeom { float4* a; float4* b; float4* n; };

#ifdef __CUDACC__
// Programmatic dependent launch.  A step is a chain of ~50 (Stress1000) to ~200 (1M boxes) small dependent kernels on one
// stream; launched this way a kernel may become resident while its predecessor drains, and waits at the
// cudaGridDependencySynchronize() EVERY kernel of this library executes first, before it reads anything another kernel wrote
// (no kernel triggers early, so the wait ends when the predecessor has completed and its writes are visible).
template <class... KArgs, class... Args>
inline cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
struct ContactDiagPart { int world; float pen, viol, lam; int contacts, manifolds, visits; };
__device__ __forceinline__ ContactDiagPart contact_diag_part(int world, float sepn, float lamN, int contacts, int manifolds, int visits) {
    ContactDiagPart p{world, 0.0f, 0.0f, 0.0f, contacts, manifolds, visits};
    if (world >= 0 && contacts) {
        p.pen = (-sepn > 0.0f) ? -sepn : 0.0f;
        float v = 0.005f - sepn;                       // PENETRATION_SLOP, solver.h:36
        p.viol = v > 0.0f ? v : 0.0f;
        p.lam = fabsf(lamN);
    }
    return p;
}
// Combines the lanes of the calling warp that share a world; returns whether this lane is its group's leader.
__device__ __forceinline__ bool combine_diag_warp(ContactDiagPart& p) {
    WorldGroup wg(p.world);
    p.pen = wg.max_nonneg(p.pen); p.viol = wg.max_nonneg(p.viol); p.lam = wg.max_nonneg(p.lam);
    p.contacts = wg.sum(p.contacts); p.manifolds = wg.sum(p.manifolds); p.visits = wg.sum(p.visits);
    return wg.leader;
}
// The maxima only grow during a step, so a plain read that already shows a value >= ours makes the atomic unnecessary
// (a stale read only costs a redundant atomic): in steady state almost every warp skips all three.
__device__ __forceinline__ void commit_diag(const ContactDiagPart& p, Diag* diag) {
    Diag* d = diag + p.world;
    if (p.pen > 0.0f && p.pen > *(volatile float*)&d->maxPenetration) atomic_max_nonneg(&d->maxPenetration, p.pen);
    if (p.viol > 0.0f && p.viol > *(volatile float*)&d->maxViolation) atomic_max_nonneg(&d->maxViolation, p.viol);
    if (p.lam > 0.0f && p.lam > *(volatile float*)&d->maxNormalImpulse) atomic_max_nonneg(&d->maxNormalImpulse, p.lam);
    if (p.contacts) atomicAdd(&d->activeContacts, p.contacts);
    if (p.manifolds) atomicAdd(&d->activeManifolds, p.manifolds);
    if (p.visits) atomicAdd(&d->contactVisits, p.visits);
}
__device__ __forceinline__ void reduce_contact_diag(int world, float sepn, float lamN, int contacts, int manifolds, int visits, Diag* diag) {
    ContactDiagPart p = contact_diag_part(world, sepn, lamN, contacts, manifolds, visits);
    bool leader = combine_diag_warp(p);
    if (leader && p.world >= 0) commit_diag(p, diag);
}
// Block-wide form (EVERY thread of a <= 1024-thread block must call it): warps whose lanes all share one world park their
// partial result in shared memory and the first warp combines them, so a block issues one set of atomics per world it
// touches instead of one per warp (a 1M-box world is ONE world: 135k warps hammering six addresses cost 0.4 ms).
__device__ __forceinline__ void reduce_contact_diag_block(int world, float sepn, float lamN, int contacts, int manifolds, int visits, Diag* diag) {
    __shared__ ContactDiagPart sPart[32];
    ContactDiagPart p = contact_diag_part(world, sepn, lamN, contacts, manifolds, visits);
    WorldGroup wg(p.world);
    bool uniform = wg.peers == 0xffffffffu;
    bool leader = combine_diag_warp(p);
    int warp = threadIdx.x >> 5, nWarps = (blockDim.x + 31) >> 5;
    if (uniform) { if (leader) sPart[warp] = p; }
    else {
        if (leader && p.world >= 0) commit_diag(p, diag);
        if ((threadIdx.x & 31) == 0) sPart[warp].world = -1;
    }
    __syncthreads();
    if (warp == 0) {
        ContactDiagPart q{-1, 0.0f, 0.0f, 0.0f, 0, 0, 0};
        if ((int)threadIdx.x < nWarps) q = sPart[threadIdx.x];
        bool lead = combine_diag_warp(q);
        if (lead && q.world >= 0) commit_diag(q, diag);
    }
}
#endif

constexpr int kThreads = 256;
constexpr int kClusterBodiesPerTile = 28;     // small-world cluster loop: bodies per CTA tile (9 lanes per body in the sum phase)

// The large-world sweep of one colour (avbd_solve.cu: primal_sweep_warp): the colour's contact visits — a contiguous run of the
// colour-ordered visit list — cut into nWarps body-aligned warp ranges (range[0 .. nWarps], launch_warp_ranges, once per graph
// build for all colours); each warp pipelines its range and solves the bodies it finishes.  primal_sweep_warps(nVisits) = warps for
// a colour of that many visits.
// biasDual >= 0: the previous iteration's dual pass is still pending and each contact's first visit applies it (deferred dual); the
// value is that pass's clamp(1 - alpha, 0, 1) (manifold.cpp:179), which lies in [0, 1] for every alpha, so a negative value can only
// mean "nothing pending" (plain primal sweep).
int primal_sweep_warps(int nVisits);
void launch_warp_ranges(cudaStream_t s, const int* vstart, const int2* colRange, int nColours, const int* nWarps, const int* off, int* range);
// freeList / nFree: bodies no contact visits and no user force touches; extra warps of THIS launch solve them (pass them with one
// colour per sweep).  staticsJustWritten: the launch right before this one wrote visits / vg / ranges (the kernel then waits for its
// predecessor before it reads them; otherwise it issues its static loads first and waits only for poses / lambda / penalty).
void launch_primal_sweep(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* range, int nWarps, SolveParams prm,
                         float alpha, float biasDual, float* dxOut, Diag* diag, const int* freeList, int nFree, bool staticsJustWritten);
// The whole iteration loop (solver.cpp:340-431, manifold rows only) in ONE cooperative launch of the same warp pipelines, a grid
// barrier between colour phases (avbd_solve.cu: solve_loop_grid).  ranges / nWarps / off as built for the per-colour launches.
// Returns false if the launch was refused (caller falls back to per-colour launches).
bool launch_solve_loop_grid(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* ranges, int nColours,
                            const int* nWarps, const int* off, SolveParams prm, Diag* diag, const int* freeList, int nFree);
// The same loop as ONE thread-block cluster of up to 16 CTAs, hardware cluster barrier between colour phases (small worlds).
bool launch_solve_loop_warps(cudaStream_t s, BodyView b, const int4* visits, VisitGeom vg, ManifoldSet ms, ForceView fv, const int* ranges, int nColours,
                             const int* nWarps, const int* off, SolveParams prm, Diag* diag, const int* freeList, int nFree);
// Dynamic bodies no contact visits that a joint / spring links to another body (listed by the graph stage), filtered to one colour
// (onlyColour < 0: no filter).
void launch_primal_free(cudaStream_t s, BodyView b, ForceView fv, const int* freeList, int nFree, const int* colour, int onlyColour, SolveParams prm,
                        float* dxOut, Diag* diag);
// Dual + penalty ramp over the nContacts live (densely stored) contacts.
// `diag` != nullptr: this is the step's last dual pass and no body moves after it, so the contact diagnostics
// (solver.cpp:472-497) are reduced here from the values already in registers instead of by a separate sweep.
// unvisitedReps > 0 (deferred-dual steps): contacts no dynamic body visits take that many passes here, and with
// onlyUnvisited all other contacts are skipped.
void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, int nContacts, SolveParams prm, float alpha, int unvisitedReps, bool onlyUnvisited, Diag* diag);
// Small worlds: the whole iteration loop (solver.cpp:340-431, manifold rows only) in ONE launch of one thread-block
// cluster (<= 16 CTAs, hardware cluster barrier between phases).  Returns false if the launch was refused (caller
// falls back to per-colour launches).
bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, int nContacts, SolveParams prm,
                       Diag* diag, bool contactDiag, bool anyUnvisited);
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm);
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6);

} // namespace avbd
