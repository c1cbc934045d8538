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

constexpr int kThreads = 256;
constexpr int kLanesPerBody = 4;

// One colour of the primal sweep: `count` bodies listed in `order`; visitStart[k] .. visitStart[k+1] is the run of
// `visits` of body order[k]; avgVisits (visits per body, whole world) picks the tile shape.
// `sums` is scratch for the split path: 28 floats per body of the colour.  Returns the number of kernels launched.
int launch_primal(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv,
                  const int* order, int count, float avgVisits, SolveParams prm, float alpha, float* sums, float* dxOut, Diag* diag);
// Dual + penalty ramp over the nContacts live contacts in contactList.
void launch_dual(cudaStream_t s, BodyView b, ManifoldSet ms, const int* contactList, int nContacts, SolveParams prm, float alpha);
// Small worlds: the whole iteration loop (solver.cpp:340-431, manifold rows only) in ONE cooperative launch.
// Returns false if the launch was refused (caller falls back to per-colour launches).
bool launch_solve_loop(cudaStream_t s, BodyView b, const int* visitStart, const int4* visits, ManifoldSet ms, ForceView fv, const int* order,
                       const int2* colRange, int nColours, int maxColourCount, const int* contactList, int nContacts, SolveParams prm,
                       Diag* diag, unsigned* barrier);
void launch_dual_user_forces(cudaStream_t s, BodyView b, ForceView fv, SolveParams prm);
void launch_solve6_batch(cudaStream_t s, const float* lhs36, const float* rhs6, int n, float* out6);

} // namespace avbd
