// avbd_world.cuh — device-resident layout of one AVBD world (or a batch of
// independent worlds packed back to back) and the launch-time view of it.
//
// Bodies (replaces the reference's 264-byte Rigid list nodes, solver.h:48-82):
//   pose[i]  {pos.xyz,radius | rot}            32 B, one sector: the only thing a NEIGHBOUR reads
//   aux[i]   {posI | rotI | mass,invMass,friction,- | Ixx,Iyy,Izz,-}   64 B, read by the body's own solve
//   vel[i]   {lin | ang}, init[i] {pos0 | rot0}, prevLin[i], size[i] {sx,sy,sz,friction}
// Manifolds (replaces 704-byte Manifold nodes, solver.h:112-143), slot m, sorted by pair key:
//   mhdr[m]  {bodyA, bodyB, numContacts, friction-bits}     16 B
//   cstart[m] first contact of manifold m; contacts are stored DENSE (ci = cstart[m] + c, live contacts only, written in place by the
//   manifold build, about 2 per
//   manifold on a box pile — a fixed 4-slot layout made every 64-byte DRAM granule half dead), one float4 per field:
//     cA {rA.xyz, C0_n}  cB {rB.xyz, C0_t.x}  cN {normal.xyz, C0_t.y}
//     lp {lambda_n, lambda_t1, lambda_t2, stick | penalty_n, penalty_t1, penalty_t2, feature-bits}   32 B = one sector: the
//        only per-contact state the iterations write, gathered by contact id from both endpoints' visits
//   cM[ci]   manifold of contact ci
//   C / fmin / fmax (solver.h:91-92) are recomputed in registers by every
//   consumer and never stored.
#pragma once
#include "avbd_math.cuh"

namespace avbd {

struct BodyPose { float4 pos; float4 rot; };
struct BodyAux  { float4 posI; float4 rotI; float4 mass; float4 inert; };
struct BodyVel  { float4 lin; float4 ang; };
struct BodyInit { float4 pos0; float4 rot0; };
struct ContactLP { float4 l; float4 p; };      // {lambda xyz, stick | penalty xyz, feature bits}

enum BodyFlag : int { kDynamic = 1, kLarge = 2 };

struct SolveParams {          // Solver fields, solver.h:147-151, re-read every step
    float dt;
    float gx, gy, gz;
    int   iterations;
    float alpha, beta, gamma;
    int   postStabilize;
};

struct ManifoldSet {          // one of the two ping-pong generations
    unsigned long long* key;  // packed pair key (A << keyShift) | B, ascending
    int4*   hdr;
    int*    cstart;           // nM + 1 entries
    int*    cM;               // per dense contact
    float4* cA; float4* cB; float4* cN; ContactLP* lp;
};

// Joint (6 rows, joint.cpp) / Spring (1 row, spring.cpp) records.  Unlike
// manifolds these are created by the user and persist, so they stay AoS.
struct JointRec {
    int a, b;                 // a == -1: world anchor
    float4 rA, rB;            // local anchors (rA = world anchor when a < 0)
    float4 rel0;              // initial relative orientation
    // the row arrays of Force (solver.h:91-97) a caller may edit between steps; stiffness == FLT_MAX: hard row
    float lambda[6], penalty[6], stiffness[6], motor[6];
};
struct SpringRec {
    int a, b;
    float4 rA, rB;
    float rest, k;            // k = stiffness[0]
    float lambda, penalty, motor;
};

struct Diag {                 // Solver::Diagnostics, solver.h:155-164 (+ sanitiser events)
    float maxPenetration, maxViolation, maxLinearSpeed, maxAngularSpeed, maxNormalImpulse;
    int activeContacts, activeManifolds, dynamicBodies;
    int nanEvents;            // bodies scrubbed by the NaN guards (solver.cpp:51-66)
    int contactVisits;        // contacts counted once per DYNAMIC endpoint (= primal contact visits per iteration)
    int pad[2];
};

struct Counters {             // device-side sizes produced by one stage, consumed by the next
    int nPairs;               // sphere-overlap pairs emitted this step (before merge with persisting manifolds)
    int nCand;                // pairs + persisting manifold keys (sort input)
    int nSurvive;             // candidates that passed the SAT cull == manifolds this step
    int nUncoloured;
    int nColours;
    int nContacts;            // live contacts this step (dense contact list length)
    int overflow;             // bit0 pairs, bit1 manifolds, bit2 colours
    int topoChanged;          // set by np_build when a manifold's slot or contact count differs from last step
    int nFree;                // dynamic bodies no contact visits and no user force touches (graph stage): solved by their own small kernel
    int nLinkedFree;          // ... no contact visits, but a joint / spring: solved in colour order by the same kernel
    int colourRounds;         // Jones-Plassmann rounds the cooperative colouring took (statistics)
    int colourKept;           // dynamic bodies that kept last step's colour (kept colouring; statistics)
    int nSphere;              // sphere-overlap pairs the fused sweep + SAT kernel tested (they never reach a list; statistics only)
};

} // namespace avbd
