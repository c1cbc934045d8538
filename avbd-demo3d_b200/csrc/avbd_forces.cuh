// avbd_forces.cuh — row math of the user-created constraints: Joint (6-row
// weld, joint.cpp:66-139) and Spring (1-row distance, spring.cpp:33-90).
// IgnoreCollision (ignorecollision.h) has no rows: on the device it is only an
// entry in the sorted pair-exclusion list the broadphase consults.
#pragma once
#include "avbd_rows.cuh"

namespace avbd {

struct ForceEval { float C[6], fmin[6], fmax[6]; };

// Joint::computeConstraint, joint.cpp:68-106
AVBD_HD void joint_constraint(const JointRec& j, bool hasA, V3 posA, Q4 rotA, V3 posB, Q4 rotB, ForceEval& e) {
    Q4 qA; V3 pA;
    if (hasA) { qA = rotA; pA = posA + qrot(qA, xyz(j.rA)); }
    else { qA = qid(); pA = xyz(j.rA); }
    V3 pB = posB + qrot(rotB, xyz(j.rB));
    V3 lc = pA - pB;
    e.C[0] = lc.x; e.C[1] = lc.y; e.C[2] = lc.z;
    Q4 cur = qmul(qconj(qA), rotB);
    Q4 dq = qmul(cur, qconj(quat(j.rel0)));
    V3 ac = mk3(dq.x, dq.y, dq.z) * 2.0f;
    e.C[3] = ac.x; e.C[4] = ac.y; e.C[5] = ac.z;
    for (int i = 0; i < 6; ++i) { e.fmin[i] = -FLT_MAX; e.fmax[i] = FLT_MAX; }
}

// Joint::computeDerivatives, joint.cpp:108-139.  `rotBody` is the orientation of the body asked about.
AVBD_HD void joint_jacobian(const JointRec& j, bool isA, Q4 rotBody, int row, V3& Jl, V3& Ja) {
    Jl = zero3(); Ja = zero3();
    float sg = isA ? 1.0f : -1.0f;
    if (isA && j.a < 0) return;
    V3 ax = zero3();
    int k = row < 3 ? row : row - 3;
    if (k == 0) ax.x = 1.0f; else if (k == 1) ax.y = 1.0f; else ax.z = 1.0f;
    if (row < 3) {
        V3 r = qrot(rotBody, isA ? xyz(j.rA) : xyz(j.rB));
        Jl = ax * sg;
        Ja = cross(r, ax) * sg;
    } else {
        Ja = ax * sg;
    }
}

AVBD_HD void spring_ends(const SpringRec& s, bool hasA, V3 posA, Q4 rotA, V3 posB, Q4 rotB, V3& pA, V3& pB) {
    Q4 qA = hasA ? rotA : qid();
    pA = hasA ? posA + qrot(qA, xyz(s.rA)) : xyz(s.rA);
    pB = posB + qrot(rotB, xyz(s.rB));
}
// Spring::computeConstraint, spring.cpp:33-56
AVBD_HD float spring_constraint(const SpringRec& s, bool hasA, V3 posA, Q4 rotA, V3 posB, Q4 rotB) {
    V3 pA, pB; spring_ends(s, hasA, posA, rotA, posB, rotB, pA, pB);
    return len(pA - pB) - s.rest;
}
// Spring::computeDerivatives, spring.cpp:59-90
AVBD_HD void spring_jacobian(const SpringRec& s, bool hasA, V3 posA, Q4 rotA, V3 posB, Q4 rotB, bool isA, V3& Jl, V3& Ja) {
    V3 pA, pB; spring_ends(s, hasA, posA, rotA, posB, rotB, pA, pB);
    V3 d = pA - pB;
    float L = len(d);
    if (L < kVecEps) { Jl = zero3(); Ja = zero3(); return; }
    V3 n = d / L;
    float sg = isA ? 1.0f : -1.0f;
    Jl = n * sg;
    V3 r = isA ? qrot(rotA, xyz(s.rA)) : qrot(rotB, xyz(s.rB));
    Ja = cross(r, n) * sg;
}

// rowPenaltyGain, solver.cpp:94-125, from the per-body Jacobian norms.
AVBD_HD float penalty_gain(float lw, float aw, float beta) {
    float tot = lw + aw;
    if (tot < 1.0e-8f) return beta;
    return (beta * lw + (beta * kAngularBetaScale) * aw) / tot;
}

// Warm-start decay of one non-manifold row, solver.cpp:281-293.
AVBD_HD void decay_row(float& lambda, float& penalty, float stiffness, const SolveParams& p) {
    if (!p.postStabilize) lambda *= p.alpha * p.gamma;
    penalty = clampf(penalty * p.gamma, kPenaltyMin, kPenaltyMax);
    if (stiffness > 0.0f && stiffness < FLT_MAX) penalty = fmin2(penalty, stiffness);
}

} // namespace avbd
