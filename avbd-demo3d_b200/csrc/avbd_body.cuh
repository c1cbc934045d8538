// avbd_body.cuh — per-body pieces of the step: predict (solver.cpp:299-337),
// the inertial part of the 6x6 system and the pose update of the primal
// solve (solver.cpp:351-369, :402-408), velocity recovery with damping
// (solver.cpp:434-469).  NaN/Inf scrubbing keeps the reference's semantics
// (solver.cpp:51-66) but counts events instead of printing from the kernel.
#pragma once
#include "avbd_rows.cuh"

namespace avbd {

AVBD_HD bool scrub3(V3& v) { if (!finite3(v)) { v = zero3(); return true; } return false; }
AVBD_HD bool scrub4(Q4& q) { if (!finite4(q)) { q = qid(); return true; } return false; }

// Returns the number of scrub events.  solver.cpp:299-337
AVBD_HD int predict_body(BodyPose& pose, BodyVel& vel, float4 prevLin4, BodyAux& aux, BodyInit& init, const SolveParams& p) {
    int events = 0;
    V3 pos = xyz(pose.pos); Q4 rot = quat(pose.rot);
    V3 lin = xyz(vel.lin), ang = xyz(vel.ang);
    float invMass = aux.mass.y;
    { float l = len(ang); if (l > kMaxAngularSpeed && l > kVecEps) ang = ang * (kMaxAngularSpeed / l); }   // solver.cpp:85-92
    init.pos0 = f4(pos, 0.0f); init.rot0 = f4(rot);
    V3 posI = pos; Q4 rotI = rot;
    if (invMass > 0.0f) {
        V3 g = mk3(p.gx, p.gy, p.gz);
        float dt = p.dt;
        events += scrub3(lin) ? 1 : 0;
        events += scrub3(ang) ? 1 : 0;
        posI = (pos + lin * dt) + g * (dt * dt);
        Q4 om = qmk(ang.x, ang.y, ang.z, 0.0f);
        rotI = qunit(qadd(rot, qscl(qmul(om, rot), 0.5f * dt)));
        float gl = len(g);
        float aw = 0.0f;
        if (gl > 1e-5f) {
            V3 acc = (lin - xyz(prevLin4)) / dt;
            float proj = dot(acc, g / gl);
            aw = clampf(proj / gl, 0.0f, 1.0f);
            if (!finite1(aw)) aw = 0.0f;
        }
        pos = pos + (lin * dt + g * (aw * dt * dt));
        rot = rotI;
        events += scrub3(pos) ? 1 : 0;
        events += scrub4(rot) ? 1 : 0;
    }
    pose.pos = f4(pos, pose.pos.w); pose.rot = f4(rot);
    vel.lin = f4(lin, 0.0f); vel.ang = f4(ang, 0.0f);
    aux.posI = f4(posI, 0.0f); aux.rotI = f4(rotI);
    return events;
}

// Inertial terms of the block system (solver.cpp:351-369).  Also returns the
// world inverse inertia used by the gyroscopic diagonal (rigid.cpp:51-54).
AVBD_HD void body_self_system(V3 pos, Q4 rot, const BodyAux& aux, float dt, BodySystem& s, M3& invIw) {
    float mass = aux.mass.x;
    V3 I = xyz(aux.inert);
    M3 R = qmat(rot);
    M3 Iw = rot_diag(R, I);
    invIw = rot_diag(R, mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
    float invDt2 = 1.0f / (dt * dt);
    s.clear();
    float m = mass * invDt2;
    s.ll[0] = m; s.ll[3] = m; s.ll[5] = m;
    s.aa[0] = Iw.c[0].x * invDt2; s.aa[1] = Iw.c[0].y * invDt2; s.aa[2] = Iw.c[0].z * invDt2;
    s.aa[3] = Iw.c[1].y * invDt2; s.aa[4] = Iw.c[1].z * invDt2; s.aa[5] = Iw.c[2].z * invDt2;
    V3 dl = (pos - xyz(aux.posI)) * invDt2;
    s.rl[0] = mass * dl.x; s.rl[1] = mass * dl.y; s.rl[2] = mass * dl.z;
    Q4 qe = qmul(rot, qconj(quat(aux.rotI)));
    V3 re = mk3(qe.x, qe.y, qe.z) * 2.0f;
    if (qe.w < 0.0f) re = -re;
    V3 ra = mv(Iw, re * invDt2);
    s.ra[0] = ra.x; s.ra[1] = ra.y; s.ra[2] = ra.z;
}

// solver.cpp:402-408
AVBD_HD int apply_body_update(V3& pos, Q4& rot, V3 dl, V3 da) {
    pos = pos - dl;
    Q4 dq = qmk(da.x, da.y, da.z, 0.0f);
    rot = qunit(qsub(rot, qscl(qmul(dq, rot), 0.5f)));
    int ev = 0;
    ev += scrub3(pos) ? 1 : 0;
    ev += scrub4(rot) ? 1 : 0;
    return ev;
}

// solver.cpp:434-469.  Outputs speeds for the diagnostics reduction.
AVBD_HD int velocity_body(const BodyPose& pose, const BodyInit& init, BodyVel& vel, float4& prevLin4, float dt, float& linSpeed, float& angSpeed) {
    V3 pos = xyz(pose.pos); Q4 rot = quat(pose.rot);
    prevLin4 = vel.lin;
    V3 lin = (pos - xyz(init.pos0)) / dt;
    Q4 dq = qmul(rot, qconj(quat(init.rot0)));
    V3 av = mk3(dq.x, dq.y, dq.z) * (2.0f / dt);
    if (dq.w < 0.0f) av = -av;
    lin = lin * kLinearDamping;
    av = av * kAngularDamping;
    int ev = 0;
    ev += scrub3(lin) ? 1 : 0;
    ev += scrub3(av) ? 1 : 0;
    vel.lin = f4(lin, 0.0f); vel.ang = f4(av, 0.0f);
    linSpeed = len(lin); angSpeed = len(av);
    return ev;
}

} // namespace avbd
