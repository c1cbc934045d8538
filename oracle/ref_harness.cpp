// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin C-ABI shim over the UNMODIFIED reference translation units, compiled
// in place from /root/reference/source by oracle/Makefile into
// oracle/_ref/libavbd_ref.so.  It gives the parity tests stage-level access to
// the real reference:
//   * whole-step stepping of an arbitrary body set           (Solver::step, solver.cpp:255)
//   * Manifold::collide on an arbitrary pair                  (collision.cpp:420)
//   * solve6x6 / rowPenaltyGain from solver.cpp's anonymous namespace
//     (reached by #including solver.cpp as this TU's solver object)
//   * dumps of body state, manifolds (contacts, feature ids, lambda, penalty)
// No reference source is copied: this file only *calls* the reference.
//
// Everything is indexed by creation order (index 0 = first body created),
// not by Rigid::id (a process-global counter, rigid.cpp:10).

#include "solver.cpp"   // the reference's solver TU, verbatim, incl. its anonymous namespace
#include "joint.h"
#include "spring.h"
#include "ignorecollision.h"
#include "scenes.h"

#include <algorithm>
#include <cstring>
#include <vector>
#include <unordered_map>

namespace {

struct RefWorld {
    Solver* solver = nullptr;
    std::vector<Rigid*> order;                       // creation order
    std::unordered_map<const Rigid*, int> index;     // Rigid* -> creation index

    void reindex() {
        order.clear();
        index.clear();
        for (Rigid* b = solver->bodies; b; b = b->next) order.push_back(b);
        // list is newest-first (rigid.cpp:19-21)
        std::reverse(order.begin(), order.end());
        for (size_t i = 0; i < order.size(); ++i) index[order[i]] = (int)i;
    }
};

static vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
static quat q4(const float* p) { return quat(p[0], p[1], p[2], p[3]); }

} // namespace

extern "C" {

void* ref_create() {
    RefWorld* w = new RefWorld();
    w->solver = new Solver();
    return w;
}

void ref_destroy(void* h) {
    RefWorld* w = (RefWorld*)h;
    delete w->solver;
    delete w;
}

void ref_set_params(void* h, float dt, const float* g, int iterations, float alpha, float beta, float gamma,
                    int postStabilize) {
    Solver* s = ((RefWorld*)h)->solver;
    s->dt = dt;
    s->gravity = v3(g);
    s->iterations = iterations;
    s->alpha = alpha;
    s->beta = beta;
    s->gamma = gamma;
    s->postStabilize = postStabilize != 0;
}

void ref_get_params(void* h, float* out8) {
    Solver* s = ((RefWorld*)h)->solver;
    out8[0] = s->dt; out8[1] = s->gravity.x; out8[2] = s->gravity.y; out8[3] = s->gravity.z;
    out8[4] = (float)s->iterations; out8[5] = s->alpha; out8[6] = s->beta; out8[7] = s->gamma;
}

// Loads a scenes.h preset by its sceneNames[] entry; returns body count or -1.
int ref_load_scene(void* h, const char* name) {
    RefWorld* w = (RefWorld*)h;
    for (int i = 0; i < sceneCount; ++i) {
        if (std::strcmp(sceneNames[i], name) == 0) {
            scenes[i](w->solver);
            w->reindex();
            return (int)w->order.size();
        }
    }
    return -1;
}

int ref_add_body(void* h, const float* size, float density, float friction, const float* pos, const float* q,
                 const float* lin, const float* ang) {
    RefWorld* w = (RefWorld*)h;
    Rigid* b = new Rigid(w->solver, v3(size), density, friction, v3(pos), q4(q), v3(lin), v3(ang));
    w->index[b] = (int)w->order.size();
    w->order.push_back(b);
    return (int)w->order.size() - 1;
}

// a < 0 => body-world joint anchored at worldAnchor = anchorA.
void ref_add_joint(void* h, int a, int b, const float* anchorA, const float* anchorB, float linK, float angK) {
    RefWorld* w = (RefWorld*)h;
    if (a < 0) new Joint(w->solver, w->order[b], v3(anchorA), linK, angK);
    else new Joint(w->solver, w->order[a], w->order[b], v3(anchorA), v3(anchorB), linK, angK);
}

void ref_add_spring(void* h, int a, int b, const float* anchorA, const float* anchorB, float k, float rest) {
    RefWorld* w = (RefWorld*)h;
    new Spring(w->solver, w->order[a], w->order[b], v3(anchorA), v3(anchorB), k, rest);
}

void ref_add_ignore(void* h, int a, int b) {
    RefWorld* w = (RefWorld*)h;
    new IgnoreCollision(w->solver, w->order[a], w->order[b]);
}

void ref_step(void* h, int n) {
    Solver* s = ((RefWorld*)h)->solver;
    for (int i = 0; i < n; ++i) s->step();
}

int ref_num_bodies(void* h) { return (int)((RefWorld*)h)->order.size(); }

// 13 floats per body, creation order: pos3 quat4 lin3 ang3.
void ref_get_state(void* h, float* out) {
    RefWorld* w = (RefWorld*)h;
    for (Rigid* b : w->order) {
        *out++ = b->position.x; *out++ = b->position.y; *out++ = b->position.z;
        *out++ = b->orientation.x; *out++ = b->orientation.y; *out++ = b->orientation.z; *out++ = b->orientation.w;
        *out++ = b->linearVelocity.x; *out++ = b->linearVelocity.y; *out++ = b->linearVelocity.z;
        *out++ = b->angularVelocity.x; *out++ = b->angularVelocity.y; *out++ = b->angularVelocity.z;
    }
}

void ref_set_state(void* h, const float* in) {
    RefWorld* w = (RefWorld*)h;
    for (Rigid* b : w->order) {
        b->position = v3(in); in += 3;
        b->orientation = q4(in); in += 4;
        b->linearVelocity = v3(in); in += 3;
        b->angularVelocity = v3(in); in += 3;
    }
}

// prevLinearVelocity feeds the adaptive gravity weight (solver.cpp:322-326).
void ref_get_prev_linvel(void* h, float* out) {
    for (Rigid* b : ((RefWorld*)h)->order) { *out++ = b->prevLinearVelocity.x; *out++ = b->prevLinearVelocity.y; *out++ = b->prevLinearVelocity.z; }
}
void ref_set_prev_linvel(void* h, const float* in) {
    for (Rigid* b : ((RefWorld*)h)->order) { b->prevLinearVelocity = v3(in); in += 3; }
}

// 10 floats per body: size3 mass invMass inertiaDiag3 friction radius.
void ref_get_body_props(void* h, float* out) {
    for (Rigid* b : ((RefWorld*)h)->order) {
        *out++ = b->size.x; *out++ = b->size.y; *out++ = b->size.z;
        *out++ = b->mass; *out++ = b->invMass;
        *out++ = b->inertiaTensor.cols[0].x; *out++ = b->inertiaTensor.cols[1].y; *out++ = b->inertiaTensor.cols[2].z;
        *out++ = b->friction; *out++ = b->radius;
    }
}

// out5f: maxPen maxViolation maxLin maxAng maxLambda ; out3i: contacts manifolds dynBodies
void ref_get_diagnostics(void* h, float* out5f, int* out3i) {
    const Solver::Diagnostics& d = ((RefWorld*)h)->solver->lastDiagnostics;
    out5f[0] = d.maxPenetration; out5f[1] = d.maxConstraintViolation; out5f[2] = d.maxLinearSpeed;
    out5f[3] = d.maxAngularSpeed; out5f[4] = d.maxNormalImpulse;
    out3i[0] = d.activeContacts; out3i[1] = d.activeManifolds; out3i[2] = d.dynamicBodies;
}

int ref_num_manifolds(void* h) {
    int n = 0;
    for (Force* f = ((RefWorld*)h)->solver->forces; f; f = f->next) n += f->isManifold() ? 1 : 0;
    return n;
}

// Manifolds in solver list order (newest first).  Per manifold:
//   ints  [3]      : idxA idxB numContacts
//   feats [4]      : feature.value per contact
//   stick [4]      : 0/1
//   flts  [1+4*17+24]: friction, per contact {rA3 rB3 n3 pen C0n C0t3 (=14)...}
// Layout per contact (14 floats): rA3 rB3 normal3 penetration C0_n C0_t.x C0_t.y C0_t.z
// then lambda[12], penalty[12].
void ref_get_manifolds(void* h, int* ints, int* feats, int* stick, float* flts) {
    RefWorld* w = (RefWorld*)h;
    for (Force* f = w->solver->forces; f; f = f->next) {
        if (!f->isManifold()) continue;
        Manifold* m = (Manifold*)f;
        *ints++ = w->index[m->bodyA]; *ints++ = w->index[m->bodyB]; *ints++ = m->numContacts;
        *flts++ = m->combinedFriction;
        for (int i = 0; i < 4; ++i) {
            const Manifold::Contact& c = m->contacts[i];
            bool live = i < m->numContacts;
            *feats++ = live ? c.feature.value : 0;
            *stick++ = live ? (c.stick ? 1 : 0) : 0;
            const float vals[14] = {c.rA.x, c.rA.y, c.rA.z, c.rB.x, c.rB.y, c.rB.z, c.normal.x, c.normal.y, c.normal.z,
                                    c.penetration, c.C0_n, c.C0_t.x, c.C0_t.y, c.C0_t.z};
            for (int k = 0; k < 14; ++k) *flts++ = live ? vals[k] : 0.0f;
        }
        for (int k = 0; k < 12; ++k) *flts++ = k < m->numContacts * 3 ? m->lambda[k] : 0.0f;
        for (int k = 0; k < 12; ++k) *flts++ = k < m->numContacts * 3 ? m->penalty[k] : 0.0f;
    }
}

// Direct narrowphase oracle.  bodies: 2 x {size3 pos3 quat4}.  Output per
// contact: feature, then 10 floats rA3 rB3 normal3 penetration.  Returns count.
int ref_collide(const float* a, const float* b, int* feats, float* out) {
    Solver tmp;
    Rigid* A = new Rigid(&tmp, v3(a), 1.0f, 0.5f, v3(a + 3), q4(a + 6));
    Rigid* B = new Rigid(&tmp, v3(b), 1.0f, 0.5f, v3(b + 3), q4(b + 6));
    Manifold::Contact c[4];
    int n = Manifold::collide(A, B, c, false);
    for (int i = 0; i < n; ++i) {
        feats[i] = c[i].feature.value;
        float* o = out + i * 10;
        o[0] = c[i].rA.x; o[1] = c[i].rA.y; o[2] = c[i].rA.z;
        o[3] = c[i].rB.x; o[4] = c[i].rB.y; o[5] = c[i].rB.z;
        o[6] = c[i].normal.x; o[7] = c[i].normal.y; o[8] = c[i].normal.z;
        o[9] = c[i].penetration;
    }
    return n;   // ~Solver deletes A and B
}

// lhs: 4 blocks ll la al aa, each 9 floats column-major (cols[c][r] at c*3+r); rhs 6; out 6.
void ref_solve6x6(const float* lhs, const float* rhs, float* out) {
    mat66 A;
    mat3* blocks[4] = {&A.ll, &A.la, &A.al, &A.aa};
    for (int b = 0; b < 4; ++b)
        for (int c = 0; c < 3; ++c)
            blocks[b]->cols[c] = vec3(lhs[b * 9 + c * 3 + 0], lhs[b * 9 + c * 3 + 1], lhs[b * 9 + c * 3 + 2]);
    vec6 r{v3(rhs), v3(rhs + 3)};
    vec6 x = solve6x6(A, r);
    out[0] = x.l.x; out[1] = x.l.y; out[2] = x.l.z; out[3] = x.a.x; out[4] = x.a.y; out[5] = x.a.z;
}

// Solver::pick (solver.cpp:145): creation index of the hit body or -1.
int ref_pick(void* h, const float* origin, const float* dir, float* local3) {
    RefWorld* w = (RefWorld*)h;
    vec3 local;
    Rigid* b = w->solver->pick(v3(origin), v3(dir), local);
    if (!b) return -1;
    local3[0] = local.x; local3[1] = local.y; local3[2] = local.z;
    return w->index[b];
}

// 3x3 LDL^T solve, maths.h:104.  A column-major.
void ref_solve3(const float* A, const float* b, float* out) {
    mat3 M(v3(A), v3(A + 3), v3(A + 6));
    vec3 x = solve(M, v3(b));
    out[0] = x.x; out[1] = x.y; out[2] = x.z;
}

} // extern "C"
