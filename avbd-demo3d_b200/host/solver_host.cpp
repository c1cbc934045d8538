// solver_host.cpp — implementation of the host mirror (solver.h, joint.h, spring.h): list bookkeeping identical in
// effect to the reference's intrusive lists (rigid.cpp:19-21, force.cpp:12-69: newest first), construction
// forwarded to the device world, and Solver::step() = sync host edits -> avbd_step -> read state back.
// The row virtuals (computeConstraint / computeDerivatives) evaluate the SAME __host__ __device__ functions the
// kernels run, on the host copy of the poses; Solver::step() never calls them.
#include "solver.h"
#include "joint.h"
#include "spring.h"
#include "ignorecollision.h"

#include "../../include/avbd_b200.h"
#include "../csrc/avbd_body.cuh"
#include "../csrc/avbd_forces.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "avbd-demo3d_b200: %s: %s\n", what, avbd_last_error());
    std::fprintf(stderr, "avbd-demo3d_b200: the solver runs on a CUDA device only; there is no CPU fallback.\n");
    std::exit(2);
}
void check(int rc, const char* what) { if (rc < 0) die(what); }
bool g_clearing = false;

avbd::ContactState to_state(const Manifold& m, int i) {
    avbd::ContactState c;
    const Manifold::Contact& k = m.contacts[i];
    c.rA = k.rA; c.rB = k.rB; c.n = k.normal; c.C0n = k.C0_n; c.C0t1 = k.C0_t.x; c.C0t2 = k.C0_t.y;
    for (int r = 0; r < 3; ++r) { c.lam[r] = m.lambda[i * 3 + r]; c.pen[r] = m.penalty[i * 3 + r]; }
    c.stick = k.stick; c.feature = k.feature.value;
    return c;
}
void from_state(Manifold& m, int i, const avbd::ContactState& c) {
    Manifold::Contact& k = m.contacts[i];
    k.rA = c.rA; k.rB = c.rB; k.normal = c.n; k.C0_n = c.C0n; k.C0_t = vec3(c.C0t1, c.C0t2, 0.0f); k.stick = c.stick; k.feature.value = c.feature;
    for (int r = 0; r < 3; ++r) { m.lambda[i * 3 + r] = c.lam[r]; m.penalty[i * 3 + r] = c.pen[r]; }
}
} // namespace

// ----------------------------------------------------------------------------------------------- Rigid
int Rigid::next_id = 1;      // process-global and never reset, as upstream (rigid.cpp:10)

Rigid::Rigid(Solver* s, const vec3& sz, float dens, float fric, const vec3& pos, const quat& orient, const vec3& linVel, const vec3& angVel)
    : solver(s), forces(nullptr), next(nullptr), id(next_id++), position(pos), orientation(orient), linearVelocity(linVel),
      angularVelocity(angVel), prevLinearVelocity(linVel), prevAngularVelocity(angVel), size(sz), friction(fric), density(dens) {
    next = s->bodies; s->bodies = this;
    index = (int)s->order.size(); s->order.push_back(this);
    mass = size.x * size.y * size.z * density;                               // rigid.cpp:24-40
    invMass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
    radius = length(size) * 0.5f;
    if (invMass > 0.0f) {
        float ixx = (1.0f / 12.0f) * mass * (size.y * size.y + size.z * size.z);
        float iyy = (1.0f / 12.0f) * mass * (size.x * size.x + size.z * size.z);
        float izz = (1.0f / 12.0f) * mass * (size.x * size.x + size.y * size.y);
        inertiaTensor = mat3::diagonal(vec3(ixx, iyy, izz));
        invInertiaTensor = mat3::diagonal(vec3(1.0f / ixx, 1.0f / iyy, 1.0f / izz));
    } else {
        inertiaTensor = mat3::diagonal(0.0f); invInertiaTensor = mat3::diagonal(0.0f);
    }
}

Rigid::~Rigid() {
    Rigid** p = &solver->bodies;
    while (*p != this) p = &(*p)->next;
    *p = next;
    if (!g_clearing) {           // a single body removed: the device world is rebuilt at the next step
        solver->order.erase(std::find(solver->order.begin(), solver->order.end(), this));
        for (size_t i = 0; i < solver->order.size(); ++i) solver->order[i]->index = (int)i;
        solver->rebuild = true;
    }
}

mat3 Rigid::getInvInertiaTensorWorld() const { mat3 R = mat3_from_quat(orientation); return R * invInertiaTensor * transpose(R); }
mat3 Rigid::getInertiaTensorWorld() const { mat3 R = mat3_from_quat(orientation); return R * inertiaTensor * transpose(R); }

bool Rigid::isConstrainedTo(Rigid* other) const {
    for (Force* f = forces; f; f = (f->bodyA == this) ? f->nextA : f->nextB)
        if (f->bodyA == other || f->bodyB == other) return true;
    return false;
}

// ----------------------------------------------------------------------------------------------- Force
Force::Force(Solver* s, Rigid* a, Rigid* b) : solver(s), bodyA(a), bodyB(b), nextA(nullptr), nextB(nullptr), next(nullptr) {
    next = s->forces; s->forces = this;
    if (a) { nextA = a->forces; a->forces = this; }
    if (b) { nextB = b->forces; b->forces = this; }
    for (int i = 0; i < MAX_CONSTRAINT_ROWS; ++i) {
        stiffness[i] = 0.0f; lambda[i] = 0.0f; penalty[i] = 0.0f; motor[i] = 0.0f; fmin[i] = -FLT_MAX; fmax[i] = FLT_MAX; C[i] = 0.0f; fracture[i] = 0.0f;
    }
}

Force::~Force() {
    Force** p = &solver->forces;
    while (*p != this) p = &(*p)->next;
    *p = next;
    for (Rigid* body : {bodyA, bodyB}) {
        if (!body) continue;
        p = &body->forces;
        while (*p != this) p = ((*p)->bodyA == body) ? &(*p)->nextA : &(*p)->nextB;
        *p = (bodyA == body) ? nextA : nextB;
    }
    auto& uf = solver->userForces;
    auto it = std::find(uf.begin(), uf.end(), this);
    if (it != uf.end()) { uf.erase(it); if (!g_clearing) solver->rebuild = true; }
}

// ----------------------------------------------------------------------------------------------- Manifold (mirror)
Manifold::Manifold(Solver* s, Rigid* a, Rigid* b) : Force(s, a, b), numContacts(0), combinedFriction(0.0f) {
    for (auto& c : contacts) { c.feature.value = 0; c.penetration = 0.0f; c.C0_n = 0.0f; c.C0_t = vec3(); c.stick = false; }
}

int Manifold::collide(Rigid* a, Rigid* b, Contact* out, bool flip) {     // collision.cpp:420-489 through the shared device functions
    int code = avbd::sat_test(avbd::make_obb(a->position, a->orientation, a->size), avbd::make_obb(b->position, b->orientation, b->size));
    if (!code) return 0;
    avbd::RawContact rc[4];
    int n = avbd::build_contacts(a->position, a->orientation, a->size, b->position, b->orientation, b->size, code, rc);
    for (int i = 0; i < n; ++i) {
        out[i].feature.value = rc[i].feature; out[i].rA = rc[i].rA; out[i].rB = rc[i].rB; out[i].normal = rc[i].normal;
        vec3 xA = a->position + rotate(a->orientation, out[i].rA), xB = b->position + rotate(b->orientation, out[i].rB);
        out[i].penetration = max(0.0f, -dot(xA - xB, out[i].normal));
        out[i].C0_n = 0.0f; out[i].C0_t = vec3(); out[i].stick = false;
        if (flip) { vec3 t = out[i].rA; out[i].rA = out[i].rB; out[i].rB = t; out[i].normal = -out[i].normal; }
    }
    return n;
}

bool Manifold::initialize() {                                            // manifold.cpp:71-175
    combinedFriction = sqrtf(bodyA->friction * bodyB->friction);
    avbd::OldManifold old; old.n = numContacts;
    for (int i = 0; i < numContacts; ++i) old.ct[i] = to_state(*this, i);
    int code = avbd::sat_test(avbd::make_obb(bodyA->position, bodyA->orientation, bodyA->size),
                              avbd::make_obb(bodyB->position, bodyB->orientation, bodyB->size));
    numContacts = 0;
    if (!code) return false;
    avbd::SolveParams noDecay{}; noDecay.alpha = 1.0f; noDecay.gamma = 1.0f; noDecay.postStabilize = 1;   // initialize() alone does not decay rows
    avbd::NewManifold nm;
    avbd::manifold_initialize(bodyA->position, bodyA->orientation, bodyA->size, bodyB->position, bodyB->orientation, bodyB->size, code, old, noDecay, nm);
    numContacts = nm.n;
    for (int i = 0; i < nm.n; ++i) {
        from_state(*this, i, nm.ct[i]);
        for (int r = 0; r < 3; ++r) { stiffness[i * 3 + r] = FLT_MAX; motor[i * 3 + r] = 0.0f; }
    }
    return numContacts > 0;
}

void Manifold::computeConstraint(float alpha) {                          // manifold.cpp:177-245
    for (int i = 0; i < numContacts; ++i) {
        avbd::ContactState c = to_state(*this, i);
        avbd::ContactEval e;
        avbd::contact_constraint(bodyA->position, bodyA->orientation, bodyA->invMass, bodyB->position, bodyB->orientation, bodyB->invMass,
                                 combinedFriction, alpha, c, e);
        from_state(*this, i, c);
        for (int r = 0; r < 3; ++r) { C[i * 3 + r] = e.C[r]; fmin[i * 3 + r] = e.fmin[r]; fmax[i * 3 + r] = e.fmax[r]; }
        vec3 d = (bodyA->position + vec3(e.wrA)) - (bodyB->position + vec3(e.wrB));
        contacts[i].penetration = max(0.0f, -dot(d, contacts[i].normal));
    }
}

void Manifold::computeDerivatives(vec3& Jl, vec3& Ja, const Rigid* body, int row) const {   // manifold.cpp:247-271
    const Contact& c = contacts[row / 3];
    avbd::V3 n, t1, t2;
    avbd::contact_basis(c.normal, n, t1, t2);
    vec3 basis = (row % 3 == 0) ? vec3(n) : ((row % 3 == 1) ? vec3(t1) : vec3(t2));
    float sign = (body == bodyA) ? 1.0f : -1.0f;
    vec3 r = (body == bodyA) ? rotate(bodyA->orientation, c.rA) : rotate(bodyB->orientation, c.rB);
    Jl = basis * sign; Ja = cross(r, basis) * sign;
}

// ----------------------------------------------------------------------------------------------- Joint / Spring
namespace {
avbd::JointRec joint_rec(const Joint& j) {
    avbd::JointRec r{};
    r.a = j.bodyA ? 0 : -1; r.b = 0;
    r.rA = avbd::f4(j.rA, 0.f); r.rB = avbd::f4(j.rB, 0.f); r.rel0 = avbd::f4((avbd::Q4)j.initialRelativeOrientation);
    return r;
}
avbd::SpringRec spring_rec(const Spring& s) {
    avbd::SpringRec r{};
    r.a = s.bodyA ? 0 : -1; r.b = 0; r.rA = avbd::f4(s.rA, 0.f); r.rB = avbd::f4(s.rB, 0.f); r.rest = s.restLength; r.k = s.springStiffness;
    return r;
}
}

Joint::Joint(Solver* s, Rigid* a, Rigid* b, const vec3& la, const vec3& lb, float linK, float angK, float motor_, float fracture_)
    : Force(s, a, b), rA(la), rB(lb) {                                   // joint.cpp:11-38
    quat qa = a ? a->orientation : quat();
    initialRelativeOrientation = conjugate(qa) * b->orientation;
    for (int i = 0; i < 6; ++i) { stiffness[i] = i < 3 ? linK : angK; lambda[i] = 0; penalty[i] = PENALTY_MIN; }
    angularStiffness = angK; angularMotor = motor_; angularFracture = fracture_;
}
Joint::Joint(Solver* s, Rigid* b, const vec3& worldAnchor, float linK, float angK, float motor_, float fracture_)
    : Force(s, nullptr, b) {                                             // joint.cpp:41-63
    rA = worldAnchor;
    rB = transpose(mat3_from_quat(b->orientation)) * (worldAnchor - b->position);
    initialRelativeOrientation = b->orientation;
    for (int i = 0; i < 6; ++i) { stiffness[i] = i < 3 ? linK : angK; lambda[i] = 0; penalty[i] = PENALTY_MIN; }
    angularStiffness = angK; angularMotor = motor_; angularFracture = fracture_;
}
void Joint::computeConstraint(float) {                                   // joint.cpp:68-106
    avbd::ForceEval e;
    avbd::joint_constraint(joint_rec(*this), bodyA != nullptr, bodyA ? (avbd::V3)bodyA->position : avbd::zero3(),
                           bodyA ? (avbd::Q4)bodyA->orientation : avbd::qid(), bodyB->position, bodyB->orientation, e);
    for (int i = 0; i < 6; ++i) { C[i] = e.C[i]; fmin[i] = e.fmin[i]; fmax[i] = e.fmax[i]; }
}
void Joint::computeDerivatives(vec3& Jl, vec3& Ja, const Rigid* body, int row) const {      // joint.cpp:108-139
    avbd::V3 l, a;
    bool isA = body == bodyA;
    avbd::joint_jacobian(joint_rec(*this), isA, body ? (avbd::Q4)body->orientation : avbd::qid(), row, l, a);
    Jl = l; Ja = a;
}

Spring::Spring(Solver* s, Rigid* a, Rigid* b, const vec3& la, const vec3& lb, float k, float rest)
    : Force(s, a, b), rA(la), rB(lb), restLength(rest), springStiffness(k) {                 // spring.cpp:10-30
    stiffness[0] = k;
    if (restLength < 0) restLength = length((a->position + rotate(a->orientation, rA)) - (b->position + rotate(b->orientation, rB)));
    lambda[0] = 0.0f; penalty[0] = PENALTY_MIN; fmin[0] = -FLT_MAX; fmax[0] = FLT_MAX;
}
void Spring::computeConstraint(float) {                                  // spring.cpp:33-56
    C[0] = avbd::spring_constraint(spring_rec(*this), bodyA != nullptr, bodyA ? (avbd::V3)bodyA->position : avbd::zero3(),
                                   bodyA ? (avbd::Q4)bodyA->orientation : avbd::qid(), bodyB->position, bodyB->orientation);
}
void Spring::computeDerivatives(vec3& Jl, vec3& Ja, const Rigid* body, int) const {          // spring.cpp:59-90
    avbd::V3 l, a;
    avbd::spring_jacobian(spring_rec(*this), bodyA != nullptr, bodyA ? (avbd::V3)bodyA->position : avbd::zero3(),
                          bodyA ? (avbd::Q4)bodyA->orientation : avbd::qid(), bodyB->position, bodyB->orientation, body == bodyA, l, a);
    Jl = l; Ja = a;
}

// ----------------------------------------------------------------------------------------------- Solver
Solver::Solver()
    : bodies(nullptr), forces(nullptr), enableDiagnostics(false), logFrequency(60), stepIndex(0), lastDiagnostics{}, world(nullptr),
      device(0), rebuild(false), uploadedBodies(0), uploadedForces(0) {
    if (const char* d = std::getenv("AVBD_DEVICE")) device = std::atoi(d);
    defaultParams();
}

Solver::~Solver() {
    clear();
    if (world) avbd_world_destroy(world);
}

void Solver::clear() {                                                   // solver.cpp:230-238
    g_clearing = true;
    while (forces) delete forces;
    while (bodies) delete bodies;
    g_clearing = false;
    bodies = nullptr; forces = nullptr; stepIndex = 0; lastDiagnostics = Diagnostics{};
    order.clear(); userForces.clear(); shadow.clear(); uploadedBodies = 0; uploadedForces = 0; rebuild = false;
    if (world) check(avbd_clear(world), "avbd_clear");
}

void Solver::defaultParams() {                                           // solver.cpp:240-253
    dt = 1.0f / 60.0f; gravity = vec3(0.0f, -10.0f, 0.0f); iterations = 10; alpha = 0.95f; beta = 100000.0f; gamma = 0.99f;
    postStabilize = false;
    if (logFrequency <= 0) logFrequency = 60;
    lastDiagnostics = Diagnostics{};
}

namespace {
void pack_body(const Rigid* b, float* o) {
    o[0] = b->position.x; o[1] = b->position.y; o[2] = b->position.z;
    o[3] = b->orientation.x; o[4] = b->orientation.y; o[5] = b->orientation.z; o[6] = b->orientation.w;
    o[7] = b->linearVelocity.x; o[8] = b->linearVelocity.y; o[9] = b->linearVelocity.z;
    o[10] = b->angularVelocity.x; o[11] = b->angularVelocity.y; o[12] = b->angularVelocity.z;
}
}

// Everything the host side created or edited since the device last saw it: parameters, moved bodies, new bodies, new
// user forces.  Runs before every step and before Solver::pick (the reference's pick works on a solver that never stepped).
void Solver::syncToDevice() {
    if (!world) { world = avbd_world_create(device); if (!world) die("avbd_world_create"); }
    if (rebuild) {               // something was deleted: start the device world over (warm-start history is lost)
        check(avbd_clear(world), "avbd_clear");
        uploadedBodies = 0; uploadedForces = 0; shadow.clear(); rebuild = false;
    }
    const float g[3] = {gravity.x, gravity.y, gravity.z};
    check(avbd_set_params(world, dt, g, iterations, alpha, beta, gamma, postStabilize ? 1 : 0), "avbd_set_params");

    int n = (int)order.size();
    // 1. host edits of already-uploaded bodies (the GUI moves / re-spins bodies between steps, main.cpp:88-142)
    if (uploadedBodies > 0) {
        std::vector<float> cur((size_t)uploadedBodies * 13);
        for (int i = 0; i < uploadedBodies; ++i) pack_body(order[i], &cur[(size_t)i * 13]);
        if (std::memcmp(cur.data(), shadow.data(), cur.size() * sizeof(float)) != 0) {
            std::memcpy(shadow.data(), cur.data(), cur.size() * sizeof(float));
            check(avbd_upload_state(world, shadow.data()), "avbd_upload_state");
        }
    }
    // 2. bodies created since the last step (append-only)
    if (n > uploadedBodies) {
        int k = n - uploadedBodies;
        std::vector<float> size(3 * (size_t)k), dens(k), fric(k), pos(3 * (size_t)k), rot(4 * (size_t)k), lin(3 * (size_t)k), ang(3 * (size_t)k);
        for (int j = 0; j < k; ++j) {
            const Rigid* b = order[uploadedBodies + j];
            for (int c = 0; c < 3; ++c) { size[3 * j + c] = b->size[c]; pos[3 * j + c] = b->position[c]; lin[3 * j + c] = b->linearVelocity[c]; ang[3 * j + c] = b->angularVelocity[c]; }
            rot[4 * j] = b->orientation.x; rot[4 * j + 1] = b->orientation.y; rot[4 * j + 2] = b->orientation.z; rot[4 * j + 3] = b->orientation.w;
            dens[j] = b->density; fric[j] = b->friction;
        }
        check(avbd_add_bodies(world, k, size.data(), dens.data(), fric.data(), pos.data(), rot.data(), lin.data(), ang.data(), nullptr), "avbd_add_bodies");
        shadow.resize((size_t)n * 13);
        for (int i = uploadedBodies; i < n; ++i) pack_body(order[i], &shadow[(size_t)i * 13]);
        uploadedBodies = n;
    }
    // 3. user forces created since the last step: every Force in the solver's list that is not a Manifold mirror and not
    //    registered yet (the list is newest first; whatever subclass it is, it is found here — header-only ones included)
    {
        std::vector<Force*> fresh;
        for (Force* f = forces; f; f = f->next)
            if (f->deviceKind() != 3 && std::find(userForces.begin(), userForces.end(), f) == userForces.end()) fresh.push_back(f);
        userForces.insert(userForces.end(), fresh.rbegin(), fresh.rend());
    }
    for (; uploadedForces < (int)userForces.size(); ++uploadedForces) {
        Force* f = userForces[uploadedForces];
        int a = f->bodyA ? f->bodyA->index : -1, b = f->bodyB->index;
        if (f->deviceKind() == 0) {
            Joint* j = static_cast<Joint*>(f);
            const float aa[3] = {j->rA.x, j->rA.y, j->rA.z}, bb[3] = {j->rB.x, j->rB.y, j->rB.z};
            check(avbd_add_joint(world, a, b, aa, bb, j->stiffness[0], j->stiffness[3]), "avbd_add_joint");
        } else if (f->deviceKind() == 1) {
            Spring* s = static_cast<Spring*>(f);
            const float aa[3] = {s->rA.x, s->rA.y, s->rA.z}, bb[3] = {s->rB.x, s->rB.y, s->rB.z};
            check(avbd_add_spring(world, a, b, aa, bb, s->springStiffness, s->restLength), "avbd_add_spring");
        } else if (f->deviceKind() == 2) {
            check(avbd_add_ignore(world, a, b), "avbd_add_ignore");
        } else {
            std::fprintf(stderr, "avbd-demo3d_b200: user-defined Force subclasses cannot run on the device (no CPU fallback)\n");
            std::exit(2);
        }
    }

}

void Solver::step() {                                                    // solver.cpp:255-514, on the device
    syncToDevice();
    int n = (int)order.size();
    ++stepIndex;
    check(avbd_step(world, 1), "avbd_step");
    if (n > 0) {
        check(avbd_download_state(world, shadow.data()), "avbd_download_state");
        for (int i = 0; i < n; ++i) {
            Rigid* b = order[i]; const float* o = &shadow[(size_t)i * 13];
            if (b->invMass > 0.0f) { b->prevLinearVelocity = b->linearVelocity; b->prevAngularVelocity = b->angularVelocity; }
            b->position = vec3(o[0], o[1], o[2]); b->orientation = quat(o[3], o[4], o[5], o[6]);
            b->linearVelocity = vec3(o[7], o[8], o[9]); b->angularVelocity = vec3(o[10], o[11], o[12]);
        }
    }
    avbd_diagnostics d;
    check(avbd_get_diagnostics(world, &d), "avbd_get_diagnostics");
    lastDiagnostics = Diagnostics{d.maxPenetration, d.maxConstraintViolation, d.maxLinearSpeed, d.maxAngularSpeed, d.maxNormalImpulse,
                                  d.activeContacts, d.activeManifolds, d.dynamicBodies};
    if (d.nanEvents > 0)         // solver.cpp:51-66 prints per body; the device counts scrubs instead
        std::printf("[Physics] Warning: %d non-finite body state value(s) were reset this step.\n", d.nanEvents);
    if (enableDiagnostics) {                                             // solver.cpp:499-512
        int fq = logFrequency > 0 ? logFrequency : 1;
        if (stepIndex % fq == 0)
            std::printf("[Physics] step %d | manifolds: %d | contacts: %d | dyn bodies: %d | maxPen: %.6f | maxDrift: %.6f | maxLin: %.3f | maxAng: %.3f | maxLambda: %.3f\n",
                        stepIndex, d.activeManifolds, d.activeContacts, d.dynamicBodies, d.maxPenetration, d.maxConstraintViolation,
                        d.maxLinearSpeed, d.maxAngularSpeed, d.maxNormalImpulse);
    }
}

void Solver::refreshManifolds() {
    for (Force* f = forces; f;) { Force* nx = f->next; if (f->isManifold()) delete f; f = nx; }
    if (!world) return;
    int slots = avbd_num_manifolds(world);
    if (slots <= 0) return;
    std::vector<int> ints(3 * (size_t)slots), feats(4 * (size_t)slots), stick(4 * (size_t)slots);
    std::vector<float> flts(81 * (size_t)slots);
    int live = avbd_download_manifolds(world, ints.data(), feats.data(), stick.data(), flts.data());
    check(live, "avbd_download_manifolds");
    for (int m = 0; m < live; ++m) {
        Manifold* mf = new Manifold(this, order[ints[3 * m]], order[ints[3 * m + 1]]);
        const float* f = &flts[81 * (size_t)m];
        mf->numContacts = ints[3 * m + 2]; mf->combinedFriction = f[0];
        for (int c = 0; c < mf->numContacts; ++c) {
            const float* g = f + 1 + 14 * c;
            Manifold::Contact& k = mf->contacts[c];
            k.feature.value = feats[4 * m + c]; k.stick = stick[4 * m + c] != 0;
            k.rA = vec3(g[0], g[1], g[2]); k.rB = vec3(g[3], g[4], g[5]); k.normal = vec3(g[6], g[7], g[8]);
            k.C0_n = g[10]; k.C0_t = vec3(g[11], g[12], 0.0f);
            vec3 xA = mf->bodyA->position + rotate(mf->bodyA->orientation, k.rA), xB = mf->bodyB->position + rotate(mf->bodyB->orientation, k.rB);
            k.penetration = max(0.0f, -dot(xA - xB, k.normal));
            for (int r = 0; r < 3; ++r) { mf->lambda[c * 3 + r] = f[57 + c * 3 + r]; mf->penalty[c * 3 + r] = f[69 + c * 3 + r]; mf->stiffness[c * 3 + r] = FLT_MAX; }
        }
    }
}

void Solver::draw() { refreshManifolds(); }     // GL drawing is out of scope; keeping the mirrors fresh is what a renderer would need

Rigid* Solver::pick(const vec3& origin, const vec3& dir, vec3& local) {      // solver.cpp:145-228, reduction on the device
    if (order.empty()) return nullptr;
    syncToDevice();
    const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {dir.x, dir.y, dir.z};
    float l[3] = {0, 0, 0};
    int hit = avbd_pick(world, o, d, l);
    if (hit < 0) return nullptr;
    local = vec3(l[0], l[1], l[2]);
    return order[hit];
}
