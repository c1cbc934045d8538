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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <sys/mman.h>

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

namespace {
// rows of the state arena as the typed fields of a Rigid (vec3 / quat are plain float triples / quadruples)
inline vec3& row_vec3(float* p) { return *reinterpret_cast<vec3*>(p); }
inline quat& row_quat(float* p) { return *reinterpret_cast<quat*>(p); }
static_assert(sizeof(vec3) == 3 * sizeof(float) && sizeof(quat) == 4 * sizeof(float), "arena rows alias vec3 / quat");
}

Rigid::Rigid(Solver* s, const vec3& sz, float dens, float fric, const vec3& pos, const quat& orient, const vec3& linVel, const vec3& angVel)
    : solver(s), forces(nullptr), next(nullptr), id(next_id++), slot(-1),
      position(row_vec3(s->claimRow(slot))), orientation(row_quat(s->arenaRows + 13 * (size_t)slot + 3)),
      linearVelocity(row_vec3(s->arenaRows + 13 * (size_t)slot + 7)), angularVelocity(row_vec3(s->arenaRows + 13 * (size_t)slot + 10)),
      prevLinearVelocity(row_vec3(s->arenaPrev + 6 * (size_t)slot)), prevAngularVelocity(row_vec3(s->arenaPrev + 6 * (size_t)slot + 3)),
      size(sz), friction(fric), deviceIndex(-1), density(dens) {
    position = pos; orientation = orient; linearVelocity = linVel; angularVelocity = angVel; prevLinearVelocity = linVel; prevAngularVelocity = angVel;
    next = s->bodies; s->bodies = this;
    index = (int)s->order.size(); s->order.push_back(this);
    mass = size.x * size.y * size.z * density;                               // rigid.cpp:24-40
    invMass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
    radius = length(size) * 0.5f;
    s->arenaDynamic[(size_t)slot] = invMass > 0.0f ? 1 : 0;
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
    if (!g_clearing) {           // a single body removed: the device world is re-created at the next step (manifolds of the others are kept)
        // what is attached to the body goes with it: its Manifold mirrors, and user forces that would otherwise point at freed
        // memory (upstream leaves them dangling, force.cpp:43-69 would then walk a dead list)
        while (forces) delete forces;
        if (deviceIndex >= 0 && deviceIndex < (int)solver->deviceOrder.size()) solver->deviceOrder[deviceIndex] = nullptr;
        solver->order.erase(std::find(solver->order.begin(), solver->order.end(), this));
        for (size_t i = 0; i < solver->order.size(); ++i) solver->order[i]->index = (int)i;
        solver->rebuild = true;
        solver->arenaDense = false;      // rows and device order differ from now on (until clear())
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
      device(0), rebuild(false), readBack(true), uploadAll(false), shadow(nullptr), shadowCap(0), arenaRows(nullptr), arenaPrev(nullptr),
      arenaCapBodies(0), arenaPinnedBodies(0), arenaCount(0), arenaDense(true), uploadedBodies(0), uploadedForces(0),
      mirrorsFresh(false), uploadedBytes(0), downloadedBytes(0), hostSyncSec(0), deviceStepSec(0), hostFetchSec(0) {
    if (const char* d = std::getenv("AVBD_DEVICE")) device = std::atoi(d);
    defaultParams();
}

Solver::~Solver() {
    clear();
    if (shadow) avbd_host_free(shadow);
    if (arenaRows) {
        if (arenaPinnedBodies) avbd_host_unregister(arenaRows);
        munmap(arenaRows, arenaCapBodies * 13 * sizeof(float)); munmap(arenaPrev, arenaCapBodies * 6 * sizeof(float));
    }
    if (world) avbd_world_destroy(world);
}

// The arena is address space first, memory later: 2^26 rows (3.5 GB + 1.6 GB of virtual range, MAP_NORESERVE) that the kernel backs
// page by page as bodies are created.  Rows never move, so the references inside every Rigid stay valid however many bodies follow.
float* Solver::claimRow(int& slot) {
    if (!arenaRows) {
        arenaCapBodies = (size_t)1 << 26;
        if (const char* e = std::getenv("AVBD_HOST_MAX_BODIES")) arenaCapBodies = std::max<size_t>(1024, (size_t)std::atoll(e));
        void* r = mmap(nullptr, arenaCapBodies * 13 * sizeof(float), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        void* q = mmap(nullptr, arenaCapBodies * 6 * sizeof(float), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (r == MAP_FAILED || q == MAP_FAILED) { std::fprintf(stderr, "avbd-demo3d_b200: cannot reserve the host state arena (AVBD_HOST_MAX_BODIES lowers it)\n"); std::exit(2); }
        arenaRows = static_cast<float*>(r); arenaPrev = static_cast<float*>(q);
    }
    if ((size_t)arenaCount >= arenaCapBodies) { std::fprintf(stderr, "avbd-demo3d_b200: host state arena is full (AVBD_HOST_MAX_BODIES)\n"); std::exit(2); }
    slot = arenaCount++;
    if (arenaDynamic.size() < (size_t)arenaCount) arenaDynamic.resize(std::max<size_t>(1024, 2 * (size_t)arenaCount), 0);
    return arenaRows + 13 * (size_t)slot;
}

// Page-locks the arena's first `bodies` rows (doubling, so a growing world re-registers O(log n) times).
void Solver::pinArena(size_t bodies) {
    if (bodies <= arenaPinnedBodies || !arenaRows) return;
    size_t want = std::min(arenaCapBodies, std::max<size_t>(2 * arenaPinnedBodies, std::max<size_t>(bodies, 4096)));
    if (arenaPinnedBodies) avbd_host_unregister(arenaRows);
    size_t bytes = (want * 13 * sizeof(float) + 4095) / 4096 * 4096;
    if (avbd_host_register(arenaRows, (long long)bytes) < 0) { arenaPinnedBodies = 0; return; }     // still correct: the copies go through the driver's staging buffer
    arenaPinnedBodies = want;
}

void Solver::clear() {                                                   // solver.cpp:230-238
    g_clearing = true;
    while (forces) delete forces;
    while (bodies) delete bodies;
    g_clearing = false;
    bodies = nullptr; forces = nullptr; stepIndex = 0; lastDiagnostics = Diagnostics{};
    order.clear(); deviceOrder.clear(); userForces.clear(); userForceSlot.clear(); rowShadow.clear(); mirrorRows.clear();
    uploadedBodies = 0; uploadedForces = 0; rebuild = false; mirrorsFresh = false;
    arenaCount = 0; arenaDense = true;
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
// Slices [begin, end) of n items over a few host threads (a million Rigid nodes are ~270 MB of heap: one thread walking them was
// most of Solver::step()'s host time).  fn(begin, end, slice).  The workers are created once and parked on a condition variable:
// spawning 16 threads per pass cost more than a small world's whole step.
class SlicePool {
public:
    static SlicePool& get() { static SlicePool p; return p; }
    int workers() const { return (int)threads_.size() + 1; }
    void run(int n, int slices, const std::function<void(int, int, int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; n_ = n; slices_ = slices; next_ = 1; pending_ = slices - 1; ++gen_;
        }
        cv_.notify_all();
        fn(0, (int)((long long)n / slices), 0);                        // the caller takes slice 0
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    SlicePool() {
        unsigned hw = std::thread::hardware_concurrency();
        int want = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
        if (const char* e = std::getenv("AVBD_HOST_THREADS")) want = std::max(1, std::atoi(e));
        for (int t = 1; t < want; ++t) threads_.emplace_back([this] { loop(); });
    }
    ~SlicePool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            while (fn_ && next_ < slices_) {
                int t = next_++;
                const std::function<void(int, int, int)>* fn = fn_;
                int b = (int)((long long)n_ * t / slices_), e = (int)((long long)n_ * (t + 1) / slices_);
                lk.unlock();
                (*fn)(b, e, t);
                lk.lock();
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int, int)>* fn_ = nullptr;
    int n_ = 0, slices_ = 0, next_ = 0, pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};
template <class Fn>
void parallel_slices(int n, int& slices, Fn fn) {
    if (n < 65536) { slices = 1; fn(0, n, 0); return; }
    SlicePool& pool = SlicePool::get();
    slices = pool.workers();
    if (slices == 1) { fn(0, n, 0); return; }
    std::function<void(int, int, int)> f = fn;
    pool.run(n, slices, f);
}
void ensure_shadow(Solver* s, size_t bodies) {
    if (bodies * 13 <= s->shadowCap) return;
    size_t cap = std::max<size_t>(bodies + bodies / 2, 1024) * 13;
    float* p = static_cast<float*>(avbd_host_alloc((long long)(cap * sizeof(float))));
    if (!p) die("avbd_host_alloc");
    if (s->shadow) { std::memcpy(p, s->shadow, s->shadowCap * sizeof(float)); avbd_host_free(s->shadow); }
    s->shadow = p; s->shadowCap = cap;
}
void rows_of(const Force* f, float* o48) {
    std::memcpy(o48, f->lambda, 12 * sizeof(float)); std::memcpy(o48 + 12, f->penalty, 12 * sizeof(float));
    std::memcpy(o48 + 24, f->motor, 12 * sizeof(float)); std::memcpy(o48 + 36, f->stiffness, 12 * sizeof(float));
}
}

// Everything the host side created or edited since the device last saw it: parameters, moved bodies, new bodies, new
// user forces, edited Force rows.  Runs before every step and before Solver::pick (the reference's pick works on a solver
// that never stepped).
void Solver::syncToDevice() {
    if (!world) { world = avbd_world_create(device); if (!world) die("avbd_world_create"); }
    std::vector<int> keepInts, keepFeats, keepStick; std::vector<float> keepFlts; int keepCount = -1;
    std::vector<Rigid*> oldDeviceOrder;
    if (rebuild) {
        // A body or a user force was deleted (force.cpp:43-69 / ~Rigid are O(degree) upstream): the device world is re-created
        // from the host mirror, and everything that carries warm-start history goes with it — the manifold set (re-indexed;
        // manifolds of a deleted body are dropped, as upstream deletes them with the body) and the user forces' rows.
        int slots = avbd_num_manifolds(world);
        if (slots > 0) {
            keepInts.resize(3 * (size_t)slots); keepFeats.resize(4 * (size_t)slots); keepStick.resize(4 * (size_t)slots); keepFlts.resize(81 * (size_t)slots);
            keepCount = avbd_download_manifolds(world, keepInts.data(), keepFeats.data(), keepStick.data(), keepFlts.data());
            check(keepCount, "avbd_download_manifolds");
        }
        oldDeviceOrder = deviceOrder;
        check(avbd_clear(world), "avbd_clear");
        uploadedBodies = 0; uploadedForces = 0; deviceOrder.clear(); userForceSlot.clear(); rowShadow.clear(); rebuild = false;
        for (Rigid* b : order) b->deviceIndex = -1;
    }
    const float g[3] = {gravity.x, gravity.y, gravity.z};
    check(avbd_set_params(world, dt, g, iterations, alpha, beta, gamma, postStabilize ? 1 : 0), "avbd_set_params");

    int n = (int)order.size();
    const bool dense = arenaDense;            // no body deleted since clear(): arena row == creation index == device index
    if (dense) {
        pinArena((size_t)n);
        if (arenaShadow.size() < 13 * (size_t)n) arenaShadow.resize(13 * std::max<size_t>((size_t)n + (size_t)n / 2, 1024));
    } else {
        ensure_shadow(this, (size_t)n);
    }
    // 1. host edits of already-uploaded bodies (the GUI moves / re-spins bodies between steps, main.cpp:88-142): each slice
    //    uploads the range between its first and last edited body — nothing at all in a headless run
    if (uploadedBodies > 0 && dense) {
        // the rows ARE the bodies' fields: one sequential compare against what the device last saw, uploads straight from the arena
        int slices = 1;
        std::vector<int> lo(64, -1), hi(64, -1);
        const bool all = uploadAll;
        parallel_slices(uploadedBodies, slices, [&](int b, int e, int t) {
            const float* rows = arenaRows + 13 * (size_t)b; float* sh = arenaShadow.data() + 13 * (size_t)b;
            int first = -1, last = -1;
            if (all) { first = b; last = e - 1; }
            else if (e > b && std::memcmp(rows, sh, 52 * (size_t)(e - b)) != 0) {
                for (int i = b; i < e; ++i) if (std::memcmp(rows + 13 * (size_t)(i - b), sh + 13 * (size_t)(i - b), 52) != 0) { if (first < 0) first = i; last = i; }
            }
            if (first >= 0) std::memcpy(sh + 13 * (size_t)(first - b), rows + 13 * (size_t)(first - b), 52 * (size_t)(last - first + 1));
            lo[t] = first; hi[t] = last;
        });
        for (int t = 0; t < slices; ++t) {
            if (lo[t] < 0) continue;
            int first = lo[t], last = hi[t];
            while (t + 1 < slices && lo[t + 1] == last + 1) last = hi[++t];      // edited ranges that touch travel as one transfer
            int cnt = last - first + 1;
            check(avbd_upload_state_range(world, first, cnt, arenaRows + 13 * (size_t)first), "avbd_upload_state_range");
            uploadedBytes += (long long)cnt * 52;
        }
    } else if (uploadedBodies > 0) {
        int slices = 1;
        std::vector<int> lo(16, -1), hi(16, -1);
        const bool all = uploadAll;
        parallel_slices(uploadedBodies, slices, [&](int b, int e, int t) {
            float cur[13];
            int first = -1, last = -1;
            for (int i = b; i < e; ++i) {
                pack_body(order[i], cur);
                float* sh = shadow + (size_t)i * 13;
                if (all || std::memcmp(cur, sh, sizeof(cur)) != 0) { std::memcpy(sh, cur, sizeof(cur)); if (first < 0) first = i; last = i; }
            }
            lo[t] = first; hi[t] = last;
        });
        for (int t = 0; t < slices; ++t)
            if (lo[t] >= 0) {
                int cnt = hi[t] - lo[t] + 1;
                check(avbd_upload_state_range(world, lo[t], cnt, shadow + (size_t)lo[t] * 13), "avbd_upload_state_range");
                uploadedBytes += (long long)cnt * 52;
            }
    }
    // 2. bodies created since the last step (append-only)
    bool reAdded = false;
    if (n > uploadedBodies) {
        int k = n - uploadedBodies;
        std::vector<float> size(3 * (size_t)k), dens(k), fric(k), pos(3 * (size_t)k), rot(4 * (size_t)k), lin(3 * (size_t)k), ang(3 * (size_t)k), prev(3 * (size_t)k);
        bool prevDiffers = false;
        for (int j = 0; j < k; ++j) {
            Rigid* b = order[uploadedBodies + j];
            for (int c = 0; c < 3; ++c) {
                size[3 * j + c] = b->size[c]; pos[3 * j + c] = b->position[c]; lin[3 * j + c] = b->linearVelocity[c]; ang[3 * j + c] = b->angularVelocity[c];
                prev[3 * j + c] = b->prevLinearVelocity[c];
                if (b->prevLinearVelocity[c] != b->linearVelocity[c]) prevDiffers = true;
            }
            rot[4 * j] = b->orientation.x; rot[4 * j + 1] = b->orientation.y; rot[4 * j + 2] = b->orientation.z; rot[4 * j + 3] = b->orientation.w;
            dens[j] = b->density; fric[j] = b->friction;
            b->deviceIndex = uploadedBodies + j;
        }
        check(avbd_add_bodies(world, k, size.data(), dens.data(), fric.data(), pos.data(), rot.data(), lin.data(), ang.data(), nullptr), "avbd_add_bodies");
        deviceOrder.insert(deviceOrder.end(), order.begin() + uploadedBodies, order.end());
        if (prevDiffers && uploadedBodies == 0) {      // a re-created world: the adaptive gravity weight (solver.cpp:318-326) needs the real previous velocities
            check(avbd_upload_prev_linvel(world, prev.data()), "avbd_upload_prev_linvel");
        }
        if (dense) std::memcpy(arenaShadow.data() + 13 * (size_t)uploadedBodies, arenaRows + 13 * (size_t)uploadedBodies, 52 * (size_t)k);
        else for (int i = uploadedBodies; i < n; ++i) pack_body(order[i], shadow + (size_t)i * 13);
        uploadedBytes += (long long)k * 52;
        reAdded = uploadedBodies == 0;
        uploadedBodies = n;
    }
    // 2b. the manifold set a re-created world inherits
    if (keepCount > 0 && reAdded) {
        int live = 0;
        for (int m = 0; m < keepCount; ++m) {
            Rigid* a = oldDeviceOrder[keepInts[3 * m]]; Rigid* b = oldDeviceOrder[keepInts[3 * m + 1]];
            if (!a || !b) continue;
            keepInts[3 * live] = a->deviceIndex; keepInts[3 * live + 1] = b->deviceIndex; keepInts[3 * live + 2] = keepInts[3 * m + 2];
            std::memmove(&keepFeats[4 * (size_t)live], &keepFeats[4 * (size_t)m], 4 * sizeof(int));
            std::memmove(&keepStick[4 * (size_t)live], &keepStick[4 * (size_t)m], 4 * sizeof(int));
            std::memmove(&keepFlts[81 * (size_t)live], &keepFlts[81 * (size_t)m], 81 * sizeof(float));
            ++live;
        }
        check(avbd_upload_manifolds(world, live, keepInts.data(), keepFeats.data(), keepStick.data(), keepFlts.data()), "avbd_upload_manifolds");
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
        int a = f->bodyA ? f->bodyA->deviceIndex : -1, b = f->bodyB->deviceIndex;
        int slot = -1;
        if (f->deviceKind() == 0) {
            // the construction-time anchor / reference orientation the Joint already holds (joint.cpp:19, :47-50), NOT values
            // re-derived from the current poses: a joint that is re-uploaded keeps the error it has accumulated
            Joint* j = static_cast<Joint*>(f);
            const float aa[3] = {j->rA.x, j->rA.y, j->rA.z}, bb[3] = {j->rB.x, j->rB.y, j->rB.z};
            const float q[4] = {j->initialRelativeOrientation.x, j->initialRelativeOrientation.y, j->initialRelativeOrientation.z, j->initialRelativeOrientation.w};
            slot = avbd_add_joint_raw(world, a, b, aa, bb, q, j->stiffness[0], j->stiffness[3]);
            check(slot, "avbd_add_joint_raw");
        } else if (f->deviceKind() == 1) {
            Spring* sp = static_cast<Spring*>(f);
            const float aa[3] = {sp->rA.x, sp->rA.y, sp->rA.z}, bb[3] = {sp->rB.x, sp->rB.y, sp->rB.z};
            slot = avbd_add_spring(world, a, b, aa, bb, sp->stiffness[0], sp->restLength);
            check(slot, "avbd_add_spring");
        } else if (f->deviceKind() == 2) {
            check(avbd_add_ignore(world, a, b), "avbd_add_ignore");
        } else {
            std::fprintf(stderr, "avbd-demo3d_b200: user-defined Force subclasses cannot run on the device (no CPU fallback)\n");
            std::exit(2);
        }
        userForceSlot.push_back(slot);
        // device rows start as the constructor left them (lambda 0, penalty PENALTY_MIN, motor 0); anything else the host holds
        // (rows edited before the first step, or the history of a re-created world) is an edit and is uploaded below
        rowShadow.resize(rowShadow.size() + 48, 0.0f);
        float* sh = &rowShadow[48 * (size_t)uploadedForces];
        int rows = f->getRowCount();
        for (int r = 0; r < 12; ++r) { sh[r] = 0.0f; sh[12 + r] = r < rows ? PENALTY_MIN : f->penalty[r]; sh[24 + r] = 0.0f; sh[36 + r] = f->stiffness[r]; }
        if (f->deviceKind() == 0) for (int r = 0; r < 6; ++r) sh[36 + r] = f->stiffness[r < 3 ? 0 : 3];
    }
    // 4. host edits of the public row arrays of user forces (solver.h:91-97): lambda, penalty, motor (enters the primal at
    //    solver.cpp:380), stiffness (hard / soft, :290, :378, :416)
    for (size_t k = 0; k < userForces.size(); ++k) {
        int slot = userForceSlot[k];
        if (slot < 0) continue;
        Force* f = userForces[k];
        int rows = f->getRowCount();
        float cur[48]; rows_of(f, cur);
        float* sh = &rowShadow[48 * k];
        bool edited = false;
        for (int part = 0; part < 4 && !edited; ++part) edited = std::memcmp(cur + 12 * part, sh + 12 * part, rows * sizeof(float)) != 0;
        if (!edited) continue;
        check(avbd_set_force_rows(world, f->deviceKind(), slot, f->lambda, f->penalty, f->motor, f->stiffness), "avbd_set_force_rows");
        std::memcpy(sh, cur, sizeof(cur));
    }
    // 5. host edits of Manifold mirrors' lambda / penalty (only meaningful while the mirrors are the device's current set)
    if (mirrorsFresh) {
        std::vector<int> ints, feats, stick; std::vector<float> flts;
        bool edited = false; size_t m = 0;
        for (Force* f = forces; f; f = f->next) {
            if (!f->isManifold()) continue;
            if (24 * (m + 1) > mirrorRows.size()) { edited = false; break; }
            if (std::memcmp(f->lambda, &mirrorRows[24 * m], 12 * sizeof(float)) || std::memcmp(f->penalty, &mirrorRows[24 * m + 12], 12 * sizeof(float))) edited = true;
            ++m;
        }
        if (edited) {
            for (Force* f = forces; f; f = f->next) {
                if (!f->isManifold()) continue;
                Manifold* mf = static_cast<Manifold*>(f);
                if (mf->numContacts <= 0) continue;
                ints.push_back(mf->bodyA->deviceIndex); ints.push_back(mf->bodyB->deviceIndex); ints.push_back(mf->numContacts);
                size_t fo = flts.size(); flts.resize(fo + 81, 0.0f);
                flts[fo] = mf->combinedFriction;
                for (int c = 0; c < 4; ++c) {
                    bool on = c < mf->numContacts; const Manifold::Contact& k = mf->contacts[c];
                    feats.push_back(on ? k.feature.value : 0); stick.push_back(on && k.stick ? 1 : 0);
                    if (!on) continue;
                    float* g = &flts[fo + 1 + 14 * c];
                    g[0] = k.rA.x; g[1] = k.rA.y; g[2] = k.rA.z; g[3] = k.rB.x; g[4] = k.rB.y; g[5] = k.rB.z; g[6] = k.normal.x; g[7] = k.normal.y; g[8] = k.normal.z;
                    g[9] = k.penetration; g[10] = k.C0_n; g[11] = k.C0_t.x; g[12] = k.C0_t.y; g[13] = 0.0f;
                    for (int r = 0; r < 3; ++r) { flts[fo + 57 + 3 * c + r] = mf->lambda[3 * c + r]; flts[fo + 69 + 3 * c + r] = mf->penalty[3 * c + r]; }
                }
            }
            check(avbd_upload_manifolds(world, (int)(ints.size() / 3), ints.data(), feats.data(), stick.data(), flts.data()), "avbd_upload_manifolds");
        }
    }
}

void Solver::fetchState() { fetchStateImpl(false); }
const float* Solver::hostState() const { return arenaDense ? arenaRows : shadow; }

// device -> Rigid fields; advancePrev: the step just taken makes the old velocities the "previous" ones (solver.cpp:457-458) — done
// in the same walk over the bodies, every Rigid is touched once.
void Solver::fetchStateImpl(bool advancePrev) {
    int n = (int)order.size();
    if (n == 0 || !world) return;
    downloadedBytes += (long long)n * 52;
    int slices = 1;
    if (arenaDense) {
        // one DMA transfer into the rows the Rigid fields alias; the previous velocities advance in a sequential pass first
        pinArena((size_t)n);
        if (arenaShadow.size() < 13 * (size_t)n) arenaShadow.resize(13 * std::max<size_t>((size_t)n + (size_t)n / 2, 1024));
        if (advancePrev)
            parallel_slices(n, slices, [&](int b0, int e0, int) {
                for (int i = b0; i < e0; ++i) if (arenaDynamic[(size_t)i]) std::memcpy(arenaPrev + 6 * (size_t)i, arenaRows + 13 * (size_t)i + 7, 6 * sizeof(float));
            });
        // four pieces (at least 128K bodies each): a piece is copied to the edit-detection shadow while the next one is still on the bus
        struct Landed { Solver* s; } ctx{this};
        check(avbd_download_state_chunked(world, arenaRows, std::max(131072, (n + 3) / 4), [](int first, int count, void* user) {
            Solver* s = static_cast<Landed*>(user)->s;
            int sl = 1;
            parallel_slices(count, sl, [&](int b0, int e0, int) {
                std::memcpy(s->arenaShadow.data() + 13 * (size_t)(first + b0), s->arenaRows + 13 * (size_t)(first + b0), 52 * (size_t)(e0 - b0));
            });
        }, &ctx), "avbd_download_state_chunked");
        return;
    }
    check(avbd_download_state(world, shadow), "avbd_download_state");
    parallel_slices(n, slices, [&](int b0, int e0, int) {
        for (int i = b0; i < e0; ++i) {
            Rigid* b = order[i]; const float* o = shadow + (size_t)i * 13;
            if (advancePrev && b->invMass > 0.0f) { b->prevLinearVelocity = b->linearVelocity; b->prevAngularVelocity = b->angularVelocity; }
            b->position = vec3(o[0], o[1], o[2]); b->orientation = quat(o[3], o[4], o[5], o[6]);
            b->linearVelocity = vec3(o[7], o[8], o[9]); b->angularVelocity = vec3(o[10], o[11], o[12]);
        }
    });
}

void Solver::step() {                                                    // solver.cpp:255-514, on the device
    const auto t0 = std::chrono::steady_clock::now();
    syncToDevice();
    const auto t1 = std::chrono::steady_clock::now();
    int n = (int)order.size();
    ++stepIndex;
    mirrorsFresh = false;
    check(avbd_step(world, 1), "avbd_step");
    const auto t2 = std::chrono::steady_clock::now();
    if (n > 0 && readBack) fetchStateImpl(true);
    const auto t3 = std::chrono::steady_clock::now();
    hostSyncSec += std::chrono::duration<double>(t1 - t0).count(); deviceStepSec += std::chrono::duration<double>(t2 - t1).count();
    hostFetchSec += std::chrono::duration<double>(t3 - t2).count();
    // lambda / penalty of the user forces' rows, as the reference leaves them in the public arrays after a step
    if (!userForces.empty()) {
        int nj = avbd_num_joints(world), ns = avbd_num_springs(world);
        std::vector<float> jr(12 * (size_t)std::max(1, nj)), sr(2 * (size_t)std::max(1, ns));
        check(avbd_download_user_rows(world, nj ? jr.data() : nullptr, ns ? sr.data() : nullptr), "avbd_download_user_rows");
        for (size_t k = 0; k < userForces.size(); ++k) {
            int slot = userForceSlot[k]; if (slot < 0) continue;
            Force* f = userForces[k]; float* sh = &rowShadow[48 * k];
            if (f->deviceKind() == 0) for (int r = 0; r < 6; ++r) { f->lambda[r] = sh[r] = jr[12 * (size_t)slot + r]; f->penalty[r] = sh[12 + r] = jr[12 * (size_t)slot + 6 + r]; }
            else { f->lambda[0] = sh[0] = sr[2 * (size_t)slot]; f->penalty[0] = sh[12] = sr[2 * (size_t)slot + 1]; }
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

std::vector<unsigned char> Solver::snapshot() {
    syncToDevice();
    long long bytes = avbd_snapshot_bytes(world);
    std::vector<unsigned char> blob((size_t)bytes);
    check(avbd_snapshot(world, blob.data(), bytes), "avbd_snapshot");
    return blob;
}

void Solver::restore(const std::vector<unsigned char>& blob) {
    if (!world) { world = avbd_world_create(device); if (!world) die("avbd_world_create"); }
    check(avbd_restore(world, blob.data(), (long long)blob.size()), "avbd_restore");
    if (avbd_num_bodies(world) != (int)order.size()) { std::fprintf(stderr, "avbd-demo3d_b200: snapshot holds a different body set\n"); std::exit(2); }
    uploadedBodies = (int)order.size(); rebuild = false; mirrorsFresh = false;
    if (!arenaDense) ensure_shadow(this, order.size());
    deviceOrder = order;
    for (size_t i = 0; i < order.size(); ++i) order[i]->deviceIndex = (int)i;
    fetchState();                                   // the host mirror shows the restored state (and is not mistaken for an edit)
    if (!order.empty()) {
        std::vector<float> prev(3 * order.size());
        check(avbd_download_prev_linvel(world, prev.data()), "avbd_download_prev_linvel");
        for (size_t i = 0; i < order.size(); ++i) order[i]->prevLinearVelocity = vec3(prev[3 * i], prev[3 * i + 1], prev[3 * i + 2]);
    }
    if (!userForces.empty()) {                      // rows as saved
        for (size_t k = 0; k < userForces.size(); ++k) {
            int slot = userForceSlot[k]; if (slot < 0) continue;
            Force* f = userForces[k];
            check(avbd_get_force_rows(world, f->deviceKind(), slot, f->lambda, f->penalty, f->motor, f->stiffness), "avbd_get_force_rows");
            rows_of(f, &rowShadow[48 * k]);
        }
    }
}

void Solver::refreshManifolds() {
    for (Force* f = forces; f;) { Force* nx = f->next; if (f->isManifold()) delete f; f = nx; }
    mirrorRows.clear(); mirrorsFresh = false;
    if (!world || rebuild) return;
    int slots = avbd_num_manifolds(world);
    if (slots <= 0) return;
    std::vector<int> ints(3 * (size_t)slots), feats(4 * (size_t)slots), stick(4 * (size_t)slots);
    std::vector<float> flts(81 * (size_t)slots);
    int live = avbd_download_manifolds(world, ints.data(), feats.data(), stick.data(), flts.data());
    check(live, "avbd_download_manifolds");
    for (int m = 0; m < live; ++m) {
        Manifold* mf = new Manifold(this, deviceOrder[ints[3 * m]], deviceOrder[ints[3 * m + 1]]);
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
    for (Force* f = forces; f; f = f->next) {          // what the mirrors hold now, in list order, to recognise host edits later
        if (!f->isManifold()) continue;
        mirrorRows.insert(mirrorRows.end(), f->lambda, f->lambda + 12);
        mirrorRows.insert(mirrorRows.end(), f->penalty, f->penalty + 12);
    }
    mirrorsFresh = true;
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
    return deviceOrder[hit];
}
