// solver.h — host-side mirror of the reference's public class API (alxspiker/avbd-demo3d source/solver.h:48-181):
// Rigid, Force (row-based interface), Manifold, Solver — same names, fields and method signatures — as a thin
// client of the C ABI in include/avbd_b200.h.  Solver::step() runs entirely on the GPU; the structs here hold
// the host-visible copy of the state (refreshed after every step) and forward construction to the device world.
// There is no CPU implementation of the step behind this API.
#pragma once
#include <vector>
#include "maths.h"

#define MAX_CONSTRAINT_ROWS 12
#define PENALTY_MIN 20000.0f
#define PENALTY_MAX 1000000000.0f
#define COLLISION_MARGIN 0.02f
#define STICK_THRESH 0.02f
#define SHOW_CONTACTS true
constexpr float PENETRATION_SLOP = 0.005f;

struct Rigid; struct Force; struct Manifold; struct Joint; struct Spring; struct IgnoreCollision; struct Solver;
struct avbd_world;

struct Rigid {                       // solver.h:48-82
    Solver* solver; Force* forces; Rigid* next;
    int id; static int next_id;
    int slot;                        // row of the solver's state arena this body's fields below live in (extension)
    // The pose / velocity fields are REFERENCES into the solver's page-locked state arena (one row of 13 floats per body, in the
    // layout of avbd_upload_state / avbd_download_state): `body->position = ...`, `body->linearVelocity.y`, `&body->position` all read
    // and write as upstream's plain members do, while Solver::step() exchanges the whole body set with the device as ONE DMA
    // transfer to / from that arena instead of walking a million heap objects twice per step.
    vec3& position; quat& orientation;
    vec3& linearVelocity; vec3& angularVelocity; vec3& prevLinearVelocity; vec3& prevAngularVelocity;
    vec3 initialPosition; quat initialOrientation; vec3 inertialPosition; quat inertialOrientation;
    vec3 size; float mass, invMass; mat3 inertiaTensor, invInertiaTensor; float friction, radius;
    int index;                       // position among the solver's live bodies, creation order (extension)
    int deviceIndex;                 // index inside the device world, -1 until uploaded (extension)
    float density;                   // kept so the body can be re-uploaded (extension)

    Rigid(Solver* solver, const vec3& size, float density, float friction, const vec3& pos, const quat& orient = quat(),
          const vec3& linVel = vec3(), const vec3& angVel = vec3());
    ~Rigid();
    mat3 getInvInertiaTensorWorld() const;
    mat3 getInertiaTensorWorld() const;
    bool isConstrainedTo(Rigid* other) const;
    void draw() const {}             // rendering is out of scope on a GPU box
};

struct Force {                       // solver.h:85-109
    Solver* solver; Rigid* bodyA; Rigid* bodyB; Force *nextA, *nextB, *next;
    float C[MAX_CONSTRAINT_ROWS], fmin[MAX_CONSTRAINT_ROWS], fmax[MAX_CONSTRAINT_ROWS], lambda[MAX_CONSTRAINT_ROWS],
        penalty[MAX_CONSTRAINT_ROWS], motor[MAX_CONSTRAINT_ROWS], stiffness[MAX_CONSTRAINT_ROWS], fracture[MAX_CONSTRAINT_ROWS];
    Force(Solver* solver, Rigid* bodyA, Rigid* bodyB);
    virtual ~Force();
    virtual int getRowCount() const = 0;
    virtual bool initialize() = 0;
    virtual void computeConstraint(float alpha) = 0;
    virtual void computeDerivatives(vec3& J_linear, vec3& J_angular, const Rigid* body, int row) const = 0;
    virtual void draw() const {}
    virtual bool isManifold() const { return false; }
    virtual int deviceKind() const { return -1; }   // 0 joint, 1 spring, 2 ignore, 3 manifold mirror (extension)
};

struct Manifold : Force {            // solver.h:112-143 — host mirror of a device manifold (see Solver::refreshManifolds)
    union FeaturePair { struct { unsigned char in_A, out_A, in_B, out_B; } e; int value; };
    struct Contact { FeaturePair feature; vec3 rA, rB, normal; float penetration, C0_n; vec3 C0_t; bool stick; };
    Contact contacts[4]; int numContacts; float combinedFriction;
    Manifold(Solver* solver, Rigid* bodyA, Rigid* bodyB);
    int getRowCount() const override { return numContacts * 3; }
    bool initialize() override;
    void computeConstraint(float alpha) override;
    void computeDerivatives(vec3& J_linear, vec3& J_angular, const Rigid* body, int row) const override;
    void draw() const override {}
    static int collide(Rigid* bodyA, Rigid* bodyB, Contact* contacts, bool flip);
    bool isManifold() const override { return true; }
    int deviceKind() const override { return 3; }
};

struct Solver {                      // solver.h:146-181
    float dt; vec3 gravity; int iterations; float alpha, beta, gamma; bool postStabilize;
    Rigid* bodies; Force* forces;
    struct Diagnostics { float maxPenetration, maxConstraintViolation, maxLinearSpeed, maxAngularSpeed, maxNormalImpulse;
                         int activeContacts, activeManifolds, dynamicBodies; };
    bool enableDiagnostics; int logFrequency; int stepIndex; Diagnostics lastDiagnostics;

    Solver();
    ~Solver();
    Rigid* pick(const vec3& origin, const vec3& dir, vec3& local);
    void clear();
    void defaultParams();
    void step();
    void draw();

    // ---- extensions (not in the reference) ----
    void refreshManifolds();         // rebuilds the Manifold mirrors in `forces` from the device (contacts, lambda, penalty)
    void syncToDevice();             // uploads parameters, host edits (bodies, Force rows, Manifold rows), new bodies and new user forces
    // Snapshot / restore of the DEVICE state for a solver whose body and force set is unchanged since the snapshot
    // (SURVEY.md section 8f-3): poses, velocities, manifolds with lambda / penalty / anchors, user-force rows.
    std::vector<unsigned char> snapshot();
    void restore(const std::vector<unsigned char>& blob);
    avbd_world* world;               // the device world behind this Solver
    int device;
    bool rebuild;                    // a body or force was deleted: the device world is re-created at the next step (manifolds and rows kept)
    bool readBack;                   // true (default): step() refreshes every Rigid from the device; false: only on fetchState()
    bool uploadAll;                  // treat every body as edited before each step (exercises the full-upload path)
    void fetchState();               // device -> Rigid fields (what step() does when readBack is set)
    void fetchStateImpl(bool advancePrev);
    const float* hostState() const;  // 13 floats per body in device order, as of the last exchange with the device
    std::vector<Rigid*> order;       // live bodies by creation order
    std::vector<Rigid*> deviceOrder; // bodies as the device world indexes them (nullptr: deleted since the upload)
    float* shadow; size_t shadowCap; // pinned host copy of the last state exchanged with the device (13 floats per body, device order);
                                     // the staging buffer of the general path (after a body was deleted rows and device order differ)
    // State arena: rows[13 * slot] = pos3 quat4 lin3 ang3 of the body created slot-th since the last clear(), prev[6 * slot] its
    // previous velocities.  Both are reserved once as large virtual ranges (they never move: Rigid fields alias them) and page-locked
    // in growing pieces.  While no body has been deleted (`arenaDense`) slot == device index, and step() uploads the edited row ranges
    // straight from the arena and downloads straight into it; `arenaShadow` is the copy of what the device last saw (edit detection).
    float* arenaRows; float* arenaPrev; size_t arenaCapBodies, arenaPinnedBodies; int arenaCount; bool arenaDense;
    std::vector<float> arenaShadow; std::vector<unsigned char> arenaDynamic;
    float* claimRow(int& slot);      // called by Rigid's constructor
    void pinArena(size_t bodies);
    int uploadedBodies, uploadedForces;
    std::vector<Force*> userForces;  // joints / springs / ignore markers in creation order
    std::vector<int> userForceSlot;  // their joint / spring index on the device (-1: no rows)
    std::vector<float> rowShadow;    // per user force 48 floats: lambda12 penalty12 motor12 stiffness12 as last exchanged with the device
    bool mirrorsFresh;               // the Manifold mirrors equal the device's set (refreshManifolds ran since the last step)
    std::vector<float> mirrorRows;   // lambda12 penalty12 of every mirror at refresh time, to detect host edits
    long long uploadedBytes, downloadedBytes;   // state exchange traffic since construction (diagnostics of the e2e path)
    double hostSyncSec, deviceStepSec, hostFetchSec;   // where step() spent its wall time: edit scan + uploads, avbd_step, read-back
};
