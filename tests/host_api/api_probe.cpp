// api_probe.cpp — ONE source, two builds: against the unmodified reference (its headers and object files, oracle/_ref/api_probe_ref)
// and against this repo's host mirror (avbd-demo3d_b200/host + libavbd_b200.so, tests/host_api/api_probe_b200).  It only uses the
// reference's public class API (solver.h:48-181, joint.h, spring.h, ignorecollision.h), the way main.cpp and scenes.h do, and
// prints "key value..." lines that tests/test_gpu_host_api.py compares: construction-derived values, list order, the row
// interface of Joint / Spring at the initial state and pick exactly; what depends on the solve order statistically.
#include <cstdio>
#include <initializer_list>
#include <cmath>
#include "solver.h"
#include "joint.h"
#include "spring.h"
#include "ignorecollision.h"

int main() {
    Solver* solver = new Solver();
    solver->defaultParams();
    printf("params %.9g %d %.9g %.9g %.9g %d\n", solver->dt, solver->iterations, solver->alpha, solver->beta, solver->gamma, (int)solver->postStabilize);

    new Rigid(solver, vec3(30, 1, 30), 0.0f, 0.5f, vec3(0, -0.5f, 0));                    // ground; owned by the solver's list like every body
    Rigid* a = new Rigid(solver, vec3(1, 2, 0.5f), 1.5f, 0.4f, vec3(0, 1.01f, 0));
    Rigid* b = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.6f, vec3(0, 2.52f, 0));
    Rigid* c = new Rigid(solver, vec3(0.5f, 0.5f, 0.5f), 2.0f, 0.5f, vec3(3, 4, 0), quat(0.1f, 0.2f, 0.3f, 0.9273618f), vec3(0.5f, 0, 0), vec3(0, 1, 0));
    Rigid* d = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(-3, 0.51f, 0));
    Rigid* e = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(-3.2f, 0.6f, 0.1f));
    Rigid* hang = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(6, 3, 0));
    Rigid* bob = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(6, 1.2f, 0));

    // rigid.cpp:12-41: mass properties; list order is newest first
    int k = 0;
    for (Rigid* r = solver->bodies; r; r = r->next, ++k)
        printf("body %d mass %.9g invMass %.9g inertia %.9g %.9g %.9g radius %.9g friction %.9g\n", k, r->mass, r->invMass,
               r->inertiaTensor.cols[0].x, r->inertiaTensor.cols[1].y, r->inertiaTensor.cols[2].z, r->radius, r->friction);
    printf("first_is_newest %d\n", (int)(solver->bodies == bob));
    mat3 Iw = c->getInertiaTensorWorld();
    printf("inertia_world %.7g %.7g %.7g %.7g %.7g %.7g\n", Iw.cols[0].x, Iw.cols[1].y, Iw.cols[2].z, Iw.cols[1].x, Iw.cols[2].x, Iw.cols[2].y);

    Joint* weld = new Joint(solver, nullptr, hang, vec3(6, 3.5f, 0), vec3(0, 0.5f, 0));
    Spring* spring = new Spring(solver, hang, bob, vec3(0, -0.5f, 0), vec3(0, 0.5f, 0), 500.0f, 0.8f);
    IgnoreCollision* ign = new IgnoreCollision(solver, d, e);
    printf("rows %d %d %d\n", weld->getRowCount(), spring->getRowCount(), ign->getRowCount());
    printf("constrained %d %d %d %d\n", (int)d->isConstrainedTo(e), (int)e->isConstrainedTo(d), (int)hang->isConstrainedTo(bob), (int)a->isConstrainedTo(b));

    // the row interface at the initial state (joint.cpp:68-139, spring.cpp:33-90)
    weld->computeConstraint(solver->alpha);
    printf("joint_C %.7g %.7g %.7g %.7g %.7g %.7g\n", weld->C[0], weld->C[1], weld->C[2], weld->C[3], weld->C[4], weld->C[5]);
    for (int row = 0; row < 6; ++row) {
        vec3 jl, ja;
        weld->computeDerivatives(jl, ja, hang, row);
        printf("joint_J %d %.7g %.7g %.7g %.7g %.7g %.7g\n", row, jl.x, jl.y, jl.z, ja.x, ja.y, ja.z);
    }
    spring->computeConstraint(solver->alpha);
    printf("spring_C %.7g\n", spring->C[0]);
    for (Rigid* body : {hang, bob}) {
        vec3 jl, ja;
        spring->computeDerivatives(jl, ja, body, 0);
        printf("spring_J %.7g %.7g %.7g %.7g %.7g %.7g\n", jl.x, jl.y, jl.z, ja.x, ja.y, ja.z);
    }

    // Solver::pick before anything moved (solver.cpp:145-228)
    const float rays[5][6] = {{0, 10, 0, 0, -1, 0}, {3, 10, 0, 0, -1, 0}, {-10, 0.5f, 0, 1, 0, 0}, {0, 10, 5, 0, -1, 0}, {6, 10, 0.2f, 0, -1, 0.01f}};
    for (const float* r : rays) {
        vec3 local;
        Rigid* hit = solver->pick(vec3(r[0], r[1], r[2]), vec3(r[3], r[4], r[5]), local);
        int idx = -1, n = 0;
        for (Rigid* q = solver->bodies; q; q = q->next) ++n;
        int pos = 0;
        for (Rigid* q = solver->bodies; q; q = q->next, ++pos) if (q == hit) idx = n - 1 - pos;        // creation index
        printf("pick %d %.5f %.5f %.5f\n", idx, hit ? local.x : 0.0f, hit ? local.y : 0.0f, hit ? local.z : 0.0f);
    }

    for (int s = 0; s < 240; ++s) solver->step();
    // order-dependent: compared with tolerances
    k = 0;
    for (Rigid* r = solver->bodies; r; r = r->next, ++k) printf("rest %d %.4f %.4f %.4f\n", k, r->position.x, r->position.y, r->position.z);
    int manifolds = 0, contacts = 0, others = 0;
    solver->draw();                                     // no-op upstream under --nogfx builds; refreshes the Manifold mirrors here
    for (Force* f = solver->forces; f; f = f->next) {
        if (f->isManifold()) { ++manifolds; contacts += f->getRowCount() / 3; } else ++others;
    }
    printf("forces %d %d %d\n", manifolds, contacts, others);
    printf("diag %d %d %d %.4f\n", solver->lastDiagnostics.activeManifolds, solver->lastDiagnostics.activeContacts, solver->lastDiagnostics.dynamicBodies,
           solver->lastDiagnostics.maxPenetration);
    printf("stepIndex %d\n", solver->stepIndex);

    delete ign;                                         // d and e may collide again
    for (int s = 0; s < 120; ++s) solver->step();
    printf("after_delete %d\n", (int)d->isConstrainedTo(e));
    printf("separated %d\n", (int)(std::fabs(d->position.x - e->position.x) > 0.9f || std::fabs(d->position.y - e->position.y) > 0.9f));

    solver->clear();
    printf("cleared %d %d\n", (int)(solver->bodies == nullptr), (int)(solver->forces == nullptr));

    // ---- the same Solver reused after clear(): host edits between steps (the GUI's drag / re-spin, main.cpp:88-142), a stacked
    //      body taken away, diagnostics logging (solver.cpp:499-512), the Manifold objects a renderer walks
    solver->enableDiagnostics = true; solver->logFrequency = 50;
    new Rigid(solver, vec3(30, 1, 30), 0.0f, 0.5f, vec3(0, -0.5f, 0));
    Rigid* p0 = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(0, 0.51f, 0));
    Rigid* p1 = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(0, 1.52f, 0));
    Rigid* p2 = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(4, 0.51f, 0));
    for (int s = 0; s < 100; ++s) solver->step();
    p2->position = vec3(4, 3.0f, 1);                    // picked up and dropped somewhere else
    p2->linearVelocity = vec3(0, 0, 0);
    for (int s = 0; s < 100; ++s) solver->step();
    printf("moved %.3f %.3f %.3f\n", p2->position.x, p2->position.y, p2->position.z);
    p1->position = vec3(-6, 0.51f, -2);                 // the top box is lifted off and put on the ground elsewhere (deleting a Rigid that
    p1->linearVelocity = vec3(0, 0, 0);                 // still has manifolds leaves them dangling upstream, rigid.cpp:43-49: not comparable)
    for (int s = 0; s < 50; ++s) solver->step();
    k = 0;
    for (Rigid* r = solver->bodies; r; r = r->next, ++k) printf("rest2 %d %.3f %.3f %.3f\n", k, r->position.x, r->position.y, r->position.z);
    solver->draw();
    for (Force* f = solver->forces; f; f = f->next) {
        if (!f->isManifold()) continue;
        Manifold* m = static_cast<Manifold*>(f);
        float lam = 0.0f;
        for (int c = 0; c < m->numContacts; ++c) lam += m->lambda[c * 3];
        // which bodies, how many contacts, the first normal, the total normal force (= weight / ... at rest)
        printf("manifold %d %d %d n %.3f %.3f %.3f mu %.4f lam %.2f\n", (int)(m->bodyA == p0 || m->bodyB == p0), (int)(m->bodyA == p2 || m->bodyB == p2),
               m->numContacts, std::fabs(m->contacts[0].normal.x), std::fabs(m->contacts[0].normal.y), std::fabs(m->contacts[0].normal.z), m->combinedFriction, lam);
    }
    printf("diag2 %d %d %d\n", solver->lastDiagnostics.activeManifolds, solver->lastDiagnostics.activeContacts, solver->lastDiagnostics.dynamicBodies);
    (void)p0;
    delete solver;
    return 0;
}
