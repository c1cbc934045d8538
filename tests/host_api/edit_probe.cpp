// edit_probe.cpp — host-mirror behaviours the reference's API implies but its demo never exercises, against the device
// world (avbd-demo3d_b200/host + libavbd_b200.so only; prints "key value..." lines for tests/test_gpu_host_api.py):
//   * deleting a body between steps keeps the warm-start history of every other manifold (force.cpp:43-69 is O(degree) upstream)
//   * a Joint re-uploaded after such a rebuild keeps its construction-time anchor (joint.cpp:19, :47-50)
//   * host edits of Force::motor / Force::stiffness reach the solver (solver.h:91-97, solver.cpp:380)
//   * Solver::snapshot / restore resume bit-identically; moving ONE body uploads one body, not the world
#include <cstdio>
#include <cmath>
#include <cstring>
#include "solver.h"
#include "joint.h"
#include "spring.h"
#include "scenes.h"

static Force* mirrorOf(Solver* s, Rigid* a, Rigid* b) {
    for (Force* f = s->forces; f; f = f->next)
        if (f->isManifold() && ((f->bodyA == a && f->bodyB == b) || (f->bodyA == b && f->bodyB == a))) return f;
    return nullptr;
}
static int mirrorCount(Solver* s) { int n = 0; for (Force* f = s->forces; f; f = f->next) n += f->isManifold() ? 1 : 0; return n; }

int main() {
    Solver* solver = new Solver();
    // ---- 1. body deletion keeps the other manifolds' history
    sceneStack(solver);
    for (int i = 0; i < 200; ++i) solver->step();
    Rigid* top = solver->bodies;                       // newest = top of the stack
    Rigid* bottom = nullptr; Rigid* ground = nullptr;
    for (Rigid* r = solver->bodies; r; r = r->next) { if (r->next && !r->next->next) bottom = r; if (!r->next) ground = r; }
    solver->refreshManifolds();
    int mirrorsBefore = mirrorCount(solver);
    float lamBefore[12], penBefore[12];
    { Force* f = mirrorOf(solver, bottom, ground); std::memcpy(lamBefore, f->lambda, sizeof(lamBefore)); std::memcpy(penBefore, f->penalty, sizeof(penBefore)); }
    long long up0 = solver->uploadedBytes;
    delete top;
    solver->syncToDevice();                            // the device world is re-created here; nothing has stepped yet
    solver->refreshManifolds();
    Force* kept = mirrorOf(solver, bottom, ground);
    int rowsSame = kept && !std::memcmp(lamBefore, kept->lambda, sizeof(lamBefore)) && !std::memcmp(penBefore, kept->penalty, sizeof(penBefore));
    printf("delete_body mirrors %d %d rows_identical %d lambda_n %.6g bodies %d\n", mirrorsBefore, mirrorCount(solver), rowsSame, lamBefore[0], (int)solver->order.size());
    solver->step();
    float worst = 0;
    for (int i = 0; i < 30; ++i) { solver->step(); worst = std::fmax(worst, solver->lastDiagnostics.maxLinearSpeed); }
    printf("delete_body settle_max_lin %.6g rebuild_upload_bytes %lld\n", worst, solver->uploadedBytes - up0);

    // ---- 2. moving one body uploads one body
    Rigid* mover = solver->bodies;
    long long up1 = solver->uploadedBytes;
    mover->position.x += 0.001f;
    solver->step();
    printf("edit_one_body upload_bytes %lld\n", solver->uploadedBytes - up1);
    long long up2 = solver->uploadedBytes;
    solver->step();
    printf("edit_nothing upload_bytes %lld\n", solver->uploadedBytes - up2);

    // ---- 3. joints: re-upload keeps the construction-time anchor; motor / stiffness edits
    solver->clear(); solver->defaultParams();
    new Rigid(solver, vec3(20, 1, 20), 0.0f, 0.5f, vec3(0, -0.5f, 0));
    Rigid* hang = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(0, 3, 0));
    Rigid* extra = new Rigid(solver, vec3(1, 1, 1), 1.0f, 0.5f, vec3(5, 0.51f, 0));
    Joint* weld = new Joint(solver, hang, vec3(0, 3, 0));
    for (int r = 0; r < 6; ++r) weld->stiffness[r] = (r == 1) ? 400.0f : FLT_MAX;       // soft vertical row: hangs m g / k = 0.025 below the anchor
    for (int i = 0; i < 240; ++i) solver->step();
    printf("soft_row offset %.6g penalty %.6g lambda %.6g\n", 3.0f - hang->position.y, weld->penalty[1], weld->lambda[1]);
    delete extra;                                      // forces a rebuild: the joint must NOT re-anchor at the sagged pose
    for (int i = 0; i < 120; ++i) solver->step();
    printf("rebuild_keeps_anchor offset %.6g rB %.6g %.6g %.6g\n", 3.0f - hang->position.y, weld->rB.x, weld->rB.y, weld->rB.z);
    weld->motor[1] = 4.0f;                             // k C + motor = m g  =>  C = (10 - 4) / 400
    for (int i = 0; i < 240; ++i) solver->step();
    printf("motor offset %.6g\n", 3.0f - hang->position.y);

    // ---- 4. snapshot / restore
    solver->clear(); solver->defaultParams();
    scenePyramid(solver);
    for (int i = 0; i < 60; ++i) solver->step();
    std::vector<unsigned char> blob = solver->snapshot();
    for (int i = 0; i < 40; ++i) solver->step();
    std::vector<float> a;
    for (Rigid* r = solver->bodies; r; r = r->next) { a.push_back(r->position.x); a.push_back(r->position.y); a.push_back(r->position.z); a.push_back(r->orientation.w); }
    solver->restore(blob);
    for (int i = 0; i < 40; ++i) solver->step();
    size_t k = 0; int same = 1;
    for (Rigid* r = solver->bodies; r; r = r->next) {
        const float b[4] = {r->position.x, r->position.y, r->position.z, r->orientation.w};
        if (std::memcmp(b, &a[k], sizeof(b)) != 0) same = 0;
        k += 4;
    }
    printf("snapshot_resume_identical %d blob_bytes %zu\n", same, blob.size());
    delete solver;
    return 0;
}
