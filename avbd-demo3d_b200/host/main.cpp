// main.cpp — headless CLI with the reference's flags and stdout format (alxspiker/avbd-demo3d source/main.cpp:189-248):
//   --nogfx | --headless, --scene | -s <name>, --steps | -n <count>
// Extensions (defaults leave the reference format untouched): --quiet (no per-step dump, one JSON timing line),
// --grid N (N^3 Stress grid instead of a named scene), --stacked (spacingY 1.01 / startY 0.51 for --grid).
// The SDL/ImGui front end is out of scope on a GPU box: without --nogfx this
// binary says so on stderr and runs headless.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "solver.h"
#include "scenes.h"

int main(int argc, char** argv) {
    bool headless = false, quiet = false, stacked = false;
    const char* requestedScene = nullptr;
    int steps = 300, grid = 0;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--nogfx") || !std::strcmp(argv[i], "--headless")) headless = true;
        else if ((!std::strcmp(argv[i], "--scene") || !std::strcmp(argv[i], "-s")) && i + 1 < argc) requestedScene = argv[++i];
        else if ((!std::strcmp(argv[i], "--steps") || !std::strcmp(argv[i], "-n")) && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) quiet = true;
        else if (!std::strcmp(argv[i], "--grid") && i + 1 < argc) grid = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--stacked")) stacked = true;
    }
    if (!headless) std::fprintf(stderr, "avbd-demo3d_b200: no graphics front end in this build; running headless (--nogfx).\n");

    Solver* solver = new Solver();
    solver->enableDiagnostics = !quiet;
    solver->logFrequency = 1;

    int sceneIdx = 0;            // unknown names silently fall back to scene 0, as upstream (main.cpp:211-219)
    if (requestedScene)
        for (int i = 0; i < sceneCount; ++i)
            if (!std::strcmp(sceneNames[i], requestedScene)) { sceneIdx = i; break; }
    const char* label = sceneNames[sceneIdx];
    if (grid > 0) {
        solver->clear();
        float w = grid * 1.15f + 20.0f; if (w < 100.0f) w = 100.0f;
        new Rigid(solver, {w, 1, w}, 0.0f, 0.5f, {0, -0.5f, 0});
        solver->iterations = 10; solver->beta = 30000.0f; solver->gamma = 0.995f;
        scene_detail::stressGrid(solver, grid, grid, grid, stacked ? 1.01f : 2.0f, stacked ? 0.51f : 20.0f);
        label = "StressGrid";
    } else {
        scenes[sceneIdx](solver);
    }

    if (!quiet) std::printf("Running in headless mode: scene '%s', steps=%d\n", label, steps);
    auto t0 = std::chrono::steady_clock::now();
    for (int step = 0; step < steps; ++step) {
        solver->step();
        if (quiet) continue;
        std::printf("Step %d:\n", step);
        for (Rigid* body = solver->bodies; body != nullptr; body = body->next) {
            std::printf("  Body %d: Pos(%.4f, %.4f, %.4f)  ", body->id, body->position.x, body->position.y, body->position.z);
            std::printf("Rot(%.4f, %.4f, %.4f, %.4f)  ", body->orientation.x, body->orientation.y, body->orientation.z, body->orientation.w);
            std::printf("LinVel(%.4f, %.4f, %.4f)  ", body->linearVelocity.x, body->linearVelocity.y, body->linearVelocity.z);
            std::printf("AngVel(%.4f, %.4f, %.4f)\n", body->angularVelocity.x, body->angularVelocity.y, body->angularVelocity.z);
        }
        const Solver::Diagnostics& st = solver->lastDiagnostics;
        std::printf("  Diagnostics: manifolds=%d contacts=%d dynBodies=%d maxPen=%.6f maxDrift=%.6f maxLin=%.3f maxAng=%.3f maxLambda=%.3f\n",
                    st.activeManifolds, st.activeContacts, st.dynamicBodies, st.maxPenetration, st.maxConstraintViolation, st.maxLinearSpeed,
                    st.maxAngularSpeed, st.maxNormalImpulse);
    }
    if (quiet) {
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const Solver::Diagnostics& st = solver->lastDiagnostics;
        std::printf("{\"scene\": \"%s\", \"steps\": %d, \"seconds\": %.6f, \"steps_per_s\": %.3f, \"manifolds\": %d, \"contacts\": %d, \"dynBodies\": %d, \"maxPen\": %.6f}\n",
                    label, steps, sec, steps / sec, st.activeManifolds, st.activeContacts, st.dynamicBodies, st.maxPenetration);
    }
    delete solver;
    return 0;
}
