// main.cpp — headless CLI with the reference's flags and stdout format (alxspiker/avbd-demo3d source/main.cpp:189-248):
//   --nogfx | --headless, --scene | -s <name>, --steps | -n <count>
// Extensions (defaults leave the reference format untouched): --quiet (no per-step dump, one JSON timing line),
// --grid N (N^3 Stress grid instead of a named scene), --stacked (spacingY 1.01 / startY 0.51 for --grid),
// --warmup W (untimed steps before the counted ones), --dump-binary FILE (compact binary trajectory instead of the text
// dump: 88 MB of text for Stress1000 x 600 upstream), --no-readback (state stays on the device; it is
// fetched once at the end), --reupload (every body treated as edited before every step: the full-upload path),
// --save-snapshot FILE / --load-snapshot FILE (device state after the run / before the first step; same scene required).
//
// Binary trajectory: header {char magic[8] = "AVBDTRJ1"; int32 bodies; int32 steps;} then per step
// {int32 step; float state[bodies][13] (creation order: pos3 quat4 lin3 ang3); float diag[5]; int32 counts[3];}.
// The SDL/ImGui front end is out of scope on a GPU box: without --nogfx this
// binary says so on stderr and runs headless.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "solver.h"
#include "scenes.h"

int main(int argc, char** argv) {
    bool headless = false, quiet = false, stacked = false, readback = true, reupload = false;
    const char* requestedScene = nullptr; const char* dumpPath = nullptr; const char* savePath = nullptr; const char* loadPath = nullptr;
    int steps = 300, grid = 0, warmup = 0;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--nogfx") || !std::strcmp(argv[i], "--headless")) headless = true;
        else if ((!std::strcmp(argv[i], "--scene") || !std::strcmp(argv[i], "-s")) && i + 1 < argc) requestedScene = argv[++i];
        else if ((!std::strcmp(argv[i], "--steps") || !std::strcmp(argv[i], "-n")) && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) quiet = true;
        else if (!std::strcmp(argv[i], "--grid") && i + 1 < argc) grid = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--stacked")) stacked = true;
        else if (!std::strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--dump-binary") && i + 1 < argc) dumpPath = argv[++i];
        else if (!std::strcmp(argv[i], "--no-readback")) readback = false;
        else if (!std::strcmp(argv[i], "--reupload")) reupload = true;
        else if (!std::strcmp(argv[i], "--save-snapshot") && i + 1 < argc) savePath = argv[++i];
        else if (!std::strcmp(argv[i], "--load-snapshot") && i + 1 < argc) loadPath = argv[++i];
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

    solver->readBack = readback; solver->uploadAll = reupload;
    if (loadPath) {
        std::FILE* f = std::fopen(loadPath, "rb");
        if (!f) { std::fprintf(stderr, "cannot open %s\n", loadPath); return 2; }
        std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> blob((size_t)sz);
        if (std::fread(blob.data(), 1, blob.size(), f) != blob.size()) { std::fprintf(stderr, "short read on %s\n", loadPath); return 2; }
        std::fclose(f);
        solver->syncToDevice();
        solver->restore(blob);
    }
    std::FILE* dump = nullptr;
    int nBodies = (int)solver->order.size();
    if (dumpPath) {
        dump = std::fopen(dumpPath, "wb");
        if (!dump) { std::fprintf(stderr, "cannot open %s\n", dumpPath); return 2; }
        const char magic[8] = {'A', 'V', 'B', 'D', 'T', 'R', 'J', '1'};
        std::fwrite(magic, 1, 8, dump); std::fwrite(&nBodies, 4, 1, dump); std::fwrite(&steps, 4, 1, dump);
    }
    if (!quiet && !dump) std::printf("Running in headless mode: scene '%s', steps=%d\n", label, steps);
    for (int step = 0; step < warmup; ++step) solver->step();
    solver->hostSyncSec = solver->deviceStepSec = solver->hostFetchSec = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    for (int step = 0; step < steps; ++step) {
        solver->step();
        if (dump) {          // the state the step left in the pinned exchange buffer, creation order — no per-body formatting
            const Solver::Diagnostics& st = solver->lastDiagnostics;
            if (!readback) solver->fetchState();
            std::fwrite(&step, 4, 1, dump);
            std::fwrite(solver->hostState(), sizeof(float), (size_t)nBodies * 13, dump);
            const float df[5] = {st.maxPenetration, st.maxConstraintViolation, st.maxLinearSpeed, st.maxAngularSpeed, st.maxNormalImpulse};
            const int di[3] = {st.activeManifolds, st.activeContacts, st.dynamicBodies};
            std::fwrite(df, 4, 5, dump); std::fwrite(di, 4, 3, dump);
            continue;
        }
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
    if (!readback) solver->fetchState();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dump) std::fclose(dump);
    if (quiet || dump) {
        const Solver::Diagnostics& st = solver->lastDiagnostics;
        std::printf("{\"scene\": \"%s\", \"steps\": %d, \"warmup\": %d, \"seconds\": %.6f, \"steps_per_s\": %.3f, \"ms_per_step\": %.4f, \"bodies\": %d, \"manifolds\": %d, \"contacts\": %d, "
                    "\"dynBodies\": %d, \"maxPen\": %.6f, \"readback\": %s, \"reupload\": %s, \"h2d_bytes\": %lld, \"d2h_bytes\": %lld, "
                    "\"ms_sync_edits\": %.4f, \"ms_device_step\": %.4f, \"ms_read_back\": %.4f}\n",
                    label, steps, warmup, sec, steps / sec, 1e3 * sec / (steps > 0 ? steps : 1), nBodies, st.activeManifolds, st.activeContacts, st.dynamicBodies, st.maxPenetration,
                    readback ? "true" : "false", reupload ? "true" : "false", solver->uploadedBytes, solver->downloadedBytes,
                    1e3 * solver->hostSyncSec / (steps > 0 ? steps : 1), 1e3 * solver->deviceStepSec / (steps > 0 ? steps : 1), 1e3 * solver->hostFetchSec / (steps > 0 ? steps : 1));
    }
    if (savePath) {
        std::vector<unsigned char> blob = solver->snapshot();
        std::FILE* f = std::fopen(savePath, "wb");
        if (!f) { std::fprintf(stderr, "cannot open %s\n", savePath); return 2; }
        std::fwrite(blob.data(), 1, blob.size(), f); std::fclose(f);
    }
    delete solver;
    return 0;
}
