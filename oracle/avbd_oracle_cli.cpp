// oracle/avbd_oracle_cli.cpp — TEST INFRASTRUCTURE ONLY.
// Headless driver over the CPU restatement with the reference CLI's flags and
// stdout format (main.cpp:189-248), so `diff`/md5 against oracle/_ref/avbd_demo3d_ref
// pins whole trajectories.  Extra flag --quiet (not in the reference) suppresses
// the per-step dump for timing runs and prints a one-line timing summary.
#include "avbd_oracle.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main(int argc, char** argv) {
    bool headless = false, quiet = false;
    const char* wanted = nullptr;
    int steps = 300;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--nogfx") || !std::strcmp(argv[i], "--headless")) headless = true;
        else if ((!std::strcmp(argv[i], "--scene") || !std::strcmp(argv[i], "-s")) && i + 1 < argc) wanted = argv[++i];
        else if ((!std::strcmp(argv[i], "--steps") || !std::strcmp(argv[i], "-n")) && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) quiet = true;
    }
    (void)headless;   // this build has no graphics mode
    void* w = orc_create();
    orc_set_logging(w, quiet ? 0 : 1, 1);
    int idx = 0;      // unknown names fall back to scene 0, main.cpp:211-219
    if (wanted) for (int i = 0; i < orc_scene_count(); ++i) if (!std::strcmp(orc_scene_name(i), wanted)) { idx = i; break; }
    orc_load_scene_index(w, idx);
    int n = orc_num_bodies(w);
    std::vector<float> st(13 * (size_t)(n > 0 ? n : 1));
    if (!quiet) std::printf("Running in headless mode: scene '%s', steps=%d\n", orc_scene_name(idx), steps);
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) {
        orc_step(w, 1);
        if (quiet) continue;
        std::printf("Step %d:\n", s);
        orc_get_state(w, st.data());
        for (int i = n - 1; i >= 0; --i) {    // list order = newest first
            const float* b = &st[13 * (size_t)i];
            std::printf("  Body %d: Pos(%.4f, %.4f, %.4f)  ", orc_body_id(w, i), b[0], b[1], b[2]);
            std::printf("Rot(%.4f, %.4f, %.4f, %.4f)  ", b[3], b[4], b[5], b[6]);
            std::printf("LinVel(%.4f, %.4f, %.4f)  ", b[7], b[8], b[9]);
            std::printf("AngVel(%.4f, %.4f, %.4f)\n", b[10], b[11], b[12]);
        }
        float f[5]; int k[3];
        orc_get_diagnostics(w, f, k);
        std::printf("  Diagnostics: manifolds=%d contacts=%d dynBodies=%d maxPen=%.6f maxDrift=%.6f maxLin=%.3f maxAng=%.3f maxLambda=%.3f\n",
                    k[1], k[0], k[2], f[0], f[1], f[2], f[3], f[4]);
    }
    if (quiet) {
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"scene\": \"%s\", \"steps\": %d, \"seconds\": %.6f, \"steps_per_s\": %.3f}\n", orc_scene_name(idx), steps, sec, steps / sec);
    }
    orc_destroy(w);
    return 0;
}
