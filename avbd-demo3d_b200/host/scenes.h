// scenes.h — scene presets, same table / names / bodies as alxspiker/avbd-demo3d source/scenes.h:23-212
// (tests/test_scenes.py pins the body sets bit for bit).  Presets only call Solver::clear and `new Rigid`.
#pragma once
#include "solver.h"
#include <cmath>

namespace scene_detail {
inline void ground(Solver* s) { new Rigid(s, {100, 1, 100}, 0.0f, 0.5f, {0, -0.5f, 0}); }
inline float hash01(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return (x & 0x00FFFFFFU) / 16777215.0f;
}
// Generalised Stress grid: NX x NY x NZ unit cubes with the Stress1000 jitter hash.
inline void stressGrid(Solver* s, int NX, int NY, int NZ, float spacingY, float startY) {
    const float spacingXZ = 1.15f, jitterXZ = 0.04f, jitterY = 0.25f;
    for (int y = 0; y < NY; ++y)
        for (int z = 0; z < NZ; ++z)
            for (int x = 0; x < NX; ++x) {
                unsigned seed = (unsigned)(x + NX * (z + NZ * y) + 1);
                float jx = (hash01(seed * 9781U) * 2.0f - 1.0f) * jitterXZ;
                float jz = (hash01(seed * 6271U) * 2.0f - 1.0f) * jitterXZ;
                float jy = hash01(seed * 3343U) * jitterY;
                new Rigid(s, {1.0f, 1.0f, 1.0f}, 1.0f, 0.5f,
                          {(x - (NX - 1) * 0.5f) * spacingXZ + jx, startY + y * spacingY + jy, (z - (NZ - 1) * 0.5f) * spacingXZ + jz});
            }
}
}

static void sceneEmpty(Solver* s) { s->clear(); }
static void sceneGround(Solver* s) { s->clear(); scene_detail::ground(s); }
static void sceneStack(Solver* s) {
    s->clear(); scene_detail::ground(s);
    for (int i = 0; i < 10; ++i) new Rigid(s, {1, 1, 1}, 1.0f, 0.5f, {0, i * 1.1f + 0.5f, 0});
}
static void scenePyramid(Solver* s) {
    s->clear(); scene_detail::ground(s);
    const int P = 10;
    for (int y = 0; y < P; ++y)
        for (int x = 0; x < P - y; ++x) new Rigid(s, {1, 1, 1}, 1.0f, 0.5f, {(x - (P - y - 1) * 0.5f) * 1.1f, y * 1.05f + 0.5f, 0});
}
static void sceneWall(Solver* s) {
    s->clear(); scene_detail::ground(s);
    const int W = 8, H = 8; const vec3 brick = {1.0f, 0.5f, 0.5f}; const float sx = 1.03f, sy = 0.52f, baseY = brick.y * 0.5f;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float xo = (i % 2 == 0) ? 0.0f : 0.5f * sx;
            new Rigid(s, brick, 1.0f, 0.4f, {(j - (W - 1) * 0.5f) * sx + xo, i * sy + baseY, -5});
        }
}
static void sceneTwoBlockDrop(Solver* s) {
    s->clear(); scene_detail::ground(s);
    new Rigid(s, {1.0f, 1.0f, 1.0f}, 1.0f, 0.5f, {0.0f, 0.5f, 0.0f});
    new Rigid(s, {1.0f, 1.0f, 1.0f}, 1.0f, 0.5f, {0.18f, 2.2f, 0.0f}, quat(vec3(0.0f, 0.0f, 1.0f), 0.45f), {0, 0, 0}, {0.0f, 0.0f, 1.0f});
}
static void sceneStress1000(Solver* s) {
    s->clear(); scene_detail::ground(s);
    s->iterations = 20; s->beta = 30000.0f; s->gamma = 0.995f;      // these overrides persist, as upstream
    scene_detail::stressGrid(s, 10, 10, 10, 2.0f, 20.0f);
}
static void sceneRod(Solver* s) {       // joints are placeholders upstream too
    s->clear();
    for (int i = 0; i < 15; ++i) new Rigid(s, {0.25f, 1, 0.25f}, (i == 0) ? 0.0f : 1.0f, 0.5f, {0, 10.0f - i * 1.0f, 0});
}
static void sceneSoftBody(Solver* s) {
    s->clear(); scene_detail::ground(s);
    const int W = 10, H = 10;
    for (int i = 0; i < W; ++i)
        for (int j = 0; j < H; ++j) new Rigid(s, {0.5f, 0.5f, 0.5f}, 1.0f, 0.3f, {i * 0.6f - W * 0.3f, j * 0.6f + 2.0f, 0});
}

static void (*scenes[])(Solver*) = {sceneEmpty, sceneGround, sceneStack, scenePyramid, sceneWall, sceneTwoBlockDrop, sceneStress1000, sceneRod, sceneSoftBody};
static const char* sceneNames[] = {"Empty", "Ground", "Stack", "Pyramid", "Wall", "TwoBlockDrop", "Stress1000", "Rod (WIP)", "Soft Body (WIP)"};
static const int sceneCount = sizeof(scenes) / sizeof(scenes[0]);
