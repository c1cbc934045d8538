/* avbd_b200.h — C ABI of the B200-native AVBD step loop (libavbd_b200.so).
 *
 * This is the drop-in boundary for the hot path of alxspiker/avbd-demo3d:
 * everything `Solver::step()` does (source/solver.cpp:255-514).  The reference
 * has no FFI of its own — its boundary is the C++ class API of
 * source/solver.h:48-181 — so each entry point below names the reference
 * interface it replaces.  The host-side C++17 mirror of those classes
 * (avbd-demo3d_b200/host/) is a thin client of this header; INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success and a negative code on failure with the text in avbd_last_error();
 * the library owns device memory, the caller owns host buffers; one CUDA
 * stream per world, a world is not thread-safe; bodies are addressed by
 * creation index (0 = first created).  There is NO CPU fallback: without a
 * CUDA device every call fails with AVBD_ERR_NO_DEVICE.
 */
#ifndef AVBD_B200_H
#define AVBD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avbd_world avbd_world;

enum {
    AVBD_OK = 0,
    AVBD_ERR_NO_DEVICE = -1,
    AVBD_ERR_CUDA = -2,
    AVBD_ERR_ARG = -3,
    AVBD_ERR_CAPACITY = -4
};

/* Solver::Diagnostics, solver.h:155-164, plus the count of NaN scrubs (solver.cpp:51-66). */
typedef struct avbd_diagnostics {
    float maxPenetration, maxConstraintViolation, maxLinearSpeed, maxAngularSpeed, maxNormalImpulse;
    int activeContacts, activeManifolds, dynamicBodies;
    int nanEvents;
} avbd_diagnostics;

/* Per-step stage timings in milliseconds (CUDA events on the world's stream) and sizes. */
typedef struct avbd_step_stats {
    float ms_broadphase, ms_narrowphase, ms_graph, ms_predict, ms_primal, ms_dual, ms_velocity, ms_total;
    int bodies, dynamicBodies, pairs, candidates, manifolds, contacts, colours, iterations;
    int contactVisits;          /* contacts counted once per dynamic endpoint = primal contact visits per iteration */
    long long kernelLaunches;   /* launches of this library's own kernels since world creation */
} avbd_step_stats;

/* Accumulated since avbd_set_profiling(w, 1): device time (CUDA events on the world's stream) of the primal
 * sweeps and dual passes, with the work they covered, so a caller can form algorithmic-bytes / time. */
typedef struct avbd_profile {
    double ms_primal, ms_dual;
    long long steps, primal_sweeps, primal_launches, primal_bodies, primal_visits, dual_launches, dual_contacts;
    long long kernel_launches, library_launches;   /* totals since world creation: own kernels / CUB passes */
    long long deferred_dual_contacts;              /* contact dual updates applied INSIDE primal sweeps (deferred dual): their time is in ms_primal */
    /* per-stage device time and the sizes SURVEY.md section 8d's byte formulas need, summed over the profiled steps */
    double ms_broadphase, ms_narrowphase, ms_graph, ms_predict, ms_solve, ms_velocity, ms_step;
    long long bodies, pairs, candidates, manifolds, manifolds_prev, contacts, visits, graph_builds;
} avbd_profile;

const char* avbd_last_error(void);
/* Page-locked host buffers for the state exchange (avbd_upload_state / avbd_download_state run at DMA speed from these;
 * any other host pointer works too, through the driver's staging copy). */
void* avbd_host_alloc(long long bytes);
void  avbd_host_free(void* p);
/* Page-lock / release a host range the caller owns (page-aligned; e.g. the C++ host mirror's state arena, whose rows the Rigid
 * objects' public fields alias — host/solver.h).  0 on success. */
int   avbd_host_register(void* p, long long bytes);
int   avbd_host_unregister(void* p);
int  avbd_device_count(void);

/* Solver::Solver / ~Solver (solver.cpp:129-143). */
avbd_world* avbd_world_create(int device);
void avbd_world_destroy(avbd_world* w);
/* Solver::clear (solver.cpp:230-238): drops bodies, forces, manifolds; keeps params. */
int  avbd_clear(avbd_world* w);
/* Solver's public tunables (solver.h:147-151), re-read every step; Solver::defaultParams = solver.cpp:240-253. */
int  avbd_set_params(avbd_world* w, float dt, const float* gravity3, int iterations, float alpha, float beta, float gamma,
                     int postStabilize);
int  avbd_default_params(avbd_world* w);

/* new Rigid(...) x count (rigid.cpp:12-41): mass properties are derived exactly as the ctor does.
 * Arrays are packed per body: size[3], density, friction, pos[3], quat[4] (x y z w), lin[3], ang[3].
 * world_ids may be NULL (single world); otherwise bodies of one world must be contiguous and
 * ids non-decreasing (ensemble batches: no pair is ever formed across worlds).
 * Returns the index of the first body added, or a negative error. */
int  avbd_add_bodies(avbd_world* w, int count, const float* size3, const float* density, const float* friction,
                     const float* pos3, const float* quat4, const float* lin3, const float* ang3, const int* world_ids);
int  avbd_num_bodies(const avbd_world* w);

/* new Joint(...) (joint.cpp:11-63).  a = -1: body-world weld at world anchor anchorA. */
int  avbd_add_joint(avbd_world* w, int a, int b, const float* anchorA3, const float* anchorB3, float linearStiffness,
                    float angularStiffness);
/* The same weld with the construction-time values the caller already holds (Joint::rB and
 * Joint::initialRelativeOrientation, joint.h:17-19, captured by the constructors at joint.cpp:19, :47-50) instead of
 * values derived from the device poses at upload time: a joint re-added after a rebuild keeps its original reference. */
int  avbd_add_joint_raw(avbd_world* w, int a, int b, const float* rA3, const float* rB3, const float* rel0_4,
                        float linearStiffness, float angularStiffness);
/* new Spring(...) (spring.cpp:10-30).  rest < 0: current distance. */
int  avbd_add_spring(avbd_world* w, int a, int b, const float* anchorA3, const float* anchorB3, float stiffness, float rest);
/* new IgnoreCollision(...) (ignorecollision.h:14-16). */
int  avbd_add_ignore(avbd_world* w, int a, int b);

/* The public row arrays of a user Force (solver.h:91-97: lambda, penalty, motor, stiffness) — kind 0 joint (6 rows),
 * 1 spring (1 row), index = the value avbd_add_joint / avbd_add_spring returned.  set: a NULL array is left alone
 * (motor enters the primal at solver.cpp:380, stiffness decides hard / soft at :378, :290, :416).  get: device values. */
int  avbd_set_force_rows(avbd_world* w, int kind, int index, const float* lambda, const float* penalty, const float* motor,
                         const float* stiffness);
int  avbd_get_force_rows(avbd_world* w, int kind, int index, float* lambda, float* penalty, float* motor, float* stiffness);
/* lambda / penalty of EVERY user force after a step, one copy: joints12 = lambda6 penalty6 per joint, springs2 = lambda penalty
 * per spring (the host mirror refreshes Force::lambda / Force::penalty from it).  Either may be NULL. */
int  avbd_download_user_rows(avbd_world* w, float* joints12, float* springs2);
int  avbd_num_joints(const avbd_world* w);
int  avbd_num_springs(const avbd_world* w);

/* Solver::step() x n (solver.cpp:255-514).  Asynchronous on the world's stream. */
int  avbd_step(avbd_world* w, int n);
int  avbd_sync(avbd_world* w);
/* n steps bracketed by CUDA events on the world's stream; *ms = device time of the n steps. */
int  avbd_step_timed(avbd_world* w, int n, float* ms);
int  avbd_set_profiling(avbd_world* w, int on);
int  avbd_get_profile(avbd_world* w, avbd_profile* out);

/* Rigid public state (solver.h:56-60): 13 floats per body pos3 quat4 lin3 ang3, creation order. */
int  avbd_download_state(avbd_world* w, float* out13);
int  avbd_upload_state(avbd_world* w, const float* in13);
/* avbd_download_state in pieces of chunkBodies bodies: `landed(first, count, user)` runs on the calling thread as each piece has
 * arrived in out13 while the following pieces are still in flight (the C++ host mirror copies each piece to its edit-detection
 * shadow under the next piece's transfer). */
typedef void (*avbd_chunk_fn)(int first, int count, void* user);
int  avbd_download_state_chunked(avbd_world* w, float* out13, int chunkBodies, avbd_chunk_fn landed, void* user);
/* Host edits of a few bodies (main.cpp:88-142 moves / re-spins bodies between steps): bodies [first, first + count). */
int  avbd_upload_state_range(avbd_world* w, int first, int count, const float* in13);
int  avbd_download_prev_linvel(avbd_world* w, float* out3);
int  avbd_upload_prev_linvel(avbd_world* w, const float* in3);
/* size3 mass invMass inertiaDiag3 friction radius (solver.h:67-72), 10 floats per body. */
int  avbd_download_body_props(avbd_world* w, float* out10);

/* Solver::lastDiagnostics (solver.h:169). */
int  avbd_get_diagnostics(avbd_world* w, avbd_diagnostics* out);
int  avbd_get_step_stats(avbd_world* w, avbd_step_stats* out);
/* per-world diagnostics of an ensemble batch: count = number of worlds. Device pointer variant for NCCL gathers. */
int  avbd_num_worlds(const avbd_world* w);
int  avbd_get_world_diagnostics(avbd_world* w, avbd_diagnostics* out, int count);
int  avbd_world_diagnostics_device_ptr(avbd_world* w, void** ptr, int* count);

/* The Manifold list (solver.h:112-143) in pair-key order.  Layout per manifold identical to the
 * oracle dumps: ints3 {idxA idxB numContacts}, feats4, stick4, flts81 {friction, 4 x (rA3 rB3 n3 pen C0n C0t3), lambda12, penalty12}. */
/* avbd_num_manifolds = slots in use (upper bound for sizing buffers); avbd_download_manifolds returns how many
 * LIVE manifolds (numContacts > 0) it wrote — pairs whose SAT passed but produced no contact are skipped,
 * as the reference deletes them (solver.cpp:274-279). */
int  avbd_num_manifolds(avbd_world* w);
int  avbd_download_manifolds(avbd_world* w, int* ints3, int* feats4, int* stick4, float* flts81);

/* Inverse of avbd_download_manifolds: replaces the manifold set (any order; idxA > idxB; dead manifolds not allowed).
 * What Manifold::initialize carries over next step (manifold.cpp:111-155) is exactly this set, so a caller can hand the
 * solver a warm-start history (parity tests, host edits of Manifold rows, avbd_restore). */
int  avbd_upload_manifolds(avbd_world* w, int count, const int* ints3, const int* feats4, const int* stick4, const float* flts81);

/* Snapshot / restore of the whole simulation state (the reference can only clear() and re-run a
 * scene, solver.cpp:230-238): parameters, bodies, user forces with their rows, the manifold set with lambda / penalty /
 * stick anchors.  A restored world continues bit-identically.  The blob is opaque and only valid for this library build. */
long long avbd_snapshot_bytes(avbd_world* w);
int  avbd_snapshot(avbd_world* w, void* buf, long long cap);
int  avbd_restore(avbd_world* w, const void* buf, long long bytes);

/* ---- per-stage entry points (parity tests drive these one at a time) ------------------------- */
/* solver.cpp:262-270: sphere-overlap pair set of the current poses, sorted (a > b); returns count. */
int  avbd_stage_broadphase(avbd_world* w);
int  avbd_download_pairs(avbd_world* w, int* pairs2, int cap);
/* solver.cpp:262-296: broadphase + Manifold::initialize + warm-start decay. */
int  avbd_stage_collide(avbd_world* w);
/* solver.cpp:299-337 */
int  avbd_stage_predict(avbd_world* w);
/* body/manifold adjacency + greedy colouring; colour_of gets one int per body (-2 static). */
int  avbd_stage_colour(avbd_world* w);
int  avbd_download_colours(avbd_world* w, int* colour_of, int* num_colours);
/* one primal sweep over all colours (solver.cpp:344-409); dx_out6 (may be NULL) gets each dynamic body's solve. */
int  avbd_stage_primal(avbd_world* w, float alpha, float* dx_out6);
/* solver.cpp:411-430 */
int  avbd_stage_dual(avbd_world* w, float alpha);
/* solver.cpp:434-497 */
int  avbd_stage_velocity(avbd_world* w);

/* ---- stand-alone kernels on caller data ------------------------------------------------------- */
/* Manifold::collide (collision.cpp:420) on n pairs; a10/b10 = size3 pos3 quat4; out: counts[n], feats[4n], geom[36n] (rA3 rB3 n3 per contact). */
int  avbd_collide_pairs(int device, int n, const float* a10, const float* b10, int* counts, int* feats4, float* geom36);
/* solve6x6 (solver.cpp:68-83) on n systems; lhs36 = ll la al aa column-major blocks (al must equal la^T). */
int  avbd_solve6x6(int device, int n, const float* lhs36, const float* rhs6, float* out6);

/* Solver::pick (solver.cpp:145-228): closest dynamic OBB hit by a ray; returns body index or -1, local hit point in local3. */
int  avbd_pick(avbd_world* w, const float* origin3, const float* dir3, float* local3);

#ifdef __cplusplus
}
#endif
#endif
