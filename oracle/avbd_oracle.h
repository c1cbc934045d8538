/* oracle/avbd_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C-ABI of the CPU restatement ("port") of the reference's per-timestep solver
 * loop (alxspiker/avbd-demo3d, source/solver.cpp:255-514 and everything it
 * calls).  It exists to CHECK the CUDA path; the product never links, imports
 * or executes it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load libavbd_oracle.so.
 *
 * Parity pin: tests/test_oracle_pin.py asserts this restatement is
 * bit-identical to the unmodified reference (oracle/_ref/libavbd_ref.so and the
 * committed fixtures under tests/golden/ generated from it) on whole
 * trajectories, manifold dumps and per-pair narrowphase outputs.
 *
 * Indices are creation order (0 = first body).  Layouts match
 * oracle/ref_harness.cpp's ref_* functions one to one so a test can drive either.
 */
#ifndef AVBD_ORACLE_H
#define AVBD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

void* orc_create(void);
void  orc_destroy(void* w);
void  orc_clear(void* w);
void  orc_default_params(void* w);
void  orc_set_params(void* w, float dt, const float* gravity3, int iterations, float alpha, float beta, float gamma,
                     int postStabilize);
void  orc_get_params(void* w, float* out8);
void  orc_set_logging(void* w, int enableDiagnostics, int logFrequency);

int   orc_load_scene(void* w, const char* name);            /* scenes.h:186-209; -1 if unknown */
int   orc_load_scene_index(void* w, int idx);
int   orc_scene_count(void);
const char* orc_scene_name(int idx);
/* generalised Stress grid (scenes.h:86-132 with NX,NY,NZ free; ground widened when wideGround!=0) */
int   orc_load_stress_grid(void* w, int nx, int ny, int nz, float spacingY, float startY, int wideGround);

int   orc_add_body(void* w, const float* size3, float density, float friction, const float* pos3, const float* quat4,
                   const float* lin3, const float* ang3);
void  orc_add_joint(void* w, int a, int b, const float* anchorA3, const float* anchorB3, float linK, float angK);
void  orc_add_spring(void* w, int a, int b, const float* anchorA3, const float* anchorB3, float k, float rest);
void  orc_add_ignore(void* w, int a, int b);

void  orc_step(void* w, int n);                             /* Solver::step, solver.cpp:255 */
/* Same step, but the primal Gauss-Seidel sweep visits bodies in `order`
 * (dynamic bodies only need appear) instead of list order — lets a test hand
 * the GPU's colour order to the CPU algorithm. */
void  orc_step_ordered(void* w, const int* order, int n);

/* individual stages of one step (solver.cpp line ranges in avbd_oracle.cpp) */
void  orc_stage_broadphase(void* w);
void  orc_stage_init(void* w);
void  orc_stage_predict(void* w);
void  orc_stage_primal(void* w, float alpha, const int* order, int n, float* dx_out6);
void  orc_stage_dual(void* w, float alpha);
void  orc_stage_velocity(void* w);
void  orc_stage_diagnostics(void* w);

int   orc_num_bodies(void* w);
int   orc_body_id(void* w, int idx);
void  orc_get_state(void* w, float* out13);                 /* pos3 quat4 lin3 ang3 per body */
void  orc_set_state(void* w, const float* in13);
void  orc_get_prev_linvel(void* w, float* out3);
void  orc_set_prev_linvel(void* w, const float* in3);
void  orc_get_body_props(void* w, float* out10);            /* size3 mass invMass I3 friction radius */
void  orc_get_diagnostics(void* w, float* out5f, int* out3i);

int   orc_num_manifolds(void* w);
void  orc_get_manifolds(void* w, int* ints3, int* feats4, int* stick4, float* flts81);
/* test seam: replace the manifold set (same layout / order as orc_get_manifolds) */
void  orc_set_manifolds(void* w, int count, const int* ints3, const int* feats4, const int* stick4, const float* flts81);
/* sphere-overlap pairs (a > b) of the CURRENT poses, reference loop order; returns count (may exceed cap) */
int   orc_overlap_pairs(void* w, int* pairs2, int cap);

int   orc_collide(const float* a10, const float* b10, int* feats4, float* out40);   /* collision.cpp:420 */
void  orc_solve6x6(const float* lhs36, const float* rhs6, float* out6);             /* solver.cpp:68 */
int   orc_pick(void* w, const float* origin3, const float* dir3, float* local3);                /* solver.cpp:145 */
void  orc_solve3(const float* A9, const float* b3, float* out3);                    /* maths.h:104 */

#ifdef __cplusplus
}
#endif
#endif
