// tests/emul/host_emul.cpp — TEST INFRASTRUCTURE ONLY (never shipped, never a fallback).
//
// Compiles the product's __host__ __device__ row/body/collision functions
// (avbd-demo3d_b200/csrc/*.cuh) for the HOST and drives them serially, so the
// device-side math can be checked against the oracle on a box without a GPU.
// It deliberately re-creates none of the GPU plumbing (sort/scan/colouring are
// brute force here); the -m gpu tests cover the real pipeline.
#include "avbd_body.cuh"
#include "avbd_forces.cuh"
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

using namespace avbd;

struct EBody { V3 pos; Q4 rot; V3 lin, ang, prevLin; V3 size; float friction; BodyAux aux; BodyInit init; float invMass; };
struct EManifold { int a, b; int n; float mu; ContactState ct[4]; };
struct EWorld {
    SolveParams prm;
    std::vector<EBody> bodies;
    std::map<std::pair<int, int>, EManifold> manifolds;   // key (a,b), a > b
    std::vector<int> order;                                // primal visiting order (dynamic bodies)
};

extern "C" {

void* emu_create() {
    EWorld* w = new EWorld();
    w->prm.dt = 1.0f / 60.0f; w->prm.gx = 0; w->prm.gy = -10.0f; w->prm.gz = 0; w->prm.iterations = 10;
    w->prm.alpha = 0.95f; w->prm.beta = 100000.0f; w->prm.gamma = 0.99f; w->prm.postStabilize = 0;
    return w;
}
void emu_destroy(void* h) { delete (EWorld*)h; }
void emu_set_params(void* h, float dt, const float* g, int it, float alpha, float beta, float gamma, int ps) {
    SolveParams& p = ((EWorld*)h)->prm;
    p.dt = dt; p.gx = g[0]; p.gy = g[1]; p.gz = g[2]; p.iterations = it; p.alpha = alpha; p.beta = beta; p.gamma = gamma; p.postStabilize = ps;
}
int emu_add_body(void* h, const float* size, float density, float friction, const float* pos, const float* q, const float* lin, const float* ang) {
    EWorld* w = (EWorld*)h;
    EBody b{};
    float sx = size[0], sy = size[1], sz = size[2];
    float mass = sx * sy * sz * density;
    b.invMass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
    float ixx = 0, iyy = 0, izz = 0;
    if (b.invMass > 0.0f) {
        ixx = (1.0f / 12.0f) * mass * (sy * sy + sz * sz);
        iyy = (1.0f / 12.0f) * mass * (sx * sx + sz * sz);
        izz = (1.0f / 12.0f) * mass * (sx * sx + sy * sy);
    }
    b.pos = mk3(pos[0], pos[1], pos[2]); b.rot = qmk(q[0], q[1], q[2], q[3]);
    b.lin = mk3(lin[0], lin[1], lin[2]); b.ang = mk3(ang[0], ang[1], ang[2]); b.prevLin = b.lin;
    b.size = mk3(sx, sy, sz); b.friction = friction;
    b.aux.mass = make_float4(mass, b.invMass, friction, sqrtf(sx * sx + sy * sy + sz * sz) * 0.5f);
    b.aux.inert = make_float4(ixx, iyy, izz, 0.0f);
    w->bodies.push_back(b);
    return (int)w->bodies.size() - 1;
}
void emu_set_order(void* h, const int* order, int n) { ((EWorld*)h)->order.assign(order, order + n); }

static void collide_stage(EWorld* w) {
    int n = (int)w->bodies.size();
    std::map<std::pair<int, int>, EManifold> next;
    for (int a = n - 1; a >= 0; --a)
        for (int b = a - 1; b >= 0; --b) {
            const EBody& A = w->bodies[a]; const EBody& B = w->bodies[b];
            auto it = w->manifolds.find({a, b});
            bool persisting = it != w->manifolds.end() && it->second.n > 0;
            V3 dp = A.pos - B.pos;
            float r = A.aux.mass.w + B.aux.mass.w;
            if (!(dot(dp, dp) <= r * r) && !persisting) continue;
            int code = sat_test(make_obb(A.pos, A.rot, A.size), make_obb(B.pos, B.rot, B.size));
            if (!code) continue;
            OldManifold om; om.n = 0;
            if (persisting) { om.n = it->second.n; for (int i = 0; i < om.n; ++i) om.ct[i] = it->second.ct[i]; }
            NewManifold nm;
            manifold_initialize(A.pos, A.rot, A.size, B.pos, B.rot, B.size, code, om, w->prm, nm);
            EManifold m; m.a = a; m.b = b; m.n = nm.n; m.mu = sqrtf(A.friction * B.friction);
            for (int i = 0; i < nm.n; ++i) m.ct[i] = nm.ct[i];
            next[{a, b}] = m;
        }
    w->manifolds.swap(next);
}

static void predict_stage(EWorld* w) {
    for (EBody& b : w->bodies) {
        BodyPose pose; pose.pos = f4(b.pos, b.invMass); pose.rot = f4(b.rot);
        BodyVel vel; vel.lin = f4(b.lin, 0); vel.ang = f4(b.ang, 0);
        predict_body(pose, vel, f4(b.prevLin, 0), b.aux, b.init, w->prm);
        b.pos = xyz(pose.pos); b.rot = quat(pose.rot); b.lin = xyz(vel.lin); b.ang = xyz(vel.ang);
    }
}

static void primal_stage(EWorld* w, float alpha, float* dxOut) {
    for (int i : w->order) {
        EBody& b = w->bodies[i];
        if (b.invMass <= 0.0f) continue;
        BodySystem sys; M3 invIw;
        body_self_system(b.pos, b.rot, b.aux, w->prm.dt, sys, invIw);
        // A-side manifolds in key order, then B-side in key order (what the kernel's adjacency gives)
        for (int side = 0; side < 2; ++side)
            for (auto& kv : w->manifolds) {
                EManifold& m = kv.second;
                bool isA = m.a == i;
                if (side == 0 ? !isA : m.b != i) continue;
                const EBody& A = w->bodies[m.a]; const EBody& B = w->bodies[m.b];
                for (int c = 0; c < m.n; ++c) {
                    ContactEval ev;
                    contact_constraint(A.pos, A.rot, A.invMass, B.pos, B.rot, B.invMass, m.mu, alpha, m.ct[c], ev);
#ifdef PARTIAL_SUMS
                    { BodySystem part; contact_system(part, m.ct[c], ev, isA, true, invIw); add_system(sys, part); }
#else
                    accumulate_contact(sys, m.ct[c], ev, isA, invIw);
#endif
                }
            }
        V3 dl, da;
        solve_body_system(sys, dl, da);
        apply_body_update(b.pos, b.rot, dl, da);
        if (dxOut) { float* o = dxOut + 6 * i; o[0] = dl.x; o[1] = dl.y; o[2] = dl.z; o[3] = da.x; o[4] = da.y; o[5] = da.z; }
    }
}

static void dual_stage(EWorld* w, float alpha) {
    for (auto& kv : w->manifolds) {
        EManifold& m = kv.second;
        const EBody& A = w->bodies[m.a]; const EBody& B = w->bodies[m.b];
        for (int c = 0; c < m.n; ++c) {
            ContactEval ev;
            contact_constraint(A.pos, A.rot, A.invMass, B.pos, B.rot, B.invMass, m.mu, alpha, m.ct[c], ev);
            dual_contact(m.ct[c], ev, w->prm.beta);
        }
    }
}

static void velocity_stage(EWorld* w) {
    for (EBody& b : w->bodies) {
        if (b.invMass <= 0.0f) continue;
        BodyPose pose; pose.pos = f4(b.pos, b.invMass); pose.rot = f4(b.rot);
        BodyVel vel; vel.lin = f4(b.lin, 0); vel.ang = f4(b.ang, 0);
        float4 pl; float ls, as;
        velocity_body(pose, b.init, vel, pl, w->prm.dt, ls, as);
        b.lin = xyz(vel.lin); b.ang = xyz(vel.ang); b.prevLin = xyz(pl);
    }
}

void emu_stage_collide(void* h) { collide_stage((EWorld*)h); }
void emu_stage_predict(void* h) { predict_stage((EWorld*)h); }
void emu_stage_primal(void* h, float alpha, float* dx) { primal_stage((EWorld*)h, alpha, dx); }
void emu_stage_dual(void* h, float alpha) { dual_stage((EWorld*)h, alpha); }
void emu_stage_velocity(void* h) { velocity_stage((EWorld*)h); }

void emu_step(void* h) {
    EWorld* w = (EWorld*)h;
    collide_stage(w); predict_stage(w);
    int total = w->prm.iterations + (w->prm.postStabilize ? 1 : 0);
    for (int it = 0; it < total; ++it) {
        float a = w->prm.postStabilize ? (it < w->prm.iterations ? 1.0f : 0.0f) : w->prm.alpha;
        primal_stage(w, a, nullptr);
        if (it < w->prm.iterations) dual_stage(w, a);
    }
    velocity_stage(w);
}

void emu_get_state(void* h, float* o) {
    for (const EBody& b : ((EWorld*)h)->bodies) {
        *o++ = b.pos.x; *o++ = b.pos.y; *o++ = b.pos.z; *o++ = b.rot.x; *o++ = b.rot.y; *o++ = b.rot.z; *o++ = b.rot.w;
        *o++ = b.lin.x; *o++ = b.lin.y; *o++ = b.lin.z; *o++ = b.ang.x; *o++ = b.ang.y; *o++ = b.ang.z;
    }
}
int emu_num_manifolds(void* h) { return (int)((EWorld*)h)->manifolds.size(); }
// same layout as orc_get_manifolds, in (a,b) ascending order
void emu_get_manifolds(void* h, int* ints, int* feats, int* stick, float* flts) {
    for (auto& kv : ((EWorld*)h)->manifolds) {
        const EManifold& m = kv.second;
        *ints++ = m.a; *ints++ = m.b; *ints++ = m.n;
        *flts++ = m.mu;
        for (int i = 0; i < 4; ++i) {
            bool live = i < m.n; const ContactState& c = m.ct[i];
            *feats++ = live ? c.feature : 0; *stick++ = live ? (c.stick ? 1 : 0) : 0;
            const float v[14] = {c.rA.x, c.rA.y, c.rA.z, c.rB.x, c.rB.y, c.rB.z, c.n.x, c.n.y, c.n.z, 0.0f, c.C0n, c.C0t1, c.C0t2, 0.0f};
            for (int k = 0; k < 14; ++k) *flts++ = live ? v[k] : 0.0f;
        }
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) *flts++ = i < m.n ? m.ct[i].lam[k] : 0.0f;
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) *flts++ = i < m.n ? m.ct[i].pen[k] : 0.0f;
    }
}

// contact_system (the solver kernels' fused three-row form) against accumulate_contact (row by row, the reference's
// association) on one contact: in[0..6] pose A, [7..13] pose B, [14..22] rA rB n, [23..25] C0, [26..28] lambda,
// [29..31] penalty, [32] stick, [33] mu, [34] alpha, [35..37] inertia diag of the visiting body, [38] isA.
// out: 27 fused then 27 row-by-row (rl ra ll la aa).
void emu_contact_system(const float* in, float* out) {
    V3 pA = mk3(in[0], in[1], in[2]); Q4 qA = qunit(qmk(in[3], in[4], in[5], in[6]));
    V3 pB = mk3(in[7], in[8], in[9]); Q4 qB = qunit(qmk(in[10], in[11], in[12], in[13]));
    ContactState c{};
    c.rA = mk3(in[14], in[15], in[16]); c.rB = mk3(in[17], in[18], in[19]); c.n = unit_or(mk3(in[20], in[21], in[22]), mk3(0, 1, 0));
    c.C0n = in[23]; c.C0t1 = in[24]; c.C0t2 = in[25];
    for (int k = 0; k < 3; ++k) { c.lam[k] = in[26 + k]; c.pen[k] = in[29 + k]; }
    c.stick = in[32] != 0.0f;
    bool isA = in[38] != 0.0f;
    V3 I = mk3(in[35], in[36], in[37]);
    M3 invIw = rot_diag(qmat(isA ? qA : qB), mk3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z));
    ContactState c1 = c, c2 = c;
    ContactEval e1, e2;
    contact_constraint(pA, qA, 1.0f, pB, qB, 1.0f, in[33], in[34], c1, e1);
    contact_constraint(pA, qA, 1.0f, pB, qB, 1.0f, in[33], in[34], c2, e2);
    BodySystem a, b; b.clear();
    contact_system(a, c1, e1, isA, true, invIw);
    accumulate_contact(b, c2, e2, isA, invIw);
    const BodySystem* two[2] = {&a, &b};
    for (int t = 0; t < 2; ++t) {
        const BodySystem& s = *two[t];
        float* o = out + 27 * t;
        for (int k = 0; k < 3; ++k) { o[k] = s.rl[k]; o[3 + k] = s.ra[k]; }
        for (int k = 0; k < 6; ++k) { o[6 + k] = s.ll[k]; o[21 + k] = s.aa[k]; }
        for (int k = 0; k < 9; ++k) o[12 + k] = s.la[k];
    }
}

int emu_collide(const float* a, const float* c, int* feats, float* out) {
    V3 sa = mk3(a[0], a[1], a[2]), pa = mk3(a[3], a[4], a[5]); Q4 qa = qmk(a[6], a[7], a[8], a[9]);
    V3 sb = mk3(c[0], c[1], c[2]), pb = mk3(c[3], c[4], c[5]); Q4 qb = qmk(c[6], c[7], c[8], c[9]);
    int code = sat_test(make_obb(pa, qa, sa), make_obb(pb, qb, sb));
    RawContact rc[4];
    int k = code ? build_contacts(pa, qa, sa, pb, qb, sb, code, rc) : 0;
    for (int j = 0; j < k; ++j) {
        feats[j] = rc[j].feature;
        float* o = out + 10 * j;
        o[0] = rc[j].rA.x; o[1] = rc[j].rA.y; o[2] = rc[j].rA.z; o[3] = rc[j].rB.x; o[4] = rc[j].rB.y; o[5] = rc[j].rB.z;
        o[6] = rc[j].normal.x; o[7] = rc[j].normal.y; o[8] = rc[j].normal.z; o[9] = 0.0f;
    }
    return k;
}

} // extern "C"
