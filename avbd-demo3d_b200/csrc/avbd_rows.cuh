// avbd_rows.cuh — per-contact row math shared by the primal, dual, warm-start
// and diagnostics kernels.  One "contact" = 3 rows (normal, tangent1, tangent2)
// of the reference's Manifold force (solver.h:112-143).
#pragma once
#include "avbd_collide.cuh"
#include "avbd_world.cuh"

namespace avbd {

struct ContactState {     // registers-resident view of one contact slot
    V3 rA, rB, n;
    float C0n, C0t1, C0t2;
    float lam[3], pen[3];
    bool stick;
    int feature;
};

AVBD_HD ContactState unpack_contact(float4 a, float4 b, float4 n, float4 l, float4 p) {
    ContactState c;
    c.rA = xyz(a); c.C0n = a.w;
    c.rB = xyz(b); c.C0t1 = b.w;
    c.n = xyz(n);  c.C0t2 = n.w;
    c.lam[0] = l.x; c.lam[1] = l.y; c.lam[2] = l.z; c.stick = l.w != 0.0f;
    c.pen[0] = p.x; c.pen[1] = p.y; c.pen[2] = p.z; c.feature = f2i(p.w);
    return c;
}
AVBD_HD float4 pack_lambda(const ContactState& c) { return make_float4(c.lam[0], c.lam[1], c.lam[2], c.stick ? 1.0f : 0.0f); }
AVBD_HD float4 pack_penalty(const ContactState& c) { return make_float4(c.pen[0], c.pen[1], c.pen[2], i2f(c.feature)); }

struct ContactEval {      // transient outputs of computeConstraint for one contact
    float C[3], fmin[3], fmax[3];
    V3 basis[3];          // normal, tangent1, tangent2
    V3 wrA, wrB;          // rotate(rot, r): world lever arms
};

// Manifold::computeConstraint for ONE contact (manifold.cpp:177-245), in two halves so a kernel that evaluates the
// constraint twice on the SAME poses (the deferred dual update followed by the primal rows) pays for the pose-dependent
// half once.  contact_geometry: basis, world lever arms, separation along the basis (manifold.cpp:183-197).
AVBD_HD void contact_geometry(V3 posA, Q4 rotA, V3 posB, Q4 rotB, const ContactState& c, ContactEval& e, float (&sep)[3]) {
    e.basis[0] = c.n;
    contact_basis_unit(c.n, e.basis[1], e.basis[2]);
    e.wrA = qrot(rotA, c.rA);
    e.wrB = qrot(rotB, c.rB);
    V3 dlt = (posA + e.wrA) - (posB + e.wrB);
    sep[0] = dot(dlt, e.basis[0]) - kNormalContactMargin;
    sep[1] = dot(dlt, e.basis[1]); sep[2] = dot(dlt, e.basis[2]);
}
// contact_limits: C, force bounds, friction cone (manifold.cpp:198-241).  Stateful exactly like the reference: clamps
// the warm tangential lambda into the current cone and re-evaluates `stick` — the caller writes c.lam / c.stick back.
AVBD_HD void contact_limits(float invMassA, float invMassB, float mu0, float alpha, const float (&sep)[3], ContactState& c, ContactEval& e) {
    float bias = clampf(1.0f - alpha, 0.0f, 1.0f);
    e.C[0] = sep[0] + bias * c.C0n;
    float ims = invMassA + invMassB;
    float mscale = (ims > 1.0e-6f) ? (1.0f / ims) : 1.0f;
    float cap = kNormalForceCap * mscale;
    e.fmin[0] = -cap; e.fmax[0] = 0.0f;
    e.C[1] = sep[1] + bias * c.C0t1;
    e.C[2] = sep[2] + bias * c.C0t2;
    float warmN = fabsf(fmin2(c.lam[0], 0.0f));
    float trial = c.pen[0] * e.C[0] + c.lam[0];
    float trialN = fabsf(fmin2(trial, 0.0f));
    float nmag = fmin2(fmax2(warmN, trialN), cap);
    float mu = mu0;
    if (!c.stick) mu *= kKineticFrictionScale;
    float lim = mu * nmag;
    float l1 = c.lam[1], l2 = c.lam[2];
    float tm = sqrtf(l1 * l1 + l2 * l2);
    if (tm > lim && tm > 1.0e-8f) { float s = lim / tm; c.lam[1] *= s; c.lam[2] *= s; }
    e.fmin[1] = -lim; e.fmax[1] = lim; e.fmin[2] = -lim; e.fmax[2] = lim;
    float slip2 = e.C[1] * e.C[1] + e.C[2] * e.C[2];
    float tl2 = c.lam[1] * c.lam[1] + c.lam[2] * c.lam[2];
    c.stick = (slip2 <= kStickThresh * kStickThresh) && (tl2 <= lim * lim + 1.0e-8f);
}
AVBD_HD void contact_constraint(V3 posA, Q4 rotA, float invMassA, V3 posB, Q4 rotB, float invMassB,
                                float mu0, float alpha, ContactState& c, ContactEval& e) {
    float sep[3];
    contact_geometry(posA, rotA, posB, rotB, c, e, sep);
    contact_limits(invMassA, invMassB, mu0, alpha, sep, c, e);
}

// 6x6 block system of one body, stored as the 27 numbers the Schur solve reads:
// rhs (6), ll lower triangle (6), la full (9, al == la^T bit for bit), aa lower (6).
struct BodySystem {
    float rl[3], ra[3];
    float ll[6];          // (0,0) (1,0) (2,0) (1,1) (2,1) (2,2)
    float la[9];          // la[r*3+c] = element (r,c)
    float aa[6];
    AVBD_HD void clear() {
        for (int i = 0; i < 3; ++i) { rl[i] = 0.0f; ra[i] = 0.0f; }
        for (int i = 0; i < 6; ++i) { ll[i] = 0.0f; aa[i] = 0.0f; }
        for (int i = 0; i < 9; ++i) la[i] = 0.0f;
    }
};

// Adds one row (solver.cpp:374-399).  `gyro` enables the manifold-only
// diagonal term diag(|Ja x I^-1 Ja| * |f|) (solver.cpp:393-397).
AVBD_HD void accumulate_row(BodySystem& s, V3 Jl, V3 Ja, float f, float pen, bool gyro, const M3& invIw) {
    s.rl[0] += Jl.x * f; s.rl[1] += Jl.y * f; s.rl[2] += Jl.z * f;
    s.ra[0] += Ja.x * f; s.ra[1] += Ja.y * f; s.ra[2] += Ja.z * f;
    if (pen > 0.0f && finite1(pen)) {
        // (J*pen) J^T: one multiply per vector entry, then one multiply-add per matrix entry
        float lp[3] = {Jl.x * pen, Jl.y * pen, Jl.z * pen}, ap[3] = {Ja.x * pen, Ja.y * pen, Ja.z * pen};
        float jl[3] = {Jl.x, Jl.y, Jl.z}, ja[3] = {Ja.x, Ja.y, Ja.z};
        s.ll[0] += lp[0] * jl[0]; s.ll[1] += lp[1] * jl[0]; s.ll[2] += lp[2] * jl[0];
        s.ll[3] += lp[1] * jl[1]; s.ll[4] += lp[2] * jl[1]; s.ll[5] += lp[2] * jl[2];
        s.aa[0] += ap[0] * ja[0]; s.aa[1] += ap[1] * ja[0]; s.aa[2] += ap[2] * ja[0];
        s.aa[3] += ap[1] * ja[1]; s.aa[4] += ap[2] * ja[1]; s.aa[5] += ap[2] * ja[2];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) s.la[r * 3 + c] += lp[r] * ja[c];
        if (gyro) {
            V3 g = vabs(cross(Ja, mv(invIw, Ja)));
            float af = fabsf(f);
            s.aa[0] += g.x * af; s.aa[3] += g.y * af; s.aa[5] += g.z * af;
        }
    }
}

// Primal contribution of one contact as seen from body side `isA`
// (solver.cpp:371-399 with Manifold::computeDerivatives, manifold.cpp:247-271).
AVBD_HD void accumulate_contact(BodySystem& s, const ContactState& c, const ContactEval& e, bool isA, const M3& invIw, bool gyro = true) {
    float sg = isA ? 1.0f : -1.0f;
    V3 wr = isA ? e.wrA : e.wrB;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V3 Jl = e.basis[r] * sg;
        V3 Ja = cross(wr, e.basis[r]) * sg;
        float f = clampf(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        accumulate_row(s, Jl, Ja, f, c.pen[r], gyro, invIw);
    }
}

// The same contribution as accumulate_contact, OVERWRITING `s` (the solver kernels' form: one visit = one partial sum
// that the kernel then adds in visit order).  The body-side sign is folded into f: (sg J) f = J (sg f) and
// pen (sg J)(sg J)^T = pen J J^T hold bit for bit, so every one of the 27 numbers equals accumulate_contact's on a cleared
// system (tests/test_abi_and_host.py).  The outer products stay per row ON PURPOSE: re-associating them through
// M = sum_r p_r b_r b_r^T is ~45 instructions cheaper but loses up to the penalty ratio (100x) in relative accuracy when the
// lever arm is nearly parallel to the stiffest row, and the Schur complement amplifies that by the system's condition.
// `w` = the visiting body's world lever arm, `sg` = +1 when it is body A of the manifold, -1 when it is body B.
AVBD_HD void contact_system_w(BodySystem& s, const ContactState& c, const ContactEval& e, V3 w, float sg, bool gyro, const M3& invIw) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V3 Jl = e.basis[r];
        V3 Ja = cross(w, e.basis[r]);
        float f0 = clampf(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        float f = f0 * sg;
        float pen = c.pen[r];
        bool stiff = pen > 0.0f && finite1(pen);                 // solver.cpp:381: otherwise the row adds no stiffness
        if (!stiff) pen = 0.0f;
        float lp[3] = {Jl.x * pen, Jl.y * pen, Jl.z * pen}, ap[3] = {Ja.x * pen, Ja.y * pen, Ja.z * pen};
        float jl[3] = {Jl.x, Jl.y, Jl.z}, ja[3] = {Ja.x, Ja.y, Ja.z};
        if (r == 0) {
            s.rl[0] = Jl.x * f; s.rl[1] = Jl.y * f; s.rl[2] = Jl.z * f;
            s.ra[0] = Ja.x * f; s.ra[1] = Ja.y * f; s.ra[2] = Ja.z * f;
            s.ll[0] = lp[0] * jl[0]; s.ll[1] = lp[1] * jl[0]; s.ll[2] = lp[2] * jl[0];
            s.ll[3] = lp[1] * jl[1]; s.ll[4] = lp[2] * jl[1]; s.ll[5] = lp[2] * jl[2];
            s.aa[0] = ap[0] * ja[0]; s.aa[1] = ap[1] * ja[0]; s.aa[2] = ap[2] * ja[0];
            s.aa[3] = ap[1] * ja[1]; s.aa[4] = ap[2] * ja[1]; s.aa[5] = ap[2] * ja[2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) s.la[i * 3 + k] = lp[i] * ja[k];
        } else {
            s.rl[0] += Jl.x * f; s.rl[1] += Jl.y * f; s.rl[2] += Jl.z * f;
            s.ra[0] += Ja.x * f; s.ra[1] += Ja.y * f; s.ra[2] += Ja.z * f;
            s.ll[0] += lp[0] * jl[0]; s.ll[1] += lp[1] * jl[0]; s.ll[2] += lp[2] * jl[0];
            s.ll[3] += lp[1] * jl[1]; s.ll[4] += lp[2] * jl[1]; s.ll[5] += lp[2] * jl[2];
            s.aa[0] += ap[0] * ja[0]; s.aa[1] += ap[1] * ja[0]; s.aa[2] += ap[2] * ja[0];
            s.aa[3] += ap[1] * ja[1]; s.aa[4] += ap[2] * ja[1]; s.aa[5] += ap[2] * ja[2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) s.la[i * 3 + k] += lp[i] * ja[k];
        }
        if (gyro && stiff) {                                     // solver.cpp:393-397; exactly zero for isotropic inertia
            V3 g = vabs(cross(Ja, mv(invIw, Ja)));
            float af = fabsf(f0);
            s.aa[0] += g.x * af; s.aa[3] += g.y * af; s.aa[5] += g.z * af;
        }
    }
}
AVBD_HD void contact_system(BodySystem& s, const ContactState& c, const ContactEval& e, bool isA, bool gyro, const M3& invIw) {
    contact_system_w(s, c, e, isA ? e.wrA : e.wrB, isA ? 1.0f : -1.0f, gyro, invIw);
}
AVBD_HD void add_system(BodySystem& s, const BodySystem& o) {
    for (int i = 0; i < 3; ++i) { s.rl[i] += o.rl[i]; s.ra[i] += o.ra[i]; }
    for (int i = 0; i < 6; ++i) { s.ll[i] += o.ll[i]; s.aa[i] += o.aa[i]; }
    for (int i = 0; i < 9; ++i) s.la[i] += o.la[i];
}

// solver.cpp:68-83 on the packed system.
AVBD_HD void solve_body_system(const BodySystem& s, V3& dl, V3& da) {
    M3 ll = m3(mk3(s.ll[0], s.ll[1], s.ll[2]), mk3(s.ll[1], s.ll[3], s.ll[4]), mk3(s.ll[2], s.ll[4], s.ll[5]));
    M3 aa = m3(mk3(s.aa[0], s.aa[1], s.aa[2]), mk3(s.aa[1], s.aa[3], s.aa[4]), mk3(s.aa[2], s.aa[4], s.aa[5]));
    // la columns: column c = (la(0,c), la(1,c), la(2,c)); al = la^T so al column c = row c of la
    M3 la = m3(mk3(s.la[0], s.la[3], s.la[6]), mk3(s.la[1], s.la[4], s.la[7]), mk3(s.la[2], s.la[5], s.la[8]));
    M3 al = m3(mk3(s.la[0], s.la[1], s.la[2]), mk3(s.la[3], s.la[4], s.la[5]), mk3(s.la[6], s.la[7], s.la[8]));
    V3 bl = mk3(s.rl[0], s.rl[1], s.rl[2]), ba = mk3(s.ra[0], s.ra[1], s.ra[2]);
    Ldl3R fl = ldl3r_factor(ll);
    M3 W = m3(ldl3r_solve(fl, la.c[0]), ldl3r_solve(fl, la.c[1]), ldl3r_solve(fl, la.c[2]));
    V3 x0 = ldl3r_solve(fl, bl);
    M3 S;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        V3 prod = mv(al, W.c[j]);
        S.c[j] = aa.c[j] - prod;
    }
    V3 rs = ba - mv(al, x0);
    da = ldl3r_solve(ldl3r_factor(S), rs);
    dl = x0 - mv(W, da);
}

// Dual + penalty ramp for one contact (solver.cpp:411-430, rowPenaltyGain :94-125).
AVBD_HD void dual_contact(ContactState& c, const ContactEval& e, float beta) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float lu = clampf(c.pen[r] * e.C[r] + c.lam[r], e.fmin[r], e.fmax[r]);
        bool active = lu > e.fmin[r] && lu < e.fmax[r];
        c.lam[r] = lu;
        if (active) {
            V3 b = e.basis[r];
            float lw = 0.0f, aw = 0.0f;
            lw += len2(b * 1.0f);  aw += len2(cross(e.wrA, b) * 1.0f);
            lw += len2(b * -1.0f); aw += len2(cross(e.wrB, b) * -1.0f);
            float tot = lw + aw;
            float br = beta;
            if (!(tot < 1.0e-8f)) br = (beta * lw + (beta * kAngularBetaScale) * aw) / tot;
            c.pen[r] = fmin2(c.pen[r] + br * fabsf(e.C[r]), kManifoldPenaltyCap);
        }
    }
}

struct OldManifold {      // last step's contacts of the same pair, for feature-id carry-over
    int n;
    ContactState ct[4];
};

struct NewManifold {
    int n;
    ContactState ct[4];
};

// One contact of Manifold::initialize (manifold.cpp:100-171) followed by the per-step warm-start decay of
// Solver::step (solver.cpp:281-293; manifold rows are hard so the stiffness cap at :290-292 never applies).
// `oldFeat[j]` (j < oldN) are last step's feature keys of the same pair, `used` the bit mask of old contacts already
// matched (first unused equal feature wins, manifold.cpp:111-119); `loadOld(j)` fetches old contact j on a match.
// `cache` (optional): the contacts of a face manifold all carry the SAME normal, bit for bit, so its unit vector and tangent frame —
// five square roots and fifteen IEEE divisions per contact — are worth remembering from one contact of a manifold to the next.  Keyed by
// the normal's bit pattern: a hit returns exactly what the computation would.
struct BasisCache {
    bool valid = false; V3 key, n, t1, t2;
    bool oldValid = false; V3 oldKey, oldFb, oldUnit;
};
AVBD_HD bool same_bits(V3 a, V3 b) { return f2i(a.x) == f2i(b.x) && f2i(a.y) == f2i(b.y) && f2i(a.z) == f2i(b.z); }
template <class LoadOld>
AVBD_HD ContactState contact_initialize(V3 posA, Q4 rotA, V3 posB, Q4 rotB, int feature, V3 rA, V3 rB, V3 normal,
                                        int oldN, const int (&oldFeat)[4], unsigned& used, LoadOld&& loadOld, const SolveParams& prm,
                                        BasisCache* cache = nullptr) {
    ContactState c;
    c.feature = feature; c.rA = rA; c.rB = rB; c.n = normal;
#pragma unroll
    for (int k = 0; k < 3; ++k) { c.lam[k] = 0.0f; c.pen[k] = kPenaltyMin; }
    c.stick = false;
    int hit = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (hit < 0 && j < oldN && !((used >> j) & 1u) && feature == oldFeat[j]) hit = j;
    if (hit >= 0) {
        used |= 1u << hit;
        ContactState o = loadOld(hit);
        if (cache && !(cache->valid && same_bits(cache->key, c.n))) {
            cache->key = c.n; contact_basis(c.n, cache->n, cache->t1, cache->t2); cache->valid = true;
        }
        V3 nn = cache ? cache->n : unit_or(c.n, mk3(0.0f, 1.0f, 0.0f));        // contact_basis' n is this very unit_or
        V3 on;
        if (cache && cache->oldValid && same_bits(cache->oldKey, o.n) && same_bits(cache->oldFb, nn)) on = cache->oldUnit;
        else {
            on = unit_or(o.n, nn);
            if (cache) { cache->oldKey = o.n; cache->oldFb = nn; cache->oldUnit = on; cache->oldValid = true; }
        }
        float nd = dot(nn, on);
        V3 oldMid = ((posA + qrot(rotA, o.rA)) + (posB + qrot(rotB, o.rB))) * 0.5f;
        V3 newMid = ((posA + qrot(rotA, c.rA)) + (posB + qrot(rotB, c.rB))) * 0.5f;
        float drift2 = len2(newMid - oldMid);
        bool warm = (nd >= kWarmNormalMinDot) && (drift2 <= kWarmMaxDrift * kWarmMaxDrift);
        if (warm) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { c.lam[k] = o.lam[k]; c.pen[k] = clampf(o.pen[k], kPenaltyMin, kManifoldPenaltyCap); }
        }
        bool reuse = false;
        if (o.stick && warm) reuse = (nd >= kStickNormalMinDot) && (drift2 <= kStickAnchorMaxDrift * kStickAnchorMaxDrift);
        c.stick = o.stick && reuse;
        if (reuse) { c.rA = o.rA; c.rB = o.rB; }
    }
    V3 n, t1, t2;
    if (cache && cache->valid && same_bits(cache->key, c.n)) { n = cache->n; t1 = cache->t1; t2 = cache->t2; }
    else {
        contact_basis(c.n, n, t1, t2);
        if (cache) { cache->key = c.n; cache->n = n; cache->t1 = t1; cache->t2 = t2; cache->valid = true; }
    }
    c.n = n;
    V3 dlt = (posA + qrot(rotA, c.rA)) - (posB + qrot(rotB, c.rB));
    c.C0n = dot(dlt, n) - kNormalContactMargin;
    c.C0t1 = dot(dlt, t1);
    c.C0t2 = dot(dlt, t2);
    // warm-start decay, solver.cpp:281-288
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!prm.postStabilize) c.lam[k] *= prm.alpha * prm.gamma;
        c.pen[k] = clampf(c.pen[k] * prm.gamma, kPenaltyMin, kPenaltyMax);
    }
    return c;
}

// Manifold::initialize on arrays (host mirror, host emulation; not compiled into the CUDA translation units): the same
// builder and per-contact routine the kernels run.
#ifndef __CUDACC__
inline void manifold_initialize(V3 posA, Q4 rotA, V3 sizeA, V3 posB, Q4 rotB, V3 sizeB, int satCode,
                                const OldManifold& old, const SolveParams& prm, NewManifold& out) {
    int oldFeat[4] = {0, 0, 0, 0};
    for (int j = 0; j < old.n && j < 4; ++j) oldFeat[j] = old.ct[j].feature;
    unsigned used = 0u;
    out.n = 0;
    auto loadOld = [&](int j) { return old.ct[j]; };
    auto emit = [&](int feature, V3 rA, V3 rB, V3 normal) {
        out.ct[out.n++] = contact_initialize(posA, rotA, posB, rotB, feature, rA, rB, normal, old.n, oldFeat, used, loadOld, prm);
    };
    PolyLocal poly;
    build_contacts_emit(posA, rotA, sizeA, posB, rotB, sizeB, satCode, poly, emit);
}
#endif

} // namespace avbd
