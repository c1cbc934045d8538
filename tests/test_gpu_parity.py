"""Parity of the CUDA path (through the C ABI) against the oracle.  Needs a GPU.

Bars (BASELINE.json north_star / SURVEY.md §8d):
  bit-exact   broadphase pair sets, contact feature ids, manifold membership, contact geometry, colouring validity
  FP32 tol    per-body 6x6 solve (dx) given identical inputs: <= 1e-4 * |dx|_inf + 1e-6
  trajectory  rest heights / counts / penetration (tests further down) — colour order differs from the serial order
"""
import os

import numpy as np
import pytest

from _libs import Oracle, add_all, copy_bodies, manifold_dict, random_pile, ref_available

pytestmark = pytest.mark.gpu

SCENES = ["Ground", "Stack", "Pyramid", "Wall", "TwoBlockDrop", "Stress1000", "Rod (WIP)", "Soft Body (WIP)"]
DX_RTOL, DX_ATOL = 1e-4, 1e-6          # stated FP32 tolerance for the 6x6 block solve
STEP_POS_TOL = 2e-5                    # one full step with the SAME colour order: summation-order rounding only


def make_pair(avbd, scene=None, bodies=None, params=None):
    """(oracle world, cuda world) holding the same bodies and params."""
    o = Oracle("port").create()
    if scene:
        o.load_scene(scene)
    else:
        add_all(o, bodies)
    if params:
        o.set_params(**params)
    p = o.params()
    w = avbd.World()
    w.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"])
    if bodies is not None:
        add_all(w, bodies)
    elif o.n:
        props, st = o.body_props(), o.state()
        size = props[:, 0:3]
        vol = size[:, 0].astype(np.float64) * size[:, 1] * size[:, 2]
        density = np.where(vol > 0, props[:, 3] / np.maximum(vol, 1e-30), 0.0).astype(np.float32)   # scenes use density 0 or 1: exact
        w.add_bodies(size, density, props[:, 8], st[:, 0:3], st[:, 3:7], st[:, 7:10], st[:, 10:13])
    return o, w


def gpu_manifolds(w):
    return manifold_dict(*w.manifolds_raw())


def pairset(arr):
    return set((int(a), int(b)) for a, b in arr)


def colour_order(w):
    col, k = w.colours()
    idx = np.arange(len(col))
    dyn = idx[col >= 0]
    return dyn[np.lexsort((dyn, col[dyn]))].astype(np.int32), col, k


# --------------------------------------------------------------------------- stand-alone kernels
def _random_pairs(rng, n):
    a = np.zeros((n, 10), np.float32)
    b = np.zeros((n, 10), np.float32)
    for t in range(n):
        for arr, c in ((a, rng.uniform(-0.2, 0.2, 3)), (b, rng.uniform(-1.2, 1.2, 3))):
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            if t % 5 == 0:
                q = np.array([0, 0, 0, 1.0])
            if t % 7 == 0:
                q = np.array([0, np.sin(0.3), 0, np.cos(0.3)])
            arr[t] = np.concatenate([rng.uniform(0.3, 1.5, 3), c, q])
    return a, b


def test_collide_pairs_bit_exact(avbd, port):
    """Manifold::collide (collision.cpp:420): counts, feature ids and contact geometry are bit-identical."""
    rng = np.random.default_rng(7)
    a, b = _random_pairs(rng, 4000)
    counts, feats, geom = avbd.collide_pairs(a, b)
    checkers = [port] + ([Oracle("ref")] if ref_available() else [])
    hits = 0
    for chk in checkers:
        for t in range(len(a)):
            k, f, g = chk.collide(a[t], b[t])
            assert counts[t] == k, (chk.kind, t, counts[t], k)
            assert (feats[t, :k] == f).all(), (chk.kind, t, feats[t], f)
            assert geom[t, :k].tobytes() == np.ascontiguousarray(g[:, :9]).tobytes(), (chk.kind, t)
            hits += k > 0
    assert hits > 1000


def test_solve6x6_tolerance(avbd, port):
    """solve6x6 (solver.cpp:68-83) on random SPD-ish systems, incl. the zero-pivot => zero-solution rule."""
    rng = np.random.default_rng(3)
    n = 2000
    lhs = np.zeros((n, 36), np.float32)
    rhs = rng.normal(size=(n, 6)).astype(np.float32) * 100
    for t in range(n):
        J = rng.normal(size=(8, 6))
        A = J.T @ np.diag(rng.uniform(1e3, 1e6, 8)) @ J + np.diag(rng.uniform(1e2, 1e4, 6))
        if t % 50 == 0:
            A[:] = 0
        ll, la, al, aa = A[:3, :3], A[:3, 3:], A[3:, :3], A[3:, 3:]
        lhs[t] = np.concatenate([blk.T.reshape(-1) for blk in (ll, la, al, aa)])   # column-major blocks
    lhs32 = lhs.copy()
    for t in range(n):    # al must equal la^T bit for bit: column-major al block = row-major la block
        lhs32[t, 18:27] = lhs32[t, 9:18].reshape(3, 3).T.reshape(-1)
    out = avbd.solve6x6(lhs32, rhs)
    for t in range(n):
        want = port.solve6x6(lhs32[t], rhs[t])
        tol = DX_RTOL * np.abs(want).max() + DX_ATOL
        assert np.abs(out[t] - want).max() <= tol, (t, out[t], want)


# --------------------------------------------------------------------------- broadphase
@pytest.mark.parametrize("scene", SCENES)
def test_broadphase_pair_set_exact(avbd, scene):
    """solver.cpp:262-266: the sphere-overlap pair set equals the reference's, pair for pair."""
    o, w = make_pair(avbd, scene=scene)
    try:
        for _ in range(3):
            got = pairset(w.stage_broadphase())
            want = pairset(o.overlap_pairs())
            assert got == want, (scene, len(got), len(want), sorted(got ^ want)[:10])
            o.step(5)
            w.set_state(o.state())
    finally:
        o.close(); w.close()


def test_broadphase_random_pile_exact(avbd):
    rng = np.random.default_rng(11)
    for n, spread in ((60, 2.0), (400, 5.0), (1500, 8.0)):
        o, w = make_pair(avbd, bodies=random_pile(rng, n, spread))
        try:
            got, want = pairset(w.stage_broadphase()), pairset(o.overlap_pairs())
            assert got == want, (n, len(got), len(want), sorted(got ^ want)[:10])
            assert len(want) > n
        finally:
            o.close(); w.close()


# --------------------------------------------------------------------------- narrowphase + warm start inside a world
def assert_manifolds_equal(got, want, exact_rows=True, ctx="", features_only=False):
    assert set(got) == set(want), (ctx, sorted(set(got) ^ set(want))[:10])
    for k in want:
        g, r = got[k], want[k]
        assert g["n"] == r["n"], (ctx, k, g["n"], r["n"])
        assert (g["feat"] == r["feat"]).all(), (ctx, k, g["feat"], r["feat"])
        assert g["mu"] == r["mu"], (ctx, k)
        if features_only:
            continue
        # geometry columns: rA3 rB3 n3 (pen skipped: draw-only) C0n C0t.x C0t.y
        cols = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12]
        assert g["geom"][:, cols].tobytes() == r["geom"][:, cols].tobytes(), (ctx, k, g["geom"], r["geom"])
        if exact_rows:
            assert g["lam"].tobytes() == r["lam"].tobytes(), (ctx, k, g["lam"], r["lam"])
            assert g["pen"].tobytes() == r["pen"].tobytes(), (ctx, k, g["pen"], r["pen"])
            assert (g["stick"] == r["stick"]).all(), (ctx, k)


@pytest.mark.parametrize("scene", ["Stack", "Pyramid", "Wall", "TwoBlockDrop"])
def test_collide_stage_bit_exact_from_reference_state(avbd, scene):
    """Manifold::initialize incl. feature-id warm start (manifold.cpp:71-175) + decay (solver.cpp:281-293):
    feed the GPU the oracle's body state each step; manifold membership, features, geometry, lambda and
    penalty carry-over must match bit for bit (same inputs, -fmad=false)."""
    o, w = make_pair(avbd, scene=scene)
    try:
        for step in range(12):
            # one full oracle step, then replay ITS collide stage on the GPU from the same pre-step state
            pre, prev = o.state(), o.prev_linvel()
            w.set_state(pre); w.set_prev_linvel(prev)
            w.stage("collide")
            o.stage("broadphase"); o.stage("init")
            want = o.manifolds()
            got = gpu_manifolds(w)
            # anchors / lambda / penalty carry over from each side's OWN history (stick anchors are reused,
            # manifold.cpp:144-155), which matches bit for bit only on the first step; afterwards the
            # history-free outputs (membership, contact count, feature ids) must still be identical
            assert_manifolds_equal(got, want, exact_rows=True, ctx=(scene, step), features_only=(step > 0))
            # finish the oracle's step
            o.stage("predict")
            p = o.params()
            for it in range(p["iterations"]):
                o.stage_primal(p["alpha"]); o.stage("dual", p["alpha"])
            o.stage("velocity"); o.stage("diagnostics")
            # and the GPU's, so its warm-start history exists next step
            w.stage("predict"); w.stage("colour")
            for it in range(p["iterations"]):
                w.stage_primal(p["alpha"]); w.stage("dual", p["alpha"])
            w.stage("velocity")
    finally:
        o.close(); w.close()


def test_collide_stage_random_pile_bit_exact(avbd):
    rng = np.random.default_rng(5)
    o, w = make_pair(avbd, bodies=random_pile(rng, 300, 2.5))
    try:
        w.stage("collide")
        o.stage("broadphase"); o.stage("init")
        want, got = o.manifolds(), gpu_manifolds(w)
        assert len(want) > 200
        assert_manifolds_equal(got, want, exact_rows=True, ctx="pile")
        kinds = {0: 0, 1: 0, 2: 0}
        for m in want.values():
            for f in m["feat"]:
                kinds[(int(f) >> 24) & 3] += 1
        assert min(kinds.values()) > 10, kinds     # face-of-A, face-of-B and edge contacts all present
    finally:
        o.close(); w.close()


# --------------------------------------------------------------------------- colouring + primal
@pytest.mark.parametrize("scene", ["Pyramid", "Wall", "Stress1000"])
def test_colouring_valid(avbd, scene):
    o, w = make_pair(avbd, scene=scene)
    try:
        o.step(40 if scene != "Stress1000" else 150)
        w.set_state(o.state()); w.set_prev_linvel(o.prev_linvel())
        w.stage("collide"); w.stage("predict"); w.stage("colour")
        col, k = w.colours()
        props = w.body_props()
        ms = gpu_manifolds(w)
        assert len(ms) > 0 and k >= 1
        for (a, b) in ms:
            if props[a, 4] > 0 and props[b, 4] > 0:
                assert col[a] != col[b], (a, b, col[a])
        assert (col[props[:, 4] > 0] >= 0).all() and (col[props[:, 4] <= 0] == -2).all()
        assert k <= 12, k
    finally:
        o.close(); w.close()


def test_incremental_recolouring_stays_valid(avbd):
    """With AVBD_KEEP_COLOUR=1 (opt-in, avbd_kernels_graph.cuh: kept_word) a body keeps last step's colour unless a higher-priority
    neighbour of the new graph holds the same one, and only the bodies that lose theirs are coloured again: it must stay a valid colouring
    of every step's graph and must not grow colours without bound while Stress1000 collapses (manifolds appear and disappear every
    step).  Stress1000 runs the one-block colouring; the second world (a 14^3 grid) the cooperative one with its kept-colour prologue."""
    from avbd_demo3d_b200 import scenes
    os.environ["AVBD_KEEP_COLOUR"] = "1"
    try:
        w = avbd.World()
        w2 = avbd.World()
    finally:
        os.environ.pop("AVBD_KEEP_COLOUR", None)
    try:
        g = scenes.stress_grid(14, 14, 14, spacing_y=1.01, start_y=0.51)
        g["params"]["iterations"] = 6
        scenes.load(w2, g)
        props2 = None
        for chunk in range(4):
            w2.step(8)
            col, k = w2.colours()
            if props2 is None:
                props2 = w2.body_props()
            dyn = props2[:, 4] > 0
            for (a, b) in gpu_manifolds(w2):
                if dyn[a] and dyn[b]:
                    assert col[a] != col[b], (chunk, a, b, col[a])
            assert (col[dyn] >= 0).all() and (col[~dyn] == -2).all() and k <= 20, (chunk, k)
    finally:
        w2.close()
    try:
        scenes.load(w, scenes.scene("Stress1000"))
        props = None
        for chunk in range(12):
            w.step(20)
            col, k = w.colours()
            if props is None:
                props = w.body_props()
            ms = gpu_manifolds(w)
            dyn = props[:, 4] > 0
            for (a, b) in ms:
                if dyn[a] and dyn[b]:
                    assert col[a] != col[b], (chunk, a, b, col[a])
            assert (col[dyn] >= 0).all() and (col[~dyn] == -2).all()
            assert k <= 16 and col.max() == k - 1, (chunk, k)
    finally:
        w.close()


@pytest.mark.parametrize("scene,warm", [("Stack", 30), ("Pyramid", 30), ("TwoBlockDrop", 40), ("Stress1000", 140)])
def test_primal_dx_matches_oracle(avbd, scene, warm):
    """Per-body 6x6 solve (solver.cpp:344-409) given identical inputs and the same visiting order."""
    o, w = make_pair(avbd, scene=scene)
    try:
        o.step(warm)
        w.set_state(o.state()); w.set_prev_linvel(o.prev_linvel())
        # both start this step with the oracle's bodies; manifolds are new on both sides => identical rows
        o2 = Oracle("port").create()
        p = o.params()
        o2.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"])
        copy_bodies(o, o2)
        o2._set_prev_linvel(o2.h, o.prev_linvel())
        w.stage("collide"); w.stage("predict"); w.stage("colour")
        o2.stage("broadphase"); o2.stage("init"); o2.stage("predict")
        assert_manifolds_equal(gpu_manifolds(w), o2.manifolds(), exact_rows=True, ctx=scene)
        order, col, k = colour_order(w)
        want = o2.stage_primal(p["alpha"], order, want_dx=True)
        got = w.stage_primal(p["alpha"], want_dx=True)
        dyn = order
        scale = np.abs(want[dyn]).max(axis=1, keepdims=True)
        err = np.abs(got[dyn] - want[dyn])
        assert (err <= DX_RTOL * scale + DX_ATOL).all(), (scene, float(err.max()), float(scale.max()))
        assert np.abs(want[dyn]).max() > 1e-5          # not a vacuous comparison
        # and the poses after the sweep
        a, b = o2.state(), w.state()
        assert np.abs(a[:, :7] - b[:, :7]).max() <= STEP_POS_TOL
        o2.close()
    finally:
        o.close(); w.close()


@pytest.mark.parametrize("scene,steps,min_exact", [("Stack", 25, 25), ("Pyramid", 8, 5), ("TwoBlockDrop", 40, 30)])
def test_full_step_tracks_oracle_with_same_colour_order(avbd, scene, steps, min_exact):
    """Whole steps, oracle driven with the GPU's colour order: only summation-order rounding may differ."""
    o, w = make_pair(avbd, scene=scene)
    try:
        p = o.params()
        worst = 0.0
        for s in range(steps):
            w.stage("collide"); w.stage("predict"); w.stage("colour")
            order, col, k = colour_order(w)
            for it in range(p["iterations"]):
                w.stage_primal(p["alpha"]); w.stage("dual", p["alpha"])
            w.stage("velocity")
            o.step_ordered(order)
            a, b = o.state(), w.state()
            d_o, d_w = o.diagnostics(), w.diagnostics()
            if (d_o["manifolds"], d_o["contacts"]) != (d_w["manifolds"], d_w["contacts"]):
                # a clipped vertex crossed the 0.02 keep threshold on one side only: from here on the two runs
                # solve different contact sets, so tracking stops (the reference shows the same sensitivity
                # between its own -O2 and FMA builds)
                assert s >= min_exact, (s, d_o, d_w)
                break
            worst = max(worst, float(np.abs(a[:, :7] - b[:, :7]).max()))
        assert worst <= 5e-4, worst
    finally:
        o.close(); w.close()
