"""Edge cases of the step path (empty / static-only / contact-free worlds, growth between steps, clear + reuse, random solver
parameters incl. postStabilize) and size-independent properties at the benchmark's scale.  Needs a GPU."""
import numpy as np
import pytest

from _libs import Oracle, random_pile, add_all

pytestmark = pytest.mark.gpu


def test_empty_world_steps(avbd):
    """Solver::step() on an empty solver is a no-op (solver.cpp:255-514 with empty lists)."""
    w = avbd.World()
    w.step(3)
    d = w.diagnostics()
    assert (d["manifolds"], d["contacts"], d["dynBodies"], d["nanEvents"]) == (0, 0, 0, 0)
    assert w.state().shape == (0, 13)
    assert w.pick((0, 5, 0), (0, -1, 0))[0] == -1
    w.close()


def test_static_only_world(avbd):
    """Two overlapping static bodies: the reference creates the manifold (no mass test in solver.cpp:262-270) and nothing moves."""
    o = Oracle("port").create()
    w = avbd.World()
    for d in (o, w):
        d.add_body((20, 1, 20), 0.0, 0.5, (0, -0.5, 0))
        d.add_body((4, 1, 4), 0.0, 0.5, (0, 0.3, 0))
    o.step(5); w.step(5)
    assert w.state().tobytes() == o.state().tobytes()
    dw, do = w.diagnostics(), o.diagnostics()
    assert (dw["manifolds"], dw["contacts"], dw["dynBodies"]) == (do["manifolds"], do["contacts"], 0)
    assert do["manifolds"] == 1
    o.close(); w.close()


def test_free_fall_matches_oracle(avbd):
    """No contacts: predict, the inertia-only block solve and the velocity update, 40 steps, against the oracle."""
    o = Oracle("port").create()
    w = avbd.World()
    for d in (o, w):
        d.add_body((1, 2, 0.5), 1.0, 0.5, (0, 50, 0), quat=(0.1, 0.2, 0.3, 0.9273618495495703), lin=(1, 0, -2), ang=(0.5, -1.0, 2.0))
        d.add_body((0.5, 0.5, 0.5), 3.0, 0.5, (5, 60, 1))
    o.step(40); w.step(40)
    a, b = o.state(), w.state()
    assert np.abs(a - b).max() <= 2e-5 * max(1.0, float(np.abs(a).max())), float(np.abs(a - b).max())
    assert w.diagnostics()["manifolds"] == 0
    o.close(); w.close()


def test_bodies_appended_between_steps(avbd):
    """The GUI's right-click path (main.cpp:139-142): bodies are added while the simulation runs; warm-started manifolds of the
    existing bodies survive the growth (same manifold / contact counts as the oracle doing the same)."""
    o = Oracle("port").create()
    w = avbd.World()
    for d in (o, w):
        d.add_body((20, 1, 20), 0.0, 0.5, (0, -0.5, 0))
        d.add_body((1, 1, 1), 1.0, 0.5, (0, 0.51, 0))
        d.add_body((1, 1, 1), 1.0, 0.5, (3, 0.51, 0))
    o.step(20); w.step(20)
    for d in (o, w):
        d.add_body((1, 1, 1), 1.0, 0.5, (0, 1.53, 0))          # lands on body 1
        d.add_body((1, 1, 1), 1.0, 0.5, (-4, 5.0, 0))          # free fall
    o.step(60); w.step(60)
    a, b = o.state(), w.state()
    assert a.shape == b.shape == (5, 13)
    assert np.abs(a[:, 1] - b[:, 1]).max() < 2e-3, np.abs(a[:, 1] - b[:, 1])
    dw, do = w.diagnostics(), o.diagnostics()
    assert (dw["manifolds"], dw["contacts"], dw["dynBodies"]) == (do["manifolds"], do["contacts"], do["dynBodies"])
    o.close(); w.close()


def test_clear_and_reuse(avbd):
    """Solver::clear (solver.cpp:230-238) drops bodies and manifolds and keeps the parameters; the world is reusable and a
    rebuilt scene steps exactly like a fresh one."""
    from avbd_demo3d_b200 import scenes
    w = avbd.World()
    scenes.load(w, scenes.scene("Pyramid"))
    w.step(30)
    w.clear()
    assert w.n == 0 and w.diagnostics()["manifolds"] == 0
    w.step(2)
    scenes.load(w, scenes.scene("Stack"))
    w.step(25)
    fresh = avbd.World()
    scenes.load(fresh, scenes.scene("Stack"))
    fresh.step(25)
    assert w.state().tobytes() == fresh.state().tobytes()
    assert w.diagnostics() == fresh.diagnostics()
    w.close(); fresh.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_solver_params_on_a_random_pile(avbd, seed):
    """Random dt / gravity / iterations / alpha / beta / gamma, with and without postStabilize, on a dense pile of tilted boxes
    with anisotropic inertia (the gyroscopic row term, solver.cpp:393-397, is live) and random initial velocities:
    * the narrowphase + warm-start stage is bit-identical to the oracle's,
    * the first primal sweep's per-body dx is within the stated 6x6 tolerance given identical inputs and order,
    * ONE whole step through the stage API (alpha schedule of solver.cpp:340-342, dual after every sweep but postStabilize's
      last) stays within 5e-3 of the oracle driven in the same colour order.  One step only, loose on purpose: bodies thrown
      together interpenetrating move 0.1 m per sweep and stick / slip decisions flip on one ulp (tools/dual_probe.py), so
      rounding-level differences grow by orders of magnitude within a few sweeps."""
    from test_gpu_parity import DX_ATOL, DX_RTOL, assert_manifolds_equal, gpu_manifolds
    rng = np.random.default_rng(seed)
    prm = dict(dt=float(rng.choice([1 / 60, 1 / 120, 1 / 30])), g=(0.0, -float(rng.uniform(5, 15)), 0.0), iterations=int(rng.integers(3, 12)),
               alpha=float(rng.uniform(0.8, 1.0)), beta=float(rng.choice([1e4, 3e4, 1e5])), gamma=float(rng.uniform(0.95, 1.0)), post=bool(seed % 2))
    bodies = random_pile(rng, 30)
    o = Oracle("port").create()
    w = avbd.World()
    o.set_params(**prm); w.set_params(**prm)
    add_all(o, bodies); add_all(w, bodies)
    try:
        w.stage("collide"); w.stage("predict"); w.stage("colour")
        o.stage("broadphase"); o.stage("init"); o.stage("predict")
        assert_manifolds_equal(gpu_manifolds(w), o.manifolds(), exact_rows=True, ctx=f"pile {seed}")
        col, k = w.colours()
        dyn = np.arange(len(col))[col >= 0]
        order = dyn[np.lexsort((dyn, col[dyn]))].astype(np.int32)
        total = prm["iterations"] + (1 if prm["post"] else 0)
        for it in range(total):
            a = (1.0 if it < prm["iterations"] else 0.0) if prm["post"] else prm["alpha"]      # solver.cpp:340-342
            want = o.stage_primal(a, order, want_dx=True)
            got = w.stage_primal(a, want_dx=True)
            if it == 0:
                scale = np.abs(want[dyn]).max(axis=1, keepdims=True)
                assert (np.abs(got[dyn] - want[dyn]) <= DX_RTOL * scale + DX_ATOL).all(), float(np.abs(got[dyn] - want[dyn]).max())
                assert np.abs(want[dyn]).max() > 1e-3
            if it < prm["iterations"]:
                o.stage("dual", a); w.stage("dual", a)
        o.stage("velocity"); w.stage("velocity")
        drift = float(np.abs(o.state()[:, :7] - w.state()[:, :7]).max())
        assert drift <= 5e-3, (drift, prm)
        assert w.diagnostics()["nanEvents"] == 0
    finally:
        o.close(); w.close()


# --------------------------------------------------------------------------- properties at benchmark scale
def _grid(avbd, n, steps):
    from avbd_demo3d_b200 import scenes
    s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True)
    s["params"]["iterations"] = 10
    w = avbd.World()
    scenes.load(w, s)
    w.step(steps)
    return w


@pytest.mark.parametrize("n", [40, 100])
def test_grid_properties_at_scale(avbd, n):
    """BASELINE.json config 3 (n = 100: the 1M-box grid of bench.py) and a 64k-box cut of it, through properties that do not
    need the O(n^2) reference: pair keys strictly ascending (sortedness, no duplicate manifold), A > B, every dynamic pair
    coloured differently and every dynamic body coloured, dense contact count = sum of the manifolds' counts = diagnostics,
    <= 4 contacts per manifold, finite state, no NaN scrub, and the pile does not sink into the ground."""
    w = _grid(avbd, n, 6)
    try:
        ints, feats, stick, flts = w.manifolds_raw()
        d = w.diagnostics()
        a, b, cnt = ints[:, 0].astype(np.int64), ints[:, 1].astype(np.int64), ints[:, 2]
        assert len(a) > 0.8 * n ** 3
        assert (a > b).all()
        key = a * (n ** 3 + 2) + b
        assert (np.diff(key) > 0).all()                                   # sorted, unique
        assert cnt.min() >= 0 and cnt.max() <= 4
        live = cnt > 0
        assert int(cnt.sum()) == d["contacts"] and int(live.sum()) == d["manifolds"]
        col, k = w.colours()
        props = w.body_props()
        dyn = props[:, 4] > 0
        assert dyn.sum() == n ** 3 == d["dynBodies"]
        assert (col[dyn] >= 0).all() and (col[~dyn] == -2).all() and k <= 16
        both = dyn[a] & dyn[b]
        assert (col[a[both]] != col[b[both]]).all()
        st = w.state()
        assert np.isfinite(st).all() and d["nanEvents"] == 0
        assert st[1:, 1].min() > 0.0                                       # no box centre below the ground's top face after 6 steps
        assert d["maxPen"] < 2.0
    finally:
        w.close()


def test_grid_pair_set_against_kdtree(avbd):
    """Broadphase at scale against an independent spatial index: the sphere-overlap pair set of a 64k-box grid equals
    scipy's cKDTree pairs within r_a + r_b, except pairs whose distance is within 1e-5 of the threshold (float32 `<=` vs float64)."""
    from scipy.spatial import cKDTree
    from avbd_demo3d_b200 import scenes
    n = 40
    s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True)
    w = avbd.World()
    scenes.load(w, s)
    try:
        pairs = w.stage_broadphase()
        props, st = w.body_props(), w.state()
        radius = props[:, 9].astype(np.float64)
        small = np.arange(1, len(st))                                      # body 0 is the ground
        r = float(radius[1])
        assert np.allclose(radius[1:], r)
        pos = st[:, :3].astype(np.float64)
        tree = cKDTree(pos[small])
        near = tree.query_pairs(2 * r + 1e-4, output_type="ndarray") + 1   # back to body ids
        dist = np.linalg.norm(pos[near[:, 0]] - pos[near[:, 1]], axis=1)
        sure = near[dist <= 2 * r - 1e-5]
        maybe = near[dist <= 2 * r + 1e-5]
        enc = lambda p: np.maximum(p[:, 0], p[:, 1]).astype(np.int64) * (len(st) + 1) + np.minimum(p[:, 0], p[:, 1])
        got = np.asarray(pairs, dtype=np.int64)
        got_small = got[(got[:, 0] != 0) & (got[:, 1] != 0)]
        g, lo, hi = set(enc(got_small).tolist()), set(enc(sure).tolist()), set(enc(maybe).tolist())
        assert lo <= g <= hi, (len(g), len(lo), len(hi))
        assert len(g) > 3 * n ** 3
        # the ground (a large body) pairs with exactly the boxes whose bounding spheres reach it
        with_ground = set(int(max(p)) for p in got if min(p) == 0)
        d0 = np.linalg.norm(pos[small] - pos[0], axis=1)
        rg = float(radius[0])
        sure0 = set((small[d0 <= r + rg - 1e-4]).tolist())
        maybe0 = set((small[d0 <= r + rg + 1e-4]).tolist())
        assert sure0 <= with_ground <= maybe0
    finally:
        w.close()


def test_live_parameter_edits(avbd):
    """The GUI edits dt / gravity / iterations / alpha / beta / gamma between steps (main.cpp:88-94, fields re-read by every
    Solver::step): the same edits applied to the oracle give the same free-fall trajectory (no contacts: order-independent) and
    the same resting state of TwoBlockDrop."""
    o = Oracle("port").create()
    w = avbd.World()
    for d in (o, w):
        d.add_body((1, 1, 1), 1.0, 0.5, (0, 100, 0), lin=(1, 0, 0), ang=(0, 0.5, 0))
    edits = [dict(), dict(g=(0, -3, 0)), dict(dt=1 / 120, g=(1, -3, 0)), dict(dt=1 / 30, iterations=3, alpha=0.9, beta=5e4, gamma=0.97), dict(post=True)]
    for e in edits:
        o.set_params(**e); w.set_params(**e)
        o.step(15); w.step(15)
        a, b = o.state(), w.state()
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, float(np.abs(a).max())), (e, float(np.abs(a - b).max()))
    o.close(); w.close()

    from avbd_demo3d_b200 import scenes
    o = Oracle("port").create(); o.load_scene("TwoBlockDrop")
    w = avbd.World(); scenes.load(w, scenes.scene("TwoBlockDrop"))
    p = o.params()
    for iters in (10, 4, 16):
        kw = dict(dt=p["dt"], g=p["g"], iterations=iters, alpha=p["alpha"], beta=p["beta"], gamma=p["gamma"])
        o.set_params(**kw); w.set_params(**kw)
        o.step(100); w.step(100)
    a, b = o.state(), w.state()
    assert np.abs(a[:, 1] - b[:, 1]).max() < 1e-3 and np.abs(b[1:, 7:]).max() < 1e-3
    assert (w.diagnostics()["manifolds"], w.diagnostics()["contacts"]) == (o.diagnostics()["manifolds"], o.diagnostics()["contacts"]) == (2, 8)
    o.close(); w.close()


def test_world_wider_than_1024_cells_pair_set_exact(avbd):
    """The broadphase identifies a cell inside a bucket by 10 bits per axis (`pack_cell`): two cells exactly 1024 cells apart share
    that identity, and share a bucket whenever their block hashes collide.  Bodies placed in cells that alias this way (clusters
    1024 and 2048 cells apart along each axis, unit cubes: cell edge 1.749) must still give exactly the reference's pair set — the
    sphere test rejects what the mix-up lets through."""
    rng = np.random.default_rng(17)
    cell = 2.02 * 0.5 * np.sqrt(3.0)               # 2.02 x the bounding radius of a unit cube (prepare())
    bodies = []
    for off in ((0, 0, 0), (1024, 0, 0), (2048, 0, 0), (0, 1024, 0), (0, 0, 1024), (1024, 1024, 1024)):
        base = np.array(off, np.float64) * cell
        for _ in range(40):
            p = base + rng.uniform(0.0, 3.0 * cell, 3)
            bodies.append(dict(size=(1, 1, 1), density=1.0, friction=0.5, pos=tuple(np.float32(p)), quat=(0, 0, 0, 1), lin=(0, 0, 0), ang=(0, 0, 0)))
    o = Oracle("port").create()
    w = avbd.World()
    add_all(o, bodies); add_all(w, bodies)
    try:
        got = set((int(a), int(b)) for a, b in w.stage_broadphase())
        want = set((int(a), int(b)) for a, b in o.overlap_pairs())
        assert got == want, (len(got), len(want), sorted(got ^ want)[:10])
        assert len(want) > 100
        far = [(a, b) for a, b in want if abs(bodies[a]["pos"][0] - bodies[b]["pos"][0]) > 100]
        assert not far
        w.step(2)                                   # and the rest of the pipeline runs at these coordinates
        assert w.diagnostics()["nanEvents"] == 0
    finally:
        o.close(); w.close()


@pytest.mark.parametrize("switch", ["AVBD_NO_SMALL_GRAPH", "AVBD_COLOUR_BLOCK_MAX"])
def test_small_world_graph_stage_forms_agree_bit_for_bit(avbd, switch, monkeypatch):
    """A small world's graph stage runs in one block (graph_small: block radix sorts, block prefix sums, colouring on shared-memory work
    words); a large world's is ~16 launches with device-wide sorts and the cooperative colouring.  Same per-element routines, stable sorts
    on both sides: adjacency order, colours, colour order and visit lists — hence whole trajectories — must be identical.  Second case:
    the multi-launch path with the cooperative grid colouring instead of the one-block colouring (AVBD_COLOUR_BLOCK_MAX=0)."""
    from avbd_demo3d_b200 import scenes
    preset = scenes.stress_grid(8, 8, 8, spacing_y=1.01, start_y=0.51)          # 512 boxes, layers start interpenetrating: a busy graph
    preset["params"]["iterations"] = 6
    def links(w):           # user forces enter the colouring's adjacency and the free / linked body lists
        w.add_joint(-1, 5, (float(preset["pos"][5][0]), float(preset["pos"][5][1]) + 0.5, float(preset["pos"][5][2])))
        w.add_spring(40, 41, (0.5, 0, 0), (-0.5, 0, 0), 800.0, 1.2)
        w.add_spring(300, 17, (0, 0.5, 0), (0, -0.5, 0), 50.0)
    ref = avbd.World(); scenes.load(ref, preset); links(ref)
    ref.step(25)
    want_state, want_colours = ref.state().copy(), ref.colours()
    dref = ref.diagnostics()
    ref.close()
    monkeypatch.setenv("AVBD_NO_SMALL_GRAPH", "1")
    if switch == "AVBD_COLOUR_BLOCK_MAX":
        monkeypatch.setenv("AVBD_COLOUR_BLOCK_MAX", "0")
    w = avbd.World(); scenes.load(w, preset); links(w)
    w.step(25)
    assert w.state().tobytes() == want_state.tobytes()
    assert np.array_equal(w.colours()[0], want_colours[0]) and w.colours()[1] == want_colours[1]
    d = w.diagnostics()
    assert (d["manifolds"], d["contacts"]) == (dref["manifolds"], dref["contacts"]) and d["manifolds"] > 200
    w.close()


@pytest.mark.parametrize("mode", ["body", "cell"])
def test_broadphase_forms_agree_bit_for_bit(avbd, mode, monkeypatch):
    """The three forms of the candidate sweep — per-cell sweep + separate SAT cull (default), the per-body sweep of round 1
    (AVBD_BROADPHASE=body) and the per-cell sweep with the cull fused in (=cell) — enumerate the same pairs and run the same SAT, in a
    different order; the survivors are sorted before anything is built from them, so whole trajectories must be identical."""
    from avbd_demo3d_b200 import scenes
    preset = scenes.stress_grid(10, 10, 10, spacing_y=1.01, start_y=0.51)
    preset["params"]["iterations"] = 6
    ref = avbd.World(); scenes.load(ref, preset)
    ref.step(20)
    want, dref = ref.state().copy(), ref.diagnostics()
    ref.close()
    monkeypatch.setenv("AVBD_BROADPHASE", mode)
    w = avbd.World(); scenes.load(w, preset)
    w.step(20)
    d = w.diagnostics()
    assert w.state().tobytes() == want.tobytes()
    assert (d["manifolds"], d["contacts"]) == (dref["manifolds"], dref["contacts"]) and d["manifolds"] > 1000
    w.close()
