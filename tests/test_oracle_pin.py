"""Pins the CPU restatement (oracle/avbd_oracle.cpp, kind 'port') to the UNMODIFIED reference: against the committed
fixtures in tests/golden/golden.npz (generated from oracle/_ref by tests/golden/make_golden.py) and, when
oracle/_ref is built in this checkout, against the live reference too.  Everything here is bit-exact.  CPU only."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from _libs import ORACLE_DIR, Oracle, add_all, build_oracle, random_pile, ref_available

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))
FLT_MAX = 3.4028234663852886e38


def _jointed(o):
    o.add_body((20, 1, 20), 0.0, 0.5, (0, -0.5, 0))
    o.add_body((1, 1, 1), 1.0, 0.5, (0, 3, 0))
    o.add_body((1, 1, 1), 1.0, 0.5, (3, 3, 0))
    o.add_body((1, 1, 1), 1.0, 0.5, (3, 5, 0))
    o.add_body((1, 1, 1), 1.0, 0.5, (-3, 0.5, 0))
    o.add_body((1, 1, 1), 1.0, 0.5, (-3.2, 0.6, 0.1))
    o.add_joint(-1, 1, (0, 3.5, 0))
    o.add_spring(2, 3, (0, 0.5, 0), (0, -0.5, 0), 1000.0, 1.0)
    o.add_joint(-1, 2, (3, 3.5, 0))
    o.add_ignore(4, 5)


def _pile40(o):
    rows = GOLD["pile40/bodies"]
    for r in rows:
        o.add_body(r[0:3], float(r[3]), float(r[4]), r[5:8], r[8:12], r[12:15], r[15:18])


CASES = {"TwoBlockDrop": (lambda o: o.load_scene("TwoBlockDrop"), 300, 25), "Stack": (lambda o: o.load_scene("Stack"), 300, 50),
         "Pyramid": (lambda o: o.load_scene("Pyramid"), 200, 50), "Wall": (lambda o: o.load_scene("Wall"), 120, 40),
         "Pile40": (_pile40, 60, 20), "Jointed": (_jointed, 120, 30)}


@pytest.mark.parametrize("scene,steps", [("TwoBlockDrop", 300), ("Pyramid", 600), ("Stack", 600), ("Wall", 300)])
def test_headless_stdout_md5_matches_reference(scene, steps):
    """Same flags, same stdout, byte for byte (main.cpp:223-247) — BASELINE.json configs 1-2 and two more scenes."""
    build_oracle()
    txt = subprocess.run([os.path.join(ORACLE_DIR, "avbd_oracle_cli"), "--nogfx", "--scene", scene, "--steps", str(steps)],
                         capture_output=True, check=True).stdout
    assert hashlib.md5(txt).hexdigest() == bytes(GOLD[f"md5/{scene}/{steps}"]).decode()


def test_known_md5s_of_the_survey():
    """The fixture itself carries the md5s SURVEY.md section 8c recorded for the reference."""
    assert bytes(GOLD["md5/TwoBlockDrop/300"]).decode() == "3f2d6a60b7a426efc49bf6735d0e740a"
    assert bytes(GOLD["md5/Pyramid/600"]).decode() == "affc80fdb5ea46fffc57a2c8ca9d1b17"
    assert bytes(GOLD["md5/Stress1000/600"]).decode() == "a0ee24c1aa382a29711208fdb7c829a6"


@pytest.mark.slow
def test_headless_stress1000_md5_matches_reference():
    """BASELINE.json config 3 (about 35 s of CPU)."""
    txt = subprocess.run([os.path.join(ORACLE_DIR, "avbd_oracle_cli"), "--nogfx", "--scene", "Stress1000", "--steps", "600"],
                         capture_output=True, check=True).stdout
    assert hashlib.md5(txt).hexdigest() == bytes(GOLD["md5/Stress1000/600"]).decode()


def test_collide_matches_golden(port):
    a, b = GOLD["collide/a"], GOLD["collide/b"]
    kinds = set()
    for t in range(len(a)):
        k, f, g = port.collide(a[t], b[t])
        assert k == GOLD["collide/count"][t], t
        assert (f == GOLD["collide/feat"][t, :k]).all(), t
        assert g.tobytes() == np.ascontiguousarray(GOLD["collide/geom"][t, :k]).tobytes(), t
        kinds.update((int(x) >> 24) & 3 for x in f)
    assert kinds == {0, 1, 2}       # face-of-A, face-of-B and edge-edge contacts are all covered
    assert (GOLD["collide/count"] == 0).sum() > 50 and (GOLD["collide/count"] == 4).sum() > 50


def test_solve6x6_matches_golden(port):
    for t in range(len(GOLD["solve6/lhs"])):
        out = port.solve6x6(GOLD["solve6/lhs"][t], GOLD["solve6/rhs"][t])
        assert out.tobytes() == GOLD["solve6/out"][t].tobytes(), t
    assert (GOLD["solve6/out"][0] == 0).all()      # zero pivot => zero solution (maths.h:104)


@pytest.mark.parametrize("name", sorted(CASES))
def test_trajectory_matches_golden(name):
    setup, steps, every = CASES[name]
    o = Oracle("port").create()
    setup(o)
    k = 0
    for s in range(steps):
        o.step(1)
        if (s + 1) % every == 0 or s == steps - 1:
            assert o.state().tobytes() == GOLD[f"traj/{name}/state"][k].tobytes(), (name, s)
            d = o.diagnostics()
            row = [d["maxPen"], d["maxViol"], d["maxLin"], d["maxAng"], d["maxLambda"], d["contacts"], d["manifolds"], d["dynBodies"]]
            assert np.array(row, np.float64).tobytes() == GOLD[f"traj/{name}/diag"][k].tobytes(), (name, s)
            k += 1
    ms = o.manifolds()
    keys = sorted(ms)
    assert np.array(keys, np.int32).reshape(-1, 2).tobytes() == GOLD[f"traj/{name}/manifold_keys"].tobytes()
    for i, key in enumerate(keys):
        n = ms[key]["n"]
        assert n == GOLD[f"traj/{name}/manifold_n"][i]
        assert (ms[key]["feat"] == GOLD[f"traj/{name}/manifold_feat"][i, :n]).all()
        assert ms[key]["geom"].tobytes() == np.ascontiguousarray(GOLD[f"traj/{name}/manifold_geom"][i, :n]).tobytes()
        assert ms[key]["lam"].tobytes() == np.ascontiguousarray(GOLD[f"traj/{name}/manifold_lam"][i, :n]).tobytes()
        assert ms[key]["pen"].tobytes() == np.ascontiguousarray(GOLD[f"traj/{name}/manifold_pen"][i, :n]).tobytes()
    o.close()


def test_reference_end_states_of_the_survey():
    """SURVEY.md section 8c anchors: TwoBlockDrop rests at y=0.5100 with 2 manifolds / 8 contacts and zero velocity."""
    st, dg = GOLD["traj/TwoBlockDrop/state"][-1], GOLD["traj/TwoBlockDrop/diag"][-1]
    assert np.allclose(st[1:, 1], 0.51, atol=1e-4) and np.abs(st[:, 7:]).max() < 1e-4
    assert dg[6] == 2 and dg[5] == 8 and dg[0] == 0.0


# ------------------------------------------------------------------ live reference (only where oracle/_ref exists)
@pytest.mark.parametrize("name", ["Pile40", "Jointed", "Wall"])
def test_port_equals_live_reference(name, ref):
    setup, steps, every = CASES[name]
    a, b = Oracle("port").create(), Oracle("ref").create()
    b.set_params()
    setup(a); setup(b)
    for s in range(min(steps, 80)):
        a.step(1); b.step(1)
        assert a.state().tobytes() == b.state().tobytes(), (name, s)
    ma, mb = a.manifolds(), b.manifolds()
    assert list(ma) == list(mb)          # same manifolds in the same (newest-first) order
    for k in ma:
        assert ma[k]["geom"].tobytes() == mb[k]["geom"].tobytes() and ma[k]["lam"].tobytes() == mb[k]["lam"].tobytes()
    a.close(); b.close()


def test_port_equals_live_reference_random_piles(ref):
    rng = np.random.default_rng(1234)
    for trial in range(3):
        bodies = random_pile(rng, 25, 1.2)
        a, b = Oracle("port").create(), Oracle("ref").create()
        b.set_params()
        params = dict(iterations=int(rng.integers(3, 12)), alpha=float(rng.uniform(0.8, 1.0)), beta=float(rng.uniform(1e4, 2e5)),
                      gamma=float(rng.uniform(0.95, 1.0)), post=bool(trial == 2))
        a.set_params(**params); b.set_params(**params)
        add_all(a, bodies); add_all(b, bodies)
        for s in range(40):
            a.step(1); b.step(1)
            assert a.state().tobytes() == b.state().tobytes(), (trial, s)
        a.close(); b.close()


def test_staged_step_equals_whole_step(port):
    """The oracle's per-stage entry points compose to exactly Solver::step()."""
    a, b = Oracle("port").create(), Oracle("port").create()
    a.load_scene("Pyramid"); b.load_scene("Pyramid")
    p = a.params()
    for s in range(15):
        a.step(1)
        b.stage("broadphase"); b.stage("init"); b.stage("predict")
        for it in range(p["iterations"]):
            b.stage_primal(p["alpha"]); b.stage("dual", p["alpha"])
        b.stage("velocity"); b.stage("diagnostics")
        assert a.state().tobytes() == b.state().tobytes(), s
        assert a.diagnostics() == b.diagnostics()
    a.close(); b.close()
