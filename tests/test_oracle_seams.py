"""CPU checks of the oracle-side test seams the GPU per-stage parity tests rely on (tests/test_gpu_stage_parity.py):
orc_set_manifolds is the inverse of orc_get_manifolds, and the scenarios those tests use are not vacuous."""
import numpy as np
import pytest

from _libs import Oracle, copy_bodies, manifold_dict


def _clone(o):
    o2 = Oracle("port").create()
    p = o.params()
    o2.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"])
    copy_bodies(o, o2)
    o2._set_prev_linvel(o2.h, o.prev_linvel())
    o2.set_manifolds(*o.manifolds_raw())
    return o2


@pytest.mark.parametrize("scene,steps", [("Stack", 40), ("Pyramid", 30), ("Wall", 25), ("TwoBlockDrop", 45)])
def test_set_manifolds_is_the_inverse_of_get(scene, steps):
    o = Oracle("port").create()
    o.load_scene(scene)
    o.step(steps)
    o2 = _clone(o)
    for a, b in zip(o.manifolds_raw(), o2.manifolds_raw()):
        assert a.tobytes() == b.tobytes()
    # and the clone continues exactly like the original: same carry-over, same trajectory
    o.stage("broadphase"); o.stage("init")
    o2.stage("broadphase"); o2.stage("init")
    new = o.manifolds()
    old = manifold_dict(*o2.manifolds_raw())
    for a, b in zip(o.manifolds_raw(), o2.manifolds_raw()):
        assert a.tobytes() == b.tobytes()
    assert sum(int((m["lam"] != 0).any()) for m in new.values()) > 0          # something was carried over
    if scene != "TwoBlockDrop":
        assert sum(int(m["stick"].any()) for m in new.values()) > 0           # and a stick anchor was re-used
    o.close(); o2.close()


def test_clone_tracks_original_for_whole_steps():
    o = Oracle("port").create()
    o.load_scene("Pyramid")
    o.step(20)
    o2 = _clone(o)
    o.step(15); o2.step(15)
    assert o.state().tobytes() == o2.state().tobytes()
    o.close(); o2.close()


def test_dual_stage_changes_rows_at_the_tested_states():
    for scene, warm, sweeps in (("Stack", 30, 3), ("Pyramid", 40, 2)):
        o = Oracle("port").create()
        o.load_scene(scene)
        o.step(warm)
        p = o.params()
        o.stage("broadphase"); o.stage("init"); o.stage("predict")
        for it in range(sweeps - 1):
            o.stage_primal(p["alpha"]); o.stage("dual", p["alpha"])
        o.stage_primal(p["alpha"])
        pre = o.manifolds()
        o.stage("dual", p["alpha"])
        post = o.manifolds()
        assert sum(int((post[k]["pen"] != pre[k]["pen"]).any()) for k in pre) > 0
        assert sum(int((post[k]["lam"] != pre[k]["lam"]).any()) for k in pre) > 0
        o.close()
