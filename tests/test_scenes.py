"""Host-side scene presets (avbd-demo3d_b200/scenes.py) against the oracle's scenes.h restatement — and, when
oracle/_ref is built, against the unmodified reference — bit for bit.  CPU only."""
import numpy as np
import pytest

from _libs import Oracle, ref_available


def _bodies(o):
    props, st = o.body_props(), o.state()
    return props, st


@pytest.mark.parametrize("kind", ["port", "ref"])
@pytest.mark.parametrize("name", ["Empty", "Ground", "Stack", "Pyramid", "Wall", "TwoBlockDrop", "Stress1000", "Rod (WIP)", "Soft Body (WIP)"])
def test_preset_matches_oracle(name, kind):
    if kind == "ref" and not ref_available():
        pytest.skip("oracle/_ref not built")
    from avbd_demo3d_b200 import scenes
    o = Oracle(kind).create()
    o.set_params()            # the ref Solver persists params across scenes; start from defaults
    o.load_scene(name)
    props, st = _bodies(o)
    s = scenes.scene(name)
    assert len(s["size"]) == o.n
    if o.n:
        assert s["size"].tobytes() == np.ascontiguousarray(props[:, 0:3]).tobytes()
        assert s["friction"].tobytes() == np.ascontiguousarray(props[:, 8]).tobytes()
        mass = (s["size"][:, 0] * s["size"][:, 1] * s["size"][:, 2] * s["density"]).astype(np.float32)
        assert mass.tobytes() == np.ascontiguousarray(props[:, 3]).tobytes()
        assert s["pos"].tobytes() == np.ascontiguousarray(st[:, 0:3]).tobytes(), np.abs(s["pos"] - st[:, :3]).max()
        assert s["quat"].tobytes() == np.ascontiguousarray(st[:, 3:7]).tobytes()
        assert s["lin"].tobytes() == np.ascontiguousarray(st[:, 7:10]).tobytes()
        assert s["ang"].tobytes() == np.ascontiguousarray(st[:, 10:13]).tobytes()
    p = o.params()
    want = dict(iterations=10, beta=1e5, gamma=np.float32(0.99))
    want.update(s["params"])
    assert p["iterations"] == want["iterations"] and p["beta"] == want["beta"] and np.float32(p["gamma"]) == np.float32(want["gamma"])
    o.close()


def test_stress_grid_generalisation(port):
    from avbd_demo3d_b200 import scenes
    o = Oracle("port").create()
    o.load_stress_grid(7, 5, 9, 1.01, 0.51, True)
    s = scenes.stress_grid(7, 5, 9, 1.01, 0.51, True)
    assert s["pos"].tobytes() == np.ascontiguousarray(o.state()[:, 0:3]).tobytes()
    assert s["size"].tobytes() == np.ascontiguousarray(o.body_props()[:, 0:3]).tobytes()
    o.close()


def test_ensemble_is_partition_invariant():
    from avbd_demo3d_b200 import scenes
    base = scenes.scene("Stack")
    whole = scenes.ensemble(base, 8)
    lo, hi = scenes.ensemble(base, 4, first_world=0), scenes.ensemble(base, 4, first_world=4)
    n = len(base["size"])
    assert whole["pos"][: 4 * n].tobytes() == lo["pos"].tobytes()
    assert whole["pos"][4 * n:].tobytes() == hi["pos"].tobytes()
    assert (whole["world_ids"] == np.repeat(np.arange(8), n)).all()
    assert len({whole["pos"][w * n + 1].tobytes() for w in range(8)}) == 8      # worlds differ
    assert whole["pos"][0].tobytes() == base["pos"][0].tobytes()                 # static ground is not moved
