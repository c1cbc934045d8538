"""Generates tests/golden/golden.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixtures pin the CPU restatement (oracle/avbd_oracle.cpp) where /root/reference does not exist (GPU box)."""
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from _libs import Oracle, add_all, random_pile  # noqa: E402

FLT_MAX = 3.4028234663852886e38


def pack_manifolds(ms):
    keys = sorted(ms)
    return dict(keys=np.array(keys, np.int32).reshape(-1, 2), n=np.array([ms[k]["n"] for k in keys], np.int32),
                feat=np.array([np.pad(ms[k]["feat"], (0, 4 - ms[k]["n"])) for k in keys], np.int32).reshape(-1, 4),
                geom=np.array([np.pad(ms[k]["geom"], ((0, 4 - ms[k]["n"]), (0, 0))) for k in keys], np.float32).reshape(-1, 4, 14),
                lam=np.array([np.pad(ms[k]["lam"], ((0, 4 - ms[k]["n"]), (0, 0))) for k in keys], np.float32).reshape(-1, 4, 3),
                pen=np.array([np.pad(ms[k]["pen"], ((0, 4 - ms[k]["n"]), (0, 0))) for k in keys], np.float32).reshape(-1, 4, 3))


def trajectory(setup, steps, every):
    o = Oracle("ref").create()
    o.set_params()
    setup(o)
    states, diags = [], []
    for s in range(steps):
        o.step(1)
        if (s + 1) % every == 0 or s == steps - 1:
            states.append(o.state())
            d = o.diagnostics()
            diags.append([d["maxPen"], d["maxViol"], d["maxLin"], d["maxAng"], d["maxLambda"], d["contacts"], d["manifolds"], d["dynBodies"]])
    m = pack_manifolds(o.manifolds())
    o.close()
    return np.array(states, np.float32), np.array(diags, np.float64), m


def main():
    out = {}
    # 1. headless stdout md5s of the three BASELINE.json anchor configs
    exe = os.path.join(ROOT, "oracle", "_ref", "avbd_demo3d_ref")
    for scene, steps in (("TwoBlockDrop", 300), ("Pyramid", 600), ("Stress1000", 600), ("Stack", 600), ("Wall", 300)):
        txt = subprocess.run([exe, "--nogfx", "--scene", scene, "--steps", str(steps)], capture_output=True, check=True).stdout
        out[f"md5/{scene}/{steps}"] = np.frombuffer(hashlib.md5(txt).hexdigest().encode(), np.uint8)
    # 2. Manifold::collide on random pairs
    rng = np.random.default_rng(2024)
    ref = Oracle("ref")
    n = 600
    a, b = np.zeros((n, 10), np.float32), np.zeros((n, 10), np.float32)
    cnt, feats, geom = np.zeros(n, np.int32), np.zeros((n, 4), np.int32), np.zeros((n, 4, 10), np.float32)
    for t in range(n):
        for arr, c in ((a, rng.uniform(-0.2, 0.2, 3)), (b, rng.uniform(-1.2, 1.2, 3))):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            if t % 5 == 0: q = np.array([0, 0, 0, 1.0])
            if t % 7 == 0: q = np.array([0, np.sin(0.3), 0, np.cos(0.3)])
            arr[t] = np.concatenate([rng.uniform(0.3, 1.5, 3), c, q])
        k, f, g = ref.collide(a[t], b[t])
        cnt[t] = k; feats[t, :k] = f; geom[t, :k] = g
    out.update({"collide/a": a, "collide/b": b, "collide/count": cnt, "collide/feat": feats, "collide/geom": geom})
    # 3. solve6x6
    m = 200
    lhs, rhs, sol = np.zeros((m, 36), np.float32), (rng.normal(size=(m, 6)) * 100).astype(np.float32), np.zeros((m, 6), np.float32)
    for t in range(m):
        J = rng.normal(size=(8, 6))
        A = J.T @ np.diag(rng.uniform(1e3, 1e6, 8)) @ J + np.diag(rng.uniform(1e2, 1e4, 6))
        if t % 40 == 0: A[:] = 0
        blocks = [A[:3, :3], A[:3, 3:], A[3:, :3], A[3:, 3:]]
        lhs[t] = np.concatenate([blk.T.reshape(-1) for blk in blocks])
        lhs[t, 18:27] = lhs[t, 9:18].reshape(3, 3).T.reshape(-1)
        sol[t] = ref.solve6x6(lhs[t], rhs[t])
    out.update({"solve6/lhs": lhs, "solve6/rhs": rhs, "solve6/out": sol})
    # 4. whole trajectories (state after selected steps, diagnostics, final manifolds)
    pile = random_pile(np.random.default_rng(99), 40, 1.5)
    cases = {
        "TwoBlockDrop": (lambda o: o.load_scene("TwoBlockDrop"), 300, 25),
        "Stack": (lambda o: o.load_scene("Stack"), 300, 50),
        "Pyramid": (lambda o: o.load_scene("Pyramid"), 200, 50),
        "Wall": (lambda o: o.load_scene("Wall"), 120, 40),
        "Pile40": (lambda o: add_all(o, pile), 60, 20),
    }

    def jointed(o):      # body-world weld holds a box; a soft spring hangs a second one; an ignored pair overlaps freely
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
    cases["Jointed"] = (jointed, 120, 30)
    for name, (setup, steps, every) in cases.items():
        st, dg, mf = trajectory(setup, steps, every)
        out[f"traj/{name}/state"] = st
        out[f"traj/{name}/diag"] = dg
        for k, v in mf.items():
            out[f"traj/{name}/manifold_{k}"] = v
    # 5. Stress1000 x600 (BASELINE.json config 3): diagnostics every 50 steps and the end state
    o = Oracle("ref").create(); o.set_params(); o.load_scene("Stress1000")
    dg = []
    for s in range(600):
        o.step(1)
        if (s + 1) % 50 == 0:
            d = o.diagnostics(); dg.append([d["maxPen"], d["maxViol"], d["maxLin"], d["maxAng"], d["maxLambda"], d["contacts"], d["manifolds"], d["dynBodies"]])
    out["traj/Stress1000/diag"] = np.array(dg, np.float64)
    out["traj/Stress1000/state_final"] = o.state()
    # 6. Solver::pick on the settled Stress1000 pile
    rays = np.concatenate([rng.uniform(-8, 8, (64, 1)), rng.uniform(5, 25, (64, 1)), rng.uniform(-8, 8, (64, 1)), rng.normal(size=(64, 3)) * 0.3 + np.array([0, -1, 0])], axis=1).astype(np.float32)
    hits, locs = np.zeros(64, np.int32), np.zeros((64, 3), np.float32)
    for t in range(64):
        hits[t], locs[t] = o.pick(rays[t, :3], rays[t, 3:])
    out.update({"pick/rays": rays, "pick/hit": hits, "pick/local": locs})
    o.close()
    out["pile40/bodies"] = np.array([list(b["size"]) + [b["density"], b["friction"]] + list(b["pos"]) + list(b["quat"]) + list(b["lin"]) + list(b["ang"]) for b in pile], np.float64)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
