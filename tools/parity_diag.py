"""Diagnostics for the per-stage parity tests (run on the GPU box): prints what differs and by how much."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from _libs import Oracle, manifold_dict, random_pile
from test_gpu_parity import make_pair, gpu_manifolds


def dual_case(scene, warm, sweeps, post, bodies=None):
    o, w = make_pair(avbd, scene=scene, bodies=bodies)
    p = o.params()
    if post:
        o.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"], True)
        w.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"], True)
    o.step(warm)
    alpha = 1.0 if post else p["alpha"]
    o.stage("broadphase"); o.stage("init"); o.stage("predict")
    for it in range(sweeps - 1):
        o.stage_primal(alpha); o.stage("dual", alpha)
    o.stage_primal(alpha)
    raw = o.manifolds_raw()
    pre = manifold_dict(*raw)
    w.set_state(o.state()); w.upload_manifolds(*raw)
    w.stage("dual", alpha); o.stage("dual", alpha)
    got, want = gpu_manifolds(w), o.manifolds()
    nl = npn = ns = nc = 0
    worst = []
    for k, r in want.items():
        g, q = got[k], pre[k]
        for c in range(r["n"]):
            nc += 1
            dl = np.abs(g["lam"][c] - r["lam"][c]); dp = np.abs(g["pen"][c] - r["pen"][c])
            bl = dl > 1e-4 * np.abs(r["lam"][c]) + 1e-3; bp = dp > 1e-4 * np.abs(r["pen"][c]) + 1e-1
            if bl.any() or bp.any() or g["stick"][c] != r["stick"][c]:
                nl += int(bl.any()); npn += int(bp.any()); ns += int(g["stick"][c] != r["stick"][c])
                if len(worst) < 6:
                    worst.append((k, c, "lam pre/gpu/ref", q["lam"][c], g["lam"][c], r["lam"][c], "pen pre/gpu/ref", q["pen"][c], g["pen"][c], r["pen"][c],
                                  "stick pre/gpu/ref", int(q["stick"][c]), int(g["stick"][c]), int(r["stick"][c]), "mu", r["mu"]))
    print(f"DUAL {scene} warm={warm} sweeps={sweeps} post={post}: contacts={nc} bad_lambda={nl} bad_penalty={npn} stick_diff={ns}")
    for x in worst:
        print("   ", x)
    o.close(); w.close()


def ensemble_case(pick, steps=400, tail=100):
    from avbd_demo3d_b200 import scenes
    base = scenes.scene("Pyramid")
    nb = len(base["size"])
    ens = scenes.ensemble(base, 8192)
    w = avbd.World(); o = Oracle("port").create()
    scenes.load(w, ens)
    sl = slice(pick * nb, (pick + 1) * nb)
    for i in range(sl.start, sl.stop):
        o.add_body(ens["size"][i], float(ens["density"][i]), float(ens["friction"][i]), ens["pos"][i], ens["quat"][i], ens["lin"][i], ens["ang"][i])
    w.step(steps - tail); o.step(steps - tail)
    ys, yo = [], []
    for _ in range(tail):
        w.step(1); o.step(1)
        ys.append(w.state()[sl, 1].copy()); yo.append(o.state()[:, 1].copy())
    y, yr = np.mean(ys, axis=0), np.mean(yo, axis=0)
    d = np.abs(y - yr)
    print(f"ENSEMBLE world {pick}: max dy {d.max():.5f} at body {d.argmax()} (y {y[d.argmax()]:.4f} vs {yr[d.argmax()]:.4f}); bodies over 1e-3: {(d > 1e-3).sum()}",
          w.world_diagnostics()[pick], o.diagnostics())
    st, so = w.state()[sl], o.state()
    print("    lateral max", np.abs(st[:, [0, 2]] - so[:, [0, 2]]).max(), "jitter", ens["pos"][sl.start + 1] - base["pos"][1])
    w.close(); o.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["dual", "ensemble"]
    if "dual" in what:
        for c in [("Stack", 30, 3, False), ("Pyramid", 40, 2, False), ("Stress1000", 150, 3, False), ("Pyramid", 25, 2, True), ("Stack", 12, 1, True)]:
            dual_case(*c)
        rng = np.random.default_rng(23)
        dual_case(None, 6, 2, False, bodies=random_pile(rng, 250, 2.5))
    if "ensemble" in what:
        for pick in (0, 1, 4097, 8191):
            ensemble_case(pick)
