"""Random-pile probe (run under gpurun): one primal sweep's dx against the oracle (same inputs, same order), then per-step drift."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from _libs import Oracle, random_pile, add_all
for seed, post, cubes in ((1, False, False), (1, False, True), (1, True, False), (2, False, False)):
    rng = np.random.default_rng(seed)
    bodies = random_pile(rng, 30)
    if cubes:
        for b in bodies[1:]: b["size"] = (1.0, 1.0, 1.0)
    o = Oracle("port").create(); w = avbd.World()
    o.set_params(post=post); w.set_params(post=post)
    add_all(o, bodies); add_all(w, bodies)
    p = o.params()
    w.stage("collide"); w.stage("predict"); w.stage("colour")
    o.stage("broadphase"); o.stage("init"); o.stage("predict")
    col, k = w.colours()
    dyn = np.arange(len(col))[col >= 0]
    order = dyn[np.lexsort((dyn, col[dyn]))].astype(np.int32)
    a0 = 1.0 if post else p["alpha"]
    want = o.stage_primal(a0, order, want_dx=True)
    got = w.stage_primal(a0, want_dx=True)
    scale = np.abs(want[dyn]).max(axis=1, keepdims=True)
    err = np.abs(got[dyn] - want[dyn])
    rel = (err / (scale + 1e-6)).max()
    print(f"seed {seed} post {post} cubes {cubes}: first sweep dx max abs err {err.max():.3e} rel {rel:.3e} (|dx| max {np.abs(want[dyn]).max():.3e}); state diff {np.abs(o.state()[:, :7] - w.state()[:, :7]).max():.3e}", flush=True)
    o.close(); w.close()
