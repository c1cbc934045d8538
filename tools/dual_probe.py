"""Deferred-dual probe (run under gpurun): state difference between the deferred and the one-pass-per-iteration dual after n steps."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
def run(steps, env):
    for k in ("AVBD_PERSISTENT_MAX_BODIES", "AVBD_SEPARATE_DUAL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    w = avbd.World(); scenes.load(w, scenes.scene(sys.argv[1] if len(sys.argv) > 1 else "Pyramid")); w.step(steps)
    st = w.state(); m = w.manifolds_raw(); w.close()
    return st, m
for n in (1, 2, 3, 4, 6, 8, 10, 12):
    a, ma = run(n, {"AVBD_PERSISTENT_MAX_BODIES": "0", "AVBD_SEPARATE_DUAL": "1"})
    b, mb = run(n, {"AVBD_PERSISTENT_MAX_BODIES": "0"})
    c, mc = run(n, {})
    same = ma[3].shape == mb[3].shape
    print(n, "flat deferred vs separate", float(np.abs(a - b).max()), "cluster vs separate", float(np.abs(a - c).max()),
          "manifold floats", float((np.abs(ma[3] - mb[3]) / (1 + np.abs(ma[3]))).max()) if same else "shape differs", flush=True)
