"""Rest-height deviation from the oracle on Stack / Pyramid (run under gpurun with env toggles)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
from _libs import Oracle
for name in ("Stack", "Pyramid"):
    o = Oracle("port").create(); o.load_scene(name)
    w = avbd.World(); scenes.load(w, scenes.scene(name))
    yo, yw = [], []
    for s in range(600):
        o.step(1); w.step(1)
        if s >= 500:
            yo.append(o.state()[:, 1].copy()); yw.append(w.state()[:, 1].copy())
    yo, yw = np.mean(yo, 0), np.mean(yw, 0)
    print(name, os.environ.get("AVBD_B200_LIB", "default")[-24:], os.environ.get("AVBD_PERSISTENT_MAX_BODIES"), os.environ.get("AVBD_PRIMAL_VARIANT"),
          "maxdev", float(np.abs(yo - yw).max()), w.diagnostics()["maxLin"], o.diagnostics()["maxLin"], flush=True)
    w.close(); o.close()
