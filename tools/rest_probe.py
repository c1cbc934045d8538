"""Rest-height probe (run under gpurun): Stack / Pyramid rest heights vs the oracle for the cluster loop and the launch path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from test_gpu_scenes import run_scene, oracle_rest
for name in ("Stack", "Pyramid"):
    o, yo = oracle_rest(name, 600)
    for tag, env in (("cluster", None), ("launch", "0")):
        w = None
        if env is not None: os.environ["AVBD_PERSISTENT_MAX_BODIES"] = env
        else: os.environ.pop("AVBD_PERSISTENT_MAX_BODIES", None)
        w, y = run_scene(avbd, name, 600)
        d = w.diagnostics()
        print(name, tag, "max|dy|", float(np.abs(y - yo).max()), "argmax", int(np.abs(y - yo).argmax()), d["manifolds"], d["contacts"], "maxLin", d["maxLin"], "oracle maxLin", o.diagnostics()["maxLin"], "tail ke", w.tail_ke, "oracle tail ke", o.tail_ke, "end ke", float((w.state()[:, 7:10] ** 2).sum()), float((o.state()[:, 7:10] ** 2).sum()))
        if name == "Stack": print("  dy", np.round(y - yo, 5))
        w.close()
    o.close()
