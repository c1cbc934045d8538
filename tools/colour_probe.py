"""Colouring probe (run under gpurun): colour histogram and step time of the 1M-box grid, incremental vs from-scratch colouring."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for full in ("0", "1"):
    os.environ["AVBD_INCREMENTAL_COLOUR"] = "0" if full == "1" else "1"
    s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
    w = avbd.World(); scenes.load(w, s); w.step(15)
    ms = w.step_timed(10)
    col, k = w.colours()
    st = w.step_stats()
    print("full" if full == "1" else "incremental", "ms/step", round(ms / 10, 3), "colours", k, "histogram", np.bincount(col[col >= 0]).tolist(),
          {x: round(st[x], 3) for x in ("ms_graph", "ms_primal")}, flush=True)
    w.close()
