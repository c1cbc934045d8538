"""Tuning aid (run under gpurun): time the 1M-box grid step for each AVBD_PRIMAL_VARIANT in fresh processes."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path[:0] = [%r, %r]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
w = avbd.World()
n = int(os.environ.get("GRID", "100"))
s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
scenes.load(w, s); w.step(5)
w.set_profiling(True)
ms = w.step_timed(10)
p = w.profile()
print(json.dumps(dict(variant=os.environ.get("AVBD_PRIMAL_VARIANT"), ms_per_step=ms/10, primal_ms_per_step=p["ms_primal"]/10, dual_ms_per_step=p["ms_dual"]/10,
      primal_launch_us=1e3*p["ms_primal"]/p["primal_launches"], primal_GBs=(100*p["primal_bodies"]+124*p["primal_visits"])/p["ms_primal"]/1e6)))
''' % (ROOT, os.path.join(ROOT, "tests"))
for v in sys.argv[1:] or ["82", "83", "42", "43", "44", "162", "163", "23", "24"]:
    env = dict(os.environ, AVBD_PRIMAL_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
