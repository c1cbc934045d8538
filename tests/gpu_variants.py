"""Tuning aid (run under gpurun): time the grid step for each AVBD_PRIMAL_VARIANT in fresh processes.
usage: gpu_variants.py [variant ...]   env GRID=100 JY=0.25"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path[:0] = [%r, %r]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
w = avbd.World()
n = int(os.environ.get("GRID", "100")); jy = float(os.environ.get("JY", "0.25"))
s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True, jitter_y=jy); s["params"]["iterations"] = 10
scenes.load(w, s); w.step(14)
w.set_profiling(True)
ms = w.step_timed(10)
p = w.profile(); st = w.step_stats()
print(json.dumps(dict(variant=os.environ.get("AVBD_PRIMAL_VARIANT"), ms_per_step=round(ms/10, 3), primal_ms_per_step=round(p["ms_primal"]/10, 3), dual_ms_per_step=round(p["ms_dual"]/10, 3),
      primal_launch_us=round(1e3*p["ms_primal"]/p["primal_launches"], 2), colours=st["colours"], contacts=st["contacts"],
      primal_GBs=round((100*p["primal_bodies"]+124*p["primal_visits"])/p["ms_primal"]/1e6, 1))))
''' % (ROOT, os.path.join(ROOT, "tests"))
for v in sys.argv[1:] or ["43", "v28", "v16", "v64"]:
    env = dict(os.environ, AVBD_PRIMAL_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-800:], flush=True)
