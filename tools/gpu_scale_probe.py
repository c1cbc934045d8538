"""Scale probe (run under gpurun): pre-stacked and falling Stress grids up to 100^3; prints per-stage timings."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes

sizes = [int(a) for a in sys.argv[1:]] or [30, 50, 100]
out = {}
for n in sizes:
    for label, kw in (("stacked", dict(spacing_y=1.01, start_y=0.51)), ("falling", dict(spacing_y=2.0, start_y=20.0))):
        w = avbd.World()
        s = scenes.stress_grid(n, n, n, wide_ground=True, **kw)
        s["params"]["iterations"] = 10
        t0 = time.time(); scenes.load(w, s); t_load = time.time() - t0
        w.step(3); w.step_stats()
        rec = []
        for k in range(3):
            t0 = time.time(); w.step(5); dt = (time.time() - t0) / 5
            st = w.step_stats(); d = w.diagnostics()
            rec.append(dict(ms_wall=1e3 * dt, **{k2: st[k2] for k2 in ("ms_broadphase", "ms_narrowphase", "ms_predict", "ms_graph", "ms_primal", "ms_velocity", "ms_total", "pairs", "candidates", "manifolds", "contacts", "colours")}, maxPen=d["maxPen"], nan=d["nanEvents"]))
        out[f"{n}^3-{label}"] = dict(load_s=t_load, bodies=w.n, steps=rec)
        print(f"{n}^3-{label}", json.dumps(out[f"{n}^3-{label}"]), flush=True)
        w.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scale_probe.json"), "w"), indent=1)
