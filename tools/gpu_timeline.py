"""Per-step timeline probe (run under gpurun): usage gpu_timeline.py <n> <jitter_y> <steps>"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
n = int(sys.argv[1]); jy = float(sys.argv[2]); steps = int(sys.argv[3])
w = avbd.World()
s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True, jitter_y=jy); s["params"]["iterations"] = 10
scenes.load(w, s); w.step(1); w.step_stats()
for k in range(steps):
    t0 = time.perf_counter(); ms = w.step_timed(1); wall = 1e3 * (time.perf_counter() - t0)
    st = w.step_stats(); d = w.diagnostics()
    print(k, f"wall={wall:.1f} dev={ms:.1f}", {a: round(st[a], 2) for a in ("ms_broadphase", "ms_narrowphase", "ms_graph", "ms_predict", "ms_primal", "ms_velocity", "ms_total")},
          st["pairs"], st["candidates"], st["manifolds"], st["contacts"], st["colours"], f"pen={d['maxPen']:.3f} lin={d['maxLin']:.2f}", flush=True)
w.close()
