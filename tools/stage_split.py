"""Per-stage CUDA-event split of a workload (run on the GPU box): python tools/stage_split.py [stress1000|grid100|ensemble|grid20] [steps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "stress1000"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = avbd.World()
if name == "stress1000":
    scenes.load(w, scenes.scene("Stress1000")); w.step(400)
elif name.startswith("ensemble"):
    scenes.load(w, scenes.ensemble(scenes.scene("Pyramid"), int(name[8:] or 8192))); w.step(30)
else:
    n = int(name[4:])
    s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
    scenes.load(w, s); w.step(15)
ms = w.step_timed(steps)
l0 = w.profile()["kernel_launches"]; lib0 = w.profile()["library_launches"]
w.set_profiling(True)
msp = w.step_timed(steps)
p = w.profile()
w.set_profiling(False)
st = w.step_stats()
out = dict(workload=name, steps=steps, ms_per_step=ms / steps, steps_per_s=steps / (ms * 1e-3), profiled_ms_per_step=msp / steps,
           stage_ms={k: p["ms_" + k] / p["steps"] for k in ("broadphase", "narrowphase", "graph", "predict", "solve", "velocity", "step")},
           launches_per_step=(p["kernel_launches"] - l0) / steps, library_launches_per_step=(p["library_launches"] - lib0) / steps,
           graph_builds_per_step=p["graph_builds"] / p["steps"], manifolds=st["manifolds"], contacts=st["contacts"], colours=st["colours"])
print(json.dumps(out))
w.close()
