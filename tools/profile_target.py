"""ncu target: build a workload, warm up, run a few steps.
usage: profile_target.py [grid100|grid50|stress1000|ensemble] [steps]
With AVBD_PROFILE_RANGE=1 the measured steps are bracketed by cudaProfilerStart/Stop (use ncu --profile-from-start off),
so -s / -c count launches of the measured steps only."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "grid100"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = avbd.World()
if name == "stress1000":
    scenes.load(w, scenes.scene("Stress1000")); w.step(400)
elif name.startswith("ensemble") and name != "ensemble":
    scenes.load(w, scenes.ensemble(scenes.scene("Pyramid"), int(name[8:]))); w.step(30)
elif name == "ensemble":
    scenes.load(w, scenes.ensemble(scenes.scene("Pyramid"), 8192)); w.step(30)
else:
    n = 100 if name == "grid100" else 50
    s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
    scenes.load(w, s); w.step(12)
rt = None
if os.environ.get("AVBD_PROFILE_RANGE"):
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    w.sync(); rt.cudaProfilerStart()
w.step(steps)
if rt:
    w.sync(); rt.cudaProfilerStop()
print(w.step_stats())
w.close()
