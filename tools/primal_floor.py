"""Where does the primal visit-sum kernel's time go?  Times one sweep over all colours of the 1M-box world in three modes:
product kernel / memory accesses only / math only (sequential indices).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
w = avbd.World(); scenes.load(w, s); w.step(14)
st = w.step_stats()
print({k: st[k] for k in ("manifolds", "contacts", "contactVisits", "colours")})
for mode, name in ((0, "product"), (1, "memory only"), (2, "math only, sequential indices")):
    print(f"mode {mode} ({name}): {w.debug_time_primal(mode, 10):.4f} ms per sweep")
w.close()
