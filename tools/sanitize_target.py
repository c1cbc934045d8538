"""compute-sanitizer target (run under gpurun): a few steps of small scenes through both solver paths, joints / springs included."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
for name, steps in (("Pyramid", 12), ("Stress1000", 4)):
    w = avbd.World(); scenes.load(w, scenes.scene(name)); w.step(steps); print(name, w.diagnostics()); w.close()
w = avbd.World()
w.add_body((20, 1, 20), 0.0, 0.5, (0, -0.5, 0)); w.add_body((1, 1, 1), 1.0, 0.5, (0, 3, 0)); w.add_body((1, 1, 1), 1.0, 0.5, (3, 3, 0)); w.add_body((1, 1, 1), 1.0, 0.5, (3, 5, 0))
w.add_joint(-1, 1, (0, 3.5, 0)); w.add_spring(2, 3, (0, 0.5, 0), (0, -0.5, 0), 1000.0, 1.0)
w.step(10); print("jointed", w.diagnostics()); w.close()
s = scenes.stress_grid(12, 12, 12, spacing_y=1.01, start_y=0.51, wide_ground=True)
w = avbd.World(); scenes.load(w, s); w.step(4); print("grid12", w.diagnostics()); w.pick((0, 50, 0), (0, -1, 0)); w.close()
# a world past the single-block colouring limit: cooperative-grid colouring + one-body-per-thread sweeps, with the state handed back
# and forth (snapshot / restore, manifold upload) in between
s = scenes.stress_grid(22, 22, 22, spacing_y=1.01, start_y=0.51, wide_ground=True)
w = avbd.World(); scenes.load(w, s); w.step(3)
blob = w.snapshot(); raw = w.manifolds_raw(); w.upload_manifolds(*raw); w.step(1); w.restore(blob); w.step(1)
print("grid22", w.diagnostics(), w.step_stats()["colours"]); w.close()
