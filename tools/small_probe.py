"""Small-world probe (run under gpurun): steps/s of Stress1000 and Pyramid through the cluster loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
for name in ("Stress1000", "Pyramid"):
    w = avbd.World(); scenes.load(w, scenes.scene(name)); w.step(400); ms = w.step_timed(300)
    print(os.environ.get("AVBD_TILE_CACHE_SLOTS", "default"), name, "steps/s", round(300 / (ms * 1e-3), 1), w.diagnostics()["contacts"], flush=True); w.close()
