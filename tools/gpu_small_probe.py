"""Small-world latency probe (run under gpurun): per-stage device time and wall time per step of Stress1000 / Pyramid.
usage: gpu_small_probe.py [scene] [settle] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "Stress1000"
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 400
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
w = avbd.World()
scenes.load(w, scenes.scene(name)); w.step(settle); w.step_stats(); w.step(3)
st = w.step_stats()
print(name, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
p0 = w.profile()
t0 = time.perf_counter(); ms = w.step_timed(steps); wall = time.perf_counter() - t0
p1 = w.profile()
print(f"{name}: {steps} steps  device {ms / steps:.3f} ms/step  wall {1e3 * wall / steps:.3f} ms/step  {steps / wall:.1f} steps/s  "
      f"own launches/step {(p1['kernel_launches'] - p0['kernel_launches']) / steps:.1f}  library launches/step {(p1['library_launches'] - p0['library_launches']) / steps:.1f}")
w.close()
