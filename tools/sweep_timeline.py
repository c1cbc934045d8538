"""What a colour phase of a small world is made of (run on the GPU box with the AVBD_TIMELINE build):
   make -C avbd-demo3d_b200 variant NAME=timeline SOLVE_DEFS=-DAVBD_TIMELINE
   AVBD_B200_LIB=avbd-demo3d_b200/variants/libavbd_b200_timeline.so python tools/sweep_timeline.py [scene] [steps]
Prints the mean of each interval between the %globaltimer stamps of warp 0 of every sweep launch."""
import ctypes as C, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "Stress1000"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = C.CDLL(os.environ["AVBD_B200_LIB"])
lib.avbd_debug_timeline.argtypes = [C.c_void_p, C.c_int]; lib.avbd_debug_timeline.restype = C.c_int
w = avbd.World(); scenes.load(w, scenes.scene(name)); w.step(400)
buf = np.zeros((16384, 6), np.uint64)
lib.avbd_debug_timeline(buf.ctypes.data, 16384)          # reset
ms = w.step_timed(steps)
rows = lib.avbd_debug_timeline(buf.ctypes.data, 16384)
t = buf[:rows].astype(np.int64)
d = {"scene": name, "steps": steps, "ms_per_step": ms / steps, "launches": rows,
     "entry_to_wait": float(np.mean(t[:, 1] - t[:, 0])), "wait_for_predecessor": float(np.mean(t[:, 2] - t[:, 1])),
     "operands_land": float(np.mean(t[:, 3] - t[:, 2])), "rows": float(np.mean(t[:, 4] - t[:, 3])), "reduce_solve_store": float(np.mean(t[:, 5] - t[:, 4])),
     "in_kernel_total": float(np.mean(t[:, 5] - t[:, 0]))}
o = np.argsort(t[:, 0]); ts = t[o]
per = np.diff(ts[:, 0]); gap = ts[1:, 0] - ts[:-1, 5]; chain = ts[1:, 2] - ts[:-1, 5]
keep = per < 50000                                            # same step (the next step's first launch comes after collision + graph)
d.update(period=float(np.mean(per[keep])), next_entry_after_prev_end=float(np.mean(gap[keep])), next_released_after_prev_end=float(np.mean(chain[keep])))
print(json.dumps(d))
w.close()
