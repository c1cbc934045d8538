"""An ensemble as K concurrent sub-batches on ONE GPU (each avbd world has its own stream; worlds are independent, so a rank's share
can be cut into K batches stepping side by side from K host threads): python tools/ensemble_streams.py <worlds on this GPU> [steps]
Prints ms per step of the whole share for K = 1, 2, 4, 8 (wall clock between two device-wide syncs, and the slowest batch's CUDA-event time)."""
import os, sys, json, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for K in (1, 2, 4, 8):
    per = total // K
    ws = []
    for k in range(K):
        w = avbd.World(); scenes.load(w, scenes.ensemble(scenes.scene("Pyramid"), per, first_world=k * per)); ws.append(w)
    ms = [0.0] * K
    def run(k, n, timed):
        if timed: ms[k] = ws[k].step_timed(n)
        else: ws[k].step(n)
    def all_(n, timed):
        th = [threading.Thread(target=run, args=(k, n, timed)) for k in range(K)]
        for t in th: t.start()
        for t in th: t.join()
    all_(20, False)
    for w in ws: w.sync()
    t0 = time.perf_counter()
    all_(steps, True)
    for w in ws: w.sync()
    wall = time.perf_counter() - t0
    print(json.dumps(dict(worlds=total, batches=K, wall_ms_per_step=1e3 * wall / steps, slowest_batch_ms_per_step=max(ms) / steps,
                          manifolds=sum(w.step_stats()["manifolds"] for w in ws))))
    for w in ws: w.close()
