"""Quick GPU sanity + timing probe (run under gpurun; writes gpurun_out/probe.json)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from _libs import Oracle

out = {}
for scene, steps in (("TwoBlockDrop", 300), ("Pyramid", 300), ("Stress1000", 300)):
    o = Oracle("port").create(); o.load_scene(scene); p = o.params()
    w = avbd.World(); w.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"])
    props, st = o.body_props(), o.state()
    vol = props[:, 0] * props[:, 1] * props[:, 2]
    w.add_bodies(props[:, 0:3], np.where(vol > 0, props[:, 3] / np.maximum(vol, 1e-30), 0), props[:, 8], st[:, 0:3], st[:, 3:7], st[:, 7:10], st[:, 10:13])
    w.step(5); w.step_stats()
    t0 = time.time(); w.step(steps); dt = time.time() - t0
    stats = w.step_stats(); d = w.diagnostics()
    t1 = time.time(); o.step(min(steps, 60)); dt_o = (time.time() - t1) / min(steps, 60)
    s = w.state()
    out[scene] = dict(steps_per_s=steps / dt, ms_per_step=1e3 * dt / steps, oracle_ms_per_step=1e3 * dt_o, stats=stats, diag=d,
                      y_min=float(s[1:, 1].min()) if len(s) > 1 else None, y_max=float(s[1:, 1].max()) if len(s) > 1 else None)
    print(scene, json.dumps(out[scene]))
    w.close(); o.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
