"""Per-colour launches vs cluster loop on the small scenes (run on the GPU box): prints steps/s for each."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import avbd_demo3d_b200 as avbd
    from avbd_demo3d_b200 import scenes
    out = {}
    for name, settle, steps in (("TwoBlockDrop", 100, 400), ("Stack", 200, 400), ("Pyramid", 200, 400), ("Wall", 200, 400), ("Stress1000", 400, 300)):
        w = avbd.World(); scenes.load(w, scenes.scene(name)); w.step(settle)
        ms = w.step_timed(steps); out[name] = round(steps / (ms * 1e-3), 1); w.close()
    for n in (12, 16, 20):
        s = scenes.stress_grid(n, n, n, spacing_y=1.01, start_y=0.51, wide_ground=True); s["params"]["iterations"] = 10
        w = avbd.World(); scenes.load(w, s); w.step(30)
        ms = w.step_timed(100); out[f"grid{n}"] = round(100 / (ms * 1e-3), 1); w.close()
    print(json.dumps(out))
else:
    for mode in ("launch", "cluster", "warps"):
        env = dict(os.environ, AVBD_LOOP=mode, AVBD_PERSISTENT_MAX_BODIES="8192")
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(mode, r.stdout.strip(), r.stderr[-300:])
