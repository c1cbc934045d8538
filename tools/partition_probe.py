"""Partition-invariance probe at flat-path sizes (run under gpurun): W worlds in one batch vs two halves."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import avbd_demo3d_b200 as avbd
from avbd_demo3d_b200 import scenes
def run(base, worlds, first, steps):
    w = avbd.World(); scenes.load(w, scenes.ensemble(base, worlds, first_world=first)); w.step(steps)
    st = w.state(); w.close(); return st
for name, W, steps in (("Stack", 1200, 40), ("Pyramid", 400, 30)):
    base = scenes.scene(name); n = len(base["size"])
    whole = run(base, W, 0, steps); lo = run(base, W // 2, 0, steps); hi = run(base, W // 2, W // 2, steps)
    both = np.concatenate([lo, hi])
    diff = np.abs(whole - both)
    bad = np.unique(np.nonzero(diff.max(axis=1) > 0)[0] // n)
    print(name, "worlds", W, "bodies", len(whole), "bit-identical", whole.tobytes() == both.tobytes(), "max diff", float(diff.max()), "worlds differing", len(bad), flush=True)
