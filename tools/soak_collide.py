"""Soak of the collision stage against the oracle on identical inputs (run on the GPU box): random tilted piles of many sizes, the
oracle advanced a few steps, then ONE collide stage on both sides from the oracle's bodies and manifold history — membership, feature
ids, anchors, C0, lambda, penalty, stick compared bit for bit (the check of tests/test_gpu_stage_parity.py, over many more inputs).
usage: python tools/soak_collide.py [first seed] [count]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import avbd_demo3d_b200 as avbd
from _libs import random_pile
from test_gpu_parity import assert_manifolds_equal, gpu_manifolds, make_pair
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
total = 0
t0 = time.time()
for seed in range(first, first + count):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([20, 60, 150, 300, 600, 1200, 2500]))
    spread = float(rng.uniform(1.2, 3.5)) * (n / 300.0) ** (1.0 / 3.0)
    o, w = make_pair(avbd, bodies=random_pile(rng, n, spread))
    try:
        o.step(int(rng.integers(2, 7)))
        raw = [a.copy() for a in o.manifolds_raw()]
        w.set_state(o.state()); w.set_prev_linvel(o.prev_linvel())
        w.upload_manifolds(*raw)
        w.stage("collide")
        o.stage("broadphase"); o.stage("init")
        want, got = o.manifolds(), gpu_manifolds(w)
        assert_manifolds_equal(got, want, exact_rows=True, ctx=("soak", seed, n))
        total += len(want)
        print(f"seed {seed}: {n} bodies, {len(want)} manifolds ok", flush=True)
    finally:
        o.close(); w.close()
print(f"soak ok: {count} piles, {total} manifolds bit-identical, {time.time() - t0:.0f} s")
