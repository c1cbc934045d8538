"""CPU-side checks of the product's boundary: the C ABI library loads and exports every symbol include/avbd_b200.h
declares (no compute calls without a GPU), fails loudly without a device, and the product's __host__ __device__
math — compiled for the host by tests/emul — agrees with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from _libs import PKG_DIR, ROOT, Emul, Oracle, copy_bodies


def test_library_exports_every_declared_symbol(avbd):
    hdr = open(os.path.join(ROOT, "include", "avbd_b200.h")).read()
    declared = set(re.findall(r"\b(avbd_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"avbd_world"}
    assert len(declared) >= 40
    L = avbd.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(avbd.ABI), declared ^ set(avbd.ABI)


def test_no_cpu_fallback(avbd):
    if avbd.lib().avbd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(avbd.AvbdError, match="no CUDA device"):
        avbd.World()
    n = np.zeros((1, 10), np.float32)
    with pytest.raises(avbd.AvbdError):
        avbd.collide_pairs(n, n)


def test_host_cli_fails_loudly_without_gpu(avbd):
    exe = os.path.join(PKG_DIR, "host", "avbd_demo3d")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG_DIR, "host")], check=True)
    if avbd.lib().avbd_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "--nogfx", "--scene", "Stack", "--steps", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_product_has_no_oracle_dependency():
    """Nothing under the package (nor the C ABI header) may include, import or link anything from oracle/."""
    for base, _, files in os.walk(PKG_DIR):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle/" not in txt and "avbd_oracle" not in txt and "_libs" not in txt, os.path.join(base, f)


def test_device_collide_math_is_bit_exact_on_host(port):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    emu = Emul()
    for t in range(len(gold["collide/a"])):
        k, f, g = emu.collide(gold["collide/a"][t], gold["collide/b"][t])
        assert k == gold["collide/count"][t] and (f == gold["collide/feat"][t, :k]).all()
        assert g[:, :9].tobytes() == np.ascontiguousarray(gold["collide/geom"][t, :k, :9]).tobytes()
    emu.close()


@pytest.mark.parametrize("scene,steps,tol", [("Stack", 60, 1e-6), ("TwoBlockDrop", 40, 1e-5), ("Pyramid", 6, 1e-6)])
def test_device_step_math_tracks_oracle_on_host(scene, steps, tol):
    """Same visiting order, same inputs: the kernels' row / body functions reproduce the oracle's step up to
    summation-order rounding (chaotic growth afterwards is the reference's own sensitivity)."""
    o = Oracle("port").create()
    o.load_scene(scene)
    p = o.params()
    e = Emul()
    e.set_params(p["dt"], p["g"], p["iterations"], p["alpha"], p["beta"], p["gamma"])
    copy_bodies(o, e)
    props = o.body_props()
    e.set_order([i for i in range(o.n - 1, -1, -1) if props[i, 4] > 0])
    for s in range(steps):
        o.step(1); e.step(1)
    assert np.abs(o.state() - e.state()).max() < tol
    mo, me = o.manifolds(), e.manifolds()
    assert set(mo) == set(me)
    o.close(); e.close()


def test_fused_contact_rows_match_row_by_row_accumulation():
    """contact_system (csrc/avbd_rows.cuh: one contact visit as the solver kernels evaluate it, body-side sign folded
    into the force) against accumulate_contact (solver.cpp:371-399 row by row) on random contacts: all 27 sums are
    bit-identical.  (Checked per contribution on purpose: after the 6x6 solve a penalty-dominated system amplifies ANY
    rounding difference — summation order included — by its condition number.)"""
    rng = np.random.default_rng(21)
    emu = Emul()
    worst = 0.0
    for t in range(3000):
        v = np.zeros(39, np.float32)
        v[0:3] = rng.uniform(-1, 1, 3); v[3:7] = rng.normal(size=4)
        v[7:10] = rng.uniform(-1, 1, 3); v[10:14] = rng.normal(size=4)
        v[14:20] = rng.uniform(-0.8, 0.8, 6); v[20:23] = rng.normal(size=3)
        v[23:26] = rng.uniform(-0.05, 0.05, 3)
        v[26] = -rng.uniform(0, 200); v[27:29] = rng.uniform(-40, 40, 2)
        v[29:32] = 10 ** rng.uniform(4.3, 6.3, 3)
        v[32] = t % 2; v[33] = 0.5; v[34] = 0.95
        v[35:38] = rng.uniform(0.05, 0.5, 3) if t % 3 else 1.0 / 6.0
        v[38] = t % 2 if t % 5 else 1 - t % 2
        fused, rows = emu.contact_system(v)
        for blk in (slice(0, 3), slice(3, 6), slice(6, 12), slice(12, 21), slice(21, 27)):
            scale = np.abs(rows[blk]).max() + 1e-20
            worst = max(worst, float(np.abs(fused[blk] - rows[blk]).max() / scale))
    emu.close()
    assert worst == 0.0, worst
