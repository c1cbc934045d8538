"""ctypes loaders shared by the tests, bench.py and __graft_entry__.

Three checkers (test infrastructure, never the product):
  * Oracle("port")  -> oracle/libavbd_oracle.so   this repo's CPU restatement   (orc_* symbols)
  * Oracle("ref")   -> oracle/_ref/libavbd_ref.so the unmodified reference behind oracle/ref_harness.cpp (ref_* symbols)
  * Emul()          -> tests/emul/libavbd_emul.so the product's __host__ __device__ math compiled for the host
and the product itself:
  * cuda_lib()      -> avbd-demo3d_b200/libavbd_b200.so (include/avbd_b200.h)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PKG_DIR = os.path.join(ROOT, "avbd-demo3d_b200")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def build_emul():
    so = os.path.join(EMUL_DIR, "libavbd_emul.so")
    src = os.path.join(EMUL_DIR, "host_emul.cpp")
    hdrs = [os.path.join(PKG_DIR, "csrc", h) for h in os.listdir(os.path.join(PKG_DIR, "csrc")) if h.endswith(".cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-w", "-shared", "-I/usr/local/cuda/include",
                        "-I" + os.path.join(PKG_DIR, "csrc"), "-x", "c++", src, "-o", so], check=True)
    return so


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libavbd_ref.so"))


class Oracle:
    """Common face of the reference harness (kind='ref') and the restatement (kind='port')."""

    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            path, self.p = os.path.join(ORACLE_DIR, "libavbd_oracle.so"), "orc_"
            if not os.path.exists(path):
                build_oracle()
        else:
            path, self.p = os.path.join(ORACLE_DIR, "_ref", "libavbd_ref.so"), "ref_"
        self.lib = C.CDLL(path)
        L = self.lib
        self._fn("create", C.c_void_p, [])
        self._fn("destroy", None, [C.c_void_p])
        self._fn("set_params", None, [C.c_void_p, C.c_float, f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int])
        self._fn("get_params", None, [C.c_void_p, f32p])
        self._fn("load_scene", C.c_int, [C.c_void_p, C.c_char_p])
        self._fn("add_body", C.c_int, [C.c_void_p, f32p, C.c_float, C.c_float, f32p, f32p, f32p, f32p])
        self._fn("add_joint", None, [C.c_void_p, C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_float])
        self._fn("add_spring", None, [C.c_void_p, C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_float])
        self._fn("add_ignore", None, [C.c_void_p, C.c_int, C.c_int])
        self._fn("step", None, [C.c_void_p, C.c_int])
        self._fn("num_bodies", C.c_int, [C.c_void_p])
        self._fn("get_state", None, [C.c_void_p, f32p])
        self._fn("set_state", None, [C.c_void_p, f32p])
        self._fn("get_prev_linvel", None, [C.c_void_p, f32p])
        self._fn("set_prev_linvel", None, [C.c_void_p, f32p])
        self._fn("get_body_props", None, [C.c_void_p, f32p])
        self._fn("get_diagnostics", None, [C.c_void_p, f32p, i32p])
        self._fn("num_manifolds", C.c_int, [C.c_void_p])
        self._fn("get_manifolds", None, [C.c_void_p, i32p, i32p, i32p, f32p])
        self._fn("collide", C.c_int, [f32p, f32p, i32p, f32p])
        self._fn("solve6x6", None, [f32p, f32p, f32p])
        self._fn("solve3", None, [f32p, f32p, f32p])
        self._fn("pick", C.c_int, [C.c_void_p, f32p, f32p, f32p])
        if kind == "port":
            self._fn("step_ordered", None, [C.c_void_p, i32p, C.c_int])
            self._fn("stage_broadphase", None, [C.c_void_p])
            self._fn("stage_init", None, [C.c_void_p])
            self._fn("stage_predict", None, [C.c_void_p])
            self._fn("stage_primal", None, [C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p])
            self._fn("stage_dual", None, [C.c_void_p, C.c_float])
            self._fn("stage_velocity", None, [C.c_void_p])
            self._fn("stage_diagnostics", None, [C.c_void_p])
            self._fn("overlap_pairs", C.c_int, [C.c_void_p, i32p, C.c_int])
            self._fn("load_stress_grid", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int])
            self._fn("set_logging", None, [C.c_void_p, C.c_int, C.c_int])
            self._fn("set_manifolds", None, [C.c_void_p, C.c_int, i32p, i32p, i32p, f32p])
        self.h = None

    def _fn(self, name, res, args):
        fn = getattr(self.lib, self.p + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, "_" + name, fn)

    # -- world lifecycle
    def create(self):
        self.h = self._create()
        return self

    def close(self):
        if self.h:
            self._destroy(self.h)
            self.h = None

    def __enter__(self):
        return self.create()

    def __exit__(self, *a):
        self.close()

    def set_params(self, dt=1 / 60, g=(0, -10, 0), iterations=10, alpha=0.95, beta=1e5, gamma=0.99, post=False):
        self._set_params(self.h, dt, _f(g), iterations, alpha, beta, gamma, int(post))

    def params(self):
        o = np.zeros(8, np.float32)
        self._get_params(self.h, o)
        return dict(dt=float(o[0]), g=tuple(float(x) for x in o[1:4]), iterations=int(o[4]), alpha=float(o[5]), beta=float(o[6]), gamma=float(o[7]))

    def load_scene(self, name):
        n = self._load_scene(self.h, name.encode())
        assert n >= 0, name
        return n

    def load_stress_grid(self, nx, ny, nz, spacing_y=2.0, start_y=20.0, wide=False):
        return self._load_stress_grid(self.h, nx, ny, nz, spacing_y, start_y, int(wide))

    def add_body(self, size, density, friction, pos, quat=(0, 0, 0, 1), lin=(0, 0, 0), ang=(0, 0, 0)):
        return self._add_body(self.h, _f(size), density, friction, _f(pos), _f(quat), _f(lin), _f(ang))

    def add_joint(self, a, b, anchor_a, anchor_b=(0, 0, 0), lin_k=3.4028234663852886e38, ang_k=3.4028234663852886e38):
        self._add_joint(self.h, a, b, _f(anchor_a), _f(anchor_b), lin_k, ang_k)

    def add_spring(self, a, b, anchor_a, anchor_b, k, rest=-1.0):
        self._add_spring(self.h, a, b, _f(anchor_a), _f(anchor_b), k, rest)

    def add_ignore(self, a, b):
        self._add_ignore(self.h, a, b)

    def step(self, n=1):
        self._step(self.h, n)

    def step_ordered(self, order):
        o = np.ascontiguousarray(order, np.int32)
        self._step_ordered(self.h, o, len(o))

    @property
    def n(self):
        return self._num_bodies(self.h)

    def state(self):
        o = np.zeros((self.n, 13), np.float32)
        if self.n:
            self._get_state(self.h, o)
        return o

    def set_state(self, s):
        self._set_state(self.h, _f(s))

    def prev_linvel(self):
        o = np.zeros((self.n, 3), np.float32)
        if self.n:
            self._get_prev_linvel(self.h, o)
        return o

    def body_props(self):
        o = np.zeros((self.n, 10), np.float32)
        if self.n:
            self._get_body_props(self.h, o)
        return o

    def diagnostics(self):
        f, i = np.zeros(5, np.float32), np.zeros(3, np.int32)
        self._get_diagnostics(self.h, f, i)
        return dict(maxPen=float(f[0]), maxViol=float(f[1]), maxLin=float(f[2]), maxAng=float(f[3]), maxLambda=float(f[4]),
                    contacts=int(i[0]), manifolds=int(i[1]), dynBodies=int(i[2]))

    def manifolds(self):
        m = self._num_manifolds(self.h)
        ints, feats, stick, flts = (np.zeros((m, 3), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 81), np.float32))
        if m:
            self._get_manifolds(self.h, ints, feats, stick, flts)
        return manifold_dict(ints, feats, stick, flts)

    def manifolds_raw(self):
        m = self._num_manifolds(self.h)
        ints, feats, stick, flts = (np.zeros((m, 3), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 81), np.float32))
        if m:
            self._get_manifolds(self.h, ints, feats, stick, flts)
        return ints, feats, stick, flts

    def set_manifolds(self, ints, feats, stick, flts):
        """port only: replace the manifold set (same layout / order manifolds_raw() returns)."""
        ints = np.ascontiguousarray(ints, np.int32).reshape(-1, 3)
        m = len(ints)
        self._set_manifolds(self.h, m, ints, np.ascontiguousarray(feats, np.int32).reshape(m, 4),
                            np.ascontiguousarray(stick, np.int32).reshape(m, 4), _f(flts).reshape(m, 81))

    def overlap_pairs(self):
        cap = max(1024, self.n * 64)
        while True:
            buf = np.zeros((cap, 2), np.int32)
            k = self._overlap_pairs(self.h, buf, cap)
            if k <= cap:
                return buf[:k]
            cap = k

    def stage(self, name, *args):
        getattr(self, "_stage_" + name)(self.h, *args)

    def stage_primal(self, alpha, order=None, want_dx=False):
        dx = np.zeros((self.n, 6), np.float32) if want_dx else None
        o = None if order is None else np.ascontiguousarray(order, np.int32)
        self._stage_primal(self.h, alpha, None if o is None else o.ctypes.data_as(C.c_void_p), 0 if o is None else len(o),
                           None if dx is None else dx.ctypes.data_as(C.c_void_p))
        return dx

    def pick(self, origin, direction):
        local = np.zeros(3, np.float32)
        i = self._pick(self.h, _f(origin), _f(direction), local)
        return i, local

    # -- stateless helpers
    def collide(self, a10, b10):
        feats, out = np.zeros(4, np.int32), np.zeros(40, np.float32)
        k = self._collide(_f(a10), _f(b10), feats, out)
        return k, feats[:k].copy(), out.reshape(4, 10)[:k].copy()

    def solve6x6(self, lhs36, rhs6):
        o = np.zeros(6, np.float32)
        self._solve6x6(_f(lhs36), _f(rhs6), o)
        return o


def manifold_dict(ints, feats, stick, flts):
    """{(a,b): dict(n, mu, feat[n], stick[n], geom[n,14], lam[n,3], pen[n,3])}"""
    out = {}
    for m in range(len(ints)):
        a, b, n = (int(x) for x in ints[m])
        f = flts[m]
        out[(a, b)] = dict(n=n, mu=float(f[0]), feat=feats[m, :n].copy(), stick=stick[m, :n].copy(),
                           geom=f[1:57].reshape(4, 14)[:n].copy(), lam=f[57:69].reshape(4, 3)[:n].copy(), pen=f[69:81].reshape(4, 3)[:n].copy())
    return out


class Emul:
    """Host build of the product's device math (tests/emul/host_emul.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(build_emul())
        L = self.lib
        L.emu_create.restype = C.c_void_p
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_set_params.argtypes = [C.c_void_p, C.c_float, f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
        L.emu_add_body.argtypes = [C.c_void_p, f32p, C.c_float, C.c_float, f32p, f32p, f32p, f32p]
        L.emu_set_order.argtypes = [C.c_void_p, i32p, C.c_int]
        for s in ("collide", "predict", "velocity"):
            getattr(L, "emu_stage_" + s).argtypes = [C.c_void_p]
        L.emu_stage_primal.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.emu_stage_dual.argtypes = [C.c_void_p, C.c_float]
        L.emu_step.argtypes = [C.c_void_p]
        L.emu_get_state.argtypes = [C.c_void_p, f32p]
        L.emu_num_manifolds.argtypes = [C.c_void_p]
        L.emu_get_manifolds.argtypes = [C.c_void_p, i32p, i32p, i32p, f32p]
        L.emu_collide.argtypes = [f32p, f32p, i32p, f32p]
        L.emu_contact_system.argtypes = [f32p, f32p]
        self.h = L.emu_create()
        self.n = 0

    def close(self):
        if self.h:
            self.lib.emu_destroy(self.h)
            self.h = None

    def set_params(self, dt=1 / 60, g=(0, -10, 0), iterations=10, alpha=0.95, beta=1e5, gamma=0.99, post=False):
        self.lib.emu_set_params(self.h, dt, _f(g), iterations, alpha, beta, gamma, int(post))

    def add_body(self, size, density, friction, pos, quat=(0, 0, 0, 1), lin=(0, 0, 0), ang=(0, 0, 0)):
        self.n += 1
        return self.lib.emu_add_body(self.h, _f(size), density, friction, _f(pos), _f(quat), _f(lin), _f(ang))

    def contact_system(self, inp):
        """(fused 27, row-by-row 27) contribution of one contact visit: contact_system vs accumulate_contact."""
        out = np.zeros(54, np.float32)
        self.lib.emu_contact_system(_f(inp), out)
        return out[:27], out[27:]

    def set_order(self, order):
        o = np.ascontiguousarray(order, np.int32)
        self.lib.emu_set_order(self.h, o, len(o))

    def step(self, n=1):
        for _ in range(n):
            self.lib.emu_step(self.h)

    def stage(self, name, *args):
        getattr(self.lib, "emu_stage_" + name)(self.h, *args)

    def stage_primal(self, alpha, want_dx=False):
        dx = np.zeros((self.n, 6), np.float32) if want_dx else None
        self.lib.emu_stage_primal(self.h, alpha, None if dx is None else dx.ctypes.data_as(C.c_void_p))
        return dx

    def state(self):
        o = np.zeros((self.n, 13), np.float32)
        if self.n:
            self.lib.emu_get_state(self.h, o)
        return o

    def manifolds(self):
        m = self.lib.emu_num_manifolds(self.h)
        ints, feats, stick, flts = (np.zeros((m, 3), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 81), np.float32))
        if m:
            self.lib.emu_get_manifolds(self.h, ints, feats, stick, flts)
        return manifold_dict(ints, feats, stick, flts)

    def collide(self, a10, b10):
        feats, out = np.zeros(4, np.int32), np.zeros(40, np.float32)
        k = self.lib.emu_collide(_f(a10), _f(b10), feats, out)
        return k, feats[:k].copy(), out.reshape(4, 10)[:k].copy()


def copy_bodies(src: Oracle, dst):
    """Re-create src's bodies (current state) in dst (Oracle / Emul / cuda World)."""
    props, st = src.body_props(), src.state()
    for i in range(src.n):
        size = props[i, 0:3]
        vol = float(size[0]) * float(size[1]) * float(size[2])
        density = float(props[i, 3]) / vol if vol > 0 else 0.0
        dst.add_body(size, density, float(props[i, 8]), st[i, 0:3], st[i, 3:7], st[i, 7:10], st[i, 10:13])


def random_pile(rng, n, spread=2.0, ground=True):
    """Bodies for a dense pile of randomly oriented boxes (exercises edge contacts and tilted faces)."""
    bodies = []
    if ground:
        bodies.append(dict(size=(30, 1, 30), density=0.0, friction=0.5, pos=(0, -0.5, 0), quat=(0, 0, 0, 1), lin=(0, 0, 0), ang=(0, 0, 0)))
    for _ in range(n):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        bodies.append(dict(size=tuple(rng.uniform(0.4, 1.4, 3)), density=float(rng.uniform(0.5, 2.0)), friction=float(rng.uniform(0.2, 0.8)),
                           pos=(rng.uniform(-spread, spread), rng.uniform(0.6, 0.6 + 2 * spread), rng.uniform(-spread, spread)),
                           quat=tuple(q), lin=tuple(rng.normal(size=3)), ang=tuple(rng.normal(size=3))))
    return bodies


def add_all(dst, bodies):
    for b in bodies:
        dst.add_body(b["size"], b["density"], b["friction"], b["pos"], b["quat"], b["lin"], b["ang"])
