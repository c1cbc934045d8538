"""avbd-demo3d_b200 — B200-native AVBD step loop behind the reference's Solver API.

This module is the thin Python face of libavbd_b200.so (include/avbd_b200.h): a
ctypes binding plus a `World` convenience class used by the tests and bench.py.
The product is the CUDA library; there is no CPU path — importing works without
a GPU (so the ABI can be inspected), but creating a World raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVBD_B200_LIB") or os.path.join(_HERE, "libavbd_b200.so")   # override: A/B builds while tuning
FLT_MAX = 3.4028234663852886e38

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Diagnostics(C.Structure):
    _fields_ = [("maxPenetration", C.c_float), ("maxConstraintViolation", C.c_float), ("maxLinearSpeed", C.c_float),
                ("maxAngularSpeed", C.c_float), ("maxNormalImpulse", C.c_float), ("activeContacts", C.c_int),
                ("activeManifolds", C.c_int), ("dynamicBodies", C.c_int), ("nanEvents", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StepStats(C.Structure):
    _fields_ = [("ms_broadphase", C.c_float), ("ms_narrowphase", C.c_float), ("ms_graph", C.c_float), ("ms_predict", C.c_float),
                ("ms_primal", C.c_float), ("ms_dual", C.c_float), ("ms_velocity", C.c_float), ("ms_total", C.c_float),
                ("bodies", C.c_int), ("dynamicBodies", C.c_int), ("pairs", C.c_int), ("candidates", C.c_int), ("manifolds", C.c_int),
                ("contacts", C.c_int), ("colours", C.c_int), ("iterations", C.c_int), ("contactVisits", C.c_int),
                ("kernelLaunches", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Profile(C.Structure):
    _fields_ = [("ms_primal", C.c_double), ("ms_dual", C.c_double), ("steps", C.c_longlong), ("primal_sweeps", C.c_longlong),
                ("primal_launches", C.c_longlong), ("primal_bodies", C.c_longlong), ("primal_visits", C.c_longlong),
                ("dual_launches", C.c_longlong), ("dual_contacts", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("library_launches", C.c_longlong), ("deferred_dual_contacts", C.c_longlong),
                ("ms_broadphase", C.c_double), ("ms_narrowphase", C.c_double), ("ms_graph", C.c_double), ("ms_predict", C.c_double),
                ("ms_solve", C.c_double), ("ms_velocity", C.c_double), ("ms_step", C.c_double),
                ("bodies", C.c_longlong), ("pairs", C.c_longlong), ("candidates", C.c_longlong), ("manifolds", C.c_longlong),
                ("manifolds_prev", C.c_longlong), ("contacts", C.c_longlong), ("visits", C.c_longlong), ("graph_builds", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/avbd_b200.h declares
ABI = {
    "avbd_last_error": (C.c_char_p, []),
    "avbd_device_count": (C.c_int, []),
    "avbd_world_create": (C.c_void_p, [C.c_int]),
    "avbd_world_destroy": (None, [C.c_void_p]),
    "avbd_clear": (C.c_int, [C.c_void_p]),
    "avbd_set_params": (C.c_int, [C.c_void_p, C.c_float, _f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "avbd_default_params": (C.c_int, [C.c_void_p]),
    "avbd_add_bodies": (C.c_int, [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "avbd_num_bodies": (C.c_int, [C.c_void_p]),
    "avbd_add_joint": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float]),
    "avbd_add_joint_raw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_float, C.c_float]),
    "avbd_set_force_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "avbd_get_force_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "avbd_download_user_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "avbd_num_joints": (C.c_int, [C.c_void_p]),
    "avbd_num_springs": (C.c_int, [C.c_void_p]),
    "avbd_host_alloc": (C.c_void_p, [C.c_longlong]),
    "avbd_host_free": (None, [C.c_void_p]),
    "avbd_host_register": (C.c_int, [C.c_void_p, C.c_longlong]),
    "avbd_download_state_chunked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "avbd_host_unregister": (C.c_int, [C.c_void_p]),
    "avbd_add_spring": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float]),
    "avbd_add_ignore": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "avbd_step": (C.c_int, [C.c_void_p, C.c_int]),
    "avbd_sync": (C.c_int, [C.c_void_p]),
    "avbd_step_timed": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "avbd_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "avbd_get_profile": (C.c_int, [C.c_void_p, C.POINTER(Profile)]),
    "avbd_download_state": (C.c_int, [C.c_void_p, _f32p]),
    "avbd_upload_state": (C.c_int, [C.c_void_p, _f32p]),
    "avbd_upload_state_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p]),
    "avbd_download_prev_linvel": (C.c_int, [C.c_void_p, _f32p]),
    "avbd_upload_prev_linvel": (C.c_int, [C.c_void_p, _f32p]),
    "avbd_download_body_props": (C.c_int, [C.c_void_p, _f32p]),
    "avbd_get_diagnostics": (C.c_int, [C.c_void_p, C.POINTER(Diagnostics)]),
    "avbd_get_step_stats": (C.c_int, [C.c_void_p, C.POINTER(StepStats)]),
    "avbd_num_worlds": (C.c_int, [C.c_void_p]),
    "avbd_get_world_diagnostics": (C.c_int, [C.c_void_p, C.POINTER(Diagnostics), C.c_int]),
    "avbd_world_diagnostics_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "avbd_num_manifolds": (C.c_int, [C.c_void_p]),
    "avbd_download_manifolds": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _f32p]),
    "avbd_upload_manifolds": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, _f32p]),
    "avbd_snapshot_bytes": (C.c_longlong, [C.c_void_p]),
    "avbd_snapshot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong]),
    "avbd_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong]),
    "avbd_stage_broadphase": (C.c_int, [C.c_void_p]),
    "avbd_download_pairs": (C.c_int, [C.c_void_p, _i32p, C.c_int]),
    "avbd_stage_collide": (C.c_int, [C.c_void_p]),
    "avbd_stage_predict": (C.c_int, [C.c_void_p]),
    "avbd_stage_colour": (C.c_int, [C.c_void_p]),
    "avbd_download_colours": (C.c_int, [C.c_void_p, _i32p, C.POINTER(C.c_int)]),
    "avbd_stage_primal": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p]),
    "avbd_stage_dual": (C.c_int, [C.c_void_p, C.c_float]),
    "avbd_stage_velocity": (C.c_int, [C.c_void_p]),
    "avbd_collide_pairs": (C.c_int, [C.c_int, C.c_int, _f32p, _f32p, _i32p, _i32p, _f32p]),
    "avbd_solve6x6": (C.c_int, [C.c_int, C.c_int, _f32p, _f32p, _f32p]),
    "avbd_pick": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
}

_lib = None


def build(force=False):
    """Compile libavbd_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-s", "-C", _HERE] + (["-B"] if force else [])
    subprocess.run(cmd, check=True)
    return LIB_PATH


def lib():
    """The loaded C ABI.  Fails loudly if the CUDA library is missing: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C {_HERE}` (or __graft_entry__.build()). "
                               "avbd-demo3d_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class AvbdError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise AvbdError(f"avbd error {rc}: {lib().avbd_last_error().decode()}")
    return rc


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class World:
    """One Solver (solver.h:146-181) on one GPU; also an ensemble batch when world ids are given."""

    def __init__(self, device=0):
        self.L = lib()
        self.h = self.L.avbd_world_create(device)
        if not self.h:
            raise AvbdError("avbd_world_create failed: " + self.L.avbd_last_error().decode())
        self.device = device
        self.params = dict(dt=1 / 60, g=(0.0, -10.0, 0.0), iterations=10, alpha=0.95, beta=1e5, gamma=0.99, post=False)

    def close(self):
        if self.h:
            self.L.avbd_world_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Solver API
    def clear(self):
        _check(self.L.avbd_clear(self.h))

    def set_params(self, dt=1 / 60, g=(0, -10, 0), iterations=10, alpha=0.95, beta=1e5, gamma=0.99, post=False):
        self.params = dict(dt=dt, g=tuple(g), iterations=iterations, alpha=alpha, beta=beta, gamma=gamma, post=post)
        _check(self.L.avbd_set_params(self.h, dt, _f(g), iterations, alpha, beta, gamma, int(post)))

    def add_bodies(self, size, density, friction, pos, quat=None, lin=None, ang=None, world_ids=None):
        size = _f(size).reshape(-1, 3)
        n = len(size)
        pos = _f(pos).reshape(n, 3)
        quat = _f(np.tile([0, 0, 0, 1], (n, 1))) if quat is None else _f(quat).reshape(n, 4)
        lin = np.zeros((n, 3), np.float32) if lin is None else _f(lin).reshape(n, 3)
        ang = np.zeros((n, 3), np.float32) if ang is None else _f(ang).reshape(n, 3)
        density = _f(np.broadcast_to(np.asarray(density, np.float32), (n,)))
        friction = _f(np.broadcast_to(np.asarray(friction, np.float32), (n,)))
        wid = None
        if world_ids is not None:
            wid_arr = np.ascontiguousarray(world_ids, np.int32)
            wid = wid_arr.ctypes.data_as(C.c_void_p)
        return _check(self.L.avbd_add_bodies(self.h, n, size, density, friction, pos, quat, lin, ang, wid))

    def add_body(self, size, density, friction, pos, quat=(0, 0, 0, 1), lin=(0, 0, 0), ang=(0, 0, 0)):
        return self.add_bodies([size], [density], [friction], [pos], [quat], [lin], [ang])

    def add_joint(self, a, b, anchor_a, anchor_b=(0, 0, 0), lin_k=FLT_MAX, ang_k=FLT_MAX):
        return _check(self.L.avbd_add_joint(self.h, a, b, _f(anchor_a), _f(anchor_b), lin_k, ang_k))

    def add_joint_raw(self, a, b, r_a, r_b, rel0, lin_k=FLT_MAX, ang_k=FLT_MAX):
        """A weld with the caller's construction-time rA / rB / initial relative orientation (joint.h:17-19)."""
        return _check(self.L.avbd_add_joint_raw(self.h, a, b, _f(r_a), _f(r_b), _f(rel0), lin_k, ang_k))

    def set_force_rows(self, kind, index, lam=None, pen=None, motor=None, stiffness=None):
        """Edit the public row arrays of a joint (kind 0, 6 rows) or spring (kind 1, 1 row), solver.h:91-97."""
        n = 6 if kind == 0 else 1
        arrs = [None if a is None else _f(np.broadcast_to(np.asarray(a, np.float32), (n,))) for a in (lam, pen, motor, stiffness)]
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
        _check(self.L.avbd_set_force_rows(self.h, kind, index, *ptrs))

    def force_rows(self, kind, index):
        n = 6 if kind == 0 else 1
        out = [np.zeros(n, np.float32) for _ in range(4)]
        _check(self.L.avbd_get_force_rows(self.h, kind, index, *[a.ctypes.data_as(C.c_void_p) for a in out]))
        return dict(lam=out[0], pen=out[1], motor=out[2], stiffness=out[3])

    def add_spring(self, a, b, anchor_a, anchor_b, k, rest=-1.0):
        return _check(self.L.avbd_add_spring(self.h, a, b, _f(anchor_a), _f(anchor_b), k, rest))

    def add_ignore(self, a, b):
        return _check(self.L.avbd_add_ignore(self.h, a, b))

    def step(self, n=1, sync=True):
        _check(self.L.avbd_step(self.h, n))
        if sync:
            _check(self.L.avbd_sync(self.h))

    def sync(self):
        _check(self.L.avbd_sync(self.h))

    def step_timed(self, n):
        """n steps; returns their device time in ms (CUDA events on the world's stream)."""
        ms = C.c_float(0)
        _check(self.L.avbd_step_timed(self.h, n, C.byref(ms)))
        return ms.value

    def set_profiling(self, on=True):
        _check(self.L.avbd_set_profiling(self.h, int(on)))

    def profile(self):
        p = Profile()
        _check(self.L.avbd_get_profile(self.h, C.byref(p)))
        return p.as_dict()

    def download_state_into(self, out):
        """Download into a caller-owned float32 [n,13] array (pin it for a single async DMA)."""
        _check(self.L.avbd_download_state(self.h, out))

    @property
    def n(self):
        return self.L.avbd_num_bodies(self.h)

    def state(self):
        o = np.zeros((self.n, 13), np.float32)
        _check(self.L.avbd_download_state(self.h, o))
        return o

    def set_state(self, s):
        _check(self.L.avbd_upload_state(self.h, _f(s)))

    def set_state_range(self, first, s):
        s = _f(s).reshape(-1, 13)
        _check(self.L.avbd_upload_state_range(self.h, first, len(s), s))

    def prev_linvel(self):
        o = np.zeros((self.n, 3), np.float32)
        _check(self.L.avbd_download_prev_linvel(self.h, o))
        return o

    def set_prev_linvel(self, v):
        _check(self.L.avbd_upload_prev_linvel(self.h, _f(v)))

    def body_props(self):
        o = np.zeros((self.n, 10), np.float32)
        _check(self.L.avbd_download_body_props(self.h, o))
        return o

    def diagnostics(self):
        d = Diagnostics()
        _check(self.L.avbd_get_diagnostics(self.h, C.byref(d)))
        return dict(maxPen=d.maxPenetration, maxViol=d.maxConstraintViolation, maxLin=d.maxLinearSpeed, maxAng=d.maxAngularSpeed,
                    maxLambda=d.maxNormalImpulse, contacts=d.activeContacts, manifolds=d.activeManifolds, dynBodies=d.dynamicBodies,
                    nanEvents=d.nanEvents)

    def world_diagnostics(self):
        k = self.L.avbd_num_worlds(self.h)
        arr = (Diagnostics * k)()
        _check(self.L.avbd_get_world_diagnostics(self.h, arr, k))
        return [a.as_dict() for a in arr]

    def step_stats(self):
        s = StepStats()
        _check(self.L.avbd_get_step_stats(self.h, C.byref(s)))
        return s.as_dict()

    def manifolds_raw(self):
        m = self.L.avbd_num_manifolds(self.h)
        ints, feats, stick, flts = (np.zeros((m, 3), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 4), np.int32), np.zeros((m, 81), np.float32))
        live = _check(self.L.avbd_download_manifolds(self.h, ints, feats, stick, flts)) if m else 0
        return ints[:live], feats[:live], stick[:live], flts[:live]

    def upload_manifolds(self, ints, feats, stick, flts):
        """Inverse of manifolds_raw(): replaces the manifold set (the warm-start history Manifold::initialize matches against)."""
        ints = np.ascontiguousarray(ints, np.int32).reshape(-1, 3)
        m = len(ints)
        _check(self.L.avbd_upload_manifolds(self.h, m, ints, np.ascontiguousarray(feats, np.int32).reshape(m, 4),
                                            np.ascontiguousarray(stick, np.int32).reshape(m, 4), _f(flts).reshape(m, 81)))

    def snapshot(self):
        """Opaque blob of the whole simulation state (bodies, user forces, manifolds with lambda / penalty / anchors, params)."""
        n = self.L.avbd_snapshot_bytes(self.h)
        buf = np.zeros(n, np.uint8)
        _check(self.L.avbd_snapshot(self.h, buf.ctypes.data_as(C.c_void_p), n))
        return buf

    def restore(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        _check(self.L.avbd_restore(self.h, blob.ctypes.data_as(C.c_void_p), len(blob)))

    # -- stages
    def stage_broadphase(self):
        k = _check(self.L.avbd_stage_broadphase(self.h))
        buf = np.zeros((max(k, 1), 2), np.int32)
        m = _check(self.L.avbd_download_pairs(self.h, buf, max(k, 1)))
        return buf[:m]

    def stage(self, name, *args):
        _check(getattr(self.L, "avbd_stage_" + name)(self.h, *args))

    def pick(self, origin, direction):
        """Solver::pick: (body index or -1, body-local hit point)."""
        local = np.zeros(3, np.float32)
        i = self.L.avbd_pick(self.h, _f(origin), _f(direction), local)
        if i < -1:
            _check(i)
        return i, local

    def colours(self):
        col = np.zeros(max(self.n, 1), np.int32)
        k = C.c_int(0)
        _check(self.L.avbd_download_colours(self.h, col, C.byref(k)))
        return col[:self.n], k.value

    def stage_primal(self, alpha, want_dx=False):
        dx = np.zeros((self.n, 6), np.float32) if want_dx else None
        _check(self.L.avbd_stage_primal(self.h, alpha, None if dx is None else dx.ctypes.data_as(C.c_void_p)))
        return dx


def collide_pairs(a10, b10, device=0):
    a10, b10 = _f(a10).reshape(-1, 10), _f(b10).reshape(-1, 10)
    n = len(a10)
    counts, feats, geom = np.zeros(n, np.int32), np.zeros((n, 4), np.int32), np.zeros((n, 4, 9), np.float32)
    _check(lib().avbd_collide_pairs(device, n, a10, b10, counts, feats, geom))
    return counts, feats, geom


def solve6x6(lhs36, rhs6, device=0):
    lhs36, rhs6 = _f(lhs36).reshape(-1, 36), _f(rhs6).reshape(-1, 6)
    out = np.zeros((len(lhs36), 6), np.float32)
    _check(lib().avbd_solve6x6(device, len(lhs36), lhs36, rhs6, out))
    return out
