"""Scene presets as body arrays (host side).

Behavioural contract: alxspiker/avbd-demo3d source/scenes.h:23-212 — same bodies, same order, same float32
arithmetic (tests/test_scenes.py checks every preset bit for bit against the oracle).  A preset is a dict of
arrays ready for World.add_bodies plus the solver overrides the reference scene applies (scenes.h:93-95).
Also: the generalised Stress grid and ensemble batches of independent worlds (§8d.5).
"""
import numpy as np

F = np.float32
SCENE_NAMES = ["Empty", "Ground", "Stack", "Pyramid", "Wall", "TwoBlockDrop", "Stress1000", "Rod (WIP)", "Soft Body (WIP)"]


def _pack(rows):
    n = len(rows)
    out = dict(size=np.zeros((n, 3), F), density=np.zeros(n, F), friction=np.zeros(n, F), pos=np.zeros((n, 3), F),
               quat=np.tile(np.array([0, 0, 0, 1], F), (n, 1)), lin=np.zeros((n, 3), F), ang=np.zeros((n, 3), F), params={})
    for i, r in enumerate(rows):
        out["size"][i], out["density"][i], out["friction"][i], out["pos"][i] = r[0], r[1], r[2], r[3]
        if len(r) > 4:
            out["quat"][i] = r[4]
        if len(r) > 5:
            out["ang"][i] = r[5]
    return out


def _ground(sx=100.0, sz=100.0):
    return ((sx, 1, sz), 0.0, 0.5, (0, F(-0.5), 0))           # scenes.h:27-31


def hash_float01(x):
    """scenes.h:108-115 on uint32 arrays."""
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return (x & np.uint32(0x00FFFFFF)).astype(F) / F(16777215.0)


def stress_grid(nx, ny, nz, spacing_y=2.0, start_y=20.0, wide_ground=False, jitter_y=0.25):
    """scenes.h:86-132 with NX/NY/NZ, spacingY, startY and the vertical jitter amplitude free (0.25 upstream; a
    pre-stacked grid with spacingY=1.01 needs 0 or neighbouring layers start up to 0.24 deep inside each other).  wide_ground widens the static ground so it
    covers the grid footprint (the stock 100x1x100 slab is too small beyond ~80 columns)."""
    gx = gz = 100.0
    if wide_ground:
        gx = max(100.0, float(F(nx) * F(1.15) + F(20.0)))
        gz = max(100.0, float(F(nz) * F(1.15) + F(20.0)))
    y, z, x = np.meshgrid(np.arange(ny), np.arange(nz), np.arange(nx), indexing="ij")
    x, y, z = x.reshape(-1), y.reshape(-1), z.reshape(-1)
    with np.errstate(over="ignore"):
        seed = (x + nx * (z + nz * y) + 1).astype(np.uint32)
        jx = (hash_float01(seed * np.uint32(9781)) * F(2.0) - F(1.0)) * F(0.04)
        jz = (hash_float01(seed * np.uint32(6271)) * F(2.0) - F(1.0)) * F(0.04)
        jy = hash_float01(seed * np.uint32(3343)) * F(jitter_y)
    px = (x.astype(F) - F(nx - 1) * F(0.5)) * F(1.15) + jx
    py = F(start_y) + y.astype(F) * F(spacing_y) + jy
    pz = (z.astype(F) - F(nz - 1) * F(0.5)) * F(1.15) + jz
    n = len(x) + 1
    out = dict(size=np.ones((n, 3), F), density=np.ones(n, F), friction=np.full(n, 0.5, F), pos=np.zeros((n, 3), F),
               quat=np.tile(np.array([0, 0, 0, 1], F), (n, 1)), lin=np.zeros((n, 3), F), ang=np.zeros((n, 3), F),
               params=dict(iterations=20, beta=30000.0, gamma=0.995))
    out["size"][0] = (gx, 1, gz)
    out["density"][0] = 0.0
    out["pos"][0] = (0, -0.5, 0)
    out["pos"][1:, 0], out["pos"][1:, 1], out["pos"][1:, 2] = px, py, pz
    return out


def scene(name):
    """A scenes.h preset by its sceneNames[] entry (unknown names fall back to Empty, main.cpp:211-219)."""
    rows = []
    if name == "Ground":
        rows = [_ground()]
    elif name == "Stack":                                      # scenes.h:33-40
        rows = [_ground()] + [((1, 1, 1), 1.0, 0.5, (0, F(i) * F(1.1) + F(0.5), 0)) for i in range(10)]
    elif name == "Pyramid":                                    # scenes.h:42-53
        rows = [_ground()]
        P = 10
        for y in range(P):
            for x in range(P - y):
                xp = (F(x) - F(P - y - 1) * F(0.5)) * F(1.1)
                yp = F(y) * F(1.05) + F(0.5)
                rows.append(((1, 1, 1), 1.0, 0.5, (xp, yp, 0)))
    elif name == "Wall":                                       # scenes.h:55-73
        rows = [_ground()]
        W = H = 8
        sx, sy = F(1.03), F(0.52)
        base = F(0.5) * F(0.5)
        for i in range(H):
            for j in range(W):
                xo = F(0.0) if i % 2 == 0 else F(0.5) * sx
                rows.append(((1.0, 0.5, 0.5), 1.0, 0.4, ((F(j) - F(W - 1) * F(0.5)) * sx + xo, F(i) * sy + base, -5)))
    elif name == "TwoBlockDrop":                               # scenes.h:75-84
        half = F(0.45) * F(0.5)
        tilt = (0.0, 0.0, F(1.0) * F(np.sin(half, dtype=F)), F(np.cos(half, dtype=F)))
        rows = [_ground(), ((1, 1, 1), 1.0, 0.5, (0, 0.5, 0)), ((1, 1, 1), 1.0, 0.5, (F(0.18), F(2.2), 0), tilt, (0, 0, 1))]
    elif name == "Stress1000":
        return stress_grid(10, 10, 10)
    elif name == "Rod (WIP)":                                  # scenes.h:138-151
        rows = [((0.25, 1, 0.25), 0.0 if i == 0 else 1.0, 0.5, (0, F(10.0) - F(i) * F(1.0), 0)) for i in range(15)]
    elif name == "Soft Body (WIP)":                            # scenes.h:153-179
        rows = [_ground()]
        W = H = 10
        for i in range(W):
            for j in range(H):
                rows.append(((0.5, 0.5, 0.5), 1.0, 0.3, (F(i) * F(0.6) - F(W) * F(0.3), F(j) * F(0.6) + F(2.0), 0)))
    return _pack(rows)


def ensemble(base, worlds, jitter=0.02, first_world=0):
    """`worlds` independent copies of preset `base`, world w (global id first_world + w) shifted in x/z by a
    hash of its GLOBAL id so a world's trajectory does not depend on which GPU or batch slot it lands in."""
    n = len(base["size"])
    gid = (np.arange(worlds) + first_world).astype(np.uint32)
    with np.errstate(over="ignore"):
        dx = (hash_float01(gid * np.uint32(7919) + np.uint32(17)) * F(2.0) - F(1.0)) * F(jitter)
        dz = (hash_float01(gid * np.uint32(104729) + np.uint32(29)) * F(2.0) - F(1.0)) * F(jitter)
    out = {k: np.tile(v, (worlds,) + (1,) * (v.ndim - 1)) for k, v in base.items() if k != "params"}
    dyn = np.tile(base["density"] > 0, worlds)
    shift = np.zeros((worlds * n, 3), F)
    shift[:, 0] = np.repeat(dx, n)
    shift[:, 2] = np.repeat(dz, n)
    out["pos"] = (out["pos"] + shift * dyn[:, None].astype(F)).astype(F)
    out["world_ids"] = np.repeat(np.arange(worlds, dtype=np.int32), n)
    out["params"] = dict(base["params"])
    return out


def load(world, preset):
    """Adds a preset's bodies to a World and applies its solver overrides (which persist, scenes.h:93-95)."""
    p = dict(world.params)
    p.update(preset["params"])
    world.set_params(**p)
    if len(preset["size"]):
        world.add_bodies(preset["size"], preset["density"], preset["friction"], preset["pos"], preset["quat"], preset["lin"],
                         preset["ang"], preset.get("world_ids"))
    return world
