"""Synthetic projection inputs, as dense arrays.

Every scene is an analytic formula evaluated with numpy, so the same input can be handed to the CUDA
path, to the dense C oracle and (through oracle/refio.py) to the unmodified reference build.
Scene geometry follows the reference's own scene files (formulas only):

  dam-break    /root/reference/resources/liquid/dambreak3.cpp:32-49
  water drop   /root/reference/resources/liquid/waterdrop3.cpp:33-61   (Container=Yes)
  plume source /root/reference/resources/smoke/plume3.cpp:37-49
  sphere test  /root/reference/src/examples/accuracytest3-example.cpp:74-121

Grid conventions are the reference's (include/shiokaze/math/shape.h:883-888): index = i + w*(j + h*k),
numpy shape (nz, ny, nx) with x fastest; cell centres dx*(i+.5), nodes dx*i, faces offset by half a
cell in the two tangential directions. Level sets are stored the way the simulators store them
(src/utility/macutility3.cpp:336-374): |phi| < band keeps its value, everything else reads +-band
(background / flood fill); `*_raw` keeps the unclamped float32 samples for the reference driver.

A scene may be generated for a z-slab only (`zrange=(k0,k1)`): cells k0..k1-1, z-faces and nodes
k0..k1. Hash noise is counter based (splitmix64 of the global face index), so any partition of the
grid sees the same values.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import os

import numpy as np

SQRT3 = float(np.sqrt(3.0))


@dataclass
class Scene:
    name: str
    nx: int
    ny: int
    nz: int                      # global z extent
    dx: float
    dt: float
    zrange: Tuple[int, int]      # cells [k0,k1) held by this object
    vel: list                    # 3 float32 arrays: (nzl,ny,nx+1), (nzl,ny+1,nx), (nzl+1,ny,nx)
    vel_active: list             # 3 uint8 arrays, same shapes
    fluid: np.ndarray            # float32 (nzl,ny,nx) dense as array3::operator() reads it
    fluid_levelset: bool         # reference's levelset_exist(fluid): any ACTIVE value < 0
    solid: Optional[np.ndarray]  # float32 nodal (nzl+1,ny+1,nx+1) as read, or None when no solid level set
    band: float                  # half band width of the level sets (absolute units)
    fluid_raw: Optional[np.ndarray] = None
    solid_raw: Optional[np.ndarray] = None
    solid_mode: int = 0          # refio: 0 nodal/empty, 1 nodal narrow band, 2 cell-shaped constant +1
    surface_tension: float = 0.0
    meta: dict = field(default_factory=dict)

    @property
    def shape(self):
        return (self.nx, self.ny, self.nz)

    @property
    def nzl(self):
        return self.zrange[1] - self.zrange[0]


# ----------------------------------------------------------------------------------------------
def _axes(n, dx, offset, lo=0, hi=None):
    hi = n if hi is None else hi
    return (np.arange(lo, hi, dtype=np.float64) + offset) * dx


def _grid(nx, ny, nz, dx, kind, zrange, dim=None):
    """Broadcastable coordinate triple (x,y,z) for cell centres / nodes / faces of `dim`."""
    k0, k1 = zrange
    if kind == "cell":
        x, y, z = _axes(nx, dx, .5), _axes(ny, dx, .5), _axes(nz, dx, .5, k0, k1)
    elif kind == "node":
        x, y, z = _axes(nx + 1, dx, 0.), _axes(ny + 1, dx, 0.), _axes(nz + 1, dx, 0., k0, k1 + 1)
    elif kind == "face":
        x = _axes(nx + 1, dx, 0.) if dim == 0 else _axes(nx, dx, .5)
        y = _axes(ny + 1, dx, 0.) if dim == 1 else _axes(ny, dx, .5)
        z = _axes(nz + 1, dx, 0., k0, k1 + 1) if dim == 2 else _axes(nz, dx, .5, k0, k1)
    else:
        raise ValueError(kind)
    return x[None, None, :], y[None, :, None], z[:, None, None]


def clamp_levelset(raw32: np.ndarray, band: float) -> np.ndarray:
    """Dense read of a narrow-band level set: active value, else flood-fill -band / background +band
    (array3.h:242-247, 796-801). `band` is stored as Real=float by the reference."""
    b32 = np.float32(band)
    active = np.abs(raw32.astype(np.float64)) < band
    return np.where(active, raw32, np.where(raw32 < 0, -b32, b32)).astype(np.float32)


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_noise(seed: int, dim: int, shape_zyx, k0: int) -> np.ndarray:
    """h in [0,1): splitmix64(seed ^ (dim<<60 | k<<40 | j<<20 | i)) >> 11 * 2^-53, global indices."""
    nzl, nyl, nxl = shape_zyx
    i = np.arange(nxl, dtype=np.uint64)[None, None, :]
    j = np.arange(nyl, dtype=np.uint64)[None, :, None]
    out = np.empty((nzl, nyl, nxl), dtype=np.float64)

    def planes(a, b):
        k = (np.arange(a, b, dtype=np.uint64) + np.uint64(k0))[:, None, None]
        key = (np.uint64(dim) << np.uint64(60)) | (k << np.uint64(40)) | (j << np.uint64(20)) | i
        h = splitmix64(np.uint64(seed) ^ key)
        out[a:b] = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    step = 8
    if nzl * nyl * nxl < (1 << 22):
        planes(0, nzl)
    else:  # counter-based, so z-chunks are independent: numpy releases the GIL inside its loops
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda a: planes(a, min(a + step, nzl)), range(0, nzl, step)))
    return out


def _face_shapes(nx, ny, nzl):
    return [(nzl, ny, nx + 1), (nzl, ny + 1, nx), (nzl + 1, ny, nx)]


def _dam(x, y, z):
    width, height, level, depth = 0.232, 0.432, 0.095, 0.2532
    v = np.maximum(np.maximum(x - width, y - height), np.abs(z - 0.5) - depth)
    return np.minimum(v, y - level)


def _finish(name, n3, dx, dt, zr, vel, act, fluid_raw, solid_raw, band, **kw):
    nx, ny, nz = n3
    fluid = clamp_levelset(fluid_raw, band)
    levelset = bool(((np.abs(fluid_raw.astype(np.float64)) < band) & (fluid_raw < 0)).any())
    solid = None
    solid_mode = 0
    if solid_raw is not None:
        s = clamp_levelset(solid_raw, band)
        has_solid = bool(((np.abs(solid_raw.astype(np.float64)) < band) & (solid_raw < 0)).any())
        solid_mode = 1
        solid = s if has_solid else None
    return Scene(name, nx, ny, nz, dx, dt, zr, [v.astype(np.float32) for v in vel],
                 [a.astype(np.uint8) for a in act], fluid, levelset, solid, band,
                 fluid_raw=fluid_raw, solid_raw=solid_raw, solid_mode=solid_mode, **kw)


# ----------------------------------------------------------------------------------------------
def dambreak(n: int, solid_obstacle: bool = False, zrange=None, dt=1.0 / 120.0) -> Scene:
    """Configs 1 and 3: dam-break column + pool; optional spherical obstacle (nodal solid SDF)."""
    nx = ny = nz = n
    dx = 1.0 / n
    zr = (0, nz) if zrange is None else tuple(zrange)
    band = SQRT3 * dx
    fluid_raw = _dam(*_grid(nx, ny, nz, dx, "cell", zr))
    fluid_raw = np.broadcast_to(fluid_raw, (zr[1] - zr[0], ny, nx)).astype(np.float64)
    solid_raw = None
    if solid_obstacle:
        def sphere(x, y, z):
            return np.sqrt((x - 0.6) ** 2 + (y - 0.12) ** 2 + (z - 0.5) ** 2) - 0.1
        solid_raw = sphere(*_grid(nx, ny, nz, dx, "node", zr)).astype(np.float32)
        # macutility3.cpp:358-364: fluid := max(fluid, -(solid(cell centre) + dx))
        fluid_raw = np.maximum(fluid_raw, -(sphere(*_grid(nx, ny, nz, dx, "cell", zr)) + dx))
    fluid_raw = fluid_raw.astype(np.float32)
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, zr[1] - zr[0])):
        f = np.broadcast_to(_dam(*_grid(nx, ny, nz, dx, "face", zr, dim)), shp)
        a = f < 2.0 * dx
        v = np.where(a, (-9.8 * dt) if dim == 1 else 0.0, 0.0)
        vel.append(v)
        act.append(a)
    return _finish("dambreak_solid" if solid_obstacle else "dambreak", (nx, ny, nz), dx, dt, zr, vel, act,
                   fluid_raw, solid_raw, band)


def smoke_plume(n: int, zrange=None, dt=1.0 / 120.0) -> Scene:
    """Config 2: all-fluid Neumann box, buoyant blob + plume source; fluid = constant -1, no actives."""
    nx = ny = nz = n
    dx = 1.0 / n
    zr = (0, nz) if zrange is None else tuple(zrange)
    nzl = zr[1] - zr[0]
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, nzl)):
        x, y, z = _grid(nx, ny, nz, dx, "face", zr, dim)
        v = np.zeros(shp, dtype=np.float64)
        if dim == 1:
            d = np.sqrt((x - 0.5) ** 2 + (y - 0.2) ** 2 + (z - 0.5) ** 2)
            v = v + 2.0 * dt * np.maximum(0.0, 10.0 * (0.1 - d) / 0.1)
        if dim == 0:
            dist = np.sqrt((x - 0.15) ** 2 + (y - 0.15) ** 2 + (z - 0.5) ** 2)
            src = 2.0 * dt * np.minimum(10.0, np.maximum(0.0, 10.0 * (0.075 - dist) / 0.075))
            v = v + np.where(dist < 0.1, src, 0.0)
        vel.append(np.broadcast_to(v, shp))
        act.append(np.ones(shp, dtype=np.uint8))
    fluid = np.full((nzl, ny, nx), -1.0, dtype=np.float32)
    return Scene("smoke_plume", nx, ny, nz, dx, dt, zr, [v.astype(np.float32) for v in vel], act, fluid, False, None,
                 SQRT3 * dx, fluid_raw=None, solid_raw=None, solid_mode=0, meta={"fluid_mode": 0})


def flip_splash(n: int, zrange=None, dt=1.0 / 120.0, seed=20260101) -> Scene:
    """Config 4: water drop over a pool inside a spherical-shell container, FLIP-like face noise."""
    nx = ny = nz = n
    dx = 1.0 / n
    zr = (0, nz) if zrange is None else tuple(zrange)
    band = SQRT3 * dx
    cx, cy, cz, rad, level = 0.5, 0.37, 0.5, 0.075, 0.245

    def fluid_f(x, y, z):
        return np.minimum(y - level, np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) - rad)

    def solid_f(x, y, z):
        return 0.5 - 0.03 - np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)

    cell = _grid(nx, ny, nz, dx, "cell", zr)
    fluid_raw = np.maximum(fluid_f(*cell), -(solid_f(*cell) + dx)).astype(np.float32)
    solid_raw = solid_f(*_grid(nx, ny, nz, dx, "node", zr)).astype(np.float32)
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, zr[1] - zr[0])):
        x, y, z = _grid(nx, ny, nz, dx, "face", zr, dim)
        a = np.broadcast_to(fluid_f(x, y, z) < 2.0 * dx, shp)
        v = np.zeros(shp, dtype=np.float64)
        if dim == 1:
            drop = np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) - rad < 0.0
            v = v + np.where(drop, -2.0, 0.0)
        v = v + 0.05 * (2.0 * hash_noise(seed, dim, shp, zr[0]) - 1.0)
        vel.append(np.where(a, v, 0.0))
        act.append(a)
    return _finish("flip_splash", (nx, ny, nz), dx, dt, zr, vel, act, fluid_raw, solid_raw, band)


def liquid_box(n: int, zrange=None, dt=1.0 / 120.0, seed=7, amplitude=0.1) -> Scene:
    """Config 5: half-filled box with an off-grid planar surface and hash-noise velocity."""
    nx = ny = nz = n
    dx = 1.0 / n
    zr = (0, nz) if zrange is None else tuple(zrange)
    band = SQRT3 * dx
    nzl = zr[1] - zr[0]

    def fluid_f(x, y, z):
        return y - (0.5 + 0.25 * dx) + 0.0 * x + 0.0 * z

    fluid_raw = np.broadcast_to(fluid_f(*_grid(nx, ny, nz, dx, "cell", zr)), (nzl, ny, nx)).astype(np.float32)
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, nzl)):
        a = np.broadcast_to(fluid_f(*_grid(nx, ny, nz, dx, "face", zr, dim)) < 2.0 * dx, shp)
        v = amplitude * (2.0 * hash_noise(seed, dim, shp, zr[0]) - 1.0)
        vel.append(np.where(a, v, 0.0))
        act.append(a)
    return _finish("liquid_box", (nx, ny, nz), dx, dt, zr, vel, act, fluid_raw, None, band)


def accuracy_sphere(n: int, q: int = 0, trial: int = 4, r0: float = 0.4) -> Scene:
    """The reference's known-answer test: liquid sphere, u = grad |x-c|^2, dt = 1, band = 2dx,
    solid = cell-shaped constant +1 (no solid level set). Exact pressure: |x-c|^2 - r^2."""
    nx = ny = nz = n
    dx = 1.0 / n
    zr = (0, nz)
    band = 2.0 * dx
    r = r0 + SQRT3 * dx / trial * q
    x, y, z = _grid(nx, ny, nz, dx, "cell", zr)
    fluid_raw = (np.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - r).astype(np.float32)
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, nz)):
        p = _grid(nx, ny, nz, dx, "face", zr, dim)
        vel.append(np.broadcast_to(-2.0 * (0.5 - p[dim]), shp))
        act.append(np.ones(shp, dtype=np.uint8))
    s = _finish("accuracy_sphere", (nx, ny, nz), dx, 1.0, zr, vel, act, fluid_raw, None, band,
                meta={"r": r, "q": q})
    s.solid_mode = 2
    return s


def random_blobs(nx: int, ny: int, nz: int, seed: int = 1, with_solid: bool = True, dt=1.0 / 120.0) -> Scene:
    """Non-cubic stress scene for parity tests: a few liquid blobs + tilted solid plane, noisy velocity,
    ragged face activity. Not one of the benchmark configurations."""
    rng = np.random.default_rng(seed)
    dx = 1.0 / max(nx, ny, nz)
    zr = (0, nz)
    band = SQRT3 * dx
    lx, ly, lz = nx * dx, ny * dx, nz * dx
    centres = rng.uniform(0.2, 0.8, size=(4, 3)) * np.array([lx, ly, lz])
    radii = rng.uniform(0.12, 0.25, size=4) * min(lx, ly, lz)

    def fluid_f(x, y, z):
        v = None
        for c, r in zip(centres, radii):
            d = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r
            v = d if v is None else np.minimum(v, d)
        return np.minimum(v, y - 0.3 * ly)

    nrm = np.array([0.3, 1.0, 0.2])
    nrm /= np.linalg.norm(nrm)

    def solid_f(x, y, z):
        return (x * nrm[0] + y * nrm[1] + z * nrm[2]) - 0.17 * ly

    cell = _grid(nx, ny, nz, dx, "cell", zr)
    fluid_raw = fluid_f(*cell)
    solid_raw = None
    if with_solid:
        fluid_raw = np.maximum(fluid_raw, -(solid_f(*cell) + dx))
        solid_raw = np.broadcast_to(solid_f(*_grid(nx, ny, nz, dx, "node", zr)), (nz + 1, ny + 1, nx + 1)).astype(np.float32)
    fluid_raw = np.broadcast_to(fluid_raw, (nz, ny, nx)).astype(np.float32)
    vel, act = [], []
    for dim, shp in enumerate(_face_shapes(nx, ny, nz)):
        a = np.broadcast_to(fluid_f(*_grid(nx, ny, nz, dx, "face", zr, dim)) < 2.0 * dx, shp)
        a = a & (hash_noise(seed + 99, dim, shp, 0) > 0.02)   # ragged: drop 2 % of the faces
        v = 0.5 * (2.0 * hash_noise(seed, dim, shp, 0) - 1.0)
        vel.append(np.where(a, v, 0.0))
        act.append(a)
    return _finish("random_blobs", (nx, ny, nz), dx, dt, zr, vel, act, fluid_raw, solid_raw, band)


BENCH_SCENES = {
    "dambreak": lambda n, **kw: dambreak(n, False, **kw),
    "smoke_plume": smoke_plume,
    "dambreak_solid": lambda n, **kw: dambreak(n, True, **kw),
    "flip_splash": flip_splash,
    "liquid_box": liquid_box,
}


def standalone(slab: Scene) -> Scene:
    """Re-label a z-slab as a whole grid of its own (nx x ny x nzl, same dx): the slab's top and bottom
    planes become walls. Used to cut bounded CPU-baseline samples out of a large workload."""
    import copy
    s = copy.copy(slab)
    s.nz = slab.nzl
    s.zrange = (0, slab.nzl)
    s.name = slab.name + "_slab"
    return s
