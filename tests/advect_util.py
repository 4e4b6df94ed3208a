"""Shared by the advection tests: the inputs ref_driver's RefAdvect modes build from a scene (oracle/ref_driver.cpp), as dense numpy grids."""
import dataclasses

import numpy as np

from shiokaze_b200 import scenes


def swirl(sc, cells: float, seed: int = 11):
    """The scene with its face VALUES replaced (masks kept) by a three-dimensional vortex plus counter-based roughness, scaled so that one time step carries
    the field over up to `cells` cells: the projection scenes' own velocities (a uniform fall, a blob) make every advection scheme agree."""
    vel = []
    amp = cells * sc.dx / sc.dt
    for d in range(3):
        nz, ny, nx = sc.vel[d].shape
        z, y, x = np.meshgrid((np.arange(nz) + 0.5 * (d != 2)) / sc.nz, (np.arange(ny) + 0.5 * (d != 1)) / sc.ny, (np.arange(nx) + 0.5 * (d != 0)) / sc.nx, indexing="ij")
        two_pi = 2.0 * np.pi
        field = [np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(np.pi * z), -np.cos(two_pi * x) * np.sin(two_pi * y) * np.cos(np.pi * z) - 0.4,
                 0.6 * np.sin(two_pi * z) * np.cos(two_pi * x) + 0.3 * np.sin(two_pi * y)][d]
        rough = 0.25 * (2.0 * scenes.hash_noise(seed, d, sc.vel[d].shape, 0) - 1.0)
        v = (amp * (field + rough)).astype(np.float32)
        v[(scenes.hash_noise(seed + 1, d, sc.vel[d].shape, 0) < 0.02)] = 0.0   # some exact zeros: vec::empty() needs all three components zero, these alone do not stop a face
        vel.append(np.where(sc.vel_active[d] != 0, v, np.float32(0)).astype(np.float32))
    still = scenes.hash_noise(seed + 2, 0, (sc.nz, sc.ny, sc.nx), 0) < 0.03   # ... and cells whose six faces all rest: the faces between two of them take the `still` branch
    for d in range(3):
        pad = [(0, 0)] * 3
        pad[2 - d] = (1, 1)
        s = np.pad(still, pad, constant_values=False)
        both = (s[:, :, 1:] | s[:, :, :-1]) if d == 0 else ((s[:, 1:, :] | s[:, :-1, :]) if d == 1 else (s[1:, :, :] | s[:-1, :, :]))
        vel[d][both] = 0.0
    return dataclasses.replace(sc, vel=vel)


def advect_scenes():
    """name -> scene (level sets and activity of the projection scenes, velocities of swirl())."""
    return {
        "dambreak_solid": lambda: swirl(scenes.dambreak(32, True), 2.5),
        "flip": lambda: swirl(scenes.flip_splash(40), 3.5, seed=5),
        "smoke": lambda: swirl(scenes.smoke_plume(24), 1.7, seed=7),              # all faces active, constant fluid grid
        "blobs": lambda: swirl(scenes.random_blobs(20, 14, 18, seed=3), 4.0, seed=9),
    }


def fluid_active(sc):
    """Activity of the liquid level set as ref_driver builds it: |raw| < band (none for a smoke scene's constant grid)."""
    if sc.fluid_raw is None:
        return np.zeros(sc.fluid.shape, dtype=np.uint8)
    return (np.abs(sc.fluid_raw.astype(np.float64)) < sc.band).astype(np.uint8)


def density_of(sc):
    """RefAdvect=density: a sparse cell grid (background 0), active where the y-face of the same index is, with that face's value."""
    act = sc.vel_active[1][:, :sc.ny, :].copy()
    val = np.where(act != 0, sc.vel[1][:, :sc.ny, :], np.float32(0)).astype(np.float32)
    return val, act


# flag combinations of the committed fixtures (tests/golden/advect_<scene>.npz, written by tests/golden/make_golden_advect.py from the reference build)
GOLDEN_FLAGS = [{}, {"MacCormack": "No"}, {"WENO": "Yes"}, {"WENO": "Yes", "MacCormack": "No"}, {"TrimNarrowBand": 3}]


def flag_key(flags):
    return "-".join(f"{k}{v}" for k, v in flags.items()) or "default"


def load_golden(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"advect_{name}.npz")
    return np.load(path)

