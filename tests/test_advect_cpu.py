"""The advection kernels' source, compiled once more for the HOST (csrc/advect.cu -DSHKZ_B200_ADVECT_HOSTCHECK -> oracle/_build/libadvect_hostcheck.so: plain
loops over the same __host__ __device__ bodies), against the unmodified reference's macadvection3 module run through its own loader (oracle/ref_driver
RefAdvect=...). Runs where the reference build exists (this container); it pins the restatement bit for bit before any GPU time is spent.
tests/test_gpu_advect.py holds the product — the CUDA kernels behind the C-ABI — to the same bar on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from advect_util import GOLDEN_FLAGS, advect_scenes, density_of, flag_key, fluid_active, load_golden
from oracle import refio
from shiokaze_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libadvect_hostcheck.so")
SCENES = advect_scenes()
FLAGS = [{}, {"MacCormack": "No"}, {"WENO": "Yes"}, {"WENO": "Yes", "MacCormack": "No"}, {"TrimNarrowBand": 3}]


def hostcheck(need_reference=True):
    if need_reference and not refio.ref_available("f32"):
        pytest.skip("oracle/_ref (the reference build) is not here")
    if not os.path.isfile("/usr/local/cuda/bin/nvcc") and not os.path.isfile(LIB):
        pytest.skip("no nvcc to build the host harness with")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)   # (no-op when __graft_entry__.build() already made it)
    return C.CDLL(LIB)


def params(flags, background=0.0):
    p = capi.AdvectParams()
    p.struct_size = C.sizeof(p)
    p.maccormack = 0 if flags.get("MacCormack") == "No" else 1
    p.weno = 1 if flags.get("WENO") == "Yes" else 0
    p.trim_narrowband = int(flags.get("TrimNarrowBand", 1))
    p.scalar_background = background
    return p


def ptrs(arrays):
    return (C.c_void_p * 3)(*[a.ctypes.data for a in arrays])


@pytest.mark.parametrize("flags", FLAGS, ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()) or "default")
@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_restatement_of_advect_vector_equals_the_reference(name, flags, real):
    L = hostcheck()
    if real == "f64" and (flags or name not in ("dambreak_solid", "smoke")):
        pytest.skip("Real=double: default flags on two scenes")
    sc = SCENES[name]()
    ref = refio.run_reference(sc, real, flags=flags, advect="vector")
    dt = np.float64 if real == "f64" else np.float32
    u = [np.ascontiguousarray(v, dtype=dt).copy() for v in sc.vel]
    act = [np.ascontiguousarray(a, dtype=np.uint8) for a in sc.vel_active]
    fluid = np.ascontiguousarray(sc.fluid, dtype=dt)
    p = params(flags)
    L.shkz_b200_hostcheck_advect_vector(sc.nx, sc.ny, sc.nz, C.c_double(sc.dx), 1 if real == "f64" else 0, C.c_double(sc.dt), ptrs(u), ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
    moved = 0
    for d in range(3):
        assert np.array_equal(ref.vel_active[d], act[d])
        on = act[d] != 0
        diff = u[d].astype(np.float64)[on] != ref.vel[d][on]
        assert not diff.any(), (name, flags, d, int(diff.sum()), int(on.sum()), float(np.abs(u[d].astype(np.float64)[on] - ref.vel[d][on]).max()))
        moved += int((u[d][on] != sc.vel[d][on].astype(dt)).sum())
    assert moved > 0


@pytest.mark.parametrize("flags", FLAGS, ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()) or "default")
@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("mode", ["density", "levelset"])
def test_restatement_of_advect_scalar_equals_the_reference(name, flags, mode):
    L = hostcheck()
    sc = SCENES[name]()
    if mode == "levelset" and sc.fluid_raw is None:
        pytest.skip("no liquid level set in a smoke scene")
    ref = refio.run_reference(sc, "f32", flags=flags, advect=mode)
    if mode == "density":
        q, qa = density_of(sc)
        background = 0.0
    else:
        q, qa = sc.fluid.astype(np.float32).copy(), fluid_active(sc)
        background = float(np.float32(sc.band))
    q = np.ascontiguousarray(q).copy()
    q_in = q.copy()
    vel = [np.ascontiguousarray(v, dtype=np.float32) for v in sc.vel]
    act = [np.ascontiguousarray(a, dtype=np.uint8) for a in sc.vel_active]
    fluid = np.ascontiguousarray(sc.fluid, dtype=np.float32)
    p = params(flags, background)
    L.shkz_b200_hostcheck_advect_scalar(sc.nx, sc.ny, sc.nz, C.c_double(sc.dx), 0, C.c_double(sc.dt), C.c_void_p(q.ctypes.data), C.c_void_p(qa.ctypes.data), ptrs(vel), ptrs(act),
                                        C.c_void_p(fluid.ctypes.data), C.byref(p))
    assert np.array_equal(ref.pressure_active, qa)
    on = qa != 0
    diff = q.astype(np.float64)[on] != ref.pressure[on]
    assert not diff.any(), (name, flags, mode, int(diff.sum()), int(on.sum()), float(np.abs(q.astype(np.float64)[on] - ref.pressure[on]).max()))
    assert (q[on] != q_in[on]).any()


@pytest.mark.parametrize("name", list(SCENES))
def test_host_build_equals_the_committed_goldens(name):
    """tests/golden/advect_<scene>.npz (outputs of the reference build, tests/golden/make_golden_advect.py) against the host build of the kernel source: every
    flag combination, velocity and both scalars. Needs no reference: this is what pins the restatement on a box that has only the repository."""
    L = hostcheck(need_reference=False)
    sc = SCENES[name]()
    G = load_golden(name)
    vel = [np.ascontiguousarray(v, dtype=np.float32) for v in sc.vel]
    act = [np.ascontiguousarray(a, dtype=np.uint8) for a in sc.vel_active]
    fluid = np.ascontiguousarray(sc.fluid, dtype=np.float32)
    for flags in GOLDEN_FLAGS:
        key = flag_key(flags)
        u = [v.copy() for v in vel]
        p = params(flags)
        L.shkz_b200_hostcheck_advect_vector(sc.nx, sc.ny, sc.nz, C.c_double(sc.dx), 0, C.c_double(sc.dt), ptrs(u), ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
        for d in range(3):
            assert np.array_equal(u[d][act[d] != 0], G[f"{key}/vector{d}"]), (name, key, d)
        for mode in ("density", "levelset"):
            if f"{key}/{mode}" not in G:
                continue
            q, qa = density_of(sc) if mode == "density" else (sc.fluid.astype(np.float32).copy(), fluid_active(sc))
            q = np.ascontiguousarray(q).copy()
            p = params(flags, 0.0 if mode == "density" else float(np.float32(sc.band)))
            L.shkz_b200_hostcheck_advect_scalar(sc.nx, sc.ny, sc.nz, C.c_double(sc.dx), 0, C.c_double(sc.dt), C.c_void_p(q.ctypes.data), C.c_void_p(qa.ctypes.data), ptrs(vel),
                                                ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
            assert np.array_equal(q[qa != 0], G[f"{key}/{mode}"]), (name, key, mode)


def test_committed_goldens_are_what_the_reference_build_gives():
    if not refio.ref_available("f32"):
        pytest.skip("oracle/_ref (the reference build) is not here")
    sc = SCENES["blobs"]()
    G = load_golden("blobs")
    for flags in ({}, {"WENO": "Yes"}):
        r = refio.run_reference(sc, "f32", flags=flags, advect="vector")
        for d in range(3):
            assert np.array_equal(r.vel[d][sc.vel_active[d] != 0].astype(np.float32), G[f"{flag_key(flags)}/vector{d}"])


@pytest.mark.parametrize("shape", [(2, 3, 2), (5, 2, 3), (9, 4, 17)])
def test_restatement_on_tiny_and_ragged_grids(shape):
    """The smallest grids the reference's interpolation is defined on (every stencil clamped at a wall), ragged activity, displacements beyond the grid:
    host build of the kernel source against the reference module run live."""
    L = hostcheck()
    from shiokaze_b200 import scenes
    nx, ny, nz = shape
    sc = scenes.random_blobs(nx, ny, nz, seed=2, with_solid=False)
    rng = np.random.default_rng(5)
    vel = [np.where(a != 0, rng.standard_normal(v.shape) * 3.0, 0).astype(np.float32) for v, a in zip(sc.vel, sc.vel_active)]
    import dataclasses
    sc = dataclasses.replace(sc, vel=vel, dt=0.37)
    for flags in ({}, {"WENO": "Yes"}):
        ref = refio.run_reference(sc, "f32", flags=flags, advect="vector")
        u = [v.copy() for v in vel]
        act = [np.ascontiguousarray(a, dtype=np.uint8) for a in sc.vel_active]
        fluid = np.ascontiguousarray(sc.fluid, dtype=np.float32)
        p = params(flags)
        L.shkz_b200_hostcheck_advect_vector(nx, ny, nz, C.c_double(sc.dx), 0, C.c_double(sc.dt), ptrs(u), ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
        for d in range(3):
            on = act[d] != 0
            assert np.array_equal(u[d].astype(np.float64)[on], ref.vel[d][on]), (shape, flags, d)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_on_random_grids_masks_flags_and_time_steps(seed):
    """Random extents (2 .. 21 per axis), activity masks unrelated to the level set, values with exact zeros, time steps of either sign that carry the field
    from a hundredth of a cell to far beyond the grid, every flag, Real = float and double, all three calls: host build of the kernel source against the
    reference module run live."""
    import dataclasses
    from shiokaze_b200 import scenes
    L = hostcheck()
    rng = np.random.default_rng(seed)
    for trial in range(6):
        nx, ny, nz = (int(rng.integers(2, 22)) for _ in range(3))
        sc = scenes.random_blobs(nx, ny, nz, seed=int(rng.integers(1, 1000)), with_solid=bool(rng.integers(0, 2)))
        act = [(rng.random(a.shape) < rng.uniform(0.2, 1.0)).astype(np.uint8) for a in sc.vel_active]
        vel = [np.where(a != 0, rng.standard_normal(a.shape) * rng.uniform(0.1, 5.0), 0).astype(np.float32) for a in act]
        for v in vel:
            v[rng.random(v.shape) < 0.05] = 0.0
        dt = float(rng.choice([-1.0, 1.0]) * rng.uniform(0.01, 2.0))
        sc = dataclasses.replace(sc, vel=vel, vel_active=act, dt=dt)
        flags = {"TrimNarrowBand": int(rng.integers(0, 4))}
        if rng.random() < 0.3:
            flags["MacCormack"] = "No"
        if rng.random() < 0.3:
            flags["WENO"] = "Yes"
        real = "f64" if rng.random() < 0.3 else "f32"
        if not refio.ref_available(real):
            continue
        npdt = np.float64 if real == "f64" else np.float32
        fa = fluid_active(sc)
        fluid = np.ascontiguousarray(sc.fluid, dtype=npdt)
        band = float(np.float32(sc.band))
        if real == "f64":   # Real=double: fill / background of the level set are +-band in double, not the float32 values of the scene object
            fluid = np.where(fa != 0, sc.fluid_raw.astype(np.float64), np.where(sc.fluid < 0, -sc.band, sc.band))
            band = float(sc.band)
        velr = [np.ascontiguousarray(v, dtype=npdt) for v in vel]
        for mode in ("vector", "density", "levelset"):
            ref = refio.run_reference(sc, real, flags=flags, advect=mode)
            p = params(flags, band if mode == "levelset" else 0.0)
            what = (seed, trial, (nx, ny, nz), mode, flags, real, dt)
            if mode == "vector":
                u = [v.copy() for v in velr]
                L.shkz_b200_hostcheck_advect_vector(nx, ny, nz, C.c_double(sc.dx), int(real == "f64"), C.c_double(dt), ptrs(u), ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
                for d in range(3):
                    assert np.array_equal(u[d].astype(np.float64)[act[d] != 0], ref.vel[d][act[d] != 0]), what
            else:
                q, qa = density_of(sc) if mode == "density" else (fluid.copy(), fa)
                q, qa = np.ascontiguousarray(q, dtype=npdt).copy(), np.ascontiguousarray(qa)
                L.shkz_b200_hostcheck_advect_scalar(nx, ny, nz, C.c_double(sc.dx), int(real == "f64"), C.c_double(dt), C.c_void_p(q.ctypes.data), C.c_void_p(qa.ctypes.data), ptrs(velr),
                                                    ptrs(act), C.c_void_p(fluid.ctypes.data), C.byref(p))
                assert np.array_equal(ref.pressure_active != 0, qa != 0), what
                assert np.array_equal(q.astype(np.float64)[qa != 0], ref.pressure[qa != 0]), what

