"""The step before the projection on the device (first part; csrc/advect.cu): advect_vector / advect_scalar through the C-ABI against the
reference's own macadvection3 module (src/advection/macadvection3.cpp), driven through the reference's loader by oracle/ref_driver (RefAdvect=...).
The floating-point expressions are restated as written and compiled without contraction: the bar is bit-exact, on every flag combination of the module
(MacCormack, WENO, TrimNarrowBand), Real=float and Real=double. Size-independent properties cover the sizes the reference does not run in seconds."""
import os

import numpy as np
import pytest
from scipy.ndimage import minimum_filter

from advect_util import GOLDEN_FLAGS, advect_scenes, density_of, flag_key, fluid_active, load_golden, swirl
from oracle import refio
from shiokaze_b200 import MacAdvection3, scenes

pytestmark = pytest.mark.gpu

SCENES = advect_scenes()
FLAGS = [{}, {"MacCormack": "No"}, {"WENO": "Yes"}, {"WENO": "Yes", "MacCormack": "No"}, {"TrimNarrowBand": 3}]
IDS = lambda f: "-".join(f"{k}{v}" for k, v in f.items()) or "default"  # noqa: E731


def need_ref(real="f32"):
    if not refio.ref_available(real):
        pytest.skip("oracle/_ref (the reference build) was not shipped to this box")


def same_bits(ours, ref64, on, what):
    diff = ours.astype(np.float64)[on] != ref64[on]
    assert not diff.any(), (what, int(diff.sum()), int(on.sum()), float(np.abs(ours.astype(np.float64)[on] - ref64[on]).max()))


@pytest.mark.parametrize("name", list(SCENES))
def test_against_the_committed_goldens(cuda_device, name):
    """tests/golden/advect_<scene>.npz — outputs of the reference build, committed (tests/golden/make_golden_advect.py) — against the CUDA kernels through the
    C-ABI: every flag combination, velocity and both scalars, bit for bit. Needs nothing but the repository."""
    sc = SCENES[name]()
    G = load_golden(name)
    for flags in GOLDEN_FLAGS:
        key = flag_key(flags)
        A = MacAdvection3(sc.shape, sc.dx, **flags)
        out = A.advect_vector(sc.vel, sc.vel_active, sc.fluid, sc.dt)
        for d in range(3):
            assert np.array_equal(out[d][sc.vel_active[d] != 0], G[f"{key}/vector{d}"]), (name, key, d)
        q, qa = density_of(sc)
        assert np.array_equal(A.advect_scalar(q, qa, sc.vel, sc.vel_active, sc.fluid, sc.dt)[qa != 0], G[f"{key}/density"]), (name, key)
        if f"{key}/levelset" in G:
            qa = fluid_active(sc)
            q = A.advect_scalar(sc.fluid, qa, sc.vel, sc.vel_active, sc.fluid, sc.dt, background=float(np.float32(sc.band)))
            assert np.array_equal(q[qa != 0], G[f"{key}/levelset"]), (name, key)
        A.close()


@pytest.mark.parametrize("flags", FLAGS, ids=IDS)
@pytest.mark.parametrize("name", list(SCENES))
def test_advect_vector_equals_the_reference_bit_for_bit(cuda_device, name, flags):
    need_ref()
    sc = SCENES[name]()
    ref = refio.run_reference(sc, "f32", flags=flags, advect="vector")
    A = MacAdvection3(sc.shape, sc.dx, **flags)
    out = A.advect_vector(sc.vel, sc.vel_active, sc.fluid, sc.dt)
    assert A.last_stats["kernel_launches"] >= 2
    A.close()
    moved = 0
    for d in range(3):
        on = sc.vel_active[d] != 0
        assert np.array_equal(ref.vel_active[d] != 0, on)
        same_bits(out[d], ref.vel[d], on, (name, flags, d))
        assert np.array_equal(out[d][~on], sc.vel[d][~on])      # inactive faces are not written
        moved += int((out[d][on] != sc.vel[d][on]).sum())
    assert moved > 0


@pytest.mark.parametrize("case", [("dambreak_solid", 128, {}), ("dambreak_solid", 96, {"WENO": "Yes"}), ("flip_splash", 128, {}), ("smoke_plume", 96, {})], ids=lambda c: f"{c[0]}{c[1]}{IDS(c[2])}")
def test_bit_exact_at_sizes_the_reference_still_runs_in_seconds(cuda_device, case):
    """128^3 / 96^3 (many blocks per plane, rows longer than one walk of a block's 64 threads, multi-cell displacements across block borders): velocity, and the
    level set carried by itself, against the reference module run live."""
    need_ref()
    workload, n, flags = case
    sc = swirl(scenes.BENCH_SCENES[workload](n), 3.0, seed=21)
    ref = refio.run_reference(sc, "f32", flags=flags, advect="vector")
    A = MacAdvection3(sc.shape, sc.dx, **flags)
    out = A.advect_vector(sc.vel, sc.vel_active, sc.fluid, sc.dt)
    for d in range(3):
        same_bits(out[d], ref.vel[d], sc.vel_active[d] != 0, (case, d))
    if sc.fluid_raw is not None:
        ref = refio.run_reference(sc, "f32", flags=flags, advect="levelset")
        qa = fluid_active(sc)
        q = A.advect_scalar(sc.fluid, qa, sc.vel, sc.vel_active, sc.fluid, sc.dt, background=float(np.float32(sc.band)))
        same_bits(q, ref.pressure, qa != 0, (case, "levelset"))
    A.close()


@pytest.mark.parametrize("name", ["dambreak_solid", "smoke"])
@pytest.mark.parametrize("flags", [{}, {"WENO": "Yes"}], ids=IDS)
def test_advect_vector_real_double(cuda_device, name, flags):
    need_ref("f64")
    sc = SCENES[name]()
    ref = refio.run_reference(sc, "f64", flags=flags, advect="vector")
    A = MacAdvection3(sc.shape, sc.dx, real="f64", **flags)
    out = A.advect_vector(sc.vel, sc.vel_active, sc.fluid, sc.dt)
    A.close()
    for d in range(3):
        same_bits(out[d], ref.vel[d], sc.vel_active[d] != 0, (name, flags, d))


@pytest.mark.parametrize("flags", FLAGS, ids=IDS)
@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("mode", ["density", "levelset"])
def test_advect_scalar_equals_the_reference_bit_for_bit(cuda_device, name, flags, mode):
    need_ref()
    sc = SCENES[name]()
    if mode == "levelset" and sc.fluid_raw is None:
        pytest.skip("no liquid level set in a smoke scene")
    ref = refio.run_reference(sc, "f32", flags=flags, advect=mode)
    if mode == "density":
        q, qa = density_of(sc)
        background = 0.0
    else:   # the level set carried by itself (maclevelsetsurfacetracker3.cpp:49-51): `fluid` is the grid q was copied from
        q, qa = sc.fluid.astype(np.float32), fluid_active(sc)
        background = float(np.float32(sc.band))
    A = MacAdvection3(sc.shape, sc.dx, **flags)
    out = A.advect_scalar(q, qa, sc.vel, sc.vel_active, sc.fluid, sc.dt, background=background)
    A.close()
    on = qa != 0
    assert np.array_equal(ref.pressure_active != 0, on)
    same_bits(out, ref.pressure, on, (name, flags, mode))
    assert np.array_equal(out[~on], q[~on])
    assert (out[on] != q[on]).any()


@pytest.mark.parametrize("name", ["dambreak_solid", "flip", "smoke"])
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_sparse_host_copies_equal_whole_array_copies(cuda_device, name, real, monkeypatch):
    """Page-locked buffers: on a liquid scene only the masks and the values of active faces cross PCIe (stats say so) and the caller's buffers end up byte for
    byte as with whole-array copies — inactive entries untouched, whatever they held; an all-active scene keeps whole-array copies."""
    from test_gpu_host_sparse import Pinned
    sc = SCENES[name]()
    dtype = np.float64 if real == "f64" else np.float32
    rng = np.random.default_rng(4)
    junk = [np.where(a != 0, v, rng.standard_normal(v.shape)).astype(dtype) for v, a in zip(sc.vel, sc.vel_active)]   # inactive entries hold anything
    P = Pinned()
    try:
        A = MacAdvection3(sc.shape, sc.dx, real=real)
        u = [P.like(v) for v in junk]
        act = [P.like(a.astype(np.uint8)) for a in sc.vel_active]
        q0, qa = density_of(sc)
        q_sparse = A.advect_scalar(q0.astype(dtype), qa, u, act, sc.fluid, sc.dt)          # the velocity of a scalar call travels the same way (and only one way)
        st_scalar = dict(A.last_stats)
        st = A.advect_vector_inplace(u, act, sc.fluid, sc.dt)
        monkeypatch.setenv("SHKZ_B200_HOST_COPIES", "dense")
        w = [v.copy() for v in junk]
        st_dense = A.advect_vector_inplace(w, [a.astype(np.uint8) for a in sc.vel_active], sc.fluid, sc.dt)
        monkeypatch.delenv("SHKZ_B200_HOST_COPIES")
        A.close()
        faces = sum(v.size for v in sc.vel)
        active = sum(int(a.sum()) for a in sc.vel_active)
        assert st_dense["host_copies"] == 0 and st_dense["d2h_bytes"] == faces * dtype().itemsize
        if 2 * active <= faces:
            assert st["host_copies"] == 1 and st["d2h_bytes"] == active * dtype().itemsize
            assert st["h2d_bytes"] == faces + active * dtype().itemsize + sc.fluid.size * dtype().itemsize
        else:
            assert st["host_copies"] == 0
        for d in range(3):
            assert np.array_equal(u[d], w[d]), (name, real, d)
            off = sc.vel_active[d] == 0
            assert np.array_equal(u[d][off], junk[d][off])
        ref = MacAdvection3(sc.shape, sc.dx, real=real)
        out = ref.advect_vector(sc.vel, sc.vel_active, sc.fluid, sc.dt)   # zeros on the inactive entries instead of junk: same active values
        assert np.array_equal(ref.advect_scalar(q0.astype(dtype), qa, sc.vel, sc.vel_active, sc.fluid, sc.dt), q_sparse)
        assert st_scalar["host_copies"] == (1 if 2 * active <= faces else 0) and ref.last_stats["host_copies"] == 0
        ref.close()
        for d in range(3):
            on = sc.vel_active[d] != 0
            assert np.array_equal(u[d][on], out[d][on])
    finally:
        P.free()


def test_module_is_a_drop_in_through_the_reference_loader(cuda_device):
    """`Advection=b200advection3` loaded by the reference's own module loader (oracle/ref_driver) against `Advection=macadvection3`: same bits, all three calls."""
    need_ref()
    if not os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200advection3.so")):
        pytest.skip("the Shiokaze module was not built (needs the reference headers)")
    for name, flags in (("dambreak_solid", {}), ("flip", {"WENO": "Yes"}), ("smoke", {"MacCormack": "No"})):
        sc = SCENES[name]()
        for mode in ("vector", "density", "levelset"):
            if mode == "levelset" and sc.fluid_raw is None:
                continue
            ref = refio.run_reference(sc, "f32", flags=flags, advect=mode)
            mod = refio.run_reference(sc, "f32", flags=flags, advect=mode, advection="b200advection3")
            for d in range(3):
                assert np.array_equal(mod.vel_active[d], ref.vel_active[d])
                assert np.array_equal(mod.vel[d], ref.vel[d]), (name, mode, d)
            assert np.array_equal(mod.pressure_active, ref.pressure_active)
            assert np.array_equal(mod.pressure, ref.pressure), (name, mode)


def test_properties_at_a_size_the_reference_does_not_run_in_seconds(cuda_device):
    """256^3, liquid scene: a field at rest stays what it is; a uniform field is carried onto itself away from the walls; the limiter keeps every value inside
    the range of the input; two calls give the same bits."""
    n = 256
    sc = scenes.dambreak(n, True)
    A = MacAdvection3(sc.shape, sc.dx)
    rest = [np.zeros_like(v) for v in sc.vel]
    out = A.advect_vector(rest, sc.vel_active, sc.fluid, 0.5)
    assert all(not o.any() for o in out)
    uni = [np.where(a != 0, np.float32(c), np.float32(0)) for a, c in zip(sc.vel_active, (0.25, -0.5, 0.125))]
    out = A.advect_vector(uni, sc.vel_active, sc.fluid, 2.0 * sc.dx)     # carried by up to one cell
    for d, c in enumerate((0.25, -0.5, 0.125)):
        on = sc.vel_active[d] != 0
        assert float(np.abs(out[d][on]).max()) <= abs(c) * (1 + 1e-6)
        inner = minimum_filter(on.astype(np.uint8), size=7, mode="constant", cval=0) != 0   # every face within three cells is active: no stencil meets a 0
        assert inner.any()
        assert float(np.abs(out[d][inner] - np.float32(c)).max()) <= 2e-6 * abs(c)
    sw = swirl(sc, 3.0)
    out1 = A.advect_vector(sw.vel, sw.vel_active, sw.fluid, sw.dt)
    out2 = A.advect_vector(sw.vel, sw.vel_active, sw.fluid, sw.dt)
    lo, hi = min(float(v.min()) for v in sw.vel), max(float(v.max()) for v in sw.vel)
    for d in range(3):
        assert np.array_equal(out1[d], out2[d])
        assert lo <= float(out1[d].min()) and float(out1[d].max()) <= hi
    A.close()


def test_edge_cases(cuda_device):
    """No active face at all; a zero time step; the smallest grid the library takes (2 x 3 x 2, every stencil clamped at a wall), ragged extents."""
    rng = np.random.default_rng(8)
    for shape in ((2, 3, 2), (5, 2, 3), (33, 7, 65)):
        nx, ny, nz = shape
        A = MacAdvection3(shape, 1.0 / max(shape))
        fs = [A.face_shape(d) for d in range(3)]
        vel = [rng.standard_normal(f).astype(np.float32) for f in fs]
        none = [np.zeros(f, np.uint8) for f in fs]
        fluid = rng.standard_normal((nz, ny, nx)).astype(np.float32)
        out = A.advect_vector(vel, none, fluid, 0.1)
        assert all(np.array_equal(o, v) for o, v in zip(out, vel))                      # nothing active: nothing written
        q = rng.standard_normal((nz, ny, nx)).astype(np.float32)
        assert np.array_equal(A.advect_scalar(q, np.zeros((nz, ny, nx), np.uint8), vel, none, fluid, 0.1), q)
        every = [np.ones(f, np.uint8) for f in fs]
        out = A.advect_vector(vel, every, fluid, 0.0)                                   # dt = 0: every position is a grid point, weight 1 on itself
        assert all(np.array_equal(o, v) for o, v in zip(out, vel))
        assert np.array_equal(A.advect_scalar(q, np.ones((nz, ny, nx), np.uint8), vel, every, fluid, 0.0), q)
        some = [(rng.random(f) < 0.6).astype(np.uint8) for f in fs]                     # ragged activity, displacements far beyond the grid: clamped, finite
        out = A.advect_vector(vel, some, fluid, 50.0)
        for d in range(3):
            assert np.isfinite(out[d]).all()
            assert np.array_equal(out[d][some[d] == 0], vel[d][some[d] == 0])
            lo, hi = min(0.0, float(vel[d].min())), max(0.0, float(vel[d].max()))
            assert lo <= float(out[d].min()) and float(out[d].max()) <= hi
        A.close()


def test_argument_errors(cuda_device):
    import ctypes as C
    from shiokaze_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    assert L.shkz_b200_advect_create(1, 8, 8, 0.1, 0, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_advect_create(8, 8, 8, 0.1, 7, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_advect_vector_host(None, 0.1, None, None, None, None, None) == capi.ERR_ARG
    A = MacAdvection3((8, 8, 8), 0.125)
    with pytest.raises(ValueError):
        A.advect_vector([np.zeros((8, 8, 8), np.float32)] * 3, [np.zeros((8, 8, 8), np.uint8)] * 3, None, 0.1)
    sc = scenes.smoke_plume(8)
    with pytest.raises(capi.ShkzError):   # MacCormack needs the liquid level set
        A.advect_vector(sc.vel, sc.vel_active, None, 0.1)
    A.configure(MacCormack="No")
    A.advect_vector(sc.vel, sc.vel_active, None, 0.1)
    A.close()
