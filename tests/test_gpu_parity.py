"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C-ABI, against
the dense oracle on the same seeded inputs, against the committed reference goldens, and — at sizes the
oracle cannot reach — through size-independent properties.

Bars: integer / mask / fraction / matrix-entry work is bit-exact; post-projection velocity agrees with
the reference within 1e-5 relative L2 for the Real=double build and 1e-3 for Real=float (BASELINE.json
north_star), both sides solving to Residual=1e-10.
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, golden_cases, load_accuracy_golden, load_golden, rel_l2
from oracle import dense_oracle
from shiokaze_b200 import MacPressureSolver3, capi, scenes

pytestmark = pytest.mark.gpu

F32_TOL = 1e-3   # north_star: fp32
F64_TOL = 1e-5   # north_star: fp64


def solver_for(sc, real="f32", test_hooks=False, **flags):
    return MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, real=real, test_hooks=test_hooks, **flags)


def apply_case(S, kw):
    flags = {k: v for k, v in kw.items() if k.startswith("SecondOrder")}
    if flags:
        S.configure(**flags)
    if "volume" in kw:
        S.set_target_volume(*kw["volume"])


def oracle_args(sc, kw):
    args = {}
    if "SecondOrderAccurateFluid" in kw:
        args.update(second_order_fluid=kw["SecondOrderAccurateFluid"], second_order_solid=kw["SecondOrderAccurateSolid"])
    if "surface_tension" in kw:
        args["surface_tension"] = kw["surface_tension"]
    if "volume" in kw:
        args["rhs_correct"] = dense_oracle.volume_correction(1.0, kw["volume"][0], kw["volume"][1], sc.dt, 0.0)[0]
    return args


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_assembly_is_bit_exact(cuda_device, name):
    """Fractions, row mask, matrix diagonal / face coefficients and right-hand side == oracle, bit for bit."""
    make, kw = golden_cases()[name]
    sc = make()
    S = solver_for(sc, Precision="fp64", Precond="none", MaxIterations=0)
    apply_case(S, kw)
    out = S.project_scene(sc, surface_tension=kw.get("surface_tension", 0.0))
    o = dense_oracle.project(sc, max_iterations=0, **oracle_args(sc, kw))
    shp = (sc.nz, sc.ny, sc.nx)
    assert np.array_equal(S.debug_fetch("in_rows").reshape(shp), o.in_rows)
    assert np.array_equal(out["pressure_active"], o.in_rows)
    for d, fs in enumerate(S.face_shapes()):
        assert np.array_equal(S.debug_fetch(f"areas{d}").view(np.float32).reshape(fs).astype(np.float64), o.areas[d])
        assert np.array_equal(S.debug_fetch(f"rhos{d}").view(np.float32).reshape(fs).astype(np.float64), o.rhos[d])
    # the diagonal is carried as (Dirichlet part) + (six couplings): the Dirichlet part is bit-exact, and the
    # reference's diagonal is recovered from the pieces to rounding
    dd = S.debug_fetch("dd").view(np.float64).reshape(shp)
    assert np.array_equal(dd, o.dirichlet)
    assert np.array_equal(S.debug_fetch("rhs").view(np.float64).reshape(shp), o.rhs)
    assert out["result"].n_rows == o.n_rows
    assert out["result"].stats["rhs_absmax"] == o.rhs_absmax
    # lower-face couplings: w = dt*A/(dx^2*theta) where both cells are rows
    names = ["wx", "wy", "wz"]
    for d in range(3):
        w = S.debug_fetch(names[d]).view(np.float64).reshape(shp)
        a, r = o.areas[d], o.rhos[d]
        with np.errstate(divide="ignore", invalid="ignore"):
            full = np.where((a != 0) & (r != 0), sc.dt * a / (sc.dx * sc.dx * r), 0.0)
        sl = [slice(None)] * 3
        sl[2 - d] = slice(0, -1)
        full = full[tuple(sl)]
        rows = o.in_rows.astype(bool)
        lower = np.zeros_like(rows)
        hi = [slice(None)] * 3
        lo = [slice(None)] * 3
        hi[2 - d], lo[2 - d] = slice(1, None), slice(0, -1)
        lower[tuple(hi)] = rows[tuple(lo)]
        assert np.array_equal(w, np.where(rows & lower, full, 0.0))
        up = np.zeros(shp)
        up[tuple(lo)] = w[tuple(hi)]
        dd = dd + w + up
    assert np.allclose(dd, o.diag, rtol=1e-14, atol=0.0)
    S.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_plain_cg_tracks_the_reference_iteration_for_iteration(cuda_device, name):
    """Precond=none is the reference's algorithm (its MIC(0) result is discarded, pcg_solver.h:383):
    same stopping rule, iteration counts within 6 % at Residual=1e-10 and 10 % at the loose default 1e-4
    (summation order differs; the oracle itself is held to the same bars in test_oracle.py), same fields."""
    make, kw = golden_cases()[name]
    sc = make()
    for residual, tag in ((1e-4, "f32_default"), (1e-10, "f32_tight")):
        S = solver_for(sc, Precision="fp64", Precond="none", Residual=residual)
        apply_case(S, kw)
        out = S.project_scene(sc, surface_tension=kw.get("surface_tension", 0.0))
        g = load_golden(name, tag)
        res = out["result"]
        assert res.converged and res.reresid <= residual
        slack = 0.10 if residual == 1e-4 else 0.06
        assert abs(res.iterations - g["iterations"]) <= max(3, slack * g["iterations"]), (res.iterations, g["iterations"])
        assert np.array_equal(out["pressure_active"], g["pressure_active"])
        for d in range(3):
            assert np.array_equal(out["vel_active"][d], g["act"][d])
        assert rel_l2(out["vel"], g["vel"]) < (2e-3 if residual == 1e-4 else 1e-5)
        S.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("precond", ["none", "mg"])
def test_velocity_parity_fp64_build(cuda_device, name, precond):
    """Real=double host, fp64 solve, Residual=1e-10 vs the Real=double reference build: <= 1e-5 rel. L2."""
    make, kw = golden_cases()[name]
    sc = make()
    S = solver_for(sc, real="f64", Precision="fp64", Precond=precond, Residual=1e-10)
    apply_case(S, kw)
    out = S.project_scene(sc, surface_tension=kw.get("surface_tension", 0.0))
    g = load_golden(name, "f64_tight")
    assert out["result"].converged
    assert np.array_equal(out["pressure_active"], g["pressure_active"])
    for d in range(3):
        assert np.array_equal(out["vel_active"][d], g["act"][d])
    assert rel_l2(out["vel"], g["vel"]) < F64_TOL
    if out["result"].stats["has_dirichlet"]:
        assert rel_l2([out["pressure"]], [g["pressure"]]) < F64_TOL
    if precond == "mg":
        assert out["result"].iterations < 0.5 * g["iterations"]
    S.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("precision", ["mixed", "fp32"])
def test_velocity_parity_fp32_build(cuda_device, name, precision):
    """Shipping configuration (Real=float host, MG preconditioner): <= 1e-3 rel. L2 vs the reference."""
    make, kw = golden_cases()[name]
    sc = make()
    residual = 1e-10 if precision == "mixed" else 1e-5   # an all-float CG cannot reach 1e-10
    S = solver_for(sc, real="f32", Precision=precision, Precond="mg", Residual=residual, MaxIterations=400)
    apply_case(S, kw)
    out = S.project_scene(sc, surface_tension=kw.get("surface_tension", 0.0))
    g = load_golden(name, "f32_tight")
    assert out["result"].converged, out["result"]
    assert np.array_equal(out["pressure_active"], g["pressure_active"])
    for d in range(3):
        assert np.array_equal(out["vel_active"][d], g["act"][d])
    assert rel_l2(out["vel"], g["vel"]) < F32_TOL
    S.close()


def test_accuracytest3_known_answer(cuda_device):
    """The reference's own test (src/examples/accuracytest3-example.cpp) through the CUDA path."""
    gold = load_accuracy_golden()
    prev = None
    for n in (8, 16, 32, 64):
        worst = 0.0
        S = MacPressureSolver3((n, n, n), 1.0 / n, Precision="fp64", Precond="mg", Residual=1e-13, EpsFluid=1e-18, MaxIterations=200)
        for q in [q for q in range(-4, 5) if q] + [0]:
            sc = scenes.accuracy_sphere(n, q)
            out = S.project_scene(sc)
            c = (np.arange(n) + .5) * sc.dx
            exact = (c[None, None, :] - .5) ** 2 + (c[None, :, None] - .5) ** 2 + (c[:, None, None] - .5) ** 2 - sc.meta["r"] ** 2
            worst = max(worst, float(np.abs(exact - out["pressure"])[out["pressure_active"] > 0].max()))
        S.close()
        assert worst == pytest.approx(gold["survey_goldens"][str(n)], rel=2e-3), (n, worst)
        if prev:
            assert 1.6 < np.log2(prev / worst) < 2.2
        prev = worst


def test_edge_cases(cuda_device):
    # no liquid at all -> no rows, zero iterations, faces keep rho = 1 (no level set)
    sc = scenes.liquid_box(8)
    sc.fluid[:] = np.float32(sc.band)
    sc.fluid_levelset = False
    S = solver_for(sc)
    out = S.project_scene(sc)
    o = dense_oracle.project(sc)
    assert out["result"].n_rows == 0 and out["result"].iterations == 0
    for d in range(3):
        assert np.array_equal(out["vel_active"][d], o.vel_active[d])
        assert np.array_equal(out["vel"][d].astype(np.float64), o.vel[d])
    S.close()
    # zero right-hand side -> zero iterations (pcg_solver.h:254-258)
    sc = scenes.dambreak(12)
    for v in sc.vel:
        v[:] = 0
    S = solver_for(sc)
    out = S.project_scene(sc)
    assert out["result"].iterations == 0 and out["result"].n_rows > 0 and out["result"].converged
    assert float(np.abs(out["pressure"]).max()) == 0.0
    S.close()
    # MaxIterations cut-off is reported as the count (pcg_solver.h:292)
    sc = scenes.smoke_plume(12)
    S = solver_for(sc, Precond="none", MaxIterations=5, CheckEvery=3)
    out = S.project_scene(sc)
    assert out["result"].iterations == 5 and not out["result"].converged
    S.close()
    # degenerate extents
    for shape in ((1, 9, 7), (5, 1, 3), (4, 6, 1), (2, 2, 2), (33, 3, 5)):
        sc = scenes.random_blobs(*shape, seed=4, with_solid=False)
        o = dense_oracle.project(sc, residual=1e-10)
        S = solver_for(sc, Precision="fp64", Residual=1e-10)
        out = S.project_scene(sc)
        assert np.array_equal(out["pressure_active"], o.in_rows), shape
        for d in range(3):
            assert np.array_equal(out["vel_active"][d], o.vel_active[d]), shape
        assert rel_l2(out["vel"], o.vel) < 1e-5, shape
        S.close()


def test_resolve_repeats_the_solve(cuda_device):
    sc = scenes.dambreak(32, True)
    S = solver_for(sc, Precision="mixed", Precond="mg")
    out = S.project_scene(sc)
    again = S.resolve()
    assert again.iterations == out["result"].iterations
    assert again.reresid == out["result"].reresid
    S.close()


def test_mg_iteration_counts_are_mesh_independent(cuda_device):
    counts = {}
    for n in (32, 64, 128):
        sc = scenes.dambreak(n, True)
        S = solver_for(sc)
        out = S.project_scene(sc)
        assert out["result"].converged
        counts[n] = out["result"].iterations
        S.close()
    assert max(counts.values()) <= 16, counts
    assert counts[128] <= counts[32] + 6, counts


@pytest.mark.parametrize("scene_name,n", [("smoke_plume", 256), ("dambreak_solid", 192)])
def test_large_grid_properties(cuda_device, scene_name, n):
    """Sizes the CPU oracle cannot solve in seconds: check what the projection must guarantee.
    (1) the weighted divergence of the projected velocity vanishes on every row to the solver tolerance
        (this IS the linear system: b - A p, recomputed independently here in numpy, float64);
    (2) plain CG and MG-preconditioned CG agree; (3) idempotence: projecting twice changes nothing."""
    sc = scenes.BENCH_SCENES[scene_name](n)
    S = solver_for(sc, Precision="mixed", Precond="mg", Residual=1e-8)
    out = S.project_scene(sc)
    assert out["result"].converged and out["result"].iterations < 40
    shp = (sc.nz, sc.ny, sc.nx)
    areas = [S.debug_fetch(f"areas{d}").view(np.float32).reshape(fs).astype(np.float64) for d, fs in enumerate(S.face_shapes())]
    rows = out["pressure_active"].astype(bool)
    bmax = out["result"].stats["rhs_absmax"]

    def divergence(vel):
        div = np.zeros(shp)
        for d in range(3):
            flux = areas[d] * vel[d].astype(np.float64)
            lo = [slice(None)] * 3
            hi = [slice(None)] * 3
            lo[2 - d], hi[2 - d] = slice(0, -1), slice(1, None)
            # walls contribute nothing (neighbour out of the grid): zero those fluxes like the assembly does
            f = flux.copy()
            first = [slice(None)] * 3
            last = [slice(None)] * 3
            first[2 - d], last[2 - d] = slice(0, 1), slice(-1, None)
            f[tuple(first)] = 0
            f[tuple(last)] = 0
            div += (f[tuple(hi)] - f[tuple(lo)]) / sc.dx
        return div

    before = np.abs(divergence(sc.vel)[rows]).max()
    after = np.abs(divergence(out["vel"])[rows]).max()
    assert before == pytest.approx(bmax, rel=1e-6)
    assert after < 5e-5 * before, (before, after)        # float storage of p and u limits this, not the solver
    # plain CG agrees (looser residual keeps the test short)
    S2 = solver_for(sc, Precision="mixed", Precond="none", Residual=1e-6, CheckEvery=25)
    out2 = S2.project_scene(sc)
    assert out2["result"].converged
    assert rel_l2(out2["vel"], out["vel"]) < 1e-4
    S2.close()
    # idempotence
    sc2 = scenes.BENCH_SCENES[scene_name](n)
    sc2.vel = [v.copy() for v in out["vel"]]
    sc2.vel_active = [a.copy() for a in out["vel_active"]]
    out3 = S.project_scene(sc2)
    assert rel_l2(out3["vel"], out["vel"]) < 1e-4
    S.close()


def test_device_entry_point_matches_host_entry_point(cuda_device):
    import torch
    sc = scenes.flip_splash(48)
    S = solver_for(sc)
    ref = S.project_scene(sc)
    dev = torch.device("cuda", 0)
    vel = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in sc.vel]
    act = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sc.vel_active]
    fluid = torch.from_numpy(sc.fluid).to(dev)
    solid = torch.from_numpy(sc.solid).to(dev)
    p = torch.zeros(sc.fluid.shape, dtype=torch.float32, device=dev)
    pa = torch.zeros(sc.fluid.shape, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    res = S.project_device(sc.dt, vel, act, solid, fluid, sc.fluid_levelset, p, pa)
    assert res.iterations == ref["result"].iterations
    for d in range(3):
        assert np.array_equal(vel[d].cpu().numpy(), ref["vel"][d])
        assert np.array_equal(act[d].cpu().numpy(), ref["vel_active"][d])
    assert np.array_equal(p.cpu().numpy(), ref["pressure"])
    S.close()


@pytest.mark.parametrize("case", [("dambreak_solid", 40, 2, 2), ("smoke", 40, 1, 1), ("flip", 48, 2, 1), ("blobs", (33, 21, 19), 2, 2),
                                  ("dambreak_solid", 96, 2, 2), ("smoke", 130, 2, 2), ("blobs", (70, 18, 37), 3, 0), ("blobs", (72, 40, 21), 2, 2),
                                  ("smoke", 256, 2, 2)])
@pytest.mark.parametrize("precision", ["mixed", "fp32"])
def test_fused_vcycle_equals_unfused_vcycle_bit_for_bit(cuda_device, case, precision):
    """The fused sweep / residual+restrict / shared-memory-tail kernels against one-launch-per-colour kernels."""
    kind, n, pre, post = case
    sc = {"dambreak_solid": lambda: scenes.dambreak(n, True), "smoke": lambda: scenes.smoke_plume(n), "flip": lambda: scenes.flip_splash(n),
          "blobs": lambda: scenes.random_blobs(*n, seed=11)}[kind]()
    S = solver_for(sc, Precision=precision, Precond="mg", MGPreSweeps=pre, MGPostSweeps=post, MaxIterations=1, test_hooks=True)
    out = S.project_scene(sc)
    fused = S.debug_vcycle(legacy=0)
    scalar = S.debug_vcycle(legacy=2)
    quad = S.debug_vcycle(legacy=3)
    legacy = S.debug_vcycle(legacy=1)
    assert np.isfinite(fused).all()
    assert np.array_equal(fused, scalar)      # TMA-staged sweep kernel == scalar sweep kernel
    assert np.array_equal(quad, scalar)       # quad (direct-load) sweep kernel == scalar sweep kernel
    assert float(np.abs(legacy).max()) > 0
    # cells without an equation are only defined after a post-sweep (the dense kernels also prolong into them)
    rows = out["pressure_active"].astype(bool)
    assert np.array_equal(fused[rows], legacy[rows])
    if post > 0:
        assert np.array_equal(fused, legacy)
    S.close()
