"""z-slab solvers on several GPUs of one box (run with `gpurun --gpus 2|4 -- python -m pytest tests -m gpu`): the slabs
are driven from this one process, one host thread per GPU, and talk through peer memory (csrc/slab_comm.cuh).
Skipped where fewer than two CUDA devices are visible."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import dense_oracle
from shiokaze_b200 import MacPressureSolver3, capi, dist, scenes

pytestmark = pytest.mark.gpu


def world_sizes():
    n = capi.lib().shkz_b200_device_count()
    return [w for w in (2, 4, 8) if w <= n]


def make_slab_solvers(sc, world, **flags):
    solvers = [MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, device=r, zrange=dist.slab_range(sc.nz, r, world), **flags) for r in range(world)]
    dist.connect_local(solvers)
    return solvers


@pytest.fixture(scope="module")
def worlds(cuda_device):
    w = world_sizes()
    if not w:
        pytest.skip("needs at least two CUDA devices")
    return w


@pytest.mark.parametrize("kind,n", [("dambreak_solid", 32), ("smoke", 32), ("flip", 64), ("box", 32)])
def test_slab_projection_matches_whole_grid_and_oracle(worlds, kind, n):
    sc = {"dambreak_solid": lambda: scenes.dambreak(n, True), "smoke": lambda: scenes.smoke_plume(n), "flip": lambda: scenes.flip_splash(n),
          "box": lambda: scenes.liquid_box(n)}[kind]()
    flags = dict(Precision="fp64", Precond="mg", Residual=1e-10)
    W = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, **flags)
    whole = W.project_scene(sc)
    W.close()
    ref = dense_oracle.project(sc, residual=1e-10) if n <= 32 else None
    for world in worlds:
        solvers = make_slab_solvers(sc, world, **flags)
        parts = dist.project_local(solvers, dist.split_dense(sc, world))
        out = dist.join_dense(parts, world)
        for p in parts:   # every rank reports the same (globally reduced) solve
            assert p["result"].iterations == parts[0]["result"].iterations and p["result"].reresid == parts[0]["result"].reresid
            assert p["result"].n_rows == whole["result"].n_rows
        assert out["result"].converged
        assert abs(out["result"].iterations - whole["result"].iterations) <= 1
        assert np.array_equal(out["pressure_active"], whole["pressure_active"])
        for d in range(3):
            assert np.array_equal(out["vel_active"][d], whole["vel_active"][d])
        assert rel_l2(out["vel"], whole["vel"]) < 1e-6
        if ref is not None:
            assert np.array_equal(out["pressure_active"], ref.in_rows)
            assert rel_l2(out["vel"], ref.vel) < 1e-5
        for s in solvers:
            s.close()


@pytest.mark.parametrize("precision", ["mixed", "fp32"])
def test_slab_vcycle_equals_whole_grid_vcycle_bit_for_bit(worlds, precision):
    """Halo planes, half-updated boundary planes and the level structure reproduce the whole-grid V-cycle exactly."""
    sc = scenes.dambreak(64, True)
    flags = dict(Precision=precision, Precond="mg", MaxIterations=1, test_hooks=True)   # debug_vcycle lives in the test-hook build of the library
    W = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, **flags)
    W.project_scene(sc)
    whole = W.debug_vcycle(0)
    W.close()
    for world in worlds:
        solvers = make_slab_solvers(sc, world, **flags)
        dist.project_local(solvers, dist.split_dense(sc, world))
        parts = dist.run_per_slab([(lambda s=s: s.debug_vcycle(0)) for s in solvers])
        assert np.array_equal(np.concatenate(parts, axis=0), whole)
        for s in solvers:
            s.close()


def test_unconnected_slab_refuses_to_project(cuda_device):
    sc = scenes.smoke_plume(16)
    S = MacPressureSolver3((16, 16, 16), sc.dx, zrange=(0, 8))
    with pytest.raises(capi.ShkzError) as e:
        S.project_scene(dist.split_dense(sc, 2)[0])
    assert e.value.code == capi.ERR_STATE
    S.close()


@pytest.mark.parametrize("scene", ["dambreak_solid", "smoke"])
def test_module_gpus_flag_through_the_reference_loader(worlds, scene):
    """`Projection=b200pressure3 GPUs=N`: the Shiokaze module cuts the grid into N z-slabs, one device and one host thread each
    (plugin/b200pressure3.cpp: project_slabs) — same host, same sparse grids, against the reference module and against GPUs=1."""
    import os
    from oracle import refio
    if not (refio.ref_available("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200pressure3.so"))):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = {"dambreak_solid": lambda: scenes.dambreak(64, True), "smoke": lambda: scenes.smoke_plume(48)}[scene]()
    flags = {"Residual": 1e-10}
    ref = refio.run_reference(sc, "f32", flags=flags)
    one = refio.run_reference(sc, "f32", flags={**flags, "Precision": "fp64"}, projection="b200pressure3")
    for world in worlds:
        if sc.nz % world:
            continue
        many = refio.run_reference(sc, "f32", flags={**flags, "Precision": "fp64", "GPUs": world}, projection="b200pressure3")
        assert np.array_equal(many.pressure_active, ref.pressure_active)
        for d in range(3):
            assert np.array_equal(many.vel_active[d], ref.vel_active[d])
        assert rel_l2(many.vel, ref.vel) < 1e-3
        assert rel_l2(many.vel, one.vel) < 1e-6
        assert abs(many.iterations - one.iterations) <= 1


@pytest.mark.parametrize("scene", ["dambreak_solid", "flip", "blobs"])
def test_module_slabs_on_the_dense_array_core(worlds, scene):
    """`GPUs=N` together with `Array=b200array3`: every slab solver is handed a z-range of the host's own page-locked grids in place (the z-face grids
    through a per-slab copy), several projections in a row on the same grids — the sparse host copies (csrc/kernels_xfer.cuh) of N solvers at once."""
    import os
    from oracle import refio
    if not (refio.ref_available("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200array3.so"))):
        pytest.skip("oracle/_ref (reference build + modules) was not shipped to this box")
    # ("blobs": liquid surfaces that cross the slab boundaries at an angle — wet cells of one slab facing dry cells of the other)
    sc = {"dambreak_solid": lambda: scenes.dambreak(64, True), "flip": lambda: scenes.flip_splash(96), "blobs": lambda: scenes.random_blobs(48, 40, 32, seed=7)}[scene]()
    flags = {"Array": "b200array3", "Residual": 1e-10, "Precision": "fp64"}
    one = refio.run_reference(sc, "f32", projection="b200pressure3", flags=flags, repeat=3)
    for world in worlds:
        if sc.nz % world:
            continue
        many = refio.run_reference(sc, "f32", projection="b200pressure3", flags={**flags, "GPUs": world}, repeat=3)
        assert np.array_equal(many.pressure_active, one.pressure_active)
        for d in range(3):
            assert np.array_equal(many.vel_active[d], one.vel_active[d])
        assert rel_l2(many.vel, one.vel) < 1e-6
        assert abs(many.iterations - one.iterations) <= 1
