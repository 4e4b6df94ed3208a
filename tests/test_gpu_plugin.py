"""Drop-in test: the SAME headless Shiokaze host (oracle/ref_driver, the reference's own libraries and module
loader) runs one project() with `Projection=macpressuresolver3` (the reference) and with
`Projection=b200pressure3` (this repository's module -> C-ABI -> CUDA), on the same sparse grids."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import refio
from shiokaze_b200 import scenes

pytestmark = pytest.mark.gpu


def have_plugin(real):
    import os
    return refio.ref_available(real) and os.path.isfile(os.path.join(refio.ref_dir(real), "libshiokaze_b200pressure3.so"))


@pytest.mark.parametrize("real,tol", [("f32", 1e-3), ("f64", 1e-5)])
@pytest.mark.parametrize("scene", ["dambreak_solid", "smoke", "flip"])
def test_module_is_a_drop_in(cuda_device, real, tol, scene):
    if not have_plugin(real):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = {"dambreak_solid": lambda: scenes.dambreak(32, True), "smoke": lambda: scenes.smoke_plume(24),
          "flip": lambda: scenes.flip_splash(40)}[scene]()
    flags = {"Residual": 1e-10}
    ref = refio.run_reference(sc, real, flags=flags)
    ours = refio.run_reference(sc, real, flags={**flags, "Precision": "fp64"}, projection="b200pressure3")
    assert "b200pressure3.so" in ours.stdout
    assert np.array_equal(ours.pressure_active, ref.pressure_active)       # get_pressure() activity == row set
    for d in range(3):
        assert np.array_equal(ours.vel_active[d], ref.vel_active[d])       # set_off() on the same faces
    assert rel_l2(ours.vel, ref.vel) < tol
    assert 0 < ours.iterations < ref.iterations                            # MG-PCG needs far fewer iterations


def test_module_reference_flags_and_volume_correction(cuda_device):
    if not have_plugin("f32"):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = scenes.dambreak(24, True)
    flags = {"Residual": 1e-10, "SecondOrderAccurateFluid": "No", "SecondOrderAccurateSolid": "No", "EpsFluid": 0.05}
    ref = refio.run_reference(sc, "f32", flags=flags, current_volume=1.05, target_volume=1.0)
    ours = refio.run_reference(sc, "f32", flags={**flags, "Precond": "none", "Precision": "fp64"}, projection="b200pressure3",
                               current_volume=1.05, target_volume=1.0)
    assert rel_l2(ours.vel, ref.vel) < 1e-5
    assert abs(ours.iterations - ref.iterations) <= max(3, 0.06 * ref.iterations)   # Precond=none is the reference's CG


def test_module_warm_start_through_the_loader(cuda_device):
    """WarmStart=Yes is a documented flag of the module this one replaces (macpressuresolver3.cpp:279, 221-242): the same host, the same
    inputs projected twice (ref_driver repeat=2), reference module vs ours."""
    if not have_plugin("f32"):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = scenes.dambreak(32, True)
    flags = {"WarmStart": "Yes"}
    ref = refio.run_reference(sc, "f32", flags=flags, repeat=2)
    ours = refio.run_reference(sc, "f32", flags={**flags, "Precond": "none", "Precision": "fp64"}, projection="b200pressure3", repeat=2)
    assert np.array_equal(ours.pressure_active, ref.pressure_active)
    for d in range(3):
        assert np.array_equal(ours.vel_active[d], ref.vel_active[d])
    assert rel_l2(ours.vel, ref.vel) < 1e-4
    assert abs(ours.iterations - ref.iterations) <= max(3, 0.1 * ref.iterations)      # second call: the correction solve
    cold = refio.run_reference(sc, "f32", projection="b200pressure3", flags={"Precond": "none", "Precision": "fp64"})
    assert ours.iterations != cold.iterations or ref.iterations == cold.iterations


def test_module_emits_the_reference_records(cuda_device):
    """console::write names of macpressuresolver3.cpp (scoped_timer::stock -> "<Arg>_<name>"): the records a Shiokaze log analyser looks for."""
    if not have_plugin("f32"):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = scenes.dambreak(24)
    sc.surface_tension = 0.05
    ref = refio.run_reference(sc, "f32", current_volume=1.05, target_volume=1.0, records=True)
    out = refio.run_reference(sc, "f32", projection="b200pressure3", current_volume=1.05, target_volume=1.0, records=True)
    assert len(ref.records) >= 9
    for name, values in ref.records.items():          # every record the reference module writes, under the same name
        assert name in out.records and len(out.records[name]) == len(values), (name, sorted(out.records))
    assert out.records["Projection_number_projection_iteration"][0] == out.iterations
    assert out.records["Projection_volume_correct_rhs"] == ref.records["Projection_volume_correct_rhs"]


@pytest.mark.parametrize("scene", ["dambreak_solid", "smoke", "flip"])
def test_dense_array_core_is_a_zero_copy_bridge(cuda_device, scene):
    """`Array=b200array3`: the host's grids live in page-locked dense buffers and b200pressure3 hands them to the C-ABI in place (no gather / scatter).
    Same host, same module, default tiledarray3 grids vs b200array3 grids: the GPU sees the same dense inputs, so the outputs are the same bits."""
    import os
    if not (have_plugin("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200array3.so"))):
        pytest.skip("oracle/_ref (reference build + modules) was not shipped to this box")
    sc = {"dambreak_solid": lambda: scenes.dambreak(40, True), "smoke": lambda: scenes.smoke_plume(32), "flip": lambda: scenes.flip_splash(40)}[scene]()
    tiled = refio.run_reference(sc, "f32", projection="b200pressure3")
    dense = refio.run_reference(sc, "f32", projection="b200pressure3", flags={"Array": "b200array3"}, repeat=2)
    assert 'Loaded "b200array3.so"' in dense.stdout
    assert dense.iterations == tiled.iterations
    for d in range(3):
        assert np.array_equal(dense.vel[d], tiled.vel[d]) and np.array_equal(dense.vel_active[d], tiled.vel_active[d])
    assert np.array_equal(dense.pressure, tiled.pressure) and np.array_equal(dense.pressure_active, tiled.pressure_active)
    ref = refio.run_reference(sc, "f32")
    assert np.array_equal(dense.pressure_active, ref.pressure_active)
    assert rel_l2(dense.vel, ref.vel) < 2e-3
