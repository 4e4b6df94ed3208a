"""The step after the projection on the device (csrc/kernels_post.cuh): velocity extrapolation + solid constraint against the reference's own
macutility3::extrapolate_and_constrain_velocity (src/utility/macutility3.cpp:61-93, include/shiokaze/array/array_extrapolator3.h:51-82), driven through the
reference's module loader by oracle/ref_driver (RefExtrapolate=<width>). Integer / mask work and the float arithmetic are restated operation for operation:
the bar is bit-exact."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import refio
from shiokaze_b200 import MacPressureSolver3, scenes

pytestmark = pytest.mark.gpu

SCENES = {"dambreak_solid": lambda: scenes.dambreak(32, True), "flip": lambda: scenes.flip_splash(40), "blobs": lambda: scenes.random_blobs(20, 14, 18, seed=3),
          "dambreak": lambda: scenes.dambreak(24), "blobs_nosolid": lambda: scenes.random_blobs(13, 21, 10, seed=5, with_solid=False)}


def need_ref():
    if not refio.ref_available("f32"):
        pytest.skip("oracle/_ref (the reference build) was not shipped to this box")


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("width", [0, 1, 3])
def test_extrapolate_and_constrain_equals_the_reference_bit_for_bit(cuda_device, name, width):
    """The operator alone, on the scene's input velocity (no projection in between): same masks, same bits."""
    need_ref()
    sc = SCENES[name]()
    ref = refio.run_reference(sc, "f32", extrapolate=width, skip_project=True)
    S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx)
    vel = [v.copy() for v in sc.vel]
    act = [a.copy() for a in sc.vel_active]
    S.extrapolate_and_constrain_velocity(sc.solid, vel, act, width)
    S.close()
    for d in range(3):
        assert np.array_equal(act[d], ref.vel_active[d]), (name, width, d)
        assert np.array_equal(vel[d].astype(np.float64), ref.vel[d]), (name, width, d, float(np.abs(vel[d] - ref.vel[d]).max()))
    if width:
        assert sum(int(a.sum()) for a in act) > sum(int(a.sum()) for a in sc.vel_active)


@pytest.mark.parametrize("name", ["dambreak_solid", "flip"])
def test_projection_followed_by_the_step_after_it(cuda_device, name):
    """project() with ExtrapolateWidth=2 (one call, everything on the device) against the reference doing project() and then its own extrapolate_and_constrain."""
    need_ref()
    sc = SCENES[name]()
    ref = refio.run_reference(sc, "f32", flags={"Residual": 1e-10}, extrapolate=2)
    S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision="fp64", Residual=1e-10, ExtrapolateWidth=2)
    out = S.project_scene(sc)
    S.close()
    for d in range(3):
        assert np.array_equal(out["vel_active"][d], ref.vel_active[d])
    assert rel_l2(out["vel"], ref.vel) < 1e-5
    if refio.ref_available("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200pressure3.so")):
        mod = refio.run_reference(sc, "f32", flags={"Residual": 1e-10, "Precision": "fp64", "ExtrapolateWidth": 2}, projection="b200pressure3")
        for d in range(3):
            assert np.array_equal(mod.vel_active[d], ref.vel_active[d])
        assert rel_l2(mod.vel, ref.vel) < 1e-5
