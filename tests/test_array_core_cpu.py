"""CPU test of the dense array core (`Array=b200array3`, shiokaze_b200/plugin/b200array3.cpp): the UNMODIFIED reference host and the
UNMODIFIED reference solver (macpressuresolver3 + pcg) run on it — every array3 / macarray3 / shared array of the run then lives in this core — and
must produce what they produce on the reference's own dense core, bit for bit (same serial iteration order, hence the same row numbering and the same
CG arithmetic), and what they produce on the default tiledarray3 to rounding. No GPU involved: without a device the core holds ordinary memory."""
import os

import numpy as np
import pytest

from oracle import refio
from shiokaze_b200 import scenes


def have_core():
    return refio.ref_available("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_b200array3.so"))


@pytest.mark.parametrize("scene", ["dambreak_solid", "flip", "smoke", "blobs"])
def test_reference_solver_on_b200array3_equals_lineararray3(scene):
    if not have_core():
        pytest.skip("oracle/_ref (reference build + b200array3) is not built here")
    sc = {"dambreak_solid": lambda: scenes.dambreak(32, True), "flip": lambda: scenes.flip_splash(32), "smoke": lambda: scenes.smoke_plume(16),
          "blobs": lambda: scenes.random_blobs(20, 14, 18, seed=3)}[scene]()
    sc.surface_tension = 0.02 if scene == "dambreak_solid" else 0.0
    ours = refio.run_reference(sc, "f32", flags={"Array": "b200array3", "Residual": 1e-10}, timeout=300)
    lin = refio.run_reference(sc, "f32", flags={"Array": "lineararray3", "Residual": 1e-10}, timeout=300)
    tiled = refio.run_reference(sc, "f32", flags={"Residual": 1e-10}, timeout=300)
    assert 'Loaded "b200array3.so"' in ours.stdout
    assert ours.iterations == lin.iterations
    for d in range(3):
        assert np.array_equal(ours.vel[d], lin.vel[d]) and np.array_equal(ours.vel_active[d], lin.vel_active[d])
        assert np.array_equal(ours.vel_active[d], tiled.vel_active[d])
        assert np.abs(ours.vel[d] - tiled.vel[d]).max() < 1e-6
    assert np.array_equal(ours.pressure, lin.pressure) and np.array_equal(ours.pressure_active, tiled.pressure_active)
    # the level sets as the host reads them back (narrow band, flood fill) are the core's doing too
    assert np.array_equal(ours.fluid, tiled.fluid) and np.array_equal(ours.solid, tiled.solid)
