import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    """name -> (scene factory, reference flags, extras) — must mirror tests/golden/make_golden.py."""
    from shiokaze_b200 import scenes
    return {
        "dambreak24": (lambda: scenes.dambreak(24), {}),
        "dambreak_solid24": (lambda: scenes.dambreak(24, True), {}),
        "smoke16": (lambda: scenes.smoke_plume(16), {}),
        "flip32": (lambda: scenes.flip_splash(32), {}),
        "box16": (lambda: scenes.liquid_box(16), {}),
        "blobs": (lambda: scenes.random_blobs(20, 14, 18, seed=3), {}),
        "blobs_nosolid": (lambda: scenes.random_blobs(13, 21, 10, seed=5, with_solid=False), {}),
        "dambreak24_firstorder": (lambda: scenes.dambreak(24, True), {"SecondOrderAccurateFluid": False, "SecondOrderAccurateSolid": False}),
        "dambreak24_tension": (lambda: scenes.dambreak(24), {"surface_tension": 0.05}),
        "dambreak24_volume": (lambda: scenes.dambreak(24), {"volume": (1.05, 1.0)}),
    }


GOLDEN_NAMES = ["dambreak24", "dambreak_solid24", "smoke16", "flip32", "box16", "blobs", "blobs_nosolid",
                "dambreak24_firstorder", "dambreak24_tension", "dambreak24_volume"]


def load_golden(name, tag):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + ".")}
    return dict(vel=[out[f"vel{d}"] for d in range(3)], act=[out[f"act{d}"] for d in range(3)], pressure=out["pressure"],
                pressure_active=out["pressure_active"], iterations=int(out["iterations"]), reresid=float(out["reresid"]))


def load_accuracy_golden():
    with open(os.path.join(GOLDEN, "accuracytest3.json")) as f:
        return json.load(f)


def rel_l2(a_list, b_list):
    num = sum(float(((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2).sum()) for a, b in zip(a_list, b_list))
    den = sum(float((np.asarray(b, dtype=np.float64) ** 2).sum()) for b in b_list)
    return (num / den) ** 0.5 if den > 0 else num ** 0.5


@pytest.fixture(scope="session")
def cuda_device():
    from shiokaze_b200 import capi
    n = capi.lib().shkz_b200_device_count()
    if n < 1:
        pytest.fail("gpu test selected but libshkz_b200 sees no CUDA device (there is no CPU fallback)")
    return 0
