import json
import os
import sys

import numpy as np
import pytest

# one process driving several GPUs from several threads (shiokaze_b200/dist.py: run_per_slab) must not load kernels lazily in the middle of a solve whose
# kernels wait for each other across devices (see shiokaze_b200/plugin/b200dense.h); set before anything initialises CUDA
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
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


# ---- fixtures at BASELINE sizes, generated from the unmodified reference build by make_golden.py --big ----------------
def _big_cases():
    from shiokaze_b200 import scenes
    return {
        "dambreak64": lambda: scenes.dambreak(64),               # configs[0]
        "flip64": lambda: scenes.flip_splash(64),                # configs[3] geometry
        "smoke64": lambda: scenes.smoke_plume(64),               # configs[1] geometry
        "dambreak_solid128": lambda: scenes.dambreak(128, True), # configs[2] geometry
        "smoke128": lambda: scenes.smoke_plume(128),
    }


class _Lazy(dict):
    def __missing__(self, key):
        self.update(_big_cases())
        return dict.__getitem__(self, key)

    def __iter__(self):
        if not len(self):
            self.update(_big_cases())
        return dict.__iter__(self)


BIG_CASES = _Lazy()
BIG_NAMES = ["dambreak64", "flip64", "smoke64", "dambreak_solid128", "smoke128"]
BIG_SAMPLE_CAP = 100_000


def big_selection(mask):
    """Flat indices of the sampled entries of a (nz, ny, nx) field: the entries where `mask` holds on every s-th z-plane,
    s chosen so that about BIG_SAMPLE_CAP values remain (all of them when the mask is that small). Deterministic in the mask."""
    m = np.asarray(mask).astype(bool)
    total = int(m.sum())
    s = max(1, -(-total // BIG_SAMPLE_CAP))
    keep = np.zeros(m.shape[0], dtype=bool)
    keep[s // 2::s] = True
    return np.flatnonzero(m & keep[:, None, None])


def load_big_golden(name, tag, scene):
    """-> dict(act[3], pressure_active: complete bool arrays; vel[3], pressure: float32 samples at big_selection(...);
    vel_sum[3], vel_sumsq[3], pressure_sumsq, n_rows, iterations, reresid)."""
    z = np.load(os.path.join(GOLDEN, "big_" + name + ".npz"))
    g = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + ".")}
    shp = (scene.nz, scene.ny, scene.nx)
    fshp = [a.shape for a in scene.vel_active]

    def unpack(bits, shape):
        return np.unpackbits(bits)[:int(np.prod(shape))].reshape(shape).astype(np.uint8)
    return dict(act=[unpack(g[f"act{d}"], fshp[d]) for d in range(3)], pressure_active=unpack(g["pressure_active"], shp),
                vel=[g[f"vel{d}"] for d in range(3)], pressure=g["pressure"],
                vel_sum=[float(g[f"vel{d}_sum"]) for d in range(3)], vel_sumsq=[float(g[f"vel{d}_sumsq"]) for d in range(3)],
                pressure_sumsq=float(g["pressure_sumsq"]), n_rows=int(g["n_rows"]), iterations=int(g["iterations"]), reresid=float(g["reresid"]))


def big_compare(out_vel, out_act, out_pact, g, scene):
    """(masks equal?, rel. L2 of the sampled velocities, rel. error of the whole-field sum of squares)."""
    masks = all(np.array_equal(np.asarray(out_act[d]).astype(np.uint8), g["act"][d]) for d in range(3)) and \
        np.array_equal(np.asarray(out_pact).astype(np.uint8), g["pressure_active"])
    samples = [np.asarray(out_vel[d]).ravel()[big_selection(scene.vel_active[d])] for d in range(3)]
    rel = rel_l2(samples, g["vel"])
    ssq = sum(float((np.asarray(v, dtype=np.float64) ** 2).sum()) for v in out_vel)
    gsq = sum(g["vel_sumsq"])
    return masks, rel, abs(ssq - gsq) / gsq if gsq > 0 else abs(ssq)


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
