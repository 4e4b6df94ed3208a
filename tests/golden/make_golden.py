"""Regenerates tests/golden/*.npz and accuracytest3.json from the UNMODIFIED reference build
(oracle/_ref, made by `make -C oracle ref` from /root/reference). Run in the build container only:

    python tests/golden/make_golden.py                 the small scenes (complete dense outputs)
    python tests/golden/make_golden.py --big [names]   BASELINE sizes: dam-break 64^3 (configs[0]), smoke 64^3 / 128^3,
                                                       dam-break + solid 128^3, FLIP splash 64^3 (compact: big_*.npz)

Every fixture is one project() call of the reference's macpressuresolver3 + pcg on a scene of
shiokaze_b200.scenes (regenerated from the formula at test time, so only OUTPUTS are stored).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402
from shiokaze_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (scene factory, kwargs)
    "dambreak24": (lambda: scenes.dambreak(24), {}),
    "dambreak_solid24": (lambda: scenes.dambreak(24, True), {}),
    "smoke16": (lambda: scenes.smoke_plume(16), {}),
    "flip32": (lambda: scenes.flip_splash(32), {}),
    "box16": (lambda: scenes.liquid_box(16), {}),
    "blobs": (lambda: scenes.random_blobs(20, 14, 18, seed=3), {}),
    "blobs_nosolid": (lambda: scenes.random_blobs(13, 21, 10, seed=5, with_solid=False), {}),
    "dambreak24_firstorder": (lambda: scenes.dambreak(24, True), {"SecondOrderAccurateFluid": "No", "SecondOrderAccurateSolid": "No"}),
    "dambreak24_tension": (lambda: scenes.dambreak(24), {"surface_tension": 0.05}),
    "dambreak24_volume": (lambda: scenes.dambreak(24), {"volume": (1.05, 1.0)}),
}


def run(name, real, residual):
    make, kw = CASES[name]
    sc = make()
    flags = {"Residual": residual}
    flags.update({k: v for k, v in kw.items() if k not in ("surface_tension", "volume")})
    sc.surface_tension = kw.get("surface_tension", 0.0)
    cur, tgt = kw.get("volume", (0.0, 0.0))
    r = refio.run_reference(sc, real, flags=flags, current_volume=cur, target_volume=tgt)
    dt = np.float32 if real == "f32" else np.float64
    out = {f"vel{d}": r.vel[d].astype(dt) for d in range(3)}
    out.update({f"act{d}": r.vel_active[d] for d in range(3)})
    out.update(pressure=r.pressure.astype(dt), pressure_active=r.pressure_active,
               iterations=np.int64(r.iterations), reresid=np.float64(r.reresid))
    return out


# ---- fixtures at BASELINE sizes (configs[0] and the 64^3 / 128^3 rungs of the others): compact storage ----------------
# masks bit-packed (complete), velocity / pressure as float32 SAMPLES on whole z-planes of the input-active faces /
# of the row set (conftest.big_selection: at most ~100 k values per array), plus float64 sum and sum of squares of every
# complete field, so that a test sees both the local values and the whole field.
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import BIG_CASES, big_selection  # noqa: E402


def run_big(name, real, residual, extra_flags=None, repeat=1):
    sc = BIG_CASES[name]()
    flags = {"Residual": residual}
    flags.update(extra_flags or {})
    r = refio.run_reference(sc, real, flags=flags, threads=os.cpu_count(), repeat=repeat)
    out = {}
    for d in range(3):
        sel = big_selection(sc.vel_active[d])
        v = r.vel[d]
        out[f"vel{d}"] = v.ravel()[sel].astype(np.float32)
        out[f"vel{d}_sum"] = np.float64(v.sum())
        out[f"vel{d}_sumsq"] = np.float64((v * v).sum())
        out[f"act{d}"] = np.packbits(r.vel_active[d].astype(bool))
    rows = r.pressure_active.astype(bool)
    out["pressure_active"] = np.packbits(rows)
    out["pressure"] = r.pressure.ravel()[big_selection(rows)].astype(np.float32)
    out["pressure_sumsq"] = np.float64((r.pressure * r.pressure).sum())
    out["n_rows"] = np.int64(rows.sum())
    out["iterations"] = np.int64(r.iterations)
    out["reresid"] = np.float64(r.reresid)
    return out


def main_big(only=None):
    assert refio.ref_available("f32") and refio.ref_available("f64")
    for name in BIG_CASES:
        if only and name not in only:
            continue
        blob = {}
        for real, residual, tag in (("f32", 1e-4, "f32_default"), ("f32", 1e-10, "f32_tight"), ("f64", 1e-10, "f64_tight")):
            for k, v in run_big(name, real, residual).items():
                blob[f"{tag}.{k}"] = v
        if name == "dambreak64":
            # a10 warm start (macpressuresolver3.cpp:221-242): the SAME inputs projected twice with WarmStart=Yes, results of the second call
            for k, v in run_big(name, "f32", 1e-4, {"WarmStart": "Yes"}, repeat=2).items():
                blob[f"f32_warm2.{k}"] = v
        np.savez_compressed(os.path.join(HERE, "big_" + name + ".npz"), **blob)
        print(name, "iterations", {t: int(blob[t + ".iterations"]) for t in sorted({k.split(".")[0] for k in blob})},
              "rows", int(blob["f32_tight.n_rows"]), flush=True)


def main():
    assert refio.ref_available("f32") and refio.ref_available("f64")
    for name in CASES:
        blob = {}
        for real, residual, tag in (("f32", 1e-4, "f32_default"), ("f32", 1e-10, "f32_tight"), ("f64", 1e-10, "f64_tight")):
            for k, v in run(name, real, residual).items():
                blob[f"{tag}.{k}"] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, "iterations", {t: int(blob[t + ".iterations"]) for t in ("f32_default", "f32_tight", "f64_tight")})
    # the reference's own known-answer test (src/examples/accuracytest3-example.cpp): 9 radii per resolution
    acc = {}
    for n in (8, 16, 32):
        worst, iters = 0.0, {}
        for q in [q for q in range(-4, 5) if q] + [0]:
            sc = scenes.accuracy_sphere(n, q)
            r = refio.run_reference(sc, "f32", flags={"Residual": 1e-18, "EpsFluid": 1e-18})
            c = (np.arange(n) + .5) * sc.dx
            exact = (c[None, None, :] - .5) ** 2 + (c[None, :, None] - .5) ** 2 + (c[:, None, None] - .5) ** 2 - sc.meta["r"] ** 2
            worst = max(worst, float(np.abs(exact - r.pressure)[r.pressure_active > 0].max()))
            iters[str(q)] = r.iterations
        acc[str(n)] = {"max_norm": worst, "iterations": iters}
        print("accuracytest3", n, worst)
    acc["survey_goldens"] = {"8": 2.418906e-03, "16": 6.949497e-04, "32": 2.086717e-04, "64": 5.556508e-05}
    with open(os.path.join(HERE, "accuracytest3.json"), "w") as f:
        json.dump(acc, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--big":
        main_big(sys.argv[2:])
    else:
        main()
