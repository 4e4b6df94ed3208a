"""Regenerates tests/golden/advect_*.npz from the UNMODIFIED reference build (oracle/_ref, made by `make -C oracle ref` from /root/reference): the outputs of
its macadvection3 module (src/advection/macadvection3.cpp), called through its own loader by oracle/ref_driver (RefAdvect=vector|density|levelset), on the
scenes of tests/advect_util.py (regenerated from their formulas at test time, so only OUTPUTS are stored — the values on the active entries, as Real=float).
Run in the build container only:    python tests/golden/make_golden_advect.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from advect_util import GOLDEN_FLAGS, advect_scenes, density_of, flag_key, fluid_active  # noqa: E402
from oracle import refio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for name, make in advect_scenes().items():
        sc = make()
        out = {}
        for flags in GOLDEN_FLAGS:
            key = flag_key(flags)
            r = refio.run_reference(sc, "f32", flags=flags, advect="vector")
            for d in range(3):
                assert np.array_equal(r.vel_active[d] != 0, sc.vel_active[d] != 0)
                out[f"{key}/vector{d}"] = r.vel[d][sc.vel_active[d] != 0].astype(np.float32)
            r = refio.run_reference(sc, "f32", flags=flags, advect="density")
            out[f"{key}/density"] = r.pressure[density_of(sc)[1] != 0].astype(np.float32)
            if sc.fluid_raw is not None:
                r = refio.run_reference(sc, "f32", flags=flags, advect="levelset")
                out[f"{key}/levelset"] = r.pressure[fluid_active(sc) != 0].astype(np.float32)
        path = os.path.join(HERE, f"advect_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
