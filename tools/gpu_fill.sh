#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_advect.py -q -x -k "drop_in or module or zero_copy" --timeout 100 2>&1 | tail -2
timeout 100 python - <<'PY' 2>&1 | tee gpurun_out/r02_module_timing_tiled_dambreak256.txt | cut -c1-200
import os, re, sys, dataclasses, importlib.util
sys.path.insert(0, ".")
from oracle import refio
from shiokaze_b200 import scenes
spec = importlib.util.spec_from_file_location("bench", "bench.py"); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
sc = scenes.BENCH_SCENES["dambreak_solid"](256)
r = refio.run_reference(sc, "f32", projection="b200pressure3", repeat=3, threads=os.cpu_count())
print("project() through b200pressure3 on tiledarray3, dam-break + obstacle 256^3:", re.search(r"project_ms_last=([0-9.]+)", r.stdout).group(1), "ms")
for line in r.stdout.splitlines()[-12:]:
    if any(k in line for k in ("Gathering", "Solving on", "Scattering")): print("   ", line.strip())
r = refio.run_reference(dataclasses.replace(sc, dt=bench.advect_dt(sc)), "f32", advect="vector", advection="b200advection3", repeat=3, threads=os.cpu_count())
print("advect_vector() through b200advection3 on tiledarray3:", re.search(r"project_ms_last=([0-9.]+)", r.stdout).group(1), "ms")
PY
