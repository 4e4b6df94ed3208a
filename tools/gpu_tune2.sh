#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python tools/gpu_tune2.py smoke_plume:256 smoke_plume:512 dambreak_solid:512 flip_splash:512 liquid_box:256 > gpurun_out/tune2.log 2>&1; echo "tune rc=$?"; cat gpurun_out/tune2.log
timeout 300 python bench.py --workload smoke_plume --n 512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/q512.json 2>gpurun_out/q512.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/q512.json") if l.startswith("{")][-1]); print(d["ms_per_step"], d["solve"]); print(d["roofline"]["by_kernel_ms"])
PY
