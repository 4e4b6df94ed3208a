#!/bin/bash
# first GPU contact: smoke, sanitizer on a tiny case, parity tests, timing probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'.')
from shiokaze_b200 import MacPressureSolver3, scenes
for prec in ('fp64','mixed','fp32'):
    sc = scenes.random_blobs(20,14,18,seed=3)
    S = MacPressureSolver3((sc.nx,sc.ny,sc.nz), sc.dx, Precision=prec, Residual=1e-5)
    out = S.project_scene(sc, surface_tension=0.01); print(prec, out['result'])
    S.close()
" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/sanitizer.log
tail -5 gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python tools/gpu_probe.py smoke_plume 256 > gpurun_out/probe_smoke256.log 2>&1; tail -12 gpurun_out/probe_smoke256.log
