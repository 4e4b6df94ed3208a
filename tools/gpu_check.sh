#!/bin/bash
# validation + timing of the current build: sanitizer on small cases, GPU tests, benches
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'.')
from shiokaze_b200 import MacPressureSolver3, scenes
import numpy as np
for prec in ('fp64','mixed','fp32'):
    for sc in (scenes.random_blobs(20,14,18,seed=3), scenes.dambreak(40, True), scenes.random_blobs(70,18,37,seed=5)):
        S = MacPressureSolver3((sc.nx,sc.ny,sc.nz), sc.dx, Precision=prec, Residual=1e-5)
        out = S.project_scene(sc, surface_tension=0.01); print(prec, sc.name, out['result'].iterations, out['result'].converged, out['result'].reresid)
        if prec != 'fp64':
            a = S.debug_vcycle(False); b = S.debug_vcycle(True); print('   vcycle equal', np.array_equal(a,b), float(np.abs(a-b).max()))
        S.close()
" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -25 gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for w in "smoke_plume 512" "dambreak_solid 512" "smoke_plume 256"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $w rc=$?"; tail -c 1800 gpurun_out/bench_$1_$2.json; tail -3 gpurun_out/bench_$1_$2.err
done
