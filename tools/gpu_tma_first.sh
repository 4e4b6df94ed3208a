#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/tma_first.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from shiokaze_b200 import MacPressureSolver3, scenes
for mk, kw in ((lambda: scenes.dambreak(40, True), {}), (lambda: scenes.smoke_plume(64), {}), (lambda: scenes.random_blobs(72,40,21,seed=11), {}), (lambda: scenes.smoke_plume(128), {})):
    sc = mk()
    S = MacPressureSolver3((sc.nx,sc.ny,sc.nz), sc.dx, Precision='mixed', Precond='mg', MaxIterations=1)
    out = S.project_scene(sc)
    print(sc.name, sc.nx, 'project ok', out['result'].iterations, out['result'].reresid, flush=True)
    a = S.debug_vcycle(0); b = S.debug_vcycle(3); c = S.debug_vcycle(1)
    print('   tma==quad', np.array_equal(a,b), 'quad==legacy', np.array_equal(b,c), 'maxdiff', float(np.abs(a-b).max()), 'nan', int(np.isnan(a).sum()), flush=True)
    S.close()
PY
echo "tma_first rc=$?"; tail -20 gpurun_out/tma_first.log
