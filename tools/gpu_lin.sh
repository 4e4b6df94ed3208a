#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linsolver.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_lin.log 2>&1; echo "pytest lin rc=$?"; tail -15 gpurun_out/pytest_lin.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_dot4|k_axpy2_norm|k_xpay" -s 3 -c 3 -o gpurun_out/prof_cg2 -f python tools/profile_step.py smoke_plume 512 > gpurun_out/prof_cg2.log 2>&1; echo "ncu cg rc=$?"
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp
from shiokaze_b200 import B200CG
# 7-point Poisson 128^3 (2.1 M rows), Dirichlet: ms per iteration of the assembled-matrix CG
n = 128; I = sp.identity(n); T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n))
A = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr(); A.sort_indices()
b = np.random.default_rng(0).standard_normal(A.shape[0])
S = B200CG(Residual=1e-6)
for _ in range(2):
    x, r = S.solve(A, None, None, b)
print("poisson128: rows", A.shape[0], "nnz", A.nnz, "iters", r.count, "converged", r.converged, {k: round(v, 3) if isinstance(v, float) else v for k, v in r.stats.items()},
      "ms/iter %.4f" % (r.stats["ms_solve"] / max(r.count, 1)), "true resid", float(np.abs(b - A @ x).max() / np.abs(b).max()))
PY
