"""Host mirror of the reference's `LinSolver` plug-in point for assembled systems
(RCMatrix_solver_interface::solve, include/shiokaze/linsolver/RCMatrix_solver.h:77; modules src/linsolver/pcg.cpp, cg.cpp):
`B200CG(Residual=..., MaxIterations=...).solve(A, b)` -> (x, Result) with the reference's flag names and the reference's
Result{count, reresid}. It only marshals CSR arrays into the C-ABI (shkz_b200_csr_*): no compute here, no CPU fallback."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class LinResult:
    count: int          # RCMatrix_solver_interface::Result::count
    reresid: float      # ... ::reresid
    converged: bool
    stats: dict = field(default_factory=dict)


class B200CG:
    """`LinSolver=b200cg`. Flags of the reference's pcg module keep their names (pcg.cpp:39-44): Residual, MaxIterations;
    ModifiedIC / MinDiagRatio are accepted and ignored, as the reference's result ignores them (pcg_solver.h:383).
    Additive: Precond=none|jacobi, GPU=<device>."""

    def __init__(self, device: int = 0, **flags):
        self.params = capi.CsrParams()
        capi.lib().shkz_b200_csr_default_params(C.byref(self.params))
        self._h = C.c_void_p()
        self.configure(**flags)
        capi.check_csr(capi.lib().shkz_b200_csr_create(int(flags.get("GPU", device)), C.byref(self._h)))

    def configure(self, **flags):
        p = self.params
        for key, value in flags.items():
            if key == "Residual": p.residual = float(value)
            elif key == "MaxIterations": p.max_iterations = int(value)
            elif key in ("ModifiedIC", "MinDiagRatio", "GPU"): pass
            elif key == "Precond":
                if value not in ("none", "jacobi"):
                    raise ValueError(f"Precond={value!r}: none or jacobi")
                p.precond = capi.CSR_PRECOND_JACOBI if value == "jacobi" else capi.CSR_PRECOND_NONE
            elif key == "CheckEvery": p.check_every = int(value)
            else:
                raise ValueError(f"unknown flag {key}")

    def solve(self, rowptr, col, val, b):
        """A in CSR (rowptr int64[n+1], col int32[nnz], val float64[nnz]) or a scipy.sparse matrix passed as `rowptr` with col = val = None."""
        if col is None and hasattr(rowptr, "tocsr"):
            A = rowptr.tocsr()
            A.sort_indices()
            rowptr, col, val = A.indptr, A.indices, A.data
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        n = b.shape[0]
        if rowptr.shape[0] != n + 1 or col.shape[0] != val.shape[0]:
            raise ValueError("inconsistent CSR arrays")
        x = np.zeros(n, dtype=np.float64)
        st = capi.CsrStats()
        capi.check_csr(capi.lib().shkz_b200_csr_solve_host(self._h, n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data,
                                                          x.ctypes.data, C.byref(self.params), C.byref(st)))
        return x, LinResult(int(st.iterations), float(st.reresid), bool(st.converged), st.asdict())

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().shkz_b200_csr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
