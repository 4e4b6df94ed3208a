"""TEST INFRASTRUCTURE — numpy restatement of the reference's linear solver on an assembled matrix, used only by tests/
as the checker of shkz_b200_csr_solve_host. Never imported by the product.

Follows local/include/pcgsolver/pcg_solver.h:246-295 (PCGSolver::solve) as it effectively runs: apply_preconditioner copies
its MIC(0) result over with z = r (pcg_solver.h:374-383), so the iteration is plain CG with
  tol = tolerance_factor * |b|_inf (:250-258), the test |r|_inf <= tol after the x / r update (:280-285),
  iterations_out = iteration + 1, residual_out = |r|_inf / |b|_inf,
  rho = z.r with the early exits rho == 0 / NaN (:263-271), s = z + beta s (:286-289).
Pinned by tests/test_oracle.py::test_csr_cg_oracle_matches_the_reference_build against iteration counts and solutions the
unmodified reference build produced (tests/golden/*.npz hold them per scene)."""
import numpy as np


def cg(A, b, residual=1e-4, max_iterations=30000, jacobi=False):
    """A: scipy.sparse CSR (SPD), b: float64[n]. Returns (x, iterations, reresid, converged)."""
    b = np.asarray(b, dtype=np.float64)
    n = b.shape[0]
    x = np.zeros(n)
    r = b.copy()                                        # :249-250
    res0 = float(np.abs(r).max()) if n else 0.0
    if res0 == 0.0:                                     # :253-256
        return x, 0, 0.0, True
    tol = max(residual, 1e-30) * res0                   # :239, :258
    invd = np.ones(n)
    if jacobi:
        d = A.diagonal()
        invd = np.where(d > 0, 1.0 / np.where(d > 0, d, 1.0), 1.0)
    z = r * invd                                        # :260-261 (MIC(0) result discarded, :383)
    rho = float(z @ r)                                  # :262
    if rho == 0.0 or rho != rho:                        # :263-271
        return x, 0, 1.0, False
    s = z.copy()                                        # :272
    it = 0
    for it in range(max_iterations):                    # :275
        q = A @ s                                       # :276
        alpha = rho / float(s @ q)                      # :277
        x += alpha * s                                  # :278
        r -= alpha * q                                  # :279
        res = float(np.abs(r).max())                    # :280
        if res <= tol:                                  # :281-285
            return x, it + 1, res / res0, True
        z = r * invd                                    # :286
        rho_new = float(z @ r)                          # :287
        s = z + (rho_new / rho) * s                     # :288-289
        rho = rho_new
    return x, max_iterations, float(np.abs(r).max()) / res0, False   # :291-293
