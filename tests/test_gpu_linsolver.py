"""`LinSolver=b200cg`: CG on an assembled CSR system through the C-ABI (shkz_b200_csr_*), checked against
the numpy restatement of the reference's solver (oracle/csr_cg_oracle.py) and — through the reference's own loader and its own
macpressuresolver3 assembly — against the reference's pcg module."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import csr_util
from conftest import golden_cases, load_golden, rel_l2
from oracle import csr_cg_oracle, refio
from shiokaze_b200 import B200CG, capi, scenes

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", ["dambreak24", "dambreak_solid24", "smoke16", "flip32", "blobs"])
@pytest.mark.parametrize("residual", [1e-4, 1e-10])
def test_csr_cg_matches_oracle_on_pressure_systems(cuda_device, name, residual):
    make, _ = golden_cases()[name]
    A, b, _ = csr_util.pressure_system(make(), real_is_double=True)
    S = B200CG(Residual=residual)
    x, res = S.solve(A, None, None, b)
    xo, it, rr, conv = csr_cg_oracle.cg(A, b, residual=residual)
    assert res.converged and conv and res.stats["ell_width"] == 7
    assert abs(res.count - it) <= max(2, 0.02 * it)                     # summation order differs, nothing else
    assert res.reresid <= residual
    assert rel(x, xo) < (1e-8 if residual < 1e-8 else 5e-3)
    assert float(np.abs(b - A @ x).max()) <= 1.0001 * residual * float(np.abs(b).max()) + 1e-300
    if residual < 1e-8:                                                 # ... and the reference build's own pressure
        g = load_golden(name, "f64_tight")
        rows = g["pressure_active"].astype(bool)
        assert rel(x, g["pressure"][rows]) < 1e-8
        assert abs(res.count - g["iterations"]) <= max(3, 0.03 * g["iterations"])
    S.close()


def test_csr_cg_wide_rows_jacobi_and_edge_cases(cuda_device):
    rng = np.random.default_rng(5)
    n = 3000
    B = sp.random(n, n, density=0.012, random_state=7, format="csr")
    A = (B @ B.T + sp.diags(1.0 + rng.random(n))).tocsr()              # SPD, ~100 entries per row -> CSR path, one warp per row
    A.sort_indices()
    b = rng.standard_normal(n)
    S = B200CG(Residual=1e-9)
    x, res = S.solve(A, None, None, b)
    xo, it, rr, conv = csr_cg_oracle.cg(A, b, residual=1e-9)
    assert res.stats["ell_width"] == 0 and res.converged and abs(res.count - it) <= max(2, 0.02 * it)
    assert rel(x, xo) < 1e-7
    # Jacobi scaling (additive flag): same solution, oracle follows with the same preconditioner
    S.configure(Precond="jacobi")
    xj, resj = S.solve(A, None, None, b)
    xjo, itj, _, _ = csr_cg_oracle.cg(A, b, residual=1e-9, jacobi=True)
    assert resj.converged and abs(resj.count - itj) <= max(2, 0.02 * itj) and rel(xj, xjo) < 1e-7
    S.configure(Precond="none")
    # pcg_solver.h:253-256: zero right-hand side -> 0 iterations, x = 0
    x0, r0 = S.solve(A, None, None, np.zeros(n))
    assert r0.count == 0 and r0.reresid == 0.0 and not x0.any()
    # MaxIterations = 0: nothing runs, reresid = 1
    S.configure(MaxIterations=0)
    x1, r1 = S.solve(A, None, None, b)
    assert r1.count == 0 and r1.reresid == 1.0 and not r1.converged and not x1.any()
    # MaxIterations reached: count == MaxIterations, not converged (pcg_solver.h:291-293)
    S.configure(MaxIterations=3)
    x3, r3 = S.solve(A, None, None, b)
    x3o, it3, rr3, conv3 = csr_cg_oracle.cg(A, b, residual=1e-9, max_iterations=3)
    assert r3.count == 3 and not r3.converged and rel(x3, x3o) < 1e-10 and abs(r3.reresid - rr3) < 1e-10 * rr3
    # 1x1 and diagonal systems, ragged rows (an empty row would be singular: rows of length 1 and 3 mixed)
    S.configure(MaxIterations=100)
    xd, rd = S.solve(sp.csr_matrix(np.array([[4.0]])), None, None, np.array([2.0]))
    assert rd.count == 1 and abs(xd[0] - 0.5) < 1e-15
    T = sp.diags([[-1.0] * 5 + [0.0] * 4, [2.0] * 10, [-1.0] * 5 + [0.0] * 4], [-1, 0, 1]).tocsr()
    T.eliminate_zeros()
    bt = np.arange(1.0, 11.0)
    xt, rt = S.solve(T, None, None, bt)
    assert rt.converged and rel(xt, np.linalg.solve(T.toarray(), bt)) < 1e-8
    S.close()
    with pytest.raises(ValueError):
        B200CG(Precond="ilu")


def have_module(real):
    return refio.ref_available(real) and os.path.isfile(os.path.join(refio.ref_dir(real), "libshiokaze_b200cg.so"))


@pytest.mark.parametrize("real,tol", [("f32", 1e-3), ("f64", 1e-5)])
@pytest.mark.parametrize("scene", ["dambreak_solid", "flip"])
def test_linsolver_module_is_a_drop_in(cuda_device, real, tol, scene):
    """The reference's own macpressuresolver3 (its assembly, its velocity update) with LinSolver=b200cg instead of pcg."""
    if not have_module(real):
        pytest.skip("oracle/_ref (reference build + module) was not shipped to this box")
    sc = scenes.dambreak(32, True) if scene == "dambreak_solid" else scenes.flip_splash(40)
    flags = {"Residual": 1e-10}
    ref = refio.run_reference(sc, real, flags=flags)
    ours = refio.run_reference(sc, real, flags={**flags, "LinSolver": "b200cg"})
    assert "b200cg.so" in ours.stdout and "macpressuresolver3.so" in ours.stdout
    assert np.array_equal(ours.pressure_active, ref.pressure_active)
    for d in range(3):
        assert np.array_equal(ours.vel_active[d], ref.vel_active[d])
    assert rel_l2(ours.vel, ref.vel) < tol
    assert abs(ours.iterations - ref.iterations) <= max(3, 0.06 * ref.iterations)     # the same CG, another summation order
