"""Test helper: the assembled pressure system of a scene as a scipy CSR matrix, rebuilt from the dense oracle's outputs
(fractions, row set, diagonal, right-hand side) with the reference's formula for the couplings
(macpressuresolver3.cpp:159-199: off-diagonal -dt*area/(dx^2*rho) between two row cells across an open, wet face).
Rows are numbered in x-fastest order over the row set — the order of the reference's index map for a dense core."""
import numpy as np
import scipy.sparse as sp

from oracle import dense_oracle


def pressure_system(sc, **kw):
    ref = dense_oracle.project(sc, max_iterations=0, **kw)
    rows = ref.in_rows.astype(bool)
    nz, ny, nx = rows.shape
    index = -np.ones(rows.shape, dtype=np.int64)
    index[rows] = np.arange(int(rows.sum()))
    scale = sc.dt / (sc.dx * sc.dx)
    I, J, V = [index[rows]], [index[rows]], [ref.diag[rows]]
    for dim, (dk, dj, di) in enumerate(((0, 0, 1), (0, 1, 0), (1, 0, 0))):
        area, rho = ref.areas[dim], ref.rhos[dim]
        hi = (slice(dk, None), slice(dj, None), slice(di, None))                       # cell c
        lo = (slice(0, nz - dk), slice(0, ny - dj), slice(0, nx - di))                 # cell c - e
        a, r = area[hi][:nz - dk, :ny - dj, :nx - di], rho[hi][:nz - dk, :ny - dj, :nx - di]   # the face between them
        both = rows[hi] & rows[lo] & (a != 0) & (r != 0)
        w = -scale * a[both] / r[both]
        I += [index[hi][both], index[lo][both]]
        J += [index[lo][both], index[hi][both]]
        V += [w, w]
    n = int(rows.sum())
    A = sp.csr_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(n, n))
    A.sort_indices()
    return A, ref.rhs[rows].copy(), ref
