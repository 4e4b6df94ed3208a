"""Design prototype (numpy, CPU): aggregation multigrid-preconditioned CG on the dense matrix-free
7-point operator. Used to choose the V-cycle layout before writing the CUDA kernels; not shipped,
not imported by the package."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from shiokaze_b200 import scenes
from oracle import dense_oracle


def build_fine(sc, o):
    nx, ny, nz = sc.nx, sc.ny, sc.nz
    R = o.in_rows.astype(bool)
    W = []
    for dim in range(3):
        a, r = o.areas[dim], o.rhos[dim]
        with np.errstate(divide='ignore', invalid='ignore'):
            w = np.where((a != 0) & (r != 0), sc.dt * a / (sc.dx * sc.dx * r), 0.0)
        # lower-face array, cell shaped: face index == cell index along dim (drop the last face)
        sl = [slice(None)] * 3
        sl[2 - dim] = slice(0, -1)
        w = w[tuple(sl)].copy()
        # zero at wall (index 0) and unless both cells in R
        lo = np.zeros_like(R)
        s_hi = [slice(None)] * 3; s_lo = [slice(None)] * 3
        s_hi[2 - dim] = slice(1, None); s_lo[2 - dim] = slice(0, -1)
        lo[tuple(s_hi)] = R[tuple(s_lo)]
        w = np.where(R & lo, w, 0.0)
        W.append(w)
    return dict(wx=W[0], wy=W[1], wz=W[2], diag=o.diag.copy())


def shift(a, dim, d):
    """a shifted so result[c] = a[c + d*e_dim], zero outside."""
    out = np.zeros_like(a)
    ax = 2 - dim
    src = [slice(None)] * 3; dst = [slice(None)] * 3
    if d == 1:
        src[ax] = slice(1, None); dst[ax] = slice(0, -1)
    else:
        src[ax] = slice(0, -1); dst[ax] = slice(1, None)
    out[tuple(dst)] = a[tuple(src)]
    return out


def offdiag(L, p):
    s = np.zeros_like(p)
    for dim, key in enumerate(('wx', 'wy', 'wz')):
        w = L[key]
        s += w * shift(p, dim, -1) + shift(w, dim, 1) * shift(p, dim, 1)
    return s


def apply(L, p):
    return L['diag'] * p - offdiag(L, p)


def coarsen(L, scale):
    nz, ny, nx = L['diag'].shape
    cz, cy, cx = (nz + 1) // 2, (ny + 1) // 2, (nx + 1) // 2
    def pad(a):
        return np.pad(a, ((0, 2 * cz - nz), (0, 2 * cy - ny), (0, 2 * cx - nx)))
    wx, wy, wz, dg = pad(L['wx']), pad(L['wy']), pad(L['wz']), pad(L['diag'])
    def blk(a):
        return a.reshape(cz, 2, cy, 2, cx, 2)
    Wx = blk(wx)[:, :, :, :, :, 0].sum(axis=(1, 3))
    Wy = blk(wy)[:, :, :, 0, :, :].sum(axis=(1, 4))
    Wz = blk(wz)[:, 0, :, :, :, :].sum(axis=(2, 4))
    internal = blk(wx)[:, :, :, :, :, 1].sum(axis=(1, 3)) + blk(wy)[:, :, :, 1, :, :].sum(axis=(1, 4)) + blk(wz)[:, 1, :, :, :, :].sum(axis=(2, 4))
    D = blk(dg).sum(axis=(1, 3, 5)) - 2.0 * internal
    return dict(wx=Wx * scale, wy=Wy * scale, wz=Wz * scale, diag=D * scale)


def restrict(r, cshape):
    nz, ny, nx = r.shape
    cz, cy, cx = cshape
    rp = np.pad(r, ((0, 2 * cz - nz), (0, 2 * cy - ny), (0, 2 * cx - nx)))
    return rp.reshape(cz, 2, cy, 2, cx, 2).sum(axis=(1, 3, 5))


def prolong(e, fshape):
    nz, ny, nx = fshape
    return np.repeat(np.repeat(np.repeat(e, 2, 0), 2, 1), 2, 2)[:nz, :ny, :nx]


def colors(shape):
    nz, ny, nx = shape
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
    return ((i + j + k) & 1) == 0


def rbgs(L, x, b, order, dinv, red, dtype):
    for c in order:
        m = red if c == 0 else ~red
        xn = (b + offdiag(L, x)) * dinv
        x = np.where(m, xn, x).astype(dtype)
    return x


class MG:
    def __init__(self, L0, scale=0.5, nu1=1, nu2=1, min_size=4, coarse_sweeps=8, dtype=np.float32, max_levels=20):
        self.levels = [L0]
        while max(self.levels[-1]['diag'].shape) > min_size and len(self.levels) < max_levels:
            self.levels.append(coarsen(self.levels[-1], scale))
        self.dtype = dtype
        for L in self.levels:
            with np.errstate(divide='ignore'):
                L['dinv'] = np.where(L['diag'] > 0, 1.0 / L['diag'], 0.0).astype(dtype)
            L['red'] = colors(L['diag'].shape)
            for key in ('wx', 'wy', 'wz', 'diag'):
                L[key] = L[key].astype(dtype)
        self.nu1, self.nu2, self.cs = nu1, nu2, coarse_sweeps

    def vcycle(self, b, l=0):
        L = self.levels[l]
        x = np.zeros_like(b)
        if l == len(self.levels) - 1:
            for _ in range(self.cs):
                x = rbgs(L, x, b, (0, 1), L['dinv'], L['red'], self.dtype)
            for _ in range(self.cs):
                x = rbgs(L, x, b, (1, 0), L['dinv'], L['red'], self.dtype)
            return x
        for _ in range(self.nu1):
            x = rbgs(L, x, b, (0, 1), L['dinv'], L['red'], self.dtype)
        r = b - apply(L, x)
        rc = restrict(r, self.levels[l + 1]['diag'].shape).astype(self.dtype)
        ec = self.vcycle(rc, l + 1)
        x = x + prolong(ec, b.shape)
        for _ in range(self.nu2):
            x = rbgs(L, x, b, (1, 0), L['dinv'], L['red'], self.dtype)
        return x


def pcg(L, b, M, tol_rel, maxit=2000, project_mean=False):
    x = np.zeros_like(b)
    r = b.copy()
    R = L['diag'] > 0
    n = R.sum()
    b0 = np.abs(b).max()
    hist = []
    def prec(r):
        if M is None:
            return r.copy()
        z = M.vcycle(r.astype(M.dtype)).astype(np.float64)
        if project_mean:
            z = np.where(R, z - z[R].mean(), 0.0)
        return z
    z = prec(r)
    rho = (z * r).sum()
    s = z.copy()
    for it in range(maxit):
        q = apply(L, s)
        alpha = rho / (s * q).sum()
        x += alpha * s
        r -= alpha * q
        res = np.abs(r).max()
        hist.append(res / b0)
        if res <= tol_rel * b0:
            return x, it + 1, hist
        z = prec(r)
        rho_new = (z * r).sum()
        s = z + (rho_new / rho) * s
        rho = rho_new
    return x, maxit, hist


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'dambreak'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
    sc = scenes.BENCH_SCENES[which](n)
    o = dense_oracle.project(sc, max_iterations=0 if n > 96 else 30000)
    L0 = build_fine(sc, o)
    Ld = {k: v.astype(np.float64) for k, v in L0.items()}
    print(which, n, 'rows', o.n_rows, 'oracle CG iters', o.iterations)
    neumann = not sc.fluid_levelset
    for scale in (0.5, 0.6, 0.75, 1.0):
        for nu in (1, 2):
            t = time.time()
            M = MG({k: v.copy() for k, v in L0.items()}, scale=scale, nu1=nu, nu2=nu)
            x, it, hist = pcg(Ld, o.rhs, M, tol, project_mean=neumann)
            print(f'  scale {scale} nu {nu} levels {len(M.levels)}: MGPCG iters {it}  final {hist[-1]:.2e}  ({time.time()-t:.1f}s)')
