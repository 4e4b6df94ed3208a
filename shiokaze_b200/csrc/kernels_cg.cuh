// Conjugate-gradient kernels on the dense matrix-free 7-point operator (replace K9/K10 of SURVEY.md 2d:
// local/include/pcgsolver/sparse_matrix.h:264-275 and the BLAS-1 calls of pcg_solver.h:277-289).
//
//   (A p)[c] = dd[c] p[c] + wx[c] (p[c]-p[c-1]) + wx[c+1] (p[c]-p[c+1]) + wy[c] (p[c]-p[c-nx]) + wy[c+nx] (p[c]-p[c+nx])
//                         + wz[c] (p[c]-p[c-plane]) + wz[c+plane] (p[c]-p[c+plane])
// ("difference form": w = face coupling between two row cells, dd = Dirichlet part of the diagonal).
// w = 0 on walls / towards non-row cells and dd = 0 off the row set, so the flat neighbour offsets
// never need bounds logic (wrapped reads hit finite values times a zero coefficient), and constants are
// annihilated exactly on pure-Neumann rows even with float coefficients.
//
// Thread layout for stencil kernels: a CTA owns a TX x TY column of cells and marches ZC planes in z
// keeping the z-neighbours in registers; x/y neighbours are re-read through L1.
#pragma once
#include "common.cuh"

namespace shkz {

constexpr int TX = 64, TY = 4, ZC = 16;

inline dim3 stencil_grid(const Dims &d) { return dim3((d.nx + TX - 1) / TX, (d.ny + TY - 1) / TY, (d.nzl + ZC - 1) / ZC); }
inline dim3 stencil_block() { return dim3(TX, TY, 1); }

// z = A s,  sz = s.z ; last block: alpha = rho / sz            (pcg_solver.h:276-277)
template <class VecT, class CoefT>
__global__ void __launch_bounds__(TX *TY) k_spmv_dot(Dims d, const CoefT *__restrict__ wx, const CoefT *__restrict__ wy, const CoefT *__restrict__ wz,
                                                    const CoefT *__restrict__ dd, const VecT *__restrict__ s, VecT *__restrict__ z, RedBuf rb, CGState *st) {
	if (st->done) return;
	const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
	const int kbeg = blockIdx.z * ZC, kend = min(kbeg + ZC, d.nzl);
	double red[1] = {0.0};
	if (i < d.nx && j < d.ny) {
		long long c = i + (long long)d.nx * (j + (long long)d.ny * kbeg);
		VecT sm = s[c - d.plane], sc = s[c];
		CoefT wzc = wz[c];
		for (int k = kbeg; k < kend; ++k, c += d.plane) {
			const VecT sp = s[c + d.plane];
			const CoefT wzp = wz[c + d.plane];
			VecT v = (VecT)dd[c] * sc;
			v += (VecT)wx[c] * (sc - s[c - 1]);
			v += (VecT)wx[c + 1] * (sc - s[c + 1]);
			v += (VecT)wy[c] * (sc - s[c - d.nx]);
			v += (VecT)wy[c + d.nx] * (sc - s[c + d.nx]);
			v += (VecT)wzc * (sc - sm);
			v += (VecT)wzp * (sc - sp);
			z[c] = v;
			red[0] += (double)sc * (double)v;
			sm = sc; sc = sp; wzc = wzp;
		}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->sz = t[0];
		st->alpha = st->rho / t[0];
	});
}

// x += alpha s ; r -= alpha z ; |r|_inf ; r.r            (pcg_solver.h:278-285)
// last block: convergence test, iteration count; for plain CG also beta and the new rho.
template <class VecT, bool PLAIN>
__global__ void __launch_bounds__(256) k_axpy2_norm(long long n, const VecT *__restrict__ s, const VecT *__restrict__ z, VecT *__restrict__ x, VecT *__restrict__ r,
                                                   RedBuf rb, CGState *st) {
	if (st->done) return;
	const VecT alpha = (VecT)st->alpha;
	double red[2] = {0.0, 0.0};
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
		x[c] += alpha * s[c];
		const VecT rv = r[c] - alpha * z[c];
		r[c] = rv;
		red[0] = fmax(red[0], fabs((double)rv));
		if (PLAIN) red[1] += (double)rv * (double)rv;
	}
	grid_reduce<2, 0x1u>(red, rb, [&](double (&t)[2]) {
		st->rnorm = t[0];
		st->iter += 1;
		if (t[0] <= st->tol) { st->done = 1; st->converged = 1; }
		else if (st->iter >= st->max_iter) st->done = 1;
		if (PLAIN) { st->beta = t[1] / st->rho; st->rho = t[1]; }
	});
}

// s = z + beta s   (pcg_solver.h:289; plain CG passes z = r)
template <class VecT>
__global__ void __launch_bounds__(256) k_xpay(long long n, const VecT *__restrict__ z, VecT *__restrict__ s, const CGState *__restrict__ st) {
	if (st->done) return;
	const VecT beta = (VecT)st->beta;
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
		s[c] = z[c] + beta * s[c];
}

// start of the solve: x = 0, s = z (= r for plain CG), rho = z.r         (pcg_solver.h:249-272)
// r already holds b. FIRST: tol and the trivial-rhs exit.
template <class VecT>
__global__ void __launch_bounds__(256) k_cg_begin(long long n, double residual, int max_iter, CGState *st) {
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const double factor = residual < 1e-30 ? 1e-30 : residual; // pcg_solver.h:239
		st->tol = factor * st->bnorm;
		st->iter = 0;
		st->max_iter = max_iter;
		st->converged = st->bnorm == 0.0 ? 1 : 0;
		st->done = (st->bnorm == 0.0 || max_iter <= 0) ? 1 : 0;
		st->rnorm = st->bnorm;
		st->alpha = st->beta = st->sz = 0.0;
	}
}

template <class VecT>
__global__ void __launch_bounds__(256) k_copy_dot(long long n, const VecT *__restrict__ z, const VecT *__restrict__ r, VecT *__restrict__ s, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
		const VecT zv = z[c];
		s[c] = zv;
		red[0] += (double)zv * (double)r[c];
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->rho = t[0];
		if (t[0] == 0.0 || t[0] != t[0]) st->done = 1; // pcg_solver.h:263-271
	});
}

// After a preconditioner application: beta = (z.r)_new / rho, rho = (z.r)_new   (pcg_solver.h:287-290)
__global__ void k_beta_from_zr(CGState *st) {
	if (st->done) return;
	st->beta = st->rr / st->rho;
	st->rho = st->rr;
}

} // namespace shkz
