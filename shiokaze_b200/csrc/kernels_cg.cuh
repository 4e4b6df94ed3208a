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
// Every kernel is persistent over the level-0 list of active tiles (common.cuh): cells of tiles without
// an unknown are never touched, so a liquid scene costs what its wet tiles cost. Block shape is
// (TX, 8): a thread owns a cell column of the tile, two y sub-rows per tile.
#pragma once
#include "common.cuh"

namespace shkz {

constexpr int CG_BY = 8; // block = (TX, CG_BY)
inline dim3 cg_block() { return dim3(TX, CG_BY, 1); }

// start of the solve: x = 0, r = b, s = 0 (and the float copy of r that feeds multigrid)    (pcg_solver.h:249-251)
template <class VecT, bool WRITE_B0>
__global__ void __launch_bounds__(TX *CG_BY) k_cg_init(Dims d, Tiles T, const VecT *__restrict__ b, VecT *__restrict__ x, VecT *__restrict__ r,
                                                      VecT *__restrict__ s, float *__restrict__ b0) {
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				const VecT bv = b[c];
				x[c] = (VecT)0;
				s[c] = (VecT)0;
				r[c] = bv;
				if (WRITE_B0) b0[c] = (float)bv;
			}
	}
}

// FIRST scalar step: tol and the trivial-rhs exit                                         (pcg_solver.h:251-258)
__global__ void k_cg_begin(double residual, int max_iter, CGState *st) {
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const double factor = residual < 1e-30 ? 1e-30 : residual; // pcg_solver.h:239
		st->tol = factor * st->bnorm;
		st->iter = 0;
		st->max_iter = max_iter;
		st->converged = st->bnorm == 0.0 ? 1 : 0;
		st->done = (st->bnorm == 0.0 || max_iter <= 0) ? 1 : 0;
		st->rnorm = st->bnorm;
		st->sum_x = 0.0;
		st->alpha = st->beta = st->sz = st->rho = 0.0;
	}
}

// plain CG only: rho = r.r before the first iteration                                      (pcg_solver.h:262)
template <class VecT>
__global__ void __launch_bounds__(TX *CG_BY) k_dot_rr(Dims d, Tiles T, const VecT *__restrict__ r, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
				const double v = (double)r[i + (long long)d.nx * (j + (long long)d.ny * k)];
				red[0] += v * v;
			}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->rho = t[0];
		st->beta = 0.0;
		if (t[0] == 0.0 || t[0] != t[0]) st->done = 1; // pcg_solver.h:263-271
	});
}

// multigrid path when the last V-cycle kernel could not fuse it: rho' = z.b0, beta = rho'/rho  (pcg_solver.h:286-288)
__device__ __forceinline__ void finish_zr(CGState *st, double zr) {
	st->beta = st->iter == 0 ? 0.0 : zr / st->rho;
	st->rho = zr;
	if (zr == 0.0 || zr != zr) st->done = 1; // pcg_solver.h:263-271
}

__global__ void __launch_bounds__(TX *CG_BY) k_dot_zb(Dims d, Tiles T, const float *__restrict__ z, const float *__restrict__ b0, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				red[0] += (double)z[c] * (double)b0[c];
			}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) { finish_zr(st, t[0]); });
}

// s = z + beta s   (pcg_solver.h:289; plain CG passes z = r; the first iteration has beta = 0, s = 0)
template <class VecT, class ZT>
// (z-slabs: the boundary planes of s go straight into the neighbours' ghost planes, SlabPush; the product kernel waits for them)
__global__ void __launch_bounds__(TX *CG_BY) k_xpay(Dims d, Tiles T, const ZT *__restrict__ z, VecT *__restrict__ s, const CGState *__restrict__ st, const SlabPush sp) {
	if (st->done) return;
	VecT *const plo = push_target_lo<VecT>(sp, d.plane, d.nzl), *const phi = push_target_hi<VecT>(sp);
	const VecT beta = (VecT)st->beta;
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		// (batching four cells per thread with all loads up front was tried: 13 % slower at 512^3 — 2048 resident threads with one cell each already
		// keep more bytes in flight than 1024 threads with four)
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				const VecT v = fma(beta, s[c], (VecT)z[c]); // (an explicit fused multiply-add: k_xpay_spmv_tma forms the same value)
				s[c] = v;
				if (k == 0 && plo) plo[c] = v;
				if (k == d.nzl - 1 && phi) phi[c - (long long)k * d.plane] = v;
			}
	}
	if (sp.cm) signal_neighbours(sp.cm, sp.seq);
}

// q = A s,  sq = s.q ; last block: alpha = rho / sq            (pcg_solver.h:276-277)
// A thread marches its cell column through the tile's planes keeping the z neighbours in registers.
template <class VecT, class CoefT>
__global__ void __launch_bounds__(TX *CG_BY) k_spmv_dot(Dims d, Tiles T, const CoefT *__restrict__ wx, const CoefT *__restrict__ wy, const CoefT *__restrict__ wz,
                                                       const CoefT *__restrict__ dd, const VecT *__restrict__ s, VecT *__restrict__ q, RedBuf rb, CGState *st,
                                                       unsigned long long wait_in) {
	if (st->done) return;
	if (wait_in) block_wait_neighbours(rb.comm, wait_in); // z-slabs: the ghost planes of s, pushed by the neighbours' k_xpay
	double red[1] = {0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
			long long c = i + (long long)d.nx * (j + (long long)d.ny * kb);
			VecT sm = s[c - d.plane], sc = s[c];
			CoefT wzc = wz[c];
			for (int k = kb; k < ke; ++k, c += d.plane) {
				const VecT sp = s[c + d.plane];
				const CoefT wzp = wz[c + d.plane];
				VecT v = (VecT)dd[c] * sc;
				v += (VecT)wx[c] * (sc - s[c - 1]);
				v += (VecT)wx[c + 1] * (sc - s[c + 1]);
				v += (VecT)wy[c] * (sc - s[c - d.nx]);
				v += (VecT)wy[c + d.nx] * (sc - s[c + d.nx]);
				v += (VecT)wzc * (sc - sm);
				v += (VecT)wzp * (sc - sp);
				q[c] = v;
				red[0] += (double)sc * (double)v;
				sm = sc; sc = sp; wzc = wzp;
			}
		}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->sz = t[0];
		st->alpha = st->rho / t[0];
	});
}

// a10 warm start (macpressuresolver3.cpp:221-230): b -= A p_prev on the active tiles before the solve, |b|_inf of the new right-hand side
// (p_prev is a DENSE per-cell field here; the reference keeps it by row number, which is the same thing as long as the row set does not move and
// an artefact of its iteration order when it does). Cells off the row set have zero coefficients: their p_prev never enters.
template <class VecT, class CoefT>
__global__ void __launch_bounds__(TX *CG_BY) k_warm_rhs(Dims d, Tiles T, const CoefT *__restrict__ wx, const CoefT *__restrict__ wy, const CoefT *__restrict__ wz,
                                                       const CoefT *__restrict__ dd, const VecT *__restrict__ p, VecT *__restrict__ b, RedBuf rb, CGState *st) {
	double red[1] = {0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
			long long c = i + (long long)d.nx * (j + (long long)d.ny * kb);
			for (int k = kb; k < ke; ++k, c += d.plane) {
				const VecT pc = p[c];
				VecT v = (VecT)dd[c] * pc;
				v += (VecT)wx[c] * (pc - p[c - 1]);
				v += (VecT)wx[c + 1] * (pc - p[c + 1]);
				v += (VecT)wy[c] * (pc - p[c - d.nx]);
				v += (VecT)wy[c + d.nx] * (pc - p[c + d.nx]);
				v += (VecT)wz[c] * (pc - p[c - d.plane]);
				v += (VecT)wz[c + d.plane] * (pc - p[c + d.plane]);
				const VecT bv = b[c] - v;
				b[c] = bv;
				red[0] = fmax(red[0], fabs((double)bv));
			}
		}
	}
	grid_reduce<1, 0x1u>(red, rb, [&](double (&t)[1]) { st->bnorm = t[0]; });
}

// ---- the same product with FOUR x-adjacent cells per thread (nx % 4 == 0): aligned 16/32-byte loads, a quarter of the
// ---- load instructions, four times the bytes in flight per thread. Block (TX/4, TY): one quad column per thread.
template <class T> struct V4 { T a, b, c, d; };
__device__ __forceinline__ V4<float> ldv4(const float *p) {
	const float4 v = *reinterpret_cast<const float4 *>(p);
	return V4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ V4<double> ldv4(const double *p) {
	const double2 u = *reinterpret_cast<const double2 *>(p), v = *reinterpret_cast<const double2 *>(p + 2);
	return V4<double>{u.x, u.y, v.x, v.y};
}
__device__ __forceinline__ void stv4(float *p, const V4<float> &v) { *reinterpret_cast<float4 *>(p) = make_float4(v.a, v.b, v.c, v.d); }
__device__ __forceinline__ void stv4(double *p, const V4<double> &v) {
	*reinterpret_cast<double2 *>(p) = make_double2(v.a, v.b);
	*reinterpret_cast<double2 *>(p + 2) = make_double2(v.c, v.d);
}
inline dim3 cg_block4() { return dim3(TX / 4, TY, 1); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class VecT, class CoefT>
__device__ __forceinline__ VecT spmv_cell(CoefT dd, CoefT w0, CoefT w1, CoefT w2, CoefT w3, CoefT w4, CoefT w5, VecT sc, VecT x0, VecT x1, VecT x2, VecT x3, VecT x4,
                                          VecT x5) {
	VecT v = (VecT)dd * sc;
	v += (VecT)w0 * (sc - x0);
	v += (VecT)w1 * (sc - x1);
	v += (VecT)w2 * (sc - x2);
	v += (VecT)w3 * (sc - x3);
	v += (VecT)w4 * (sc - x4);
	v += (VecT)w5 * (sc - x5);
	return v;
}

template <class VecT, class CoefT>
__global__ void __launch_bounds__((TX / 4) * TY, 3) k_spmv_dot4(Dims d, Tiles T, const CoefT *__restrict__ wx, const CoefT *__restrict__ wy, const CoefT *__restrict__ wz,
                                                               const CoefT *__restrict__ dd, const VecT *__restrict__ s, VecT *__restrict__ q, RedBuf rb, CGState *st,
                                                               unsigned long long wait_in) {
	if (st->done) return;
	if (wait_in) block_wait_neighbours(rb.comm, wait_in); // z-slabs: the ghost planes of s, pushed by the neighbours' k_xpay
	double red[1] = {0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + 4 * threadIdx.x, j = j0 + threadIdx.y;
		if (i >= d.nx || j >= d.ny) continue;
		long long c = i + nx * (j + (long long)d.ny * kb);
		V4<VecT> sm = ldv4(s + c - d.plane), sc = ldv4(s + c);
		V4<CoefT> wzc = ldv4(wz + c);
		for (int k = kb; k < ke; ++k, c += d.plane) {
			// the product waits on DRAM for the operands nobody fetched ahead (ncu: 85 % long-scoreboard stalls at 35 % occupancy, no registers left
			// to load them early): ask L2 for the next plane's wx, wy, dd and for s, wz two planes up now, the loads below then find them there
			if (k + 1 < ke) { prefetch_l2(wx + c + d.plane); prefetch_l2(wy + c + d.plane); prefetch_l2(dd + c + d.plane); }
			if (k + 2 <= ke) { prefetch_l2(s + c + 2 * d.plane); prefetch_l2(wz + c + 2 * d.plane); }
			const V4<VecT> sp = ldv4(s + c + d.plane), sd = ldv4(s + c - nx), su = ldv4(s + c + nx);
			const V4<CoefT> wzp = ldv4(wz + c + d.plane), wxq = ldv4(wx + c), wyq = ldv4(wy + c), wyu = ldv4(wy + c + nx), ddq = ldv4(dd + c);
			const CoefT wx4 = wx[c + 4];
			const VecT sl = s[c - 1], sr = s[c + 4];
			V4<VecT> v;
			v.a = spmv_cell<VecT, CoefT>(ddq.a, wxq.a, wxq.b, wyq.a, wyu.a, wzc.a, wzp.a, sc.a, sl, sc.b, sd.a, su.a, sm.a, sp.a);
			v.b = spmv_cell<VecT, CoefT>(ddq.b, wxq.b, wxq.c, wyq.b, wyu.b, wzc.b, wzp.b, sc.b, sc.a, sc.c, sd.b, su.b, sm.b, sp.b);
			v.c = spmv_cell<VecT, CoefT>(ddq.c, wxq.c, wxq.d, wyq.c, wyu.c, wzc.c, wzp.c, sc.c, sc.b, sc.d, sd.c, su.c, sm.c, sp.c);
			v.d = spmv_cell<VecT, CoefT>(ddq.d, wxq.d, wx4, wyq.d, wyu.d, wzc.d, wzp.d, sc.d, sc.c, sr, sd.d, su.d, sm.d, sp.d);
			stv4(q + c, v);
			red[0] += (double)sc.a * (double)v.a + (double)sc.b * (double)v.b + (double)sc.c * (double)v.c + (double)sc.d * (double)v.d;
			sm = sc; sc = sp; wzc = wzp;
		}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->sz = t[0];
		st->alpha = st->rho / t[0];
	});
}

// x += alpha s ; r -= alpha q ; |r|_inf ; (plain CG: r.r) ; sum of x ; float copy of r for multigrid      (pcg_solver.h:278-285)
// (sum of x: a pure-Neumann system loses its mean when the pressure is stored; x is zero off the row set, so the sum over the active tiles is the sum over rows)
// last block: convergence test, iteration count; for plain CG also beta and the new rho.
template <class VecT, bool PLAIN, bool WRITE_B0>
__global__ void __launch_bounds__(TX *CG_BY) k_axpy2_norm(Dims d, Tiles T, const VecT *__restrict__ s, const VecT *__restrict__ q, VecT *__restrict__ x,
                                                         VecT *__restrict__ r, float *__restrict__ b0, RedBuf rb, CGState *st) {
	if (st->done) return;
	const VecT alpha = (VecT)st->alpha;
	double red[3] = {0.0, 0.0, 0.0};
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += CG_BY) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				const VecT xv = x[c] + alpha * s[c];
				x[c] = xv;
				red[2] += (double)xv;
				const VecT rv = r[c] - alpha * q[c];
				r[c] = rv;
				if (WRITE_B0) b0[c] = (float)rv;
				red[0] = fmax(red[0], fabs((double)rv));
				if (PLAIN) red[1] += (double)rv * (double)rv;
			}
	}
	grid_reduce<3, 0x1u>(red, rb, [&](double (&t)[3]) {
		st->rnorm = t[0];
		st->sum_x = t[2];
		st->iter += 1;
		if (t[0] <= st->tol) { st->done = 1; st->converged = 1; }
		else if (st->iter >= st->max_iter) st->done = 1;
		if (PLAIN) { st->beta = t[1] / st->rho; st->rho = t[1]; }
	});
}

} // namespace shkz
