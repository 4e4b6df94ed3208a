// Geometric (aggregation) multigrid V-cycle on the dense 7-point operator, fp32 — the GPU-parallel
// replacement of the reference's serial MIC(0) sweeps (pcg_solver.h:89-223, K8/K11 of SURVEY.md 2d).
//
// Hierarchy: 2x2x2 cell aggregates, piecewise-constant prolongation P, restriction P^T, coarse
// operator scale * P^T A P. With this P the Galerkin product of a face-coefficient 7-point operator is
// again one (coarse face coefficient = sum of the 4 fine faces crossing the coarse face), so every
// level uses the same four arrays (wx, wy, wz, dd) and the same kernels; cut cells, ghost-fluid
// Dirichlet faces and Neumann walls are carried algebraically (coarse dd = sum of the children's dd,
// couplings inside an aggregate drop out). scale = 1/2 restores the h^-2 scaling
// that piecewise-constant transfer loses (the usual over-correction of unsmoothed aggregation).
// Smoother: red-black Gauss-Seidel, colours by (i+j+k_global) parity, reversed order after the
// coarse correction so that the V-cycle is a symmetric operator for CG.
#pragma once
#include "common.cuh"
#include "kernels_cg.cuh"

namespace shkz {

struct MGLevel {
	Dims d;
	float *wx, *wy, *wz, *dd; // with ghost planes, pointing at plane 0
	float *x, *b, *r;
};

// One colour of a Gauss-Seidel sweep. Each thread owns a pair of x-adjacent cells and updates the
// one whose parity matches. x == 0 on entry of the very first half sweep is exploited by ZERO_X.
template <bool ZERO_X>
__global__ void __launch_bounds__(256) k_rbgs(Dims d, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                             const float *__restrict__ dd, const float *__restrict__ b, float *__restrict__ x, int color,
                                             const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int ip = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	const int i = 2 * ip + ((j + k + d.k0 + color) & 1);
	if (i >= d.nx || j >= d.ny) return;
	const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
	const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + d.nx], w4 = wz[c], w5 = wz[c + d.plane];
	const float dg = dd[c] + ((w0 + w1) + (w2 + w3) + (w4 + w5));
	float v = 0.f;
	if (dg > 0.f) {
		float acc = b[c];
		if (!ZERO_X) {
			acc += w0 * x[c - 1];
			acc += w1 * x[c + 1];
			acc += w2 * x[c - d.nx];
			acc += w3 * x[c + d.nx];
			acc += w4 * x[c - d.plane];
			acc += w5 * x[c + d.plane];
		}
		v = __fdividef(acc, dg);
	}
	x[c] = v;
}

// r = b - A x
__global__ void __launch_bounds__(TX *TY) k_residual(Dims d, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                    const float *__restrict__ dd, const float *__restrict__ b, const float *__restrict__ x,
                                                    float *__restrict__ r, const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
	const int kbeg = blockIdx.z * ZC, kend = min(kbeg + ZC, d.nzl);
	if (i >= d.nx || j >= d.ny) return;
	long long c = i + (long long)d.nx * (j + (long long)d.ny * kbeg);
	float xm = x[c - d.plane], xc = x[c], wzc = wz[c];
	for (int k = kbeg; k < kend; ++k, c += d.plane) {
		const float xp = x[c + d.plane], wzp = wz[c + d.plane];
		float v = b[c] - dd[c] * xc;
		v += wx[c] * (x[c - 1] - xc);
		v += wx[c + 1] * (x[c + 1] - xc);
		v += wy[c] * (x[c - d.nx] - xc);
		v += wy[c + d.nx] * (x[c + d.nx] - xc);
		v += wzc * (xm - xc);
		v += wzp * (xp - xc);
		r[c] = v;
		xm = xc; xc = xp; wzc = wzp;
	}
}

// coarse b = P^T r (sum over the 2x2x2 children that exist)
__global__ void __launch_bounds__(256) k_restrict(Dims df, Dims dc, const float *__restrict__ r, float *__restrict__ bc, const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int I = blockIdx.x * blockDim.x + threadIdx.x;
	const int J = blockIdx.y * blockDim.y + threadIdx.y;
	const int K = blockIdx.z;
	if (I >= dc.nx || J >= dc.ny) return;
	float acc = 0.f;
#pragma unroll
	for (int dk = 0; dk < 2; ++dk)
#pragma unroll
		for (int dj = 0; dj < 2; ++dj)
#pragma unroll
			for (int di = 0; di < 2; ++di) {
				const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
				if (i < df.nx && j < df.ny && k < df.nzl) acc += r[i + (long long)df.nx * (j + (long long)df.ny * k)];
			}
	bc[I + (long long)dc.nx * (J + (long long)dc.ny * K)] = acc;
}

// x += P e
__global__ void __launch_bounds__(256) k_prolong_add(Dims df, Dims dc, const float *__restrict__ ec, float *__restrict__ x, const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i >= df.nx || j >= df.ny) return;
	x[i + (long long)df.nx * (j + (long long)df.ny * k)] += ec[(i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1))];
}

// Coarse operator = scale * P^T A P (setup, once per projection): coarse face coupling = sum of the four
// fine couplings crossing the coarse face, coarse dd = sum of the children's dd.
__global__ void __launch_bounds__(256) k_coarsen_operator(Dims df, Dims dc, float scale, const float *__restrict__ wx, const float *__restrict__ wy,
                                                         const float *__restrict__ wz, const float *__restrict__ dd, float *__restrict__ cwx,
                                                         float *__restrict__ cwy, float *__restrict__ cwz, float *__restrict__ cdd) {
	const int I = blockIdx.x * blockDim.x + threadIdx.x;
	const int J = blockIdx.y * blockDim.y + threadIdx.y;
	const int K = blockIdx.z;
	if (I >= dc.nx || J >= dc.ny) return;
	float sx = 0.f, sy = 0.f, sz = 0.f, sd = 0.f;
#pragma unroll
	for (int dk = 0; dk < 2; ++dk)
#pragma unroll
		for (int dj = 0; dj < 2; ++dj)
#pragma unroll
			for (int di = 0; di < 2; ++di) {
				const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
				if (i >= df.nx || j >= df.ny || k >= df.nzl) continue;
				const long long c = i + (long long)df.nx * (j + (long long)df.ny * k);
				sd += dd[c];
				if (!di) sx += wx[c];
				if (!dj) sy += wy[c];
				if (!dk) sz += wz[c];
			}
	const long long C = I + (long long)dc.nx * (J + (long long)dc.ny * K);
	cwx[C] = scale * sx;
	cwy[C] = scale * sy;
	cwz[C] = scale * sz;
	cdd[C] = scale * sd;
}

// Hand-off CG -> MG: b0 = float(r)
template <class VecT>
__global__ void __launch_bounds__(256) k_to_mg(long long n, const VecT *__restrict__ r, float *__restrict__ b, const CGState *__restrict__ st) {
	if (st->done) return;
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) b[c] = (float)r[c];
}

// Hand-off MG -> CG: z = x0, rr = z.r                                  (pcg_solver.h:286-287)
template <class VecT>
__global__ void __launch_bounds__(256) k_from_mg(long long n, const float *__restrict__ x0, const VecT *__restrict__ r, VecT *__restrict__ z, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
		const VecT zv = (VecT)x0[c];
		if ((const void *)x0 != (const void *)z) z[c] = zv;
		red[0] += (double)zv * (double)r[c];
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) { st->rr = t[0]; });
}

} // namespace shkz
