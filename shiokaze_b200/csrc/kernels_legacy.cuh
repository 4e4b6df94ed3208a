// Unfused multigrid kernels (one launch per colour / transfer step, dense launch geometry). They are NOT on
// the product path: shkz_b200_debug_vcycle(mode=legacy) runs them so that tests can check, bit for bit, that
// the fused tile kernels of kernels_mg.cuh compute the same V-cycle.
#pragma once
#include "common.cuh"
#include "kernels_mg.cuh"

namespace shkz {

constexpr int LTX = 64, LTY = 4, LZC = 16;
inline dim3 legacy_stencil_grid(const Dims &d) { return dim3((d.nx + LTX - 1) / LTX, (d.ny + LTY - 1) / LTY, (d.nzl + LZC - 1) / LZC); }
inline dim3 legacy_stencil_block() { return dim3(LTX, LTY, 1); }

// One colour of a Gauss-Seidel sweep. Each thread owns a pair of x-adjacent cells and updates the
// one whose parity matches. x == 0 on entry of the very first half sweep is exploited by ZERO_X.
template <bool ZERO_X>
__global__ void __launch_bounds__(256) k_legacy_rbgs(Dims d, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
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
	const float v = ZERO_X ? gs_relax0(w0, w1, w2, w3, w4, w5, dd[c], b[c])
	                       : gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], x[c - 1], x[c + 1], x[c - d.nx], x[c + d.nx], x[c - d.plane], x[c + d.plane], x[c]);
	x[c] = v;
}

// r = b - A x
__global__ void __launch_bounds__(LTX *LTY) k_legacy_residual(Dims d, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                    const float *__restrict__ dd, const float *__restrict__ b, const float *__restrict__ x,
                                                    float *__restrict__ r, const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int i = blockIdx.x * LTX + threadIdx.x, j = blockIdx.y * LTY + threadIdx.y;
	const int kbeg = blockIdx.z * LZC, kend = min(kbeg + LZC, d.nzl);
	if (i >= d.nx || j >= d.ny) return;
	long long c = i + (long long)d.nx * (j + (long long)d.ny * kbeg);
	float xm = x[c - d.plane], xc = x[c], wzc = wz[c];
	for (int k = kbeg; k < kend; ++k, c += d.plane) {
		const float xp = x[c + d.plane], wzp = wz[c + d.plane];
		const float v = residual7(wx[c], wx[c + 1], wy[c], wy[c + d.nx], wzc, wzp, dd[c], b[c], xc, x[c - 1], x[c + 1], x[c - d.nx], x[c + d.nx], xm, xp);
		r[c] = v;
		xm = xc; xc = xp; wzc = wzp;
	}
}

// coarse b = P^T r (sum over the 2x2x2 children that exist)
__global__ void __launch_bounds__(256) k_legacy_restrict(Dims df, Dims dc, const float *__restrict__ r, float *__restrict__ bc, const CGState *__restrict__ st) {
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
__global__ void __launch_bounds__(256) k_legacy_prolong_add(Dims df, Dims dc, const float *__restrict__ ec, float *__restrict__ x, const CGState *__restrict__ st) {
	if (st && st->done) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i >= df.nx || j >= df.ny) return;
	x[i + (long long)df.nx * (j + (long long)df.ny * k)] += ec[(i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1))];
}

} // namespace shkz
