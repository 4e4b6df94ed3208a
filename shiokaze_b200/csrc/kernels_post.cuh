// The step right after the projection: velocity extrapolation + solid constraint, on the velocity that is still on the device.
//   macarray_extrapolator3::extrapolate        include/shiokaze/array/macarray_extrapolator3.h:49-53 -> array_extrapolator3.h:51-82 -> src/array/dilate3.h
//   macutility3::constrain_velocity            src/utility/macutility3.cpp:61-88
//   (macutility3::extrapolate_and_constrain_velocity = the two in a row, :89-93; called by macliquid3.cpp:309-319, macsmoke3.cpp:294)
// Arithmetic follows the reference operation for operation (Real sums, double interpolation weights, the float <- double accumulations of
// array_interpolator3 / array_derivative3), with explicit round-to-nearest intrinsics so that nothing is contracted: results are the reference's bits.
#pragma once
#include "kernels_assemble.cuh"

namespace shkz {

__device__ __forceinline__ float real_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double real_div(double a, double b) { return __ddiv_rn(a, b); }

// One round of array_extrapolator3::extrapolate on one face grid (w x h x dz): every INACTIVE face with an active in-bounds neighbour becomes active with
// the mean of its active neighbours, all rounds reading the state before the round (dilate3.h evaluates on a scratch copy and applies afterwards). Values
// are written in place — only faces that were inactive are written, only faces that were active are read —, the new mask goes to act_out.
template <class RealT>
__global__ void __launch_bounds__(256) k_extrapolate_round(int w, int h, int dz, RealT *__restrict__ v, const uint8_t *__restrict__ act_in, uint8_t *__restrict__ act_out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
	if (i >= w || j >= h) return;
	const long long sx = 1, sy = w, sz = (long long)w * h, n = i + sy * j + sz * k;
	uint8_t a = act_in[n];
	if (!a) {
		// query order of array_extrapolator3.h:60-61: +x -x +y -y -z +z; T sum, int weight, sum / weight
		RealT sum = (RealT)0;
		int weight = 0;
		if (i + 1 < w && act_in[n + sx]) { sum = real_add(sum, v[n + sx]); ++weight; }
		if (i > 0 && act_in[n - sx]) { sum = real_add(sum, v[n - sx]); ++weight; }
		if (j + 1 < h && act_in[n + sy]) { sum = real_add(sum, v[n + sy]); ++weight; }
		if (j > 0 && act_in[n - sy]) { sum = real_add(sum, v[n - sy]); ++weight; }
		if (k > 0 && act_in[n - sz]) { sum = real_add(sum, v[n - sz]); ++weight; }
		if (k + 1 < dz && act_in[n + sz]) { sum = real_add(sum, v[n + sz]); ++weight; }
		if (weight) {
			v[n] = real_div(sum, (RealT)weight);
			a = 1;
		}
	}
	act_out[n] = a;
}

// array_interpolator3::interpolate_coef (array_interpolator3.h:52-76): the cell (i,j,k) of the 2x2x2 stencil and its eight weights at index-space position p
struct TrilinearStencil {
	int i, j, k;
	double coef[8]; // order (i,j,k) (i+1,j,k) (i,j+1,k) (i+1,j+1,k) (i,j,k+1) (i+1,j,k+1) (i,j+1,k+1) (i+1,j+1,k+1)
	double x, y, z;
};
__device__ __forceinline__ void trilinear_stencil(int w, int h, int dz, double px, double py, double pz, TrilinearStencil &S) {
	const double x = fmax(0.0, fmin((double)w - 1., px)), y = fmax(0.0, fmin((double)h - 1., py)), z = fmax(0.0, fmin((double)dz - 1., pz));
	const int i = (int)fmin(x, (double)w - 2.), j = (int)fmin(y, (double)h - 2.), k = (int)fmin(z, (double)dz - 2.);
	S.i = i; S.j = j; S.k = k; S.x = x; S.y = y; S.z = z;
	const double xa = __dsub_rn((double)(i + 1), x), xb = __dsub_rn(x, (double)i), ya = __dsub_rn((double)(j + 1), y), yb = __dsub_rn(y, (double)j);
	const double za = __dsub_rn((double)(k + 1), z), zb = __dsub_rn(z, (double)k);
	S.coef[0] = __dmul_rn(__dmul_rn(za, xa), ya);
	S.coef[1] = __dmul_rn(__dmul_rn(za, xb), ya);
	S.coef[2] = __dmul_rn(__dmul_rn(za, xa), yb);
	S.coef[3] = __dmul_rn(__dmul_rn(za, xb), yb);
	S.coef[4] = __dmul_rn(__dmul_rn(zb, xa), ya);
	S.coef[5] = __dmul_rn(__dmul_rn(zb, xb), ya);
	S.coef[6] = __dmul_rn(__dmul_rn(zb, xa), yb);
	S.coef[7] = __dmul_rn(__dmul_rn(zb, xb), yb);
}
// array_interpolator3::interpolate (:89-105, only_actives = false): T value; value += array(index) * coef for every non-zero coef — a T <- double accumulation.
// `act` == nullptr: every entry of `a` is what operator() reads; otherwise inactive entries read as the background value 0.
template <class RealT>
__device__ __forceinline__ RealT trilinear(const RealT *__restrict__ a, const uint8_t *__restrict__ act, int w, int h, const TrilinearStencil &S) {
	RealT value = (RealT)0;
#pragma unroll
	for (int n = 0; n < 8; ++n) {
		if (S.coef[n] == 0.0) continue;
		const long long idx = (S.i + (n & 1)) + (long long)w * ((S.j + ((n >> 1) & 1)) + (long long)h * (S.k + (n >> 2)));
		const RealT s = (act == nullptr || act[idx]) ? a[idx] : (RealT)0;
		value = (RealT)__dadd_rn((double)value, __dmul_rn((double)s, S.coef[n]));
	}
	return value;
}

// macutility3::constrain_velocity (src/utility/macutility3.cpp:61-88) on the faces of direction `dim`; vs / as = the velocity BEFORE the constraint
// (the reference's velocity_save). solid: nodal (nx+1, ny+1, nz+1). Whole grids only.
template <class RealT>
__global__ void __launch_bounds__(256) k_constrain_velocity(Dims d, double dx, int dim, const RealT *__restrict__ solid, ConstFaceGrids<RealT> vs, FaceMasks as,
                                                           RealT *__restrict__ vel, const uint8_t *__restrict__ act) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
	const int w = d.nx + (dim == 0), h = d.ny + (dim == 1);
	if (i >= w || j >= h) return;
	const long long f = i + (long long)w * (j + (long long)h * k);
	if (!act[f]) return;
	// p = vec3i(i,j,k).face(dim): the face centre in the index space of the nodal grid (include/shiokaze/math/vec.h:377-381)
	const double px = i + 0.5 * (dim != 0), py = j + 0.5 * (dim != 1), pz = k + 0.5 * (dim != 2);
	RealT value = vel[f];
	TrilinearStencil S;
	trilinear_stencil(d.nx + 1, d.ny + 1, d.nzl + 1, px, py, pz, S);
	if ((double)trilinear<RealT>(solid, nullptr, d.nx + 1, d.ny + 1, S) < 0.0) {
		// array_derivative3::derivative (array_derivative3.h:49-107): result[dim] += coef[dim][n] * array(index), T <- double accumulations, no zero skip
		const double xa = __dsub_rn((double)(S.i + 1), S.x), xb = __dsub_rn(S.x, (double)S.i), ya = __dsub_rn((double)(S.j + 1), S.y), yb = __dsub_rn(S.y, (double)S.j);
		const double za = __dsub_rn((double)(S.k + 1), S.z), zb = __dsub_rn(S.z, (double)S.k);
		const double cx[8] = {-__dmul_rn(za, ya), __dmul_rn(za, ya), -__dmul_rn(za, yb), __dmul_rn(za, yb), -__dmul_rn(zb, ya), __dmul_rn(zb, ya), -__dmul_rn(zb, yb), __dmul_rn(zb, yb)};
		const double cy[8] = {-__dmul_rn(za, xa), -__dmul_rn(za, xb), __dmul_rn(za, xa), __dmul_rn(za, xb), -__dmul_rn(zb, xa), -__dmul_rn(zb, xb), __dmul_rn(zb, xa), __dmul_rn(zb, xb)};
		const double cz[8] = {-__dmul_rn(xa, ya), -__dmul_rn(xb, ya), -__dmul_rn(xa, yb), -__dmul_rn(xb, yb), __dmul_rn(xa, ya), __dmul_rn(xb, ya), __dmul_rn(xa, yb), __dmul_rn(xb, yb)};
		RealT g[3] = {(RealT)0, (RealT)0, (RealT)0};
		const long long sw = d.nx + 1, sh = d.ny + 1;
#pragma unroll
		for (int n = 0; n < 8; ++n) {
			const double s = (double)solid[(S.i + (n & 1)) + sw * ((S.j + ((n >> 1) & 1)) + sh * (long long)(S.k + (n >> 2)))];
			g[0] = (RealT)__dadd_rn((double)g[0], __dmul_rn(cx[n], s));
			g[1] = (RealT)__dadd_rn((double)g[1], __dmul_rn(cy[n], s));
			g[2] = (RealT)__dadd_rn((double)g[2], __dmul_rn(cz[n], s));
		}
		const double nrm[3] = {__ddiv_rn((double)g[0], dx), __ddiv_rn((double)g[1], dx), __ddiv_rn((double)g[2], dx)};
		const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(nrm[0], nrm[0])), __dmul_rn(nrm[1], nrm[1])), __dmul_rn(nrm[2], nrm[2]));
		if (n2 != 0.0) {
			// u = macarray_interpolator3::interpolate(velocity_save, p) (macarray_interpolator3.h:48-55): component e sampled at p - 0.5 (e != axis)
			double u[3];
#pragma unroll
			for (int e = 0; e < 3; ++e) {
				TrilinearStencil V;
				const int we = d.nx + (e == 0), he = d.ny + (e == 1), de = d.nzl + (e == 2);
				trilinear_stencil(we, he, de, px - 0.5 * (e != 0), py - 0.5 * (e != 1), pz - 0.5 * (e != 2), V);
				u[e] = (double)trilinear<RealT>(vs.p[e], as.p[e], we, he, V);
			}
			const double un = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(u[0], nrm[0])), __dmul_rn(u[1], nrm[1])), __dmul_rn(u[2], nrm[2]));
			if (un < 0.0) value = (RealT)__dsub_rn(u[dim], __dmul_rn(nrm[dim], un));
		}
	}
	const int pd = dim == 0 ? i : (dim == 1 ? j : k), n_dim = dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzl);
	if (pd == 0 && (double)value < 0.0) value = (RealT)0;
	if (pd == n_dim && (double)value > 0.0) value = (RealT)0;
	vel[f] = value;
}

} // namespace shkz
