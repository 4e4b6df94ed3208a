// libshkz_b200 — the step BEFORE the projection (first part): semi-Lagrangian / MacCormack advection on the MAC grid.
// Replaces macadvection3_interface::advect_vector / advect_scalar as implemented by the reference's `macadvection3` module
//     src/advection/macadvection3.cpp:40-55      the two entry points
//     :69-141                                     advect_semiLagrangian_u   (+ the min / max / narrow-band record of the MacCormack limiter)
//     :143-187                                    advect_u                  (forward, backward with -dt, limiter)
//     :189-237, :239-283                          the same for cell-centred scalars
//     include/shiokaze/array/macarray3.h:305-372  convert_to_full (cell-centred and face-centred full velocity)
//     include/shiokaze/array/array_interpolator3.h:50-105   trilinear weights and the T <- double accumulation
//     include/shiokaze/math/WENO3.h:51-143, WENO.h:54-98    WENO=Yes: sixth-order WENO interpolation (Macdonald & Ruuth 2008), dimension by dimension
// called by every simulator right before project(): src/liquid/macliquid3.cpp:343-346 (level set through maclevelsetsurfacetracker3.cpp:51, then velocity),
// src/smoke/macsmoke3.cpp:274,281 (density, velocity).
//
// THIS TRANSLATION UNIT IS COMPILED WITH --fmad=false (csrc/Makefile): every + - * / below is one IEEE round-to-nearest operation, never contracted, so the
// expressions can be written as the reference writes them and give the reference's bits (g++ on x86-64 does not contract either). The bar is bit-exact
// (tests/test_gpu_advect.py, against the unmodified reference's own module driven through its loader).
//
// Quirks of the reference that are kept because they are its results:
//   * advect_vector ignores its `velocity` argument: the field is traced with ITSELF (macadvection3.cpp:79 converts v_in, not v).
//   * the value pass of the face version forms the back-traced position through vec3<Real> (two roundings to Real: dt*u, then /dx, :86), the min / max pass of
//     the same face through vec3d (:105): two slightly different positions.
//   * the upper bound of the limiter starts from numeric_limits<double>::min() — the smallest POSITIVE double (:111; Real for scalars, :216): where all eight
//     corner values are <= 0 the bound is that tiny positive number (0 after rounding to float), not their maximum.
// Layouts: include/shkz_b200.h (x fastest). Inactive faces of a velocity grid read as 0 (the background value of the simulators' velocity grids) whatever the
// buffer holds there; cell grids (the advected scalar, the liquid level set) are read densely: every entry is what array3::operator() returns.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cfloat>
#include <cmath>
#include <string>

#include "../../include/shkz_b200.h"

namespace {

thread_local std::string g_adv_error;

int fail(int code, const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_adv_error = buf;
	return code;
}
#define CK(call)                                                                                              \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if (e_ != cudaSuccess) return fail(SHKZ_B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)
#define CKR(call)                     \
	do {                              \
		int r_ = (call);              \
		if (r_ != SHKZ_B200_OK) return r_; \
	} while (0)

constexpr int ADV_THREADS = 256;
#define HD __host__ __device__ __forceinline__

struct Grid {
	int nx, ny, nz;
	double dx;
	double inv_dx; // != 0: dx is a power of two and x / dx == x * inv_dx bit for bit (scaling by a power of two is exact); 0: divide
};
__host__ __device__ __forceinline__ double over_dx(const Grid &g, double x) { return g.inv_dx != 0.0 ? x * g.inv_dx : x / g.dx; }
inline Grid make_grid(int nx, int ny, int nz, double dx) {
	Grid g{nx, ny, nz, dx, 0.0};
	int e = 0;
	if (frexp(dx, &e) == 0.5 && e > -500 && e < 500) g.inv_dx = ldexp(1.0, 1 - e);
	return g;
}
__host__ __device__ __forceinline__ int fw(const Grid &g, int dim) { return g.nx + (dim == 0); }
__host__ __device__ __forceinline__ int fh(const Grid &g, int dim) { return g.ny + (dim == 1); }
__host__ __device__ __forceinline__ int fd(const Grid &g, int dim) { return g.nz + (dim == 2); }

// a MAC field: three face grids and their activity bytes
template <class RealT> struct Mac {
	const RealT *v[3];
	const uint8_t *a[3];
};
// macarray3::operator() on a face inside the grid: the active value, else the background value 0
template <class RealT, bool MASKED = true> HD RealT face_read(const Mac<RealT> &F, const Grid &g, int dim, int i, int j, int k) {
	const long long n = i + (long long)fw(g, dim) * (long long)(j + fh(g, dim) * k); // (rows x planes fits an int: one widening multiply)
	const RealT v = F.v[dim][n]; // (requested together with the mask byte, not after it: the entry exists whether the face is active or not)
	if (!MASKED) return v;
	return F.a[dim][n] ? v : (RealT)0;
}
HD int clampi(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }
// the same read at a flat index (the callers below form a base index once and add constant strides: the index products were a fifth of the instructions)
template <class RealT, bool MASKED> HD RealT face_at(const Mac<RealT> &F, int dim, long long n) {
	const RealT v = F.v[dim][n];
	if (!MASKED) return v;
	return F.a[dim][n] ? v : (RealT)0;
}
// std::min / std::max as the reference's library defines them (the first argument wins a tie: signed zeros keep their place)
HD double std_min(double a, double b) { return b < a ? b : a; }
HD double std_max(double a, double b) { return a < b ? b : a; }

// array_interpolator3::interpolate_coef (array_interpolator3.h:50-76) on a w x h x d grid
struct Stencil {
	int i, j, k;
	double x, y, z;
	double coef[8]; // (i,j,k) (i+1,j,k) (i,j+1,k) (i+1,j+1,k) (i,j,k+1) (i+1,j,k+1) (i,j+1,k+1) (i+1,j+1,k+1)
};
HD void clamp_position(int w, int h, int d, double px, double py, double pz, Stencil &S) {
	S.x = std_max(0.0, std_min(w - 1., px));
	S.y = std_max(0.0, std_min(h - 1., py));
	S.z = std_max(0.0, std_min(d - 1., pz));
	S.i = (int)std_min(S.x, w - 2.);
	S.j = (int)std_min(S.y, h - 2.);
	S.k = (int)std_min(S.z, d - 2.);
}
HD void make_stencil(int w, int h, int d, double px, double py, double pz, Stencil &S) {
	clamp_position(w, h, d, px, py, pz, S);
	const int i = S.i, j = S.j, k = S.k;
	const double x = S.x, y = S.y, z = S.z;
	S.coef[0] = (k + 1 - z) * (i + 1 - x) * (j + 1 - y);
	S.coef[1] = (k + 1 - z) * (x - i) * (j + 1 - y);
	S.coef[2] = (k + 1 - z) * (i + 1 - x) * (y - j);
	S.coef[3] = (k + 1 - z) * (x - i) * (y - j);
	S.coef[4] = (z - k) * (i + 1 - x) * (j + 1 - y);
	S.coef[5] = (z - k) * (x - i) * (j + 1 - y);
	S.coef[6] = (z - k) * (i + 1 - x) * (y - j);
	S.coef[7] = (z - k) * (x - i) * (y - j);
}
// array_interpolator3::interpolate, only_actives = false (:89-105): T value; value += array(index) * coef over the non-zero weights.
// read(i,j,k) = array3::operator()
template <class RealT, class Read> HD RealT trilinear(const Stencil &S, Read read) {
	RealT value = (RealT)0;
#pragma unroll
	for (int n = 0; n < 8; ++n)
		if (S.coef[n]) value = (RealT)((double)value + (double)read(S.i + (n & 1), S.j + ((n >> 1) & 1), S.k + (n >> 2)) * S.coef[n]);
	return value;
}

// (the same accumulation over eight values already in registers, in stencil order)
template <class RealT> HD RealT trilinear_corners(const Stencil &S, const RealT corner[8]) {
	RealT value = (RealT)0;
#pragma unroll
	for (int n = 0; n < 8; ++n)
		if (S.coef[n]) value = (RealT)((double)value + (double)corner[n] * S.coef[n]);
	return value;
}

// WENO::interp6 (include/shiokaze/math/WENO.h:54-98): v = the values at -2 .. 3, x in [0,1]
HD double sqr(double x) { return x * x; }
__host__ __device__ __noinline__ double weno6(double x, const double v[6]) {
	const double eps = DBL_EPSILON;
	const double f_m2 = v[0], f_m1 = v[1], f_p0 = v[2], f_p1 = v[3], f_p2 = v[4], f_p3 = v[5];
	const double x_m2 = -2.0, x_m1 = -1.0, x_p0 = 0.0, x_p1 = 1.0, x_p2 = 2.0, x_p3 = 3.0;
	double C[3], S[3], P[3];
	C[0] = (x_p2 - x) * (x_p3 - x) / 20.0;
	C[1] = (x_p3 - x) * (x - x_m2) / 10.0;
	C[2] = (x - x_m2) * (x - x_m1) / 20.0;
	S[0] = ((814. * sqr(f_p1)) + (4326. * sqr(f_p0)) + (2976. * sqr(f_m1)) + (244. * sqr(f_m2)) - (3579. * f_p0 * f_p1) - (6927. * f_p0 * f_m1) + (1854. * f_p0 * f_m2) +
	        (2634. * f_p1 * f_m1) - (683. * f_p1 * f_m2) - (1659. * f_m1 * f_m2)) / 180.0;
	S[1] = ((1986. * sqr(f_p1)) + (1986. * sqr(f_p0)) + (244. * sqr(f_m1)) + (244. * sqr(f_p2)) + (1074. * f_p0 * f_p2) - (3777. * f_p0 * f_p1) - (1269. * f_p0 * f_m1) +
	        (1074. * f_p1 * f_m1) - (1269. * f_p2 * f_p1) - (293. * f_p2 * f_m1)) / 180.0;
	S[2] = ((814. * sqr(f_p0)) + (4326. * sqr(f_p1)) + (2976. * sqr(f_p2)) + (244. * sqr(f_p3)) - (683. * f_p0 * f_p3) + (2634. * f_p0 * f_p2) - (3579. * f_p0 * f_p1) -
	        (6927. * f_p1 * f_p2) + (1854. * f_p1 * f_p3) - (1659. * f_p2 * f_p3)) / 180.0;
	P[0] = f_m2 + (f_m1 - f_m2) * (x - x_m2) + (f_p0 - 2. * f_m1 + f_m2) * (x - x_m2) * (x - x_m1) / 2.0 +
	       (f_p1 - 3. * f_p0 + 3. * f_m1 - f_m2) * (x - x_m2) * (x - x_m1) * (x - x_p0) / 6.0;
	P[1] = f_m1 + (f_p0 - f_m1) * (x - x_m1) + (f_p1 - 2. * f_p0 + f_m1) * (x - x_m1) * (x - x_p0) / 2.0 +
	       (f_p2 - 3. * f_p1 + 3. * f_p0 - f_m1) * (x - x_m1) * (x - x_p0) * (x - x_p1) / 6.0;
	P[2] = f_p0 + (f_p1 - f_p0) * (x - x_p0) + (f_p2 - 2. * f_p1 + f_p0) * (x - x_p0) * (x - x_p1) / 2.0 +
	       (f_p3 - 3. * f_p2 + 3. * f_p1 - f_p0) * (x - x_p0) * (x - x_p1) * (x - x_p2) / 6.0;
	double a[3], sum = 0.0;
	for (int i = 0; i < 3; ++i) {
		a[i] = C[i] / (eps + sqr(S[i]));
		sum += a[i];
	}
	double w[3];
	for (int i = 0; i < 3; ++i) w[i] = a[i] / sum;
	return w[0] * P[0] + w[1] * P[1] + w[2] * P[2];
}
// WENO3::interpolate, order 6 (WENO3.h:103-121): x, then y, then z; indices clamped into the grid. read = array3::operator()
template <class Read> HD double weno3d(int w, int h, int d, double px, double py, double pz, Read read) {
	Stencil S;
	clamp_position(w, h, d, px, py, pz, S);
	const double tx = S.x - S.i, ty = S.y - S.j, tz = S.z - S.k;
	double vvv[6];
#pragma unroll 1
	for (int kk = 0; kk < 6; ++kk) {
		double vv[6];
		const int k = clampi(S.k + kk - 2, d);
#pragma unroll 1
		for (int jj = 0; jj < 6; ++jj) {
			double v[6];
			const int j = clampi(S.j + jj - 2, h);
#pragma unroll
			for (int ii = 0; ii < 6; ++ii) v[ii] = (double)read(clampi(S.i + ii - 2, w), j, k);
			vv[jj] = weno6(tx, v);
		}
		vvv[kk] = weno6(ty, vv);
	}
	return weno6(tz, vvv);
}

// macarray3::convert_to_full, face version (macarray3.h:350-372), at the ACTIVE face (DIM; i,j,k): the face's own component, the mean of the four surrounding
// faces for the two others (indices clamped into that component's grid, inactive faces contribute the background 0), summed in double in the order
// (0,0) (0,1) (1,0) (1,1) of (step along DIM, step along the component), /4, rounded to Real.
// MASKED = false: the field holds zeros on its inactive faces (the forward result, written by this file), no mask lookups. DIM is a template parameter so
// that no array of pointers is ever indexed at run time (that would move the kernel parameters into local memory: the first version ran at the speed of L1).
template <class RealT, int DIM, bool MASKED> HD void face_full_velocity(const Mac<RealT> &F, const Grid &g, int i, int j, int k, RealT ur[3]) {
	const int p[3] = {i - (DIM == 0), j - (DIM == 1), k - (DIM == 2)};
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		double u = 0.0;
		if (c == DIM) u = (double)F.v[c][i + (long long)fw(g, c) * (long long)(j + fh(g, c) * k)];
		else {
			// the four faces differ by one step along DIM (outer, ii) and one along c (inner, jj); the third axis is fixed. Every coordinate is clamped into the
			// component's grid on its own, so the flat index is a sum of three per-axis terms, each formed once.
			const int T = 3 - DIM - c;
			const int ext[3] = {fw(g, c), fh(g, c), fd(g, c)};
			const long long stride[3] = {1, ext[0], (long long)ext[0] * ext[1]};
			const long long a0 = clampi(p[DIM], ext[DIM]) * stride[DIM], a1 = clampi(p[DIM] + 1, ext[DIM]) * stride[DIM];
			const long long b0 = clampi(p[c], ext[c]) * stride[c], b1 = clampi(p[c] + 1, ext[c]) * stride[c];
			const long long t0 = clampi(p[T], ext[T]) * stride[T];
			u += (double)face_at<RealT, MASKED>(F, c, t0 + a0 + b0);
			u += (double)face_at<RealT, MASKED>(F, c, t0 + a0 + b1);
			u += (double)face_at<RealT, MASKED>(F, c, t0 + a1 + b0);
			u += (double)face_at<RealT, MASKED>(F, c, t0 + a1 + b1);
			u /= 4.0;
		}
		ur[c] = (RealT)u;
	}
}
// macarray3::convert_to_full, cell version (macarray3.h:305-343), at cell (i,j,k): 0.5 * (the two faces) per component when both are active; the cell
// carries a velocity only when all three components do, else it reads as the background 0.
template <class RealT> HD void cell_full_velocity(const Mac<RealT> &F, const Grid &g, int i, int j, int k, RealT ur[3]) {
	double v[3] = {0.0, 0.0, 0.0};
	int valid = 0;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const int w = fw(g, c), h = fh(g, c);
		const long long n0 = i + (long long)w * (long long)(j + h * k), n1 = (i + (c == 0)) + (long long)w * (long long)((j + (c == 1)) + h * (k + (c == 2)));
		int wsum = 0;
		double value = 0.0;
		if (F.a[c][n0]) { value += (double)F.v[c][n0]; ++wsum; }
		if (F.a[c][n1]) { value += (double)F.v[c][n1]; ++wsum; }
		if (wsum == 2) { v[c] = 0.5 * value; ++valid; }
	}
#pragma unroll
	for (int c = 0; c < 3; ++c) ur[c] = valid == 3 ? (RealT)v[c] : (RealT)0;
}

template <class RealT> struct Limits;
template <> struct Limits<float> {
	static HD double max() { return (double)FLT_MAX; }
	static HD double min() { return (double)FLT_MIN; }
};
template <> struct Limits<double> {
	static HD double max() { return DBL_MAX; }
	static HD double min() { return DBL_MIN; }
};

// ---- faces ------------------------------------------------------------------------------------------------------------------------------------------
// advect_face: the work of ONE active face of direction DIM (k_advect_faces below maps threads to faces).
//   COMBINE = false: advect_semiLagrangian_u(in = F, dt) -> out; RECORD: also the limiter's record (min, max over the eight corners, narrow-band flag)
//   COMBINE = true : the backward pass (F = the forward result, called with -dt) and the limiter (macadvection3.cpp:163-178) in one go: out is the final
//                    field; `orig` = the field before the advection (only this face of it is read, so `out` may be `orig`)
template <class RealT> struct FaceRecord {
	RealT *mn[3], *mx[3];
	uint8_t *nb[3];
};
template <class RealT, bool WENO, bool RECORD, bool COMBINE, int DIM>
HD void advect_face(const Grid &g, const Mac<RealT> &F, double dt, const RealT *__restrict__ fluid, double band, const FaceRecord<RealT> &R, const Mac<RealT> &orig,
                    RealT *out, int i, int j, int k) {
	constexpr bool MASKED = !COMBINE; // the backward pass reads the forward result, which carries zeros on inactive faces
	const int w = fw(g, DIM), h = fh(g, DIM), d = fd(g, DIM);
	const long long plane = (long long)w * h;
	const long long n = i + (long long)w * (long long)(j + h * k);
	RealT ur[3];
	face_full_velocity<RealT, DIM, MASKED>(F, g, i, j, k, ur);
	const bool still = ur[0] == (RealT)0 && ur[1] == (RealT)0 && ur[2] == (RealT)0; // vec::empty()
	auto read = [&](int a, int b, int c) -> RealT { return face_read<RealT, MASKED>(F, g, DIM, a, b, c); };
	const RealT own = F.v[DIM][n];
	// (the limiter's operands do not depend on anything computed here: requested up front, they arrive under the interpolation)
	RealT rec_lo = (RealT)0, rec_hi = (RealT)0, before = (RealT)0;
	uint8_t rec_nb = 0;
	if (COMBINE) { rec_nb = R.nb[DIM][n]; rec_lo = R.mn[DIM][n]; rec_hi = R.mx[DIM][n]; before = orig.v[DIM][n]; }
	RealT value;
	RealT corner[8]; // the eight values of the trilinear stencil, kept for the limiter's min / max when its (double-precision) position lands in the same cell
	int ci = -1, cj = -1, ck = -1;
	if (!still) {
		// p = vec3d(i,j,k) - dt*u/dx with u a vec3<Real>: (Real)(u*dt), then (Real)(that/dx), subtracted in double (macadvection3.cpp:86)
		RealT tx = (RealT)((double)ur[0] * dt), ty = (RealT)((double)ur[1] * dt), tz = (RealT)((double)ur[2] * dt);
		tx = (RealT)over_dx(g, (double)tx); ty = (RealT)over_dx(g, (double)ty); tz = (RealT)over_dx(g, (double)tz);
		const double px = (double)i - (double)tx, py = (double)j - (double)ty, pz = (double)k - (double)tz;
		if (WENO) value = (RealT)weno3d(w, h, d, px, py, pz, read);
		else {
			Stencil S;
			make_stencil(w, h, d, px, py, pz, S);
			ci = S.i; cj = S.j; ck = S.k;
			const long long base = S.i + (long long)w * (long long)(S.j + h * S.k);
#pragma unroll
			for (int e = 0; e < 8; ++e) corner[e] = face_at<RealT, MASKED>(F, DIM, base + (e & 1) + ((e >> 1) & 1) * (long long)w + (e >> 2) * plane);
			value = trilinear_corners<RealT>(S, corner);
		}
	} else value = own;
	if (RECORD) {
		// (macadvection3.cpp:99-139) here u is read back as a vec3d: the position is formed in double throughout
		double min_value, max_value;
		double fx = (double)i + 0.5 * (DIM != 0), fy = (double)j + 0.5 * (DIM != 1), fz = (double)k + 0.5 * (DIM != 2);
		if (!still) {
			const double tx = over_dx(g, (double)ur[0] * dt), ty = over_dx(g, (double)ur[1] * dt), tz = over_dx(g, (double)ur[2] * dt);
			fx = fx - tx; fy = fy - ty; fz = fz - tz;
			Stencil S;
			clamp_position(w, h, d, (double)i - tx, (double)j - ty, (double)k - tz, S);
			const bool same = S.i == ci && S.j == cj && S.k == ck;
			min_value = DBL_MAX;
			max_value = DBL_MIN;
			const long long base = S.i + (long long)w * (long long)(S.j + h * S.k);
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				const double v = (double)(same ? corner[e] : face_at<RealT, MASKED>(F, DIM, base + (e & 1) + ((e >> 1) & 1) * (long long)w + (e >> 2) * plane));
				min_value = std_min(min_value, v);
				max_value = std_max(max_value, v);
			}
		} else min_value = max_value = (double)own;
		Stencil Sf;
		make_stencil(g.nx, g.ny, g.nz, fx - 0.5, fy - 0.5, fz - 0.5, Sf);
		RealT fc[8];
		{
			const long long base = Sf.i + (long long)g.nx * (long long)(Sf.j + g.ny * Sf.k), cplane = (long long)g.nx * g.ny;
#pragma unroll
			for (int e = 0; e < 8; ++e) fc[e] = fluid[base + (e & 1) + ((e >> 1) & 1) * (long long)g.nx + (e >> 2) * cplane];
		}
		const bool within_narrowband = (double)trilinear_corners<RealT>(Sf, fc) > band;
		R.mn[DIM][n] = (RealT)min_value;
		R.mx[DIM][n] = (RealT)max_value;
		R.nb[DIM][n] = within_narrowband ? 1 : 0;
	}
	if (COMBINE) {
		// `value` is velocity_1 (the forward result traced back), F is velocity_0, orig the field before (macadvection3.cpp:163-178)
		if (rec_nb) value = own;
		else {
			const double lo = (double)rec_lo, hi = (double)rec_hi;
			const double vel0 = (double)own;
			const RealT diff = before - value;
			const double correction = 0.5 * (double)diff;
			if (vel0 + correction < lo) value = (RealT)lo;
			else if (vel0 + correction > hi) value = (RealT)hi;
			else value = (RealT)(vel0 + correction);
		}
	}
	out[n] = value;
}
// An inactive face of a forward result is written as 0, which is what lets the backward pass read that field without mask lookups. `orig.a`: the activity.
template <class RealT, bool WENO, bool RECORD, bool COMBINE, int DIM>
HD void advect_face_or_zero(const Grid &g, const Mac<RealT> &F, double dt, const RealT *__restrict__ fluid, double band, const FaceRecord<RealT> &R,
                            const Mac<RealT> &orig, RealT *out, int i, int j, int k) {
	const long long n = i + (long long)fw(g, DIM) * (long long)(j + fh(g, DIM) * k);
	if (orig.a[DIM][n]) advect_face<RealT, WENO, RECORD, COMBINE, DIM>(g, F, dt, fluid, band, R, orig, out, i, j, k);
	else if (!COMBINE) out[n] = (RealT)0;
}
// A block takes ADV_TY rows of one plane of one face grid (blockIdx.y walks the planes of the x-, then y-, then z-faces), 64 threads walk along each row:
// no index divisions, coalesced mask bytes, and a row without an active face costs its mask bytes (and, in a forward pass, its zeros) only — on a liquid scene
// seven faces of eight are inactive, and with one thread per face their index arithmetic was most of the kernel.
constexpr int ADV_TX = 64, ADV_TY = 4;
template <class RealT, bool WENO, bool RECORD, bool COMBINE>
__global__ void __launch_bounds__(ADV_TX *ADV_TY, 4) k_advect_faces(Grid g, Mac<RealT> F, double dt, const RealT *__restrict__ fluid, double band, FaceRecord<RealT> R,
                                                                  Mac<RealT> orig, RealT *out0, RealT *out1, RealT *out2) {
	int kz = blockIdx.y, dim = 0;
	if (kz >= g.nz) { kz -= g.nz; dim = 1; if (kz >= g.nz) { kz -= g.nz; dim = 2; } }
	const int j = blockIdx.x * ADV_TY + threadIdx.y;
	if (j >= fh(g, dim)) return;
	if (dim == 0) for (int i = threadIdx.x; i < g.nx + 1; i += ADV_TX) advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 0>(g, F, dt, fluid, band, R, orig, out0, i, j, kz);
	else if (dim == 1) for (int i = threadIdx.x; i < g.nx; i += ADV_TX) advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 1>(g, F, dt, fluid, band, R, orig, out1, i, j, kz);
	else for (int i = threadIdx.x; i < g.nx; i += ADV_TX) advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 2>(g, F, dt, fluid, band, R, orig, out2, i, j, kz);
}

// ---- cells ------------------------------------------------------------------------------------------------------------------------------------------
// advect_semiLagrangian_cell (macadvection3.cpp:189-237) and the limiter of advect_cell (:257-271); one thread per cell, inactive cells of q leave at once.
// q: the advected grid (dense read), qa its activity; V: the velocity that carries it.
template <class RealT, bool WENO, bool RECORD, bool COMBINE>
HD void advect_cell(const Grid &g, const RealT *__restrict__ q, const uint8_t *__restrict__ qa, const Mac<RealT> &V, double dt, const RealT *__restrict__ fluid,
                    double band, RealT *__restrict__ mn, RealT *__restrict__ mx, uint8_t *__restrict__ nb, const RealT *orig, RealT *out, RealT background,
                    int i, int j, int k) {
	const long long n = i + (long long)g.nx * (long long)(j + g.ny * k);
	if (!qa[n]) {
		// the forward result q_0 is a freshly borrowed grid of q_in's type (macadvection3.cpp:245): off the active set it reads its background value — not
		// the flood-fill value a level set reads inside the liquid —, and the backward pass interpolates in it
		if (!COMBINE) out[n] = background;
		return;
	}
	RealT ur[3];
	cell_full_velocity(V, g, i, j, k, ur);
	const bool still = ur[0] == (RealT)0 && ur[1] == (RealT)0 && ur[2] == (RealT)0;
	auto read = [&](int a, int b, int c) -> RealT { return q[a + (long long)g.nx * (long long)(b + g.ny * c)]; };
	const double p[3] = {(double)i - over_dx(g, (double)ur[0] * dt), (double)j - over_dx(g, (double)ur[1] * dt), (double)k - over_dx(g, (double)ur[2] * dt)};
	RealT value;
	Stencil S;
	RealT corner[8];
	const long long cplane = (long long)g.nx * g.ny;
	long long cbase = 0;
	if (!still) {
		if (!WENO || RECORD) {
			make_stencil(g.nx, g.ny, g.nz, p[0], p[1], p[2], S);
			cbase = S.i + (long long)g.nx * (long long)(S.j + g.ny * S.k);
#pragma unroll
			for (int e = 0; e < 8; ++e) corner[e] = q[cbase + (e & 1) + ((e >> 1) & 1) * (long long)g.nx + (e >> 2) * cplane];
		}
		if (WENO) value = (RealT)weno3d(g.nx, g.ny, g.nz, p[0], p[1], p[2], read);
		else value = trilinear_corners<RealT>(S, corner);
	} else value = q[n];
	if (RECORD) {
		double min_value, max_value;
		bool within_narrowband;
		if (!still) {
			min_value = Limits<RealT>::max();
			max_value = Limits<RealT>::min();
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				const double v = (double)corner[e];
				min_value = std_min(min_value, v);
				max_value = std_max(max_value, v);
			}
			RealT fc[8]; // (same shape, same position: same weights, same flat indices)
#pragma unroll
			for (int e = 0; e < 8; ++e) fc[e] = fluid[cbase + (e & 1) + ((e >> 1) & 1) * (long long)g.nx + (e >> 2) * cplane];
			within_narrowband = (double)trilinear_corners<RealT>(S, fc) > band;
		} else {
			min_value = max_value = (double)q[n];
			within_narrowband = (double)fluid[n] > band;
		}
		mn[n] = (RealT)min_value;
		mx[n] = (RealT)max_value;
		nb[n] = within_narrowband ? 1 : 0;
	}
	if (COMBINE) {
		if (nb[n]) value = q[n];
		else {
			const double lo = (double)mn[n], hi = (double)mx[n];
			const double q0 = (double)q[n];
			const RealT diff = orig[n] - value;
			const double correction = 0.5 * (double)diff;
			if (q0 + correction < lo) value = (RealT)lo;
			else if (q0 + correction > hi) value = (RealT)hi;
			else value = (RealT)(q0 + correction);
		}
	}
	out[n] = value;
}
template <class RealT, bool WENO, bool RECORD, bool COMBINE>
__global__ void __launch_bounds__(ADV_TX *ADV_TY) k_advect_cells(Grid g, const RealT *__restrict__ q, const uint8_t *__restrict__ qa, Mac<RealT> V, double dt,
                                                                  const RealT *__restrict__ fluid, double band, RealT *__restrict__ mn, RealT *__restrict__ mx,
                                                                  uint8_t *__restrict__ nb, const RealT *orig, RealT *out, RealT background) {
	const int j = blockIdx.x * ADV_TY + threadIdx.y;
	if (j >= g.ny) return;
	for (int i = threadIdx.x; i < g.nx; i += ADV_TX)
		advect_cell<RealT, WENO, RECORD, COMBINE>(g, q, qa, V, dt, fluid, band, mn, mx, nb, orig, out, background, i, j, (int)blockIdx.y);
}

// out[n] = src[n] on active entries (the plain semi-Lagrangian result goes back into the caller's grid)
template <class RealT> __global__ void __launch_bounds__(ADV_THREADS) k_copy_active(long long count, const RealT *__restrict__ src, const uint8_t *__restrict__ act, RealT *__restrict__ dst) {
	const long long n = (long long)blockIdx.x * ADV_THREADS + threadIdx.x;
	if (n < count && act[n]) dst[n] = src[n];
}

// ---- sparse host copies of advect_vector_host (page-locked buffers: the GPU addresses them directly) -------------------------------------------------
// An advection reads the VALUES of active faces only (an inactive face reads as 0 whatever its entry holds) and writes active faces only, so on a liquid scene
// — one face in eight active — the masks travel whole and the values by these kernels: four faces per thread, 16-byte accesses where all four are active.
__global__ void __launch_bounds__(256) k_count_active(long long n, const uint8_t *__restrict__ act, unsigned long long *__restrict__ total) {
	unsigned c = 0;
	for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < (n + 3) >> 2; q += (long long)gridDim.x * blockDim.x) {
		const long long f = q << 2;
		if (f + 3 < n) { const uchar4 m = *reinterpret_cast<const uchar4 *>(act + f); c += (m.x != 0) + (m.y != 0) + (m.z != 0) + (m.w != 0); }
		else for (long long e = f; e < n; ++e) c += act[e] != 0;
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, (unsigned long long)c);
}
// dst[f] = src[f] on active faces (pull: src = host, dst = device staging; push: the other way round)
template <class RealT>
__global__ void __launch_bounds__(256) k_move_active(long long n, const uint8_t *__restrict__ act, const RealT *__restrict__ src, RealT *__restrict__ dst) {
	const bool vec_ok = sizeof(RealT) == 4 && ((reinterpret_cast<unsigned long long>(src) | reinterpret_cast<unsigned long long>(dst)) & 15ull) == 0ull;
	for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < (n + 3) >> 2; q += (long long)gridDim.x * blockDim.x) {
		const long long f = q << 2;
		if (f + 3 < n) {
			const uchar4 m = *reinterpret_cast<const uchar4 *>(act + f);
			if (!(m.x | m.y | m.z | m.w)) continue;
			if (vec_ok && m.x && m.y && m.z && m.w) { *reinterpret_cast<float4 *>(dst + f) = *reinterpret_cast<const float4 *>(src + f); continue; }
			if (m.x) dst[f] = src[f];
			if (m.y) dst[f + 1] = src[f + 1];
			if (m.z) dst[f + 2] = src[f + 2];
			if (m.w) dst[f + 3] = src[f + 3];
		} else for (long long e = f; e < n; ++e) if (act[e]) dst[e] = src[e];
	}
}

struct Buf {
	void *p = nullptr;
	size_t bytes = 0;
	int need(size_t n) {
		if (n <= bytes) return SHKZ_B200_OK;
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
		CK(cudaMalloc(&p, n));
		CK(cudaMemset(p, 0, n));
		bytes = n;
		return SHKZ_B200_OK;
	}
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
	}
};

} // namespace

struct shkz_b200_advect {
	int device = 0;
	Grid g{};
	int real = 0;
	size_t rb = 4;
	// work arrays of the MacCormack scheme (device): forward result, limiter record — faces [0..2], cells [3]
	Buf fwd[4], mn[4], mx[4], nb[4];
	// staging of the `_host` entry points
	Buf st_val[4], st_act[4], st_fluid, st_count;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	uint64_t launches = 0;
};

namespace {

size_t face_count(const Grid &g, int dim) { return (size_t)fw(g, dim) * fh(g, dim) * fd(g, dim); }
size_t cell_count(const Grid &g) { return (size_t)g.nx * g.ny * g.nz; }

// device-visible address of a page-locked host buffer; nullptr for pageable memory
void *mapped_host(const void *p) {
	if (!p) return nullptr;
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int device_ready(int device) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return fail(SHKZ_B200_ERR_NO_DEVICE, "no CUDA device available; libshkz_b200 has no CPU fallback");
	}
	if (device < 0 || device >= n) return fail(SHKZ_B200_ERR_ARG, "device %d out of range (%d visible)", device, n);
	return SHKZ_B200_OK;
}

int check(const shkz_b200_advect_params *in, shkz_b200_advect_params &P) {
	shkz_b200_advect_default_params(&P);
	if (in) {
		if (in->struct_size != sizeof(shkz_b200_advect_params)) return fail(SHKZ_B200_ERR_ARG, "params.struct_size %u != %zu", in->struct_size, sizeof(shkz_b200_advect_params));
		P = *in;
	}
	return SHKZ_B200_OK;
}

template <class RealT>
int vector_device(shkz_b200_advect *A, double dt, void *const u[3], const uint8_t *const act[3], const void *fluid, const shkz_b200_advect_params &P, cudaStream_t s) {
	const Grid &g = A->g;
	Mac<RealT> U, F0;
	FaceRecord<RealT> R;
	for (int dim = 0; dim < 3; ++dim) {
		const size_t nf = face_count(g, dim);
		CKR(A->fwd[dim].need(nf * sizeof(RealT)));
		if (P.maccormack) { CKR(A->mn[dim].need(nf * sizeof(RealT))); CKR(A->mx[dim].need(nf * sizeof(RealT))); CKR(A->nb[dim].need(nf)); }
		U.v[dim] = static_cast<const RealT *>(u[dim]); U.a[dim] = act[dim];
		F0.v[dim] = static_cast<const RealT *>(A->fwd[dim].p); F0.a[dim] = act[dim];
		R.mn[dim] = static_cast<RealT *>(A->mn[dim].p); R.mx[dim] = static_cast<RealT *>(A->mx[dim].p); R.nb[dim] = static_cast<uint8_t *>(A->nb[dim].p);
	}
	RealT *f0 = static_cast<RealT *>(A->fwd[0].p), *f1 = static_cast<RealT *>(A->fwd[1].p), *f2 = static_cast<RealT *>(A->fwd[2].p);
	const dim3 grid((unsigned)((g.ny + 1 + ADV_TY - 1) / ADV_TY), (unsigned)(3 * g.nz + 1)), block(ADV_TX, ADV_TY);
	const RealT *fl = static_cast<const RealT *>(fluid);
	const double band = -g.dx * (double)P.trim_narrowband;
	if (P.maccormack) {
		if (P.weno) {
			k_advect_faces<RealT, true, true, false><<<grid, block, 0, s>>>(g, U, dt, fl, band, R, U, f0, f1, f2);
			k_advect_faces<RealT, true, false, true><<<grid, block, 0, s>>>(g, F0, -dt, fl, band, R, U, (RealT *)u[0], (RealT *)u[1], (RealT *)u[2]);
		} else {
			k_advect_faces<RealT, false, true, false><<<grid, block, 0, s>>>(g, U, dt, fl, band, R, U, f0, f1, f2);
			k_advect_faces<RealT, false, false, true><<<grid, block, 0, s>>>(g, F0, -dt, fl, band, R, U, (RealT *)u[0], (RealT *)u[1], (RealT *)u[2]);
		}
		A->launches += 2;
	} else {
		if (P.weno) k_advect_faces<RealT, true, false, false><<<grid, block, 0, s>>>(g, U, dt, fl, band, R, U, f0, f1, f2);
		else k_advect_faces<RealT, false, false, false><<<grid, block, 0, s>>>(g, U, dt, fl, band, R, U, f0, f1, f2);
		for (int dim = 0; dim < 3; ++dim) {
			const long long nf = (long long)face_count(g, dim);
			k_copy_active<RealT><<<(unsigned)((nf + ADV_THREADS - 1) / ADV_THREADS), ADV_THREADS, 0, s>>>(nf, F0.v[dim], act[dim], (RealT *)u[dim]);
		}
		A->launches += 4;
	}
	CK(cudaGetLastError());
	return SHKZ_B200_OK;
}

template <class RealT>
int scalar_device(shkz_b200_advect *A, double dt, void *q, const uint8_t *qact, const void *const vel[3], const uint8_t *const vact[3], const void *fluid,
                  const shkz_b200_advect_params &P, cudaStream_t s) {
	const Grid &g = A->g;
	const size_t nc = cell_count(g);
	CKR(A->fwd[3].need(nc * sizeof(RealT)));
	if (P.maccormack) { CKR(A->mn[3].need(nc * sizeof(RealT))); CKR(A->mx[3].need(nc * sizeof(RealT))); CKR(A->nb[3].need(nc)); }
	Mac<RealT> V;
	for (int dim = 0; dim < 3; ++dim) { V.v[dim] = static_cast<const RealT *>(vel[dim]); V.a[dim] = vact[dim]; }
	RealT *q0 = static_cast<RealT *>(A->fwd[3].p), *mn = static_cast<RealT *>(A->mn[3].p), *mx = static_cast<RealT *>(A->mx[3].p), *qq = static_cast<RealT *>(q);
	uint8_t *nb = static_cast<uint8_t *>(A->nb[3].p);
	const RealT *fl = static_cast<const RealT *>(fluid);
	const double band = -g.dx * (double)P.trim_narrowband;
	const dim3 grid((unsigned)((g.ny + ADV_TY - 1) / ADV_TY), (unsigned)g.nz), block(ADV_TX, ADV_TY);
	const RealT bg = (RealT)P.scalar_background;
	if (P.maccormack) {
		if (P.weno) {
			k_advect_cells<RealT, true, true, false><<<grid, block, 0, s>>>(g, qq, qact, V, dt, fl, band, mn, mx, nb, qq, q0, bg);
			k_advect_cells<RealT, true, false, true><<<grid, block, 0, s>>>(g, q0, qact, V, -dt, fl, band, mn, mx, nb, qq, qq, bg);
		} else {
			k_advect_cells<RealT, false, true, false><<<grid, block, 0, s>>>(g, qq, qact, V, dt, fl, band, mn, mx, nb, qq, q0, bg);
			k_advect_cells<RealT, false, false, true><<<grid, block, 0, s>>>(g, q0, qact, V, -dt, fl, band, mn, mx, nb, qq, qq, bg);
		}
		A->launches += 2;
	} else {
		if (P.weno) k_advect_cells<RealT, true, false, false><<<grid, block, 0, s>>>(g, qq, qact, V, dt, fl, band, mn, mx, nb, qq, q0, bg);
		else k_advect_cells<RealT, false, false, false><<<grid, block, 0, s>>>(g, qq, qact, V, dt, fl, band, mn, mx, nb, qq, q0, bg);
		k_copy_active<RealT><<<(unsigned)((nc + ADV_THREADS - 1) / ADV_THREADS), ADV_THREADS, 0, s>>>((long long)nc, q0, qact, qq);
		A->launches += 2;
	}
	CK(cudaGetLastError());
	return SHKZ_B200_OK;
}

// Face grids of a `_host` call -> device staging: the masks whole; the values whole through the copy engines, or — page-locked buffers, at most half of the
// faces active, SHKZ_B200_HOST_COPIES != dense — only those of the active faces, by a kernel that reads the host buffers directly. *sparse tells which it was.
int upload_faces(shkz_b200_advect *A, const void *const val[3], const uint8_t *const act[3], cudaStream_t s, void *dval[3], const uint8_t *dact[3], void *hmap[3],
                 bool *sparse, uint64_t *n_active, uint64_t *h2d) {
	const Grid &g = A->g;
	const char *mode = getenv("SHKZ_B200_HOST_COPIES");
	*sparse = !(mode && !strcmp(mode, "dense"));
	for (int dim = 0; dim < 3; ++dim) {
		hmap[dim] = mapped_host(val[dim]);
		if (!hmap[dim]) *sparse = false;
	}
	for (int dim = 0; dim < 3; ++dim) {
		const size_t nf = face_count(g, dim);
		CKR(A->st_val[dim].need(nf * A->rb));
		CKR(A->st_act[dim].need(nf));
		CK(cudaMemcpyAsync(A->st_act[dim].p, act[dim], nf, cudaMemcpyHostToDevice, s));
		dval[dim] = A->st_val[dim].p;
		dact[dim] = static_cast<const uint8_t *>(A->st_act[dim].p);
		*h2d += nf;
	}
	uint64_t n_faces = 0;
	*n_active = 0;
	if (*sparse) { // how many faces are active decides: above half of them the copy engines move whole arrays faster than kernels move the active entries
		CKR(A->st_count.need(sizeof(unsigned long long)));
		CK(cudaMemsetAsync(A->st_count.p, 0, sizeof(unsigned long long), s));
		for (int dim = 0; dim < 3; ++dim) {
			const long long nf = (long long)face_count(g, dim);
			k_count_active<<<148 * 8, 256, 0, s>>>(nf, dact[dim], static_cast<unsigned long long *>(A->st_count.p));
			n_faces += nf;
		}
		A->launches += 3;
		unsigned long long c = 0;
		CK(cudaMemcpyAsync(&c, A->st_count.p, sizeof c, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		*n_active = c;
		if (2 * c > n_faces) *sparse = false;
	}
	for (int dim = 0; dim < 3; ++dim) {
		const long long nf = (long long)face_count(g, dim);
		if (*sparse) {
			if (A->real == SHKZ_B200_REAL_F64) k_move_active<double><<<148 * 8, 256, 0, s>>>(nf, dact[dim], static_cast<const double *>(hmap[dim]), static_cast<double *>(dval[dim]));
			else k_move_active<float><<<148 * 8, 256, 0, s>>>(nf, dact[dim], static_cast<const float *>(hmap[dim]), static_cast<float *>(dval[dim]));
			++A->launches;
		} else {
			CK(cudaMemcpyAsync(dval[dim], val[dim], nf * A->rb, cudaMemcpyHostToDevice, s));
			*h2d += nf * A->rb;
		}
	}
	if (*sparse) *h2d += *n_active * A->rb;
	CK(cudaGetLastError());
	return SHKZ_B200_OK;
}

} // namespace

extern "C" {

const char *shkz_b200_advect_last_error(void) { return g_adv_error.c_str(); }

void shkz_b200_advect_default_params(shkz_b200_advect_params *P) {
	if (!P) return;
	memset(P, 0, sizeof *P);
	P->struct_size = sizeof *P;
	P->maccormack = 1;      // macadvection3.cpp:286
	P->weno = 0;            // :287
	P->trim_narrowband = 1; // :288
	P->scalar_background = 0.0;
}

int shkz_b200_advect_create(int nx, int ny, int nz, double dx, int real, int device, shkz_b200_advect **out) {
	if (!out) return fail(SHKZ_B200_ERR_ARG, "out is NULL");
	*out = nullptr;
	if (nx < 2 || ny < 2 || nz < 2) return fail(SHKZ_B200_ERR_ARG, "grid %dx%dx%d: every extent must be >= 2", nx, ny, nz);
	if (!(dx > 0.0)) return fail(SHKZ_B200_ERR_ARG, "dx must be positive");
	if (real != SHKZ_B200_REAL_F32 && real != SHKZ_B200_REAL_F64) return fail(SHKZ_B200_ERR_ARG, "real must be SHKZ_B200_REAL_F32 or _F64");
	CKR(device_ready(device));
	CK(cudaSetDevice(device));
	shkz_b200_advect *A = new shkz_b200_advect;
	A->device = device;
	A->g = make_grid(nx, ny, nz, dx);
	A->real = real;
	A->rb = real == SHKZ_B200_REAL_F64 ? 8 : 4;
	for (auto &e : A->ev)
		if (cudaEventCreate(&e) != cudaSuccess) {
			shkz_b200_advect_destroy(A);
			return fail(SHKZ_B200_ERR_CUDA, "cudaEventCreate failed");
		}
	*out = A;
	return SHKZ_B200_OK;
}

void shkz_b200_advect_destroy(shkz_b200_advect *A) {
	if (!A) return;
	cudaSetDevice(A->device);
	for (int n = 0; n < 4; ++n) { A->fwd[n].release(); A->mn[n].release(); A->mx[n].release(); A->nb[n].release(); A->st_val[n].release(); A->st_act[n].release(); }
	A->st_fluid.release(); A->st_count.release();
	for (auto &e : A->ev) if (e) cudaEventDestroy(e);
	delete A;
}

int shkz_b200_advect_vector_device(shkz_b200_advect *A, double dt, void *const u[3], const uint8_t *const u_active[3], const void *fluid,
                                   const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats, void *cuda_stream) {
	if (!A) return fail(SHKZ_B200_ERR_ARG, "advect handle is NULL");
	shkz_b200_advect_params P;
	CKR(check(params, P));
	if (!u || !u_active) return fail(SHKZ_B200_ERR_ARG, "u / u_active must not be NULL");
	for (int dim = 0; dim < 3; ++dim) if (!u[dim] || !u_active[dim]) return fail(SHKZ_B200_ERR_ARG, "u[%d] / u_active[%d] is NULL", dim, dim);
	if (P.maccormack && !fluid) return fail(SHKZ_B200_ERR_ARG, "MacCormack needs the liquid level set (pass the constant grid of a smoke solver)");
	CKR(device_ready(A->device));
	CK(cudaSetDevice(A->device));
	cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
	const uint64_t before = A->launches;
	CK(cudaEventRecord(A->ev[0], s));
	if (A->real == SHKZ_B200_REAL_F64) CKR(vector_device<double>(A, dt, u, u_active, fluid, P, s));
	else CKR(vector_device<float>(A, dt, u, u_active, fluid, P, s));
	CK(cudaEventRecord(A->ev[1], s));
	CK(cudaStreamSynchronize(s));
	if (stats) {
		memset(stats, 0, sizeof *stats);
		stats->kernel_launches = A->launches - before;
		CK(cudaEventElapsedTime(&stats->ms_advect, A->ev[0], A->ev[1]));
	}
	return SHKZ_B200_OK;
}

int shkz_b200_advect_scalar_device(shkz_b200_advect *A, double dt, void *q, const uint8_t *q_active, const void *const vel[3], const uint8_t *const vel_active[3],
                                   const void *fluid, const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats, void *cuda_stream) {
	if (!A) return fail(SHKZ_B200_ERR_ARG, "advect handle is NULL");
	shkz_b200_advect_params P;
	CKR(check(params, P));
	if (!q || !q_active) return fail(SHKZ_B200_ERR_ARG, "q / q_active must not be NULL");
	if (!vel || !vel_active) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active must not be NULL");
	for (int dim = 0; dim < 3; ++dim) if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
	if (P.maccormack && !fluid) return fail(SHKZ_B200_ERR_ARG, "MacCormack needs the liquid level set (pass the constant grid of a smoke solver)");
	CKR(device_ready(A->device));
	CK(cudaSetDevice(A->device));
	cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
	const uint64_t before = A->launches;
	CK(cudaEventRecord(A->ev[0], s));
	if (A->real == SHKZ_B200_REAL_F64) CKR(scalar_device<double>(A, dt, q, q_active, vel, vel_active, fluid, P, s));
	else CKR(scalar_device<float>(A, dt, q, q_active, vel, vel_active, fluid, P, s));
	CK(cudaEventRecord(A->ev[1], s));
	CK(cudaStreamSynchronize(s));
	if (stats) {
		memset(stats, 0, sizeof *stats);
		stats->kernel_launches = A->launches - before;
		CK(cudaEventElapsedTime(&stats->ms_advect, A->ev[0], A->ev[1]));
	}
	return SHKZ_B200_OK;
}

int shkz_b200_advect_vector_host(shkz_b200_advect *A, double dt, void *const u[3], const uint8_t *const u_active[3], const void *fluid,
                                 const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats) {
	if (!A) return fail(SHKZ_B200_ERR_ARG, "advect handle is NULL");
	if (!u || !u_active) return fail(SHKZ_B200_ERR_ARG, "u / u_active must not be NULL");
	for (int dim = 0; dim < 3; ++dim) if (!u[dim] || !u_active[dim]) return fail(SHKZ_B200_ERR_ARG, "u[%d] / u_active[%d] is NULL", dim, dim);
	CKR(device_ready(A->device));
	CK(cudaSetDevice(A->device));
	cudaStream_t s = nullptr;
	const Grid &g = A->g;
	void *du[3], *hmap[3];
	const uint8_t *da[3];
	uint64_t h2d = 0, d2h = 0, n_active = 0;
	bool sparse = false;
	CK(cudaEventRecord(A->ev[2], s));
	CKR(upload_faces(A, u, u_active, s, du, da, hmap, &sparse, &n_active, &h2d));
	if (fluid) {
		CKR(A->st_fluid.need(cell_count(g) * A->rb));
		CK(cudaMemcpyAsync(A->st_fluid.p, fluid, cell_count(g) * A->rb, cudaMemcpyHostToDevice, s));
		h2d += cell_count(g) * A->rb;
	}
	CK(cudaEventRecord(A->ev[3], s));
	shkz_b200_advect_stats st;
	CKR(shkz_b200_advect_vector_device(A, dt, du, da, fluid ? A->st_fluid.p : nullptr, params, &st, s));
	float ms_h2d = 0.f, ms_d2h = 0.f;
	CK(cudaEventElapsedTime(&ms_h2d, A->ev[2], A->ev[3]));
	CK(cudaEventRecord(A->ev[2], s));
	for (int dim = 0; dim < 3; ++dim) {
		const long long nf = (long long)face_count(g, dim);
		if (sparse) {
			if (A->real == SHKZ_B200_REAL_F64) k_move_active<double><<<148 * 8, 256, 0, s>>>(nf, da[dim], static_cast<const double *>(du[dim]), static_cast<double *>(hmap[dim]));
			else k_move_active<float><<<148 * 8, 256, 0, s>>>(nf, da[dim], static_cast<const float *>(du[dim]), static_cast<float *>(hmap[dim]));
			++A->launches;
		} else {
			CK(cudaMemcpyAsync(u[dim], du[dim], nf * A->rb, cudaMemcpyDeviceToHost, s));
			d2h += nf * A->rb;
		}
	}
	if (sparse) d2h += n_active * A->rb;
	CK(cudaGetLastError());
	CK(cudaEventRecord(A->ev[3], s));
	CK(cudaStreamSynchronize(s));
	CK(cudaEventElapsedTime(&ms_d2h, A->ev[2], A->ev[3]));
	if (stats) {
		*stats = st;
		stats->kernel_launches += sparse ? 9 : 0;
		stats->host_copies = sparse ? 1 : 0;
		stats->ms_h2d = ms_h2d; stats->ms_d2h = ms_d2h; stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
	}
	return SHKZ_B200_OK;
}

int shkz_b200_advect_scalar_host(shkz_b200_advect *A, double dt, void *q, const uint8_t *q_active, const void *const vel[3], const uint8_t *const vel_active[3],
                                 const void *fluid, const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats) {
	if (!A) return fail(SHKZ_B200_ERR_ARG, "advect handle is NULL");
	if (!q || !q_active) return fail(SHKZ_B200_ERR_ARG, "q / q_active must not be NULL");
	if (!vel || !vel_active) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active must not be NULL");
	for (int dim = 0; dim < 3; ++dim) if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
	CKR(device_ready(A->device));
	CK(cudaSetDevice(A->device));
	cudaStream_t s = nullptr;
	const Grid &g = A->g;
	const size_t nc = cell_count(g);
	void *dv[3], *hmap[3];
	const uint8_t *da[3];
	uint64_t h2d = 0, n_active = 0;
	bool sparse = false;
	CK(cudaEventRecord(A->ev[2], s));
	CKR(upload_faces(A, vel, vel_active, s, dv, da, hmap, &sparse, &n_active, &h2d)); // (the velocity is only read: nothing of it travels back)
	CKR(A->st_val[3].need(nc * A->rb));
	CKR(A->st_act[3].need(nc));
	CK(cudaMemcpyAsync(A->st_val[3].p, q, nc * A->rb, cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(A->st_act[3].p, q_active, nc, cudaMemcpyHostToDevice, s));
	h2d += nc * (A->rb + 1);
	const void *dfluid = nullptr;
	if (fluid) { // (also when fluid == q, a level set carried by itself: the advection overwrites q's copy, the narrow-band test reads the grid before)
		CKR(A->st_fluid.need(nc * A->rb));
		CK(cudaMemcpyAsync(A->st_fluid.p, fluid, nc * A->rb, cudaMemcpyHostToDevice, s));
		h2d += nc * A->rb;
		dfluid = A->st_fluid.p;
	}
	CK(cudaEventRecord(A->ev[3], s));
	shkz_b200_advect_stats st;
	CKR(shkz_b200_advect_scalar_device(A, dt, A->st_val[3].p, static_cast<const uint8_t *>(A->st_act[3].p), dv, da, dfluid, params, &st, s));
	float ms_h2d = 0.f, ms_d2h = 0.f;
	CK(cudaEventElapsedTime(&ms_h2d, A->ev[2], A->ev[3]));
	CK(cudaEventRecord(A->ev[2], s));
	CK(cudaMemcpyAsync(q, A->st_val[3].p, nc * A->rb, cudaMemcpyDeviceToHost, s));
	CK(cudaEventRecord(A->ev[3], s));
	CK(cudaStreamSynchronize(s));
	CK(cudaEventElapsedTime(&ms_d2h, A->ev[2], A->ev[3]));
	if (stats) {
		*stats = st;
		stats->kernel_launches += sparse ? 6 : 0;
		stats->host_copies = sparse ? 1 : 0;
		stats->ms_h2d = ms_h2d; stats->ms_d2h = ms_d2h; stats->h2d_bytes = h2d; stats->d2h_bytes = nc * A->rb;
	}
	return SHKZ_B200_OK;
}

} // extern "C"

#ifdef SHKZ_B200_ADVECT_HOSTCHECK
// ---- TEST INFRASTRUCTURE (never in libshkz_b200.so: tests/ build this file once more with -DSHKZ_B200_ADVECT_HOSTCHECK into oracle/_build/) -------------------
// The per-face / per-cell bodies above are __host__ __device__: here plain host loops run the SAME source over a grid, so that the CPU suite can hold it
// against the unmodified reference module where the reference exists (this container), before any GPU time is spent. It checks the restatement, not the
// product: nothing in shiokaze_b200/ loads it.
#include <vector>
namespace {
template <class RealT, bool WENO, bool RECORD, bool COMBINE>
void advect_face_any(const Grid &g, const Mac<RealT> &F, double dt, const RealT *fluid, double band, const FaceRecord<RealT> &R, const Mac<RealT> &orig, RealT *out0,
                     RealT *out1, RealT *out2, int dim, int i, int j, int k) {
	if (dim == 0) advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 0>(g, F, dt, fluid, band, R, orig, out0, i, j, k);
	else if (dim == 1) advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 1>(g, F, dt, fluid, band, R, orig, out1, i, j, k);
	else advect_face_or_zero<RealT, WENO, RECORD, COMBINE, 2>(g, F, dt, fluid, band, R, orig, out2, i, j, k);
}
template <class RealT, bool WENO>
void hostcheck_vector(const Grid &g, double dt, RealT *const u[3], const uint8_t *const act[3], const RealT *fluid, const shkz_b200_advect_params &P) {
	std::vector<RealT> fwd[3], mn[3], mx[3];
	std::vector<uint8_t> nb[3];
	Mac<RealT> U, F0;
	FaceRecord<RealT> R;
	for (int dim = 0; dim < 3; ++dim) {
		const size_t nf = face_count(g, dim);
		fwd[dim].assign(nf, (RealT)0); mn[dim].assign(nf, (RealT)0); mx[dim].assign(nf, (RealT)0); nb[dim].assign(nf, 0);
		U.v[dim] = u[dim]; U.a[dim] = act[dim]; F0.v[dim] = fwd[dim].data(); F0.a[dim] = act[dim];
		R.mn[dim] = mn[dim].data(); R.mx[dim] = mx[dim].data(); R.nb[dim] = nb[dim].data();
	}
	const double band = -g.dx * (double)P.trim_narrowband;
	auto sweep = [&](auto body) {
		for (int dim = 0; dim < 3; ++dim)
			for (int k = 0; k < fd(g, dim); ++k)
				for (int j = 0; j < fh(g, dim); ++j)
					for (int i = 0; i < fw(g, dim); ++i)
						body(dim, i, j, k);
	};
	if (P.maccormack) {
		sweep([&](int dim, int i, int j, int k) { advect_face_any<RealT, WENO, true, false>(g, U, dt, fluid, band, R, U, fwd[0].data(), fwd[1].data(), fwd[2].data(), dim, i, j, k); });
		sweep([&](int dim, int i, int j, int k) { advect_face_any<RealT, WENO, false, true>(g, F0, -dt, fluid, band, R, U, u[0], u[1], u[2], dim, i, j, k); });
	} else {
		sweep([&](int dim, int i, int j, int k) { advect_face_any<RealT, WENO, false, false>(g, U, dt, fluid, band, R, U, fwd[0].data(), fwd[1].data(), fwd[2].data(), dim, i, j, k); });
		sweep([&](int dim, int i, int j, int k) { const size_t n = i + (size_t)fw(g, dim) * (j + (size_t)fh(g, dim) * k); if (act[dim][n]) u[dim][n] = fwd[dim][n]; });
	}
}
template <class RealT, bool WENO>
void hostcheck_scalar(const Grid &g, double dt, RealT *q, const uint8_t *qa, const RealT *const vel[3], const uint8_t *const vact[3], const RealT *fluid,
                      const shkz_b200_advect_params &P) {
	const size_t nc = cell_count(g);
	std::vector<RealT> q0(nc, (RealT)0), mn(nc, (RealT)0), mx(nc, (RealT)0);
	std::vector<uint8_t> nb(nc, 0);
	Mac<RealT> V;
	for (int dim = 0; dim < 3; ++dim) { V.v[dim] = vel[dim]; V.a[dim] = vact[dim]; }
	const double band = -g.dx * (double)P.trim_narrowband;
	const RealT bg = (RealT)P.scalar_background;
	auto sweep = [&](auto body) {
		for (int k = 0; k < g.nz; ++k)
			for (int j = 0; j < g.ny; ++j)
				for (int i = 0; i < g.nx; ++i) body(i, j, k);
	};
	if (P.maccormack) {
		sweep([&](int i, int j, int k) { advect_cell<RealT, WENO, true, false>(g, q, qa, V, dt, fluid, band, mn.data(), mx.data(), nb.data(), q, q0.data(), bg, i, j, k); });
		sweep([&](int i, int j, int k) { advect_cell<RealT, WENO, false, true>(g, q0.data(), qa, V, -dt, fluid, band, mn.data(), mx.data(), nb.data(), q, q, bg, i, j, k); });
	} else {
		sweep([&](int i, int j, int k) { advect_cell<RealT, WENO, false, false>(g, q, qa, V, dt, fluid, band, mn.data(), mx.data(), nb.data(), q, q0.data(), bg, i, j, k); });
		for (size_t n = 0; n < nc; ++n) if (qa[n]) q[n] = q0[n];
	}
}
} // namespace
extern "C" int shkz_b200_hostcheck_advect_vector(int nx, int ny, int nz, double dx, int real, double dt, void *const u[3], const uint8_t *const act[3], const void *fluid,
                                                 const shkz_b200_advect_params *params) {
	const Grid g = make_grid(nx, ny, nz, dx);
	const shkz_b200_advect_params &P = *params;
	if (real == SHKZ_B200_REAL_F64) {
		double *uu[3] = {(double *)u[0], (double *)u[1], (double *)u[2]};
		if (P.weno) hostcheck_vector<double, true>(g, dt, uu, act, (const double *)fluid, P); else hostcheck_vector<double, false>(g, dt, uu, act, (const double *)fluid, P);
	} else {
		float *uu[3] = {(float *)u[0], (float *)u[1], (float *)u[2]};
		if (P.weno) hostcheck_vector<float, true>(g, dt, uu, act, (const float *)fluid, P); else hostcheck_vector<float, false>(g, dt, uu, act, (const float *)fluid, P);
	}
	return 0;
}
extern "C" int shkz_b200_hostcheck_advect_scalar(int nx, int ny, int nz, double dx, int real, double dt, void *q, const uint8_t *qa, const void *const vel[3],
                                                 const uint8_t *const vact[3], const void *fluid, const shkz_b200_advect_params *params) {
	const Grid g = make_grid(nx, ny, nz, dx);
	const shkz_b200_advect_params &P = *params;
	if (real == SHKZ_B200_REAL_F64) {
		const double *vv[3] = {(const double *)vel[0], (const double *)vel[1], (const double *)vel[2]};
		if (P.weno) hostcheck_scalar<double, true>(g, dt, (double *)q, qa, vv, vact, (const double *)fluid, P); else hostcheck_scalar<double, false>(g, dt, (double *)q, qa, vv, vact, (const double *)fluid, P);
	} else {
		const float *vv[3] = {(const float *)vel[0], (const float *)vel[1], (const float *)vel[2]};
		if (P.weno) hostcheck_scalar<float, true>(g, dt, (float *)q, qa, vv, vact, (const float *)fluid, P); else hostcheck_scalar<float, false>(g, dt, (float *)q, qa, vv, vact, (const float *)fluid, P);
	}
	return 0;
}
#endif
