// Coefficient assembly and velocity update kernels (replace the reference's CPU loops K1-K6, K12-K14
// of SURVEY.md section 2d). Arithmetic that must agree bit-for-bit with the reference (fractions,
// matrix entries, right-hand side) uses the explicit round-to-nearest intrinsics so that nvcc never
// contracts it into FMAs; the operation order is the reference's.
#pragma once
#include <float.h>
#include "common.cuh"

namespace shkz {

struct AsmParams {
	double dt, dx;
	double eps_fluid, eps_solid;
	double surface_tension;
	double rhs_correct;
	int second_order_fluid, second_order_solid;
	int have_solid, fluid_levelset;
	int apply_rhs_correct;
	int dx_pow2;   // dx is an exact power of two: x / dx == x * inv_dx bit for bit
	double inv_dx;
};

// x / dx with the reference's rounding; a multiplication when that is exact
__device__ __forceinline__ double div_dx(const AsmParams &P, double x) { return P.dx_pow2 ? __dmul_rn(x, P.inv_dx) : __ddiv_rn(x, P.dx); }

// Without a solid level set the area fractions are in closed form — 0 on the domain walls, 1 elsewhere (macutility3.cpp:149-164) —, without a
// liquid level set (a smoke solver) rho = 1 (:191-193): exactly what k_face_fractions stores, so the kernels that consume the fractions
// skip those face-array reads (12 B per cell and array triple in each of k_label_rows, k_build_system and k_update_velocity).
__device__ __forceinline__ bool closed_form_area(const AsmParams &P) { return !P.have_solid; }
__device__ __forceinline__ bool closed_form_rho(const AsmParams &P) { return !P.fluid_levelset; }
template <class RealT>
__device__ __forceinline__ RealT wall_area(int pd, int n_dim) { return (pd == 0 || pd == n_dim) ? (RealT)0 : (RealT)1; }

template <class RealT>
struct FaceGrids { // face-shaped, no ghost planes: x (nx+1,ny,nzl)  y (nx,ny+1,nzl)  z (nx,ny,nzl+1)
	RealT *p[3];
};
template <class RealT>
struct ConstFaceGrids {
	const RealT *p[3];
};
struct FaceMasks {
	uint8_t *p[3];
};

__device__ __forceinline__ long long face_index(const Dims &d, int dim, int i, int j, int k) {
	const long long w = d.nx + (dim == 0), h = d.ny + (dim == 1);
	return i + w * (j + h * (long long)k);
}
__device__ __forceinline__ int clampi(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }

// include/shiokaze/utility/utility.h:162-170
__device__ __forceinline__ double fraction(double phi0, double phi1) {
	if (__dmul_rn(phi0, phi1) >= 0.0) {
		if (phi0 < 0.0 || phi1 < 0.0) return 1.0;
		return 0.0;
	}
	const double denom = fmax(fabs(__dsub_rn(phi1, phi0)), DBL_MIN);
	return __ddiv_rn(-fmin(phi0, phi1), denom);
}

// include/shiokaze/utility/utility.h:179-214 — polygon of {phi<0} on the unit square, shoelace area.
// Corner order (0,0),(1,0),(1,1),(0,1).
__device__ __forceinline__ double get_area(double v0, double v1, double v2, double v3) {
	const double qx[4] = {0.0, 1.0, 1.0, 0.0}, qy[4] = {0.0, 0.0, 1.0, 1.0};
	const double v[4] = {v0, v1, v2, v3};
	double px[8], py[8];
	int pnum = 0;
#pragma unroll
	for (int n = 0; n < 4; ++n) {
		const int m = (n + 1) & 3;
		if (v[n] < 0.0) {
			px[pnum] = qx[n];
			py[pnum] = qy[n];
			pnum++;
		}
		if (__dmul_rn(v[n], v[m]) < 0.0) {
			const double y0 = v[n], y1 = v[m];
			const double den = __dsub_rn(y0, y1);
			if (den != 0.0) {
				const double a = __ddiv_rn(y0, den);
				const double na = __dsub_rn(1.0, a);
				px[pnum] = __dadd_rn(__dmul_rn(na, qx[n]), __dmul_rn(a, qx[m]));
				py[pnum] = __dadd_rn(__dmul_rn(na, qy[n]), __dmul_rn(a, qy[m]));
				pnum++;
			}
		}
	}
	double sum = 0.0;
	for (int m = 0; m < pnum; ++m) {
		const int m1 = (m + 1 == pnum) ? 0 : m + 1;
		sum = __dadd_rn(sum, __dsub_rn(__dmul_rn(px[m], py[m1]), __dmul_rn(py[m], px[m1])));
	}
	return __dmul_rn(0.5, sum);
}

// K1 + K2 (+ first-order toggles): src/utility/macutility3.cpp:94-194, macpressuresolver3.cpp:70-80.
// One thread per (i,j,k) in [0,nx] x [0,ny] x [0,nzl]; it owns the three lower faces of that slot.
template <class RealT>
__global__ void __launch_bounds__(256) k_face_fractions(Dims d, AsmParams P, const RealT *__restrict__ solid,
                                                       const RealT *__restrict__ phi, FaceGrids<RealT> areas, FaceGrids<RealT> rhos) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i > d.nx || j > d.ny) return;
	const int kg = k + d.k0;
	const long long sw = d.nx + 1, sh = d.ny + 1;
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		if (i >= d.nx + (dim == 0) || j >= d.ny + (dim == 1) || k >= d.nzl + (dim == 2)) continue;
		const int pd = dim == 0 ? i : (dim == 1 ? j : kg);
		const int n_dim = dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzg);
		double area;
		if (!P.have_solid) {
			area = (pd == 0 || pd == n_dim) ? 0.0 : 1.0; // macutility3.cpp:149-164
		} else {
			if (pd == 0) area = 0.0; // :120 (a nodal solid never matches the far wall)
			else {
#define SOLID(a, b, c) (double)solid[(a) + sw * ((b) + sh * (long long)(c))]
				double q00, q10, q11, q01;
				if (dim == 0) { q00 = SOLID(i, j, k); q10 = SOLID(i, j + 1, k); q11 = SOLID(i, j + 1, k + 1); q01 = SOLID(i, j, k + 1); }
				else if (dim == 1) { q00 = SOLID(i, j, k); q10 = SOLID(i + 1, j, k); q11 = SOLID(i + 1, j, k + 1); q01 = SOLID(i, j, k + 1); }
				else { q00 = SOLID(i, j, k); q10 = SOLID(i + 1, j, k); q11 = SOLID(i + 1, j + 1, k); q01 = SOLID(i, j + 1, k); }
#undef SOLID
				// all four corners open / solid: the polygon is empty / the unit square, exactly (no arithmetic needed)
				if (q00 >= 0.0 && q10 >= 0.0 && q11 >= 0.0 && q01 >= 0.0) area = 1.0;
				else if (q00 < 0.0 && q10 < 0.0 && q11 < 0.0 && q01 < 0.0) area = 0.0;
				else area = __dsub_rn(1.0, get_area(q00, q10, q11, q01));
			}
			if (area != 0.0 && area < P.eps_solid) area = P.eps_solid; // :141
		}
		RealT area_r = (RealT)area;
		if (!P.second_order_solid && area_r != (RealT)0) area_r = (RealT)1;
		double rho;
		if (!P.fluid_levelset) rho = 1.0; // :191-193
		else {
			// :179-182 with shape3::clamp (shape.h:790-798); the z clamp is global, ghost planes carry the neighbour slab
			const int ia = clampi(i, d.nx), ja = clampi(j, d.ny), ka = clampi(kg, d.nzg) - d.k0;
			const int ib = clampi(i - (dim == 0), d.nx), jb = clampi(j - (dim == 1), d.ny), kb = clampi(kg - (dim == 2), d.nzg) - d.k0;
			const double a = (double)phi[ia + (long long)d.nx * (ja + (long long)d.ny * ka)];
			const double b = (double)phi[ib + (long long)d.nx * (jb + (long long)d.ny * kb)];
			rho = fraction(a, b);
			if (rho != 0.0 && rho < P.eps_fluid) rho = P.eps_fluid; // :183
		}
		RealT rho_r = (RealT)rho;
		if (!P.second_order_fluid && rho_r != (RealT)0) rho_r = (RealT)1;
		const long long f = face_index(d, dim, i, j, k);
		areas.p[dim][f] = area_r;
		rhos.p[dim][f] = rho_r;
	}
}

__device__ __forceinline__ float real_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double real_add(double a, double b) { return __dadd_rn(a, b); }

// K3a: curvature = 7-point Laplacian of phi with clamped neighbours / dx^2 (macpressuresolver3.cpp:92-102)
template <class RealT>
__global__ void __launch_bounds__(256) k_curvature(Dims d, AsmParams P, const RealT *__restrict__ phi, RealT *__restrict__ curv) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i >= d.nx || j >= d.ny) return;
	const int kg = k + d.k0;
	// the reference's expression (macpressuresolver3.cpp:93-98) adds the six neighbour values as Real — float on the shipping build —, left to right,
	// and only the "- 6.0*fluid" term promotes to double
#define PHI(a, b, c) phi[clampi(a, d.nx) + (long long)d.nx * (clampi(b, d.ny) + (long long)d.ny * (clampi(c, d.nzg) - d.k0))]
	RealT sr = PHI(i - 1, j, kg);
	sr = real_add(sr, PHI(i + 1, j, kg));
	sr = real_add(sr, PHI(i, j - 1, kg));
	sr = real_add(sr, PHI(i, j + 1, kg));
	sr = real_add(sr, PHI(i, j, kg - 1));
	sr = real_add(sr, PHI(i, j, kg + 1));
	const double s = __dsub_rn((double)sr, __dmul_rn(6.0, (double)PHI(i, j, kg)));
#undef PHI
	curv[i + (long long)d.nx * (j + (long long)d.ny * k)] = (RealT)__ddiv_rn(s, __dmul_rn(P.dx, P.dx));
}

// K3b: surface-tension increment on ACTIVE faces with 0 < rho < 1 (macpressuresolver3.cpp:104-113)
template <class RealT>
__global__ void __launch_bounds__(256) k_surface_tension(Dims d, AsmParams P, const RealT *__restrict__ phi, const RealT *__restrict__ curv,
                                                        ConstFaceGrids<RealT> rhos, FaceGrids<RealT> vel, FaceMasks active) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i > d.nx || j > d.ny) return;
	const int kg = k + d.k0;
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		if (i >= d.nx + (dim == 0) || j >= d.ny + (dim == 1) || k >= d.nzl + (dim == 2)) continue;
		const long long f = face_index(d, dim, i, j, k);
		if (!active.p[dim][f]) continue;
		const double rho = (double)rhos.p[dim][f];
		if (rho != 0.0 && rho < 1.0) {
			const long long ca = clampi(i, d.nx) + (long long)d.nx * (clampi(j, d.ny) + (long long)d.ny * (clampi(kg, d.nzg) - d.k0));
			const long long cb = clampi(i - (dim == 0), d.nx) + (long long)d.nx * (clampi(j - (dim == 1), d.ny) + (long long)d.ny * (clampi(kg - (dim == 2), d.nzg) - d.k0));
			const double sgn = (double)phi[ca] < 0.0 ? -1.0 : 1.0;
			const double theta = sgn < 0 ? __dsub_rn(1.0, rho) : rho;
			const double face_c = __dadd_rn(__dmul_rn(theta, (double)curv[ca]), __dmul_rn(__dsub_rn(1.0, theta), (double)curv[cb]));
			// -sgn * dt / (dx*rho) * kappa * face_c, left to right
			double inc = __ddiv_rn(__dmul_rn(-sgn, P.dt), __dmul_rn(P.dx, rho));
			inc = __dmul_rn(__dmul_rn(inc, P.surface_tension), face_c);
			vel.p[dim][f] = vel.p[dim][f] + (RealT)inc; // array3::increment in Real arithmetic (array3.h:631-639)
		}
	}
}

// K5 + K6: matrix-free coefficients and right-hand side (macpressuresolver3.cpp:163-217).
//   w{x,y,z}[c] = dt*A/(dx^2*theta) of the LOWER face of c when both cells are rows, else 0
//   dd[c]       = the part of the reference's diagonal that comes from AIR neighbours (ghost-fluid
//                 Dirichlet faces); the full diagonal is dd + the six couplings, which the solver
//                 kernels re-form on the fly so that the operator annihilates constants exactly on
//                 pure-Neumann rows whatever the coefficient precision ("difference form")
//   rhs[c]      = sum -sgn*A*u/dx (+ volume-correction constant); 0 outside the row set
// CoefT/VecT copies feed the CG operator; the float copies feed multigrid level 0 (NULL when CoefT is
// float and level 0 shares the operator arrays). tile_flags (zeroed by the caller) marks the level-0 tiles
// that hold at least one unknown.
template <class RealT, class CoefT, class VecT>
__global__ void __launch_bounds__(256, 4) k_build_system(Dims d, AsmParams P, const RealT *__restrict__ phi, uint8_t *__restrict__ in_rows,
                                                     ConstFaceGrids<RealT> areas, ConstFaceGrids<RealT> rhos, ConstFaceGrids<RealT> vel,
                                                     CoefT *__restrict__ wx, CoefT *__restrict__ wy, CoefT *__restrict__ wz, CoefT *__restrict__ dd,
                                                     float *__restrict__ mwx, float *__restrict__ mwy, float *__restrict__ mwz, float *__restrict__ mdd,
                                                     VecT *__restrict__ rhs, Tiles T, unsigned char *__restrict__ tile_flags, RedBuf rb, CGState *st) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	double red[3] = {0.0, 0.0, 0.0}; // |b|_inf, row count, dirichlet flag
	// a block walks planes blockIdx.z, blockIdx.z + gridDim.z, ...: a few thousand blocks in all, so the grid-wide
	// reduction at the end stays cheap (one counter atomic per block)
	if (i < d.nx && j < d.ny) for (int k = blockIdx.z; k < d.nzl; k += gridDim.z) {
		const int kg = k + d.k0;
		const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
		double dirichlet = 0.0, b = 0.0, lower[3] = {0.0, 0.0, 0.0};
		// K4, the row labelling (macpressuresolver3.cpp:121-156), happens here: a cell is a row iff phi(c) < 0 and some in-grid neighbour q has
		// phi(q) < 0 across a face with area != 0 and rho != 0 — the very operands the assembly of that row loads anyway
		bool is_row = false;
		if (phi[c] < (RealT)0) {
			const int qo[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
			const double dx2 = __dmul_rn(P.dx, P.dx);
			const double w_unit = __ddiv_rn(P.dt, dx2); // value of a fully open, fully wet face: (dt*1)/(dx2*1)
			// all loads first (independent, in flight together), then the reference's arithmetic in its order
			RealT ar[6], rh[6], uf[6], pq[6];
			bool in_grid[6];
#pragma unroll
			for (int nq = 0; nq < 6; ++nq) {
				const int dim = nq >> 1;
				const int qi = i + qo[nq][0], qj = j + qo[nq][1], qkg = kg + qo[nq][2];
				in_grid[nq] = !(qi < 0 || qj < 0 || qkg < 0 || qi >= d.nx || qj >= d.ny || qkg >= d.nzg);
				const int up = (nq & 1) ? 0 : 1;
				const long long f = face_index(d, dim, i + (dim == 0) * up, j + (dim == 1) * up, k + (dim == 2) * up);
				const long long q = c + qo[nq][0] + (long long)d.nx * qo[nq][1] + d.plane * qo[nq][2];
				ar[nq] = closed_form_area(P) ? (in_grid[nq] ? (RealT)1 : (RealT)0) : areas.p[dim][f]; // (0 exactly on the faces whose neighbour is outside the grid)
				rh[nq] = closed_form_rho(P) ? (RealT)1 : rhos.p[dim][f];
				uf[nq] = vel.p[dim][f];
				pq[nq] = in_grid[nq] ? phi[q] : (RealT)1;
			}
#pragma unroll
			for (int nq = 0; nq < 6; ++nq)
				if (in_grid[nq] && pq[nq] < (RealT)0 && ar[nq] != (RealT)0 && rh[nq] != (RealT)0) is_row = true;
#pragma unroll
			for (int nq = 0; nq < 6; ++nq) {
				const int dim = nq >> 1;
				if (!is_row || !in_grid[nq]) continue;
				const int up = (nq & 1) ? 0 : 1;
				const double area = (double)ar[nq];
				if (area != 0.0) {
					const double rho = (double)rh[nq];
					if (rho != 0.0) {
						const double value = (area == 1.0 && rho == 1.0) ? w_unit : __ddiv_rn(__dmul_rn(P.dt, area), __dmul_rn(dx2, rho));
						if (pq[nq] < (RealT)0) {
							if (!up) lower[dim] = value;
						} else {
							dirichlet = __dadd_rn(dirichlet, value);
							red[2] = 1.0;
						}
					}
					const double sgn = up ? -1.0 : 1.0; // -sgn[nq]
					b = __dadd_rn(b, div_dx(P, __dmul_rn(__dmul_rn(sgn, area), (double)uf[nq])));
				}
			}
			if (is_row) {
				if (P.apply_rhs_correct) b = __dadd_rn(b, P.rhs_correct);
				red[0] = fmax(red[0], fabs(b));
				red[1] += 1.0;
				tile_flags[tile_of(T, i, j, k)] = 1; // the solve only visits tiles that hold an unknown
			}
		}
		in_rows[c] = is_row ? 1 : 0;
		wx[c] = (CoefT)lower[0]; wy[c] = (CoefT)lower[1]; wz[c] = (CoefT)lower[2]; dd[c] = (CoefT)dirichlet;
		if (mwx != nullptr) { mwx[c] = (float)lower[0]; mwy[c] = (float)lower[1]; mwz[c] = (float)lower[2]; mdd[c] = (float)dirichlet; }
		rhs[c] = (VecT)b;
	}
	grid_reduce<3, 0x5u>(red, rb, [&](double (&t)[3]) {
		st->bnorm = t[0];
		st->n_rows = (unsigned long long)(t[1] + 0.5);
		st->has_dirichlet = t[2] > 0.0 ? 1 : 0;
	});
}

// K12: pressure scatter to the Real grid (macpressuresolver3.cpp:245-248); singular systems lose their mean.
// (the caller's pressure / activity grids, when given, are written by the same pass)
// warm start (macpressuresolver3.cpp:239-242): the solve ran on b - A p_prev, so the pressure is x + p_prev, which also becomes the next p_prev
// (zero off the row set: a cell that joins the row set later starts from nothing, like a new row of the reference's resized vector).
template <class RealT, class VecT>
__global__ void __launch_bounds__(256) k_store_pressure(Dims d, const VecT *__restrict__ x, const uint8_t *__restrict__ in_rows,
                                                       const CGState *__restrict__ st, RealT *__restrict__ pressure, RealT *__restrict__ pressure_out,
                                                       uint8_t *__restrict__ active_out, VecT *__restrict__ p_prev) {
	__shared__ double shift_sh; // one fp64 division per block, not per cell
	if (threadIdx.x == 0) shift_sh = (!st->has_dirichlet && st->n_rows) ? st->sum_x / (double)st->n_rows : 0.0;
	__syncthreads();
	const double shift = shift_sh;
	const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (c >= d.ncell) return;
	const uint8_t row = in_rows[c];
	double pv = row ? (double)x[c] - shift : 0.0;
	if (p_prev) {
		if (row) pv += (double)p_prev[c];
		p_prev[c] = (VecT)pv;
	}
	const RealT p = (RealT)pv;
	pressure[c] = p;
	if (pressure_out) pressure_out[c] = p;
	if (active_out) active_out[c] = row;
}

// sum of x over the row set by itself: only for MGPostSweeps = 0, where the prolongation leaves values on cells without an equation
// (every other configuration gets the sum from k_axpy2_norm)
template <class VecT>
__global__ void __launch_bounds__(256) k_sum_rows(Dims d, const VecT *__restrict__ x, const uint8_t *__restrict__ in_rows, RedBuf rb, CGState *st) {
	double red[1] = {0.0};
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < d.ncell; c += (long long)gridDim.x * blockDim.x)
		if (in_rows[c]) red[0] += (double)x[c];
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) { st->sum_x = t[0]; });
}

// K13: velocity update on ACTIVE faces (macpressuresolver3.cpp:252-268). pressure has ghost planes.
template <class RealT>
__global__ void __launch_bounds__(256) k_update_velocity(Dims d, AsmParams P, const RealT *__restrict__ phi, const RealT *__restrict__ pressure,
                                                        ConstFaceGrids<RealT> areas, ConstFaceGrids<RealT> rhos, FaceGrids<RealT> vel, FaceMasks active) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i > d.nx || j > d.ny) return;
	const int kg = k + d.k0;
	const long long c = i + (long long)d.nx * (j + (long long)d.ny * k); // may be a ghost / out-of-row slot for the far faces
	// first the three activity bytes, then every other operand of the active faces at once (independent loads)
	bool act[3];
	long long fi[3];
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		const bool exists = !(i >= d.nx + (dim == 0) || j >= d.ny + (dim == 1) || k >= d.nzl + (dim == 2));
		fi[dim] = face_index(d, dim, i, j, k);
		act[dim] = exists && active.p[dim][fi[dim]] != 0;
	}
	RealT ar[3], rh[3], uf[3], pc = (RealT)0, pm[3], phc = (RealT)0, phm[3];
	if (act[0] || act[1] || act[2]) { pc = pressure[c]; phc = phi[c]; }
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		ar[dim] = rh[dim] = uf[dim] = pm[dim] = phm[dim] = (RealT)0;
		if (!act[dim]) continue;
		const long long cm = c - (dim == 0 ? 1 : (dim == 1 ? d.nx : d.plane));
		ar[dim] = closed_form_area(P) ? wall_area<RealT>(dim == 0 ? i : (dim == 1 ? j : kg), dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzg)) : areas.p[dim][fi[dim]];
		rh[dim] = closed_form_rho(P) ? (RealT)1 : rhos.p[dim][fi[dim]];
		uf[dim] = vel.p[dim][fi[dim]];
		pm[dim] = pressure[cm]; phm[dim] = phi[cm];
	}
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		if (!act[dim]) continue;
		const long long f = fi[dim];
		const int pd = dim == 0 ? i : (dim == 1 ? j : kg);
		const int n_dim = dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzg);
		const RealT rho = rh[dim];
		if (ar[dim] != (RealT)0 && rho != (RealT)0) {
			if (pd == 0 || pd == n_dim) vel.p[dim][f] = (RealT)0;
			else {
				const RealT diff = pc - pm[dim]; // Real arithmetic, as the reference's expression
				const double num = __dmul_rn(P.dt, (double)diff);
				const RealT delta = (RealT)(rho == (RealT)1 ? div_dx(P, num) : __ddiv_rn(num, __dmul_rn((double)rho, P.dx)));
				vel.p[dim][f] = uf[dim] - delta; // array3::subtract (array3.h:663-671)
			}
		} else {
			if (pd == 0 && phc < (RealT)0) vel.p[dim][f] = (RealT)0;
			else if (pd == n_dim && phm[dim] < (RealT)0) vel.p[dim][f] = (RealT)0;
			else { active.p[dim][f] = 0; vel.p[dim][f] = (RealT)0; } // set_off(): reads back as the background 0
		}
	}
}

} // namespace shkz
