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
	int off_mark;  // what k_update_velocity stores into the mask of a face it switches off: 0, or kernels_xfer.cuh's marker when the host copies are sparse
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
__device__ __noinline__ double get_area(double v0, double v1, double v2, double v3) {
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

// K1 (+ first-order toggle): open-area fraction of face (dim; i,j,k) — k local, kg global — as the reference stores it (Real), from the nodal solid level set:
// src/utility/macutility3.cpp:94-165, macpressuresolver3.cpp:75-80. Every consumer recomputes it from the four nodes instead of reading a face array: the values
// are the same bits (same operations, same rounding to Real), the face arrays (6 x 4 B per cell written, then read twice) are gone.
template <class RealT>
__device__ __forceinline__ RealT face_area(const Dims &d, const AsmParams &P, const RealT *__restrict__ solid, int dim, int i, int j, int k) {
	const int kg = k + d.k0;
	const int pd = dim == 0 ? i : (dim == 1 ? j : kg);
	const int n_dim = dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzg);
	double area;
	if (!P.have_solid) {
		area = (pd == 0 || pd == n_dim) ? 0.0 : 1.0; // macutility3.cpp:149-164
	} else {
		if (pd == 0) area = 0.0; // :120 (a nodal solid never matches the far wall)
		else {
			const long long sw = d.nx + 1, sh = d.ny + 1;
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
	return area_r;
}

// K2 (+ first-order toggle): liquid fraction of a face from the level-set values of its two cells (macutility3.cpp:166-194, macpressuresolver3.cpp:70-74);
// phi_a, phi_b = fluid at shape3::clamp of the face index and of the index one step down in `dim` (:179-182) — a wall face sees the same cell twice.
template <class RealT>
__device__ __forceinline__ RealT face_rho_of(const AsmParams &P, RealT phi_a, RealT phi_b) {
	if (!P.fluid_levelset) return (RealT)1; // :191-193
	double rho = fraction((double)phi_a, (double)phi_b);
	if (rho != 0.0 && rho < P.eps_fluid) rho = P.eps_fluid; // :183
	RealT rho_r = (RealT)rho;
	if (!P.second_order_fluid && rho_r != (RealT)0) rho_r = (RealT)1;
	return rho_r;
}
template <class RealT>
__device__ __forceinline__ RealT face_rho(const Dims &d, const AsmParams &P, const RealT *__restrict__ phi, int dim, int i, int j, int k) {
	if (!P.fluid_levelset) return (RealT)1;
	// the z clamp is global, ghost planes carry the neighbour slab
	const int kg = k + d.k0;
	const int ia = clampi(i, d.nx), ja = clampi(j, d.ny), ka = clampi(kg, d.nzg) - d.k0;
	const int ib = clampi(i - (dim == 0), d.nx), jb = clampi(j - (dim == 1), d.ny), kb = clampi(kg - (dim == 2), d.nzg) - d.k0;
	return face_rho_of<RealT>(P, phi[ia + (long long)d.nx * (ja + (long long)d.ny * ka)], phi[ib + (long long)d.nx * (jb + (long long)d.ny * kb)]);
}

// The six face arrays themselves, materialised only when somebody asks for them (shkz_b200_debug_fetch: the bit-exactness tests).
// One thread per (i,j,k) in [0,nx] x [0,ny] x [0,nzl]; it owns the three lower faces of that slot.
template <class RealT>
__global__ void __launch_bounds__(256) k_face_fractions(Dims d, AsmParams P, const RealT *__restrict__ solid,
                                                       const RealT *__restrict__ phi, FaceGrids<RealT> areas, FaceGrids<RealT> rhos) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i > d.nx || j > d.ny) return;
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		if (i >= d.nx + (dim == 0) || j >= d.ny + (dim == 1) || k >= d.nzl + (dim == 2)) continue;
		const long long f = face_index(d, dim, i, j, k);
		areas.p[dim][f] = face_area<RealT>(d, P, solid, dim, i, j, k);
		rhos.p[dim][f] = face_rho<RealT>(d, P, phi, dim, i, j, k);
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
                                                        FaceGrids<RealT> vel, FaceMasks active) {
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
		const double rho = (double)face_rho<RealT>(d, P, phi, dim, i, j, k);
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
// K1 from the four node values of a face already in registers: the operations of face_area, without the loads. pd = face coordinate along its direction.
template <class RealT>
__device__ __forceinline__ RealT area_from_nodes(const AsmParams &P, int pd, RealT s00, RealT s10, RealT s11, RealT s01) {
	double area;
	if (pd == 0) area = 0.0; // macutility3.cpp:120 (a nodal solid never matches the far wall)
	else {
		const double q00 = (double)s00, q10 = (double)s10, q11 = (double)s11, q01 = (double)s01;
		if (q00 >= 0.0 && q10 >= 0.0 && q11 >= 0.0 && q01 >= 0.0) area = 1.0;
		else if (q00 < 0.0 && q10 < 0.0 && q11 < 0.0 && q01 < 0.0) area = 0.0;
		else area = __dsub_rn(1.0, get_area(q00, q10, q11, q01));
	}
	if (area != 0.0 && area < P.eps_solid) area = P.eps_solid; // :141
	RealT area_r = (RealT)area;
	if (!P.second_order_solid && area_r != (RealT)0) area_r = (RealT)1;
	return area_r;
}

// One cell of K4-K6 (labelling, couplings, Dirichlet diagonal, right-hand side), the reference's arithmetic in its order.
struct CellSystem {
	double lower[3], dirichlet, b;
	bool is_row, has_dirichlet;
};
// HAVE_SOLID / LEVELSET: compile-time copies of P.have_solid / P.fluid_levelset (without them the fractions are the closed forms 1 and wall-0).
template <class RealT, bool HAVE_SOLID, bool LEVELSET>
__device__ __forceinline__ void assemble_cell(const Dims &d, const AsmParams &P, const RealT *__restrict__ phi, const RealT *__restrict__ solid,
                                              const ConstFaceGrids<RealT> &vel, int i, int j, int k, long long c, RealT pc, CellSystem &out) {
	out.lower[0] = out.lower[1] = out.lower[2] = 0.0;
	out.dirichlet = 0.0; out.b = 0.0;
	out.is_row = false; out.has_dirichlet = false;
	// K4, the row labelling (macpressuresolver3.cpp:121-156), happens here: a cell is a row iff phi(c) < 0 and some in-grid neighbour q has
	// phi(q) < 0 across a face with area != 0 and rho != 0 — the very operands the assembly of that row loads anyway
	if (!(pc < (RealT)0)) return;
	const int kg = k + d.k0;
	const long long nx = d.nx;
	const double dx2 = __dmul_rn(P.dx, P.dx);
	const double w_unit = __ddiv_rn(P.dt, dx2); // value of a fully open, fully wet face: (dt*1)/(dx2*1)
	// neighbour order of the reference (:163-165): +x -x +y -y +z -z. All loads first (independent, in flight together), then the reference's arithmetic in its
	// order. The fractions of the six faces come straight from the level sets (K1, K2: the eight solid nodes of the cell, the seven fluid values the labelling
	// needs anyway), never from memory.
	const bool in_grid[6] = {i + 1 < d.nx, i > 0, j + 1 < d.ny, j > 0, kg + 1 < d.nzg, kg > 0};
	const long long qoff[6] = {1, -1, nx, -nx, d.plane, -d.plane};
	const long long fx = i + (nx + 1) * (j + (long long)d.ny * k), fy = i + nx * (j + (long long)(d.ny + 1) * k);
	const RealT uf[6] = {vel.p[0][fx + 1], vel.p[0][fx], vel.p[1][fy + nx], vel.p[1][fy], vel.p[2][c + d.plane], vel.p[2][c]};
	RealT pq[6], ar[6], rh[6];
#pragma unroll
	for (int nq = 0; nq < 6; ++nq) pq[nq] = in_grid[nq] ? phi[c + qoff[nq]] : (RealT)1;
	if (HAVE_SOLID) {
		const long long sw = d.nx + 1, ss = sw * (d.ny + 1), n = i + sw * (j + (long long)(d.ny + 1) * k);
		const RealT s000 = solid[n], s100 = solid[n + 1], s010 = solid[n + sw], s110 = solid[n + sw + 1];
		const RealT s001 = solid[n + ss], s101 = solid[n + ss + 1], s011 = solid[n + ss + sw], s111 = solid[n + ss + sw + 1];
		// corner order of macutility3.cpp:122-139 — x face: (j,k) (j+1,k) (j+1,k+1) (j,k+1); y face: (i,k) (i+1,k) (i+1,k+1) (i,k+1); z face: (i,j) (i+1,j) (i+1,j+1) (i,j+1)
		ar[0] = !in_grid[0] ? (RealT)0 : area_from_nodes<RealT>(P, i + 1, s100, s110, s111, s101);
		ar[1] = !in_grid[1] ? (RealT)0 : area_from_nodes<RealT>(P, i, s000, s010, s011, s001);
		ar[2] = !in_grid[2] ? (RealT)0 : area_from_nodes<RealT>(P, j + 1, s010, s110, s111, s011);
		ar[3] = !in_grid[3] ? (RealT)0 : area_from_nodes<RealT>(P, j, s000, s100, s101, s001);
		ar[4] = !in_grid[4] ? (RealT)0 : area_from_nodes<RealT>(P, kg + 1, s001, s101, s111, s011);
		ar[5] = !in_grid[5] ? (RealT)0 : area_from_nodes<RealT>(P, kg, s000, s100, s110, s010);
	} else {
		// (a face whose neighbour is outside the grid contributes nothing, macpressuresolver3.cpp:166: its fractions are never looked at; inside, :149-164 gives 1)
#pragma unroll
		for (int nq = 0; nq < 6; ++nq) ar[nq] = in_grid[nq] ? (RealT)1 : (RealT)0;
	}
#pragma unroll
	for (int nq = 0; nq < 6; ++nq) {
		// in-grid neighbour: the face's two cells are c and q (the clamp of :179-182 does nothing), upper cell first
		rh[nq] = (!LEVELSET || !in_grid[nq]) ? (RealT)1 : ((nq & 1) ? face_rho_of<RealT>(P, pc, pq[nq]) : face_rho_of<RealT>(P, pq[nq], pc));
	}
	bool is_row = false;
#pragma unroll
	for (int nq = 0; nq < 6; ++nq)
		if (in_grid[nq] && pq[nq] < (RealT)0 && ar[nq] != (RealT)0 && rh[nq] != (RealT)0) is_row = true;
	out.is_row = is_row;
	if (!is_row) return;
	double dirichlet = 0.0, b = 0.0;
#pragma unroll
	for (int nq = 0; nq < 6; ++nq) {
		const int dim = nq >> 1;
		if (!in_grid[nq]) continue;
		const int up = (nq & 1) ? 0 : 1;
		const double area = (double)ar[nq];
		if (area != 0.0) {
			const double rho = (double)rh[nq];
			if (rho != 0.0) {
				const double value = (area == 1.0 && rho == 1.0) ? w_unit : __ddiv_rn(__dmul_rn(P.dt, area), __dmul_rn(dx2, rho));
				if (pq[nq] < (RealT)0) {
					if (!up) out.lower[dim] = value;
				} else {
					dirichlet = __dadd_rn(dirichlet, value);
					out.has_dirichlet = true;
				}
			}
			const double sgn = up ? -1.0 : 1.0; // -sgn[nq]
			b = __dadd_rn(b, div_dx(P, __dmul_rn(__dmul_rn(sgn, area), (double)uf[nq])));
		}
	}
	if (P.apply_rhs_correct) b = __dadd_rn(b, P.rhs_correct);
	out.dirichlet = dirichlet;
	out.b = b;
}

// The assembly kernel: persistent over (tile footprint, plane) units, plane-major so that the CTAs in flight sit side by side in one plane.
// A unit is the TX x TY footprint of a level-0 tile in one plane; block (TX, BS_ROWS): a thread owns TY / BS_ROWS cells of a column.
//   * a unit without a wet cell costs one read of phi — and NOT EVEN ITS STORES when its tile held no unknown in the previous projection
//     (`dirty` = last call's tile flags): what it would write is all zeros, and that is what those arrays still hold. The arrays start
//     zeroed, so the argument holds from the first call on. A dam-break at 512^3 writes 23 M of its 134 M cells.
//   * tile_flags (zeroed by the caller) marks the level-0 tiles that hold at least one unknown.
constexpr int BS_ROWS = 4;
template <class RealT, class CoefT, class VecT, bool HAVE_SOLID, bool LEVELSET>
__global__ void __launch_bounds__(TX *BS_ROWS, 3) k_build_system(Dims d, AsmParams P, const RealT *__restrict__ phi, const RealT *__restrict__ solid, uint8_t *__restrict__ in_rows,
                                                     ConstFaceGrids<RealT> vel,
                                                     CoefT *__restrict__ wx, CoefT *__restrict__ wy, CoefT *__restrict__ wz, CoefT *__restrict__ dd,
                                                     float *__restrict__ mwx, float *__restrict__ mwy, float *__restrict__ mwz, float *__restrict__ mdd,
                                                     VecT *__restrict__ rhs, Tiles T, unsigned char *__restrict__ tile_flags, const unsigned char *__restrict__ dirty,
                                                     RedBuf rb, CGState *st) {
	double red[3] = {0.0, 0.0, 0.0}; // |b|_inf, row count, dirichlet flag
	const int per_plane = T.ntx * T.nty;
	const long long units = (long long)per_plane * d.nzl;
	constexpr int CELLS = TY / BS_ROWS;
	for (long long u = blockIdx.x; u < units; u += gridDim.x) {
		const int k = (int)(u / per_plane), r = (int)(u - (long long)k * per_plane);
		const int i0 = (r % T.ntx) * TX, j0 = (r / T.ntx) * TY;
		const int tile = (r % T.ntx) + T.ntx * ((r / T.ntx) + T.nty * (k / T.slice)); // flag slot (slice) of the unit
		const int i = i0 + threadIdx.x;
		RealT pc[CELLS];
		bool wet = false;
#pragma unroll
		for (int m = 0; m < CELLS; ++m) {
			const int j = j0 + threadIdx.y + BS_ROWS * m;
			pc[m] = (i < d.nx && j < d.ny) ? phi[i + (long long)d.nx * (j + (long long)d.ny * k)] : (RealT)1;
			wet = wet || pc[m] < (RealT)0;
		}
		const bool any_wet = __syncthreads_or(wet);
		if (!any_wet && !dirty[tile]) continue; // nothing but zeros to write over zeros
		bool row_here = false;
#pragma unroll 1
		for (int m = 0; m < CELLS; ++m) {
			const int j = j0 + threadIdx.y + BS_ROWS * m;
			if (i >= d.nx || j >= d.ny) continue;
			const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
			CellSystem cs;
			assemble_cell<RealT, HAVE_SOLID, LEVELSET>(d, P, phi, solid, vel, i, j, k, c, pc[m], cs);
			if (cs.is_row) {
				red[0] = fmax(red[0], fabs(cs.b));
				red[1] += 1.0;
				if (cs.has_dirichlet) red[2] = 1.0;
				row_here = true;
			}
			in_rows[c] = cs.is_row ? 1 : 0;
			wx[c] = (CoefT)cs.lower[0]; wy[c] = (CoefT)cs.lower[1]; wz[c] = (CoefT)cs.lower[2]; dd[c] = (CoefT)cs.dirichlet;
			if (mwx != nullptr) { mwx[c] = (float)cs.lower[0]; mwy[c] = (float)cs.lower[1]; mwz[c] = (float)cs.lower[2]; mdd[c] = (float)cs.dirichlet; }
			rhs[c] = (VecT)cs.b;
		}
		if (row_here) tile_flags[tile] = 1; // the solve only visits tiles that hold an unknown
	}
	grid_reduce<3, 0x5u>(red, rb, [&](double (&t)[3]) {
		st->bnorm = t[0];
		st->n_rows = (unsigned long long)(t[1] + 0.5);
		st->has_dirichlet = t[2] > 0.0 ? 1 : 0;
	});
}

// K12: pressure scatter to the Real grid (macpressuresolver3.cpp:245-248); singular systems lose their mean.
// Persistent over the UNION list of level-0 tiles (tiles that hold an unknown now or held one in the previous projection): everywhere else the
// internal pressure grid already is zero, and the caller's pressure / activity grids are cleared by a memset before this kernel.
// warm start (macpressuresolver3.cpp:239-242): the solve ran on b - A p_prev, so the pressure is x + p_prev, which also becomes the next p_prev
// (zero off the row set: a cell that joins the row set later starts from nothing, like a new row of the reference's resized vector).
template <class RealT, class VecT>
__global__ void __launch_bounds__(TX * 8) k_store_pressure(Dims d, Tiles U, const VecT *__restrict__ x, const uint8_t *__restrict__ in_rows,
                                                          const CGState *__restrict__ st, RealT *__restrict__ pressure, RealT *__restrict__ pressure_out,
                                                          uint8_t *__restrict__ active_out, VecT *__restrict__ p_prev) {
	const double shift = (!st->has_dirichlet && st->n_rows) ? st->sum_x / (double)st->n_rows : 0.0;
	resolve_tiles(U);
	const int ntiles = *U.count;
	int i0, j0, kb, ke;
	for (TileWalk w(U, ntiles, true); w.next(U, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += 8) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
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
	}
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

// K13: velocity update on ACTIVE faces (macpressuresolver3.cpp:252-268). pressure and phi have ghost planes.
// Finish one face from operands already in registers. pha / phb = fluid at the clamped cells a (upper) and b (lower) of macutility3.cpp:179-182;
// s00.. = the four solid nodes of the face.
template <class RealT, bool HAVE_SOLID, bool LEVELSET>
__device__ __forceinline__ void finish_face(const AsmParams &P, int pd, int n_dim, RealT u, RealT pc, RealT pm, RealT pha, RealT phb, RealT s00, RealT s10, RealT s11, RealT s01,
                                            RealT *__restrict__ vel, uint8_t *__restrict__ active, long long f) {
	RealT ar, rho;
	if (HAVE_SOLID) {
		double area;
		if (pd == 0) area = 0.0; // macutility3.cpp:120
		else {
			const double q00 = (double)s00, q10 = (double)s10, q11 = (double)s11, q01 = (double)s01;
			if (q00 >= 0.0 && q10 >= 0.0 && q11 >= 0.0 && q01 >= 0.0) area = 1.0;
			else if (q00 < 0.0 && q10 < 0.0 && q11 < 0.0 && q01 < 0.0) area = 0.0;
			else area = __dsub_rn(1.0, get_area(q00, q10, q11, q01));
		}
		if (area != 0.0 && area < P.eps_solid) area = P.eps_solid; // :141
		ar = (RealT)area;
		if (!P.second_order_solid && ar != (RealT)0) ar = (RealT)1;
	} else ar = (pd == 0 || pd == n_dim) ? (RealT)0 : (RealT)1; // :149-164
	rho = LEVELSET ? face_rho_of<RealT>(P, pha, phb) : (RealT)1;
	if (ar != (RealT)0 && rho != (RealT)0) {
		if (pd == 0 || pd == n_dim) vel[f] = (RealT)0;
		else {
			const RealT diff = pc - pm; // Real arithmetic, as the reference's expression
			const double num = __dmul_rn(P.dt, (double)diff);
			const RealT delta = (RealT)(rho == (RealT)1 ? div_dx(P, num) : __ddiv_rn(num, __dmul_rn((double)rho, P.dx)));
			vel[f] = u - delta; // array3::subtract (array3.h:663-671)
		}
	} else {
		// (:263-264: fluid at the face's own cell for the lower wall = the clamped cell a; at the cell below for the upper wall = the clamped cell b)
		if (pd == 0 && pha < (RealT)0) vel[f] = (RealT)0;
		else if (pd == n_dim && phb < (RealT)0) vel[f] = (RealT)0;
		else { active[f] = (uint8_t)P.off_mark; vel[f] = (RealT)0; } // set_off(): reads back as the background 0
	}
}

// One thread owns the slots (i, j0 + r, k), r = 0..UV_ROWS-1, and for each of them the three lower faces, one direction after the other: first the
// UV_ROWS activity bytes of the direction, and only where one is set every operand of those faces at once (independent loads in flight together), then the
// arithmetic. The clamped cells of :179-182 need no clamp arithmetic: away from the two walls of the face's own direction they are the face's two cells,
// on the lower wall both are the face's cell, on the upper wall both are the cell below. (History: one slot per thread was bound by load latency at a few
// percent of the bandwidth; walking the masks sixteen bytes at a time with shuffles and index divisions needed 230 instructions per face and was
// issue-bound. This form needs a fifth of that.)
constexpr int UV_ROWS = 4;
template <class RealT, bool HAVE_SOLID, bool LEVELSET>
__global__ void __launch_bounds__(256, 3) k_update_velocity(Dims d, AsmParams P, const RealT *__restrict__ phi, const RealT *__restrict__ solid, const RealT *__restrict__ pressure,
                                                        FaceGrids<RealT> vel, FaceMasks active) {
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int j0 = (blockIdx.y * 8 + threadIdx.y) * UV_ROWS;
	const int k = blockIdx.z;
	if (i > d.nx || j0 > d.ny) return;
	const int kg = k + d.k0;
	const long long nx = d.nx, c0 = i + nx * (j0 + (long long)d.ny * k);
	const long long sw = d.nx + 1, sh = d.ny + 1, n0 = i + sw * (j0 + sh * (long long)k); // node (i, j0, k)
	// every activity byte of the thread first: a dry slot costs ONE memory round trip, not one per direction
	uint8_t on_all[3][UV_ROWS];
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		const bool exists = !(i >= d.nx + (dim == 0) || k >= d.nzl + (dim == 2));
		const long long w = d.nx + (dim == 0), h = d.ny + (dim == 1);
		const long long f0 = i + w * (j0 + h * (long long)k);
#pragma unroll
		for (int r = 0; r < UV_ROWS; ++r) on_all[dim][r] = (exists && j0 + r < d.ny + (dim == 1)) ? active.p[dim][f0 + w * r] : (uint8_t)0;
	}
#pragma unroll
	for (int dim = 0; dim < 3; ++dim) {
		const long long w = d.nx + (dim == 0), h = d.ny + (dim == 1);
		const long long f0 = i + w * (j0 + h * (long long)k);
		const long long back = dim == 0 ? 1 : (dim == 1 ? nx : d.plane);
		uint8_t *act = active.p[dim];
		RealT *v = vel.p[dim];
		bool on[UV_ROWS];
		bool any = false;
#pragma unroll
		for (int r = 0; r < UV_ROWS; ++r) {
			on[r] = on_all[dim][r] != 0;
			any = any || on[r];
		}
		if (!any) continue;
		RealT u[UV_ROWS], pc[UV_ROWS], pm[UV_ROWS], pha[UV_ROWS], phb[UV_ROWS], s00[UV_ROWS], s10[UV_ROWS], s11[UV_ROWS], s01[UV_ROWS];
		const int n_dim = dim == 0 ? d.nx : (dim == 1 ? d.ny : d.nzg);
#pragma unroll
		for (int r = 0; r < UV_ROWS; ++r) {
			u[r] = pc[r] = pm[r] = pha[r] = phb[r] = (RealT)0;
			s00[r] = s10[r] = s11[r] = s01[r] = (RealT)1;
			if (!on[r]) continue;
			const int pd = dim == 0 ? i : (dim == 1 ? j0 + r : kg);
			const long long c = c0 + nx * r, cm = c - back;
			u[r] = v[f0 + w * r];
			if (pd != 0 && pd != n_dim) { pc[r] = pressure[c]; pm[r] = pressure[cm]; }
			if (LEVELSET || pd == 0 || pd == n_dim) { // (without a level set only the wall rule of :263-264 looks at the fluid)
				const RealT hi = pd == n_dim ? phi[cm] : phi[c], lo = pd == 0 ? hi : (pd == n_dim ? hi : phi[cm]);
				pha[r] = hi; phb[r] = lo;
			}
			if (HAVE_SOLID && pd != 0) {
				const long long n = n0 + sw * r;
				s00[r] = solid[n];
				if (dim == 0) { s10[r] = solid[n + sw]; s11[r] = solid[n + sw + sw * sh]; s01[r] = solid[n + sw * sh]; }
				else if (dim == 1) { s10[r] = solid[n + 1]; s11[r] = solid[n + 1 + sw * sh]; s01[r] = solid[n + sw * sh]; }
				else { s10[r] = solid[n + 1]; s11[r] = solid[n + 1 + sw]; s01[r] = solid[n + sw]; }
			}
		}
#pragma unroll
		for (int r = 0; r < UV_ROWS; ++r) {
			if (!on[r]) continue;
			const int pd = dim == 0 ? i : (dim == 1 ? j0 + r : kg);
			finish_face<RealT, HAVE_SOLID, LEVELSET>(P, pd, n_dim, u[r], pc[r], pm[r], pha[r], phb[r], s00[r], s10[r], s11[r], s01[r], v, act, f0 + w * r);
		}
	}
}

} // namespace shkz
