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
//
// Kernels (all persistent over the level's list of active tiles, common.cuh):
//   k_sweep             ONE launch = one full red-black sweep (both colours), out of place x_old -> x_new.
//                       A CTA marches a TX x TY tile through its planes: the first colour of plane p+1 is
//                       relaxed from x_old into a three-slot shared-memory ring of half-updated planes (with a
//                       one-cell halo relaxed redundantly), then the second colour of plane p is relaxed from
//                       the half-updated planes p-1, p, p+1. Every array is read once and x written once per
//                       sweep (28 B/cell) instead of once per colour (56 B/cell). Options fold the neighbouring
//                       V-cycle steps in: ZERO_X (first pre-sweep, x_old = 0 not read), PROLONG (first
//                       post-sweep reads x_old + P e_coarse), DOT (last sweep on level 0 also reduces z.r).
//   k_residual_restrict coarse b = P^T (b - A x) in one pass, no residual array.
//   k_vcycle_tail       all levels small enough to live in shared memory run their part of the V-cycle in ONE
//                       single-CTA launch (dozens of launch-latency-bound kernels otherwise).
#pragma once
#include "common.cuh"

namespace shkz {

// ---- the two arithmetic atoms; every smoother / residual in the library goes through these, with explicit
// ---- fused multiply-adds so that fused, tail and legacy kernels agree bit for bit
__device__ __forceinline__ float gs_diag(float w0, float w1, float w2, float w3, float w4, float w5, float dd) {
	return dd + ((w0 + w1) + (w2 + w3) + (w4 + w5));
}
// relaxation factor of the red-black sweeps (MGOmega; 1 = Gauss-Seidel, > 1 = SOR); set per device before every solve
__constant__ float c_mg_omega = 1.f;
// xc = current value of the cell (only used when omega != 1)
__device__ __forceinline__ float gs_relax(float w0, float w1, float w2, float w3, float w4, float w5, float dd, float b, float x0, float x1, float x2,
                                          float x3, float x4, float x5, float xc) {
	const float dg = gs_diag(w0, w1, w2, w3, w4, w5, dd);
	float acc = b;
	acc = __fmaf_rn(w0, x0, acc);
	acc = __fmaf_rn(w1, x1, acc);
	acc = __fmaf_rn(w2, x2, acc);
	acc = __fmaf_rn(w3, x3, acc);
	acc = __fmaf_rn(w4, x4, acc);
	acc = __fmaf_rn(w5, x5, acc);
	if (!(dg > 0.f)) return 0.f;
	const float g = __fdividef(acc, dg);
	return c_mg_omega == 1.f ? g : __fmaf_rn(c_mg_omega, __fsub_rn(g, xc), xc); // (intrinsics: the compiler must not contract the division's multiply into this)
}
__device__ __forceinline__ float gs_relax0(float w0, float w1, float w2, float w3, float w4, float w5, float dd, float b) { // ... from x = 0
	const float dg = gs_diag(w0, w1, w2, w3, w4, w5, dd);
	if (!(dg > 0.f)) return 0.f;
	const float g = __fdividef(b, dg);
	return c_mg_omega == 1.f ? g : __fmul_rn(c_mg_omega, g);
}
// b - A x at one cell; 0 on cells without an equation (their b may be stale)
__device__ __forceinline__ float residual7(float w0, float w1, float w2, float w3, float w4, float w5, float dd, float b, float xc, float x0, float x1,
                                           float x2, float x3, float x4, float x5) {
	float v = __fmaf_rn(-dd, xc, b);
	v = __fmaf_rn(w0, x0 - xc, v);
	v = __fmaf_rn(w1, x1 - xc, v);
	v = __fmaf_rn(w2, x2 - xc, v);
	v = __fmaf_rn(w3, x3 - xc, v);
	v = __fmaf_rn(w4, x4 - xc, v);
	v = __fmaf_rn(w5, x5 - xc, v);
	return gs_diag(w0, w1, w2, w3, w4, w5, dd) > 0.f ? v : 0.f;
}

struct MGLevel {
	Dims d;
	Tiles tiles;
	float *wx, *wy, *wz, *dd; // with ghost planes, pointing at plane 0
	float *b;                 // right-hand side of the level
	float *xa, *xb;           // ping-pong solution buffers
};

constexpr int SWEEP_THREADS = (TX / 2) * TY; // a thread owns two x-adjacent cells of the tile footprint
constexpr int RING_CELLS = 2 * TX + 2 * TY;
static_assert(RING_CELLS <= SWEEP_THREADS && (RING_CELLS % 32) == 0, "ring threads must be whole warps");
inline dim3 sweep_block() { return dim3(TX / 2, TY, 1); }

// One full red-black Gauss-Seidel sweep, colour FIRST then the other one, x_old -> x_new.
template <int FIRST, bool ZERO_X, bool PROLONG, bool DOT>
__global__ void __launch_bounds__(SWEEP_THREADS, 2) k_sweep(Dims d, Tiles T, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                        const float *__restrict__ dd, const float *__restrict__ b, const float *__restrict__ xo,
                                                        float *__restrict__ xn, const float *__restrict__ ec, Dims dc, int slab_ghosts, RedBuf rb, CGState *st) {
	if (st && st->done) return;
	__shared__ float H[3][TY + 2][TX + 2];
	const int px = threadIdx.x, ty = threadIdx.y;
	const int tid = px + (TX / 2) * ty;
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx, ny = d.ny, plane = d.plane;
	double red[1] = {0.0};

	// x_old at (i,j,k) / flat index c (k may be a ghost plane; i, j may be one step outside the grid: such reads
	// wrap to another finite slot of the allocation and only ever meet a zero coefficient)
	// (z-slab solvers: with ZERO_X the ghost planes still hold what the neighbours stored there, slab_ghosts)
	auto XO = [&](long long c, int i, int j, int k) -> float {
		if (ZERO_X && !(slab_ghosts && (k < 0 || k >= d.nzl))) return 0.f;
		float v = xo[c];
		if (PROLONG) v += ec[(i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1))];
		return v;
	};
	// value of cell (i,j) in the half-updated plane p: first-colour cells of in-slab planes relaxed from x_old
	auto half_update = [&](int i, int j, int p, bool in_slab) -> float {
		const long long c = i + nx * (j + ny * p);
		if (in_slab && ((i + j + p + d.k0) & 1) == FIRST) {
			const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + nx], w4 = wz[c], w5 = wz[c + plane];
			if (ZERO_X) return gs_relax0(w0, w1, w2, w3, w4, w5, dd[c], b[c]);
			return gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], XO(c - 1, i - 1, j, p), XO(c + 1, i + 1, j, p), XO(c - nx, i, j - 1, p),
			                XO(c + nx, i, j + 1, p), XO(c - plane, i, j, p - 1), XO(c + plane, i, j, p + 1), XO(c, i, j, p));
		}
		return XO(c, i, j, p);
	};

	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + 2 * px, j = j0 + ty;
		const bool v0 = i < d.nx && j < d.ny, v1 = i + 1 < d.nx && j < d.ny;
		// halo ring cell of this thread (first RING_CELLS threads = whole warps)
		const bool ring = tid < RING_CELLS;
		int ri, rj, rlx, rly;
		if (tid < TX) { ri = i0 + tid; rj = j0 - 1; rlx = tid + 1; rly = 0; }
		else if (tid < 2 * TX) { ri = i0 + tid - TX; rj = j0 + TY; rlx = tid - TX + 1; rly = TY + 1; }
		else if (tid < 2 * TX + TY) { ri = i0 - 1; rj = j0 + tid - 2 * TX; rlx = 0; rly = tid - 2 * TX + 1; }
		else { ri = i0 + TX; rj = j0 + tid - 2 * TX - TY; rlx = TX + 1; rly = tid - 2 * TX - TY + 1; }
		const bool rv = ring && ri >= 0 && ri < d.nx && rj >= 0 && rj < d.ny;

		float hm0 = 0.f, hm1 = 0.f, hc0 = 0.f, hc1 = 0.f;
		for (int p = kb - 1; p <= ke; ++p) {
			const int slot = (p + 3) % 3;
			const bool in_slab = p >= 0 && p < d.nzl;
			// phase 1: half-updated plane p (own pair + ring)
			const float hp0 = v0 ? half_update(i, j, p, in_slab) : 0.f;
			const float hp1 = v1 ? half_update(i + 1, j, p, in_slab) : 0.f;
			H[slot][ty + 1][2 * px + 1] = hp0;
			H[slot][ty + 1][2 * px + 2] = hp1;
			if (ring) H[slot][rly][rlx] = rv ? half_update(ri, rj, p, in_slab) : 0.f;
			__syncthreads();
			// phase 2: finish plane k = p - 1 from the half-updated planes k-1 (hm), k (hc, shared memory), k+1 (hp)
			const int k = p - 1;
			if (k >= kb) {
				const int ks = (k + 3) % 3;
				const int par = (i + j + k + d.k0) & 1; // colour of the even cell of the pair
#pragma unroll
				for (int e = 0; e < 2; ++e) {
					if (!(e ? v1 : v0)) continue;
					const long long c = (i + e) + nx * (j + ny * k);
					const int lx = 2 * px + e + 1, ly = ty + 1;
					float xnew = e ? hc1 : hc0;
					if (((par + e) & 1) != FIRST) {
						const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + nx], w4 = wz[c], w5 = wz[c + plane];
						xnew = gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], H[ks][ly][lx - 1], H[ks][ly][lx + 1], H[ks][ly - 1][lx], H[ks][ly + 1][lx],
						                e ? hm1 : hm0, e ? hp1 : hp0, xnew);
					}
					xn[c] = xnew;
					if (DOT) red[0] += (double)xnew * (double)b[c];
				}
			}
			hm0 = hc0; hm1 = hc1; hc0 = hp0; hc1 = hp1;
		}
		__syncthreads(); // the next tile's first slot may be one this tile's last phase 2 still reads
	}
	if (DOT) {
		grid_reduce<1, 0u>(red, rb, [&](double (&tot)[1]) {
			const double zr = tot[0];
			st->beta = st->iter == 0 ? 0.0 : zr / st->rho; // pcg_solver.h:286-288
			st->rho = zr;
			if (zr == 0.0 || zr != zr) st->done = 1;        // pcg_solver.h:263-271
		});
	}
}

// ---- the same sweep, vectorised: requires nx % 4 == 0 (every level of a power-of-two grid down to 4) -----------
// A thread owns FOUR x-adjacent cells (one aligned float4 per array and plane) of one tile row; the block holds
// TY+2 row slots (the two extra ones relax the halo rows in phase 1 only) plus one warp for the two halo
// columns. All global loads of a plane happen once, at the top of the step; the coefficients of the plane are
// carried in registers to the next step where the second colour needs them, so phase 2 touches shared memory only.
// Rows r and r+2 share a warp, hence which cells of the quad carry the first colour is warp-uniform.
constexpr int S4_ROWS = TY + 2;
constexpr int S4_ROW_WARPS = S4_ROWS / 2;          // 9
constexpr int S4_THREADS = 32 * (S4_ROW_WARPS + 1); // + the halo-column warp
constexpr int S4_PITCH = TX + 8;                    // own quads start at column 4 (16-byte aligned), halo columns at 3 and TX+4
static_assert(TY == 16 && TX == 64, "k_sweep4 thread mapping assumes 64 x 16 tiles");

struct Quad { // operator data of four x-adjacent cells in one plane
	float4 wx;  // lower x faces of the four cells
	float wx4;  // ... and of the cell after them
	float4 wy, wyu, wz, dd, b; // lower y faces, upper y faces (next row's wy), lower z faces, Dirichlet diagonal, rhs
};

template <int M> __device__ __forceinline__ float q_get(const float4 &v) { return M == 0 ? v.x : (M == 1 ? v.y : (M == 2 ? v.z : v.w)); }
template <int M> __device__ __forceinline__ void q_set(float4 &v, float f) { if (M == 0) v.x = f; else if (M == 1) v.y = f; else if (M == 2) v.z = f; else v.w = f; }

// relax cell M of the quad: x = values of the quad itself, xl / xr = the cells left and right of it, xd / xu = the rows below and
// above, zm / zp = the planes below and above, wzu = upper z faces
template <int M, bool ZERO>
__device__ __forceinline__ float relax_cell(const Quad &Q, const float4 &wzu, const float4 &x, float xl, float xr, const float4 &xd, const float4 &xu,
                                            const float4 &zm, const float4 &zp) {
	const float w0 = q_get<M>(Q.wx), w1 = M == 3 ? Q.wx4 : q_get<(M + 1) & 3>(Q.wx);
	const float w2 = q_get<M>(Q.wy), w3 = q_get<M>(Q.wyu), w4 = q_get<M>(Q.wz), w5 = q_get<M>(wzu);
	if (ZERO) return gs_relax0(w0, w1, w2, w3, w4, w5, q_get<M>(Q.dd), q_get<M>(Q.b));
	const float x0 = M == 0 ? xl : q_get<(M + 3) & 3>(x), x1 = M == 3 ? xr : q_get<(M + 1) & 3>(x);
	return gs_relax(w0, w1, w2, w3, w4, w5, q_get<M>(Q.dd), q_get<M>(Q.b), x0, x1, q_get<M>(xd), q_get<M>(xu), q_get<M>(zm), q_get<M>(zp), q_get<M>(x));
}
// relax cells A and A+2 of the quad, keep the other two
template <int A, bool ZERO>
__device__ __forceinline__ float4 relax_quad(const Quad &Q, const float4 &wzu, const float4 &x, float xl, float xr, const float4 &xd, const float4 &xu,
                                             const float4 &zm, const float4 &zp) {
	float4 out = x;
	q_set<A>(out, relax_cell<A, ZERO>(Q, wzu, x, xl, xr, xd, xu, zm, zp));
	q_set<A + 2>(out, relax_cell<A + 2, ZERO>(Q, wzu, x, xl, xr, xd, xu, zm, zp));
	return out;
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ld4cg(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); } // L2: data a peer GPU may have just stored

// z-slab solvers: the halo traffic of ONE red-black sweep, carried by the sweep kernel itself (cm == nullptr on a whole grid).
//   start    wait for exchange `wait_in` (the ghost planes of x_old, pushed by the previous sweep of the neighbours);
//            relax the first colour of the own boundary planes and store those half-updated planes into the neighbours'
//            ghost planes of x_old; the block that finishes last publishes exchange `seq_half`
//   tiles    first every tile that touches no ghost plane, then — once the neighbours' `seq_half` is here — the rest;
//            the boundary planes of x_new go to the neighbours' ghost planes of x_new as they are computed
//   end      the block that finishes last publishes exchange `seq_out` (0: nothing reads those ghost planes)
// so one launch replaces three (half-plane push + wait, sweep, halo push + wait) and the NVLink round trip is hidden
// behind the interior tiles. The ghost plane then holds exactly what both phases of the neighbour's sweep need (second-colour
// cells still old for phase 1, first-colour cells new for phase 2), which is why the slab V-cycle equals the whole-grid one.
struct SlabSweep {
	const CommDev *cm;
	size_t off_xo, off_xn;                 // arena offsets of the lower ghost planes of x_old / x_new
	unsigned long long wait_in, seq_half, seq_out;
	const float *wx, *wy, *wz, *dd, *b;    // the level's arrays (the TMA kernel otherwise only holds tensor maps)
};

// start of a fused slab sweep: every thread of the grid takes quads of the two boundary planes (nx % 4 == 0, 1-D blocks)
// PROLONG: x_old is "x_old + P e_coarse" everywhere it is read — on the boundary planes, their in-slab neighbours and the ghost planes beyond them (the coarse
// correction `ec` has valid ghost planes: pushed by the coarse level's last sweep, or simply the next planes of the gathered global level). The half-updated
// planes that leave here carry relaxed values on the first colour and the PLAIN x_old, without the correction, on the second: a ghost plane must read the same on
// its second-colour cells before and after the neighbour's half-updated plane lands on it (that is what makes reading it while it lands race-free), so the
// correction of those cells is added by whoever reads them — here for the planes beyond the boundary, in the stage for the tiles (k_sweep_tma: add_quad).
template <int FIRST, bool ZERO_X, bool PROLONG>
__device__ __forceinline__ void slab_push_half_planes(const Dims &d, const SlabSweep &sl, const float *__restrict__ xo, const float *__restrict__ ec, const Dims &dc) {
	const CommDev *cm = sl.cm;
	const long long nx = d.nx, plane = d.plane, quads = plane >> 2, nth = (long long)gridDim.x * blockDim.x;
	const int qx = d.nx >> 2;
	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	auto EC = [&](int i, int j, int k) -> long long { return (i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1)); };
	auto add4 = [&](float4 &v, int i, int j, int k) { // + P e on an aligned quad inside the grid's x / y extent (k may be a ghost plane)
		if (j < 0 || j >= d.ny) return;
		const float2 e = __ldcg(reinterpret_cast<const float2 *>(ec + EC(i, j, k)));
		v.x += e.x; v.y += e.x; v.z += e.y; v.w += e.y;
	};
	for (int side = 0; side < 2; ++side) {
		char *peer = side ? cm->hi : cm->lo;
		if (!peer) continue;
		const int p = side ? d.nzl - 1 : 0;
		float *dst = reinterpret_cast<float *>(peer + sl.off_xo) + (side ? 0 : (long long)(d.nzl + 1) * plane);
		for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < quads; e += nth) {
			const int j = (int)(e / qx), i = 4 * (int)(e - (long long)j * qx);
			const long long c = i + nx * j + plane * p;
			Quad Q;
			Q.wx = ld4(sl.wx + c); Q.wx4 = sl.wx[c + 4];
			Q.wy = ld4(sl.wy + c); Q.wyu = ld4(sl.wy + c + nx);
			Q.wz = ld4(sl.wz + c); Q.dd = ld4(sl.dd + c); Q.b = ld4(sl.b + c);
			const float4 wzu = ld4(sl.wz + c + plane);
			float4 x = zero4, xd = zero4, xu = zero4, zm = zero4, zp = zero4, x_plain = zero4;
			float xl = 0.f, xr = 0.f;
			if (!ZERO_X) { // (the plane beyond the boundary is a ghost plane: read it where the neighbour's stores land, in L2)
				x = ld4cg(xo + c); xl = __ldcg(xo + c - 1); xr = __ldcg(xo + c + 4);
				xd = ld4cg(xo + c - nx); xu = ld4cg(xo + c + nx); zm = ld4cg(xo + c - plane); zp = ld4cg(xo + c + plane);
				x_plain = x;
				if (PROLONG) {
					add4(x, i, j, p); add4(xd, i, j - 1, p); add4(xu, i, j + 1, p); add4(zm, i, j, p - 1); add4(zp, i, j, p + 1);
					if (i > 0) xl += __ldcg(ec + EC(i - 1, j, p));
					if (i + 4 < d.nx) xr += __ldcg(ec + EC(i + 4, j, p));
				}
			}
			const int a1 = (FIRST + j + p + d.k0) & 1;
			float4 h = a1 == 0 ? relax_quad<0, ZERO_X>(Q, wzu, x, xl, xr, xd, xu, zm, zp) : relax_quad<1, ZERO_X>(Q, wzu, x, xl, xr, xd, xu, zm, zp);
			if (PROLONG) { // the cells of the OTHER colour leave as they are in memory, without the correction (see above)
				if (a1 == 0) { h.y = x_plain.y; h.w = x_plain.w; }
				else { h.x = x_plain.x; h.z = x_plain.z; }
			}
			*reinterpret_cast<float4 *>(dst + (c - plane * p)) = h;
		}
	}
	signal_neighbours(cm, sl.seq_half);
}

template <int FIRST, bool ZERO_X, bool PROLONG, bool DOT>
__global__ void __launch_bounds__(S4_THREADS, 2) k_sweep4(Dims d, Tiles T, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                         const float *__restrict__ dd, const float *__restrict__ b, const float *__restrict__ xo,
                                                         float *__restrict__ xn, const float *__restrict__ ec, Dims dc, int slab_ghosts, RedBuf rb, CGState *st) {
	if (st && st->done) return;
	__shared__ __align__(16) float H[3][S4_ROWS][S4_PITCH];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool colwarp = warp == S4_ROW_WARPS;
	const int tx = lane & 15, half = lane >> 4;
	const int r = warp < S4_ROW_WARPS - 1 ? ((warp >> 1) * 4 + (warp & 1) + 2 * half) : TY + half; // row slot of a row-warp thread
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx, ny = d.ny, plane = d.plane;
	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	double red[1] = {0.0};

	auto EC = [&](int i, int j, int k) -> long long { return (i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1)); };
	auto XO = [&](long long c, int i, int j, int k) -> float { // one cell of x_old (see k_sweep)
		if (ZERO_X && !(slab_ghosts && (k < 0 || k >= d.nzl))) return 0.f;
		float v = xo[c];
		if (PROLONG) v += ec[EC(i, j, k)];
		return v;
	};
	auto XO4 = [&](long long c0, int i, int j, int k) -> float4 { // an aligned quad of x_old
		if (ZERO_X && !(slab_ghosts && (k < 0 || k >= d.nzl))) return zero4;
		float4 v = ld4(xo + c0);
		if (PROLONG) {
			const float2 e = *reinterpret_cast<const float2 *>(ec + EC(i, j, k));
			v.x += e.x; v.y += e.x; v.z += e.y; v.w += e.y;
		}
		return v;
	};
	auto scalar_half_update = [&](int i, int j, int p, bool in_slab) -> float { // halo columns
		const long long c = i + nx * (j + ny * p);
		if (in_slab && ((i + j + p + d.k0) & 1) == FIRST) {
			const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + nx], w4 = wz[c], w5 = wz[c + plane];
			if (ZERO_X) return gs_relax0(w0, w1, w2, w3, w4, w5, dd[c], b[c]);
			return gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], XO(c - 1, i - 1, j, p), XO(c + 1, i + 1, j, p), XO(c - nx, i, j - 1, p),
			                XO(c + nx, i, j + 1, p), XO(c - plane, i, j, p - 1), XO(c + plane, i, j, p + 1), XO(c, i, j, p));
		}
		return XO(c, i, j, p);
	};

	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		if (colwarp) {
			// ---- halo columns i0-1 and i0+TX of the TY tile rows: one cell per lane and plane
			const int side = lane >> 4, ci = side ? i0 + TX : i0 - 1, cj = j0 + (lane & 15);
			const bool cv = ci >= 0 && ci < d.nx && cj < d.ny;
			for (int p = kb - 1; p <= ke; ++p) {
				H[(p + 3) % 3][(lane & 15) + 1][side ? TX + 4 : 3] = cv ? scalar_half_update(ci, cj, p, p >= 0 && p < d.nzl) : 0.f;
				__syncthreads();
			}
			__syncthreads();
			continue;
		}
		const int i = i0 + 4 * tx, j = j0 - 1 + r;
		const bool valid = i < d.nx && j >= 0 && j < d.ny;
		const bool finish = valid && r >= 1 && r <= TY;
		const long long row = i + nx * j; // flat index of the quad in plane 0
		// x_old of the own quad in planes p-1, p (p+1 is loaded in the step), lower z faces of plane p
		float4 xm = zero4, xc = zero4, wz_cur = zero4;
		if (valid) {
			if (kb - 2 >= -1) xm = XO4(row + plane * (kb - 2), i, j, kb - 2);
			xc = XO4(row + plane * (kb - 1), i, j, kb - 1);
			wz_cur = ld4(wz + row + plane * (kb - 1));
		}
		float4 hm = zero4, hc = zero4;
		Quad prv; // operator data of plane p-1
		prv.wx = prv.wy = prv.wyu = prv.wz = prv.dd = prv.b = zero4;
		prv.wx4 = 0.f;
		for (int p = kb - 1; p <= ke; ++p) {
			const int slot = (p + 3) % 3;
			const bool in_slab = p >= 0 && p < d.nzl;
			const long long c0 = row + plane * p;
			// ---- loads of the step
			Quad cur;
			cur.wz = wz_cur;
			float4 wz_next = zero4, xp = zero4, xd = zero4, xu = zero4;
			float xl = 0.f, xr = 0.f;
			const bool relaxing = valid && in_slab;
			if (valid && p + 1 <= d.nzl) {
				wz_next = ld4(wz + c0 + plane);
				xp = XO4(c0 + plane, i, j, p + 1);
			}
			if (relaxing) {
				cur.wx = ld4(wx + c0); cur.wx4 = wx[c0 + 4];
				cur.wy = ld4(wy + c0); cur.wyu = ld4(wy + c0 + nx);
				cur.dd = ld4(dd + c0); cur.b = ld4(b + c0);
				if (!ZERO_X) {
					xl = XO(c0 - 1, i - 1, j, p); xr = XO(c0 + 4, i + 4, j, p);
					xd = XO4(c0 - nx, i, j - 1, p); xu = XO4(c0 + nx, i, j + 1, p);
				}
			} else {
				cur.wx = cur.wy = cur.wyu = cur.dd = cur.b = zero4;
				cur.wx4 = 0.f;
			}
			// ---- phase 1: half-updated plane p
			float4 hp = xc;
			const int a1 = (FIRST + j + p + d.k0) & 1; // first relaxed cell of the quad in this row and plane (i is a multiple of 4)
			if (relaxing) {
				if (a1 == 0) hp = relax_quad<0, ZERO_X>(cur, wz_next, xc, xl, xr, xd, xu, xm, xp);
				else hp = relax_quad<1, ZERO_X>(cur, wz_next, xc, xl, xr, xd, xu, xm, xp);
			}
			*reinterpret_cast<float4 *>(&H[slot][r][4 + 4 * tx]) = valid ? hp : zero4;
			__syncthreads();
			// ---- phase 2: finish plane k = p - 1 (second colour) from half-updated planes k-1 (hm), k (hc + shared memory), k+1 (hp)
			const int k = p - 1;
			if (finish && k >= kb) {
				const int ks = (k + 3) % 3;
				const float hl = H[ks][r][3 + 4 * tx], hr = H[ks][r][8 + 4 * tx];
				const float4 hd = *reinterpret_cast<const float4 *>(&H[ks][r - 1][4 + 4 * tx]);
				const float4 hu = *reinterpret_cast<const float4 *>(&H[ks][r + 1][4 + 4 * tx]);
				const int a2 = (FIRST + 1 + j + k + d.k0) & 1; // first cell of the other colour
				float4 xnew;
				if (a2 == 0) xnew = relax_quad<0, false>(prv, cur.wz, hc, hl, hr, hd, hu, hm, hp);
				else xnew = relax_quad<1, false>(prv, cur.wz, hc, hl, hr, hd, hu, hm, hp);
				*reinterpret_cast<float4 *>(xn + c0 - plane) = xnew;
				if (DOT) red[0] += (double)xnew.x * (double)prv.b.x + (double)xnew.y * (double)prv.b.y + (double)xnew.z * (double)prv.b.z + (double)xnew.w * (double)prv.b.w;
			}
			hm = hc; hc = hp;
			xm = xc; xc = xp;
			wz_cur = wz_next;
			prv = cur;
		}
		__syncthreads(); // the next tile's first slot may be one this tile's last phase 2 still reads
	}
	if (DOT) {
		grid_reduce<1, 0u>(red, rb, [&](double (&tot)[1]) {
			const double zr = tot[0];
			st->beta = st->iter == 0 ? 0.0 : zr / st->rho; // pcg_solver.h:286-288
			st->rho = zr;
			if (zr == 0.0 || zr != zr) st->done = 1;        // pcg_solver.h:263-271
		});
	}
}

// z-slab solvers: the half-updated version (first colour relaxed) of the own boundary planes, stored straight into the
// neighbours' ghost planes of the same buffer. A neighbour's sweep kernel then finds in its ghost plane exactly what
// both of its phases need (second-colour cells still old for phase 1, first-colour cells new for phase 2), which is
// why the sweep kernels need no slab logic beyond reading their ghost planes. blockIdx.y: 0 = plane 0 to the lower
// neighbour, 1 = plane nzl-1 to the upper one.
template <int FIRST, bool ZERO_X>
__global__ void __launch_bounds__(256) k_boundary_half_push(Dims d, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                           const float *__restrict__ dd, const float *__restrict__ b, const float *__restrict__ xo,
                                                           const CommDev *cm, size_t off, unsigned long long seq) {
	const int p = blockIdx.y ? d.nzl - 1 : 0;
	char *peer = blockIdx.y ? cm->hi : cm->lo;
	if (peer) {
		float *dst = reinterpret_cast<float *>(peer + off) + (blockIdx.y ? 0 : (long long)(d.nzl + 1) * d.plane);
		const long long nx = d.nx, plane = d.plane;
		for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < plane; e += (long long)gridDim.x * blockDim.x) {
			const int j = (int)(e / nx), i = (int)(e - (long long)j * nx);
			const long long c = e + plane * p;
			float v;
			if (((i + j + p + d.k0) & 1) == FIRST) {
				const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + nx], w4 = wz[c], w5 = wz[c + plane];
				// with ZERO_X the in-slab x_old is zero but the ghost planes already hold the neighbours' values: none yet, x_old = 0 everywhere
				v = ZERO_X ? gs_relax0(w0, w1, w2, w3, w4, w5, dd[c], b[c])
				           : gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], xo[c - 1], xo[c + 1], xo[c - nx], xo[c + nx], xo[c - plane], xo[c + plane], xo[c]);
			} else v = ZERO_X ? 0.f : xo[c];
			dst[e] = v;
		}
	}
	signal_neighbours(cm, seq);
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) wait_neighbours(cm, seq);
}

// coarse b = P^T (b - A x): block (TX/2, TY/2), one coarse column per thread, aggregates never straddle tiles
inline dim3 restrict_block() { return dim3(TX / 2, TY / 2, 1); }
__global__ void __launch_bounds__((TX / 2) * (TY / 2)) k_residual_restrict(Dims d, Tiles T, const float *__restrict__ wx, const float *__restrict__ wy,
                                                                          const float *__restrict__ wz, const float *__restrict__ dd, const float *__restrict__ b,
                                                                          const float *__restrict__ x, Dims dc, float *__restrict__ bc,
                                                                          const CGState *__restrict__ st, const CommDev *cm, unsigned long long wait_in) {
	if (st && st->done) return;
	if (cm && wait_in) block_wait_neighbours(cm, wait_in); // z-slabs: the ghost planes of x come from the neighbours' last fused sweep
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx, ny = d.ny, plane = d.plane;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int I = (i0 >> 1) + threadIdx.x, J = (j0 >> 1) + threadIdx.y;
		if (I >= dc.nx || J >= dc.ny) continue;
		for (int k = kb; k < ke; k += 2) {
			float acc = 0.f;
#pragma unroll
			for (int dk = 0; dk < 2; ++dk)
#pragma unroll
				for (int dj = 0; dj < 2; ++dj)
#pragma unroll
					for (int di = 0; di < 2; ++di) {
						const int i = 2 * I + di, j = 2 * J + dj, kk = k + dk;
						if (i < d.nx && j < d.ny && kk < d.nzl) {
							const long long c = i + nx * (j + ny * kk);
							acc += residual7(wx[c], wx[c + 1], wy[c], wy[c + nx], wz[c], wz[c + plane], dd[c], b[c], x[c], x[c - 1], x[c + 1], x[c - nx],
							                 x[c + nx], x[c - plane], x[c + plane]);
						}
					}
			bc[I + (long long)dc.nx * (J + (long long)dc.ny * (k >> 1))] = acc;
		}
	}
}

// x_new = x_old + P e (only used when MGPostSweeps = 0; otherwise the first post-sweep folds it in)
// (z-slabs: the boundary planes of x_new go straight into the neighbours' ghost planes, SlabPush; the first post-sweep waits for them)
__global__ void __launch_bounds__(TX *8) k_prolong_add(Dims d, Tiles T, Dims dc, const float *__restrict__ ec, const float *__restrict__ xo, float *__restrict__ xn,
                                                      const CGState *__restrict__ st, const SlabPush sp) {
	if (st && st->done) return;
	float *const plo = push_target_lo<float>(sp, d.plane, d.nzl), *const phi = push_target_hi<float>(sp);
	resolve_tiles(T);
	const int ntiles = *T.count;
	int i0, j0, kb, ke;
	for (TileWalk w(T, ntiles, true); w.next(T, d.nzl, i0, j0, kb, ke);) {
		const int i = i0 + threadIdx.x, je = min(j0 + TY, d.ny);
		if (i >= d.nx) continue;
		for (int k = kb; k < ke; ++k)
			for (int j = j0 + threadIdx.y; j < je; j += 8) {
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				const float v = xo[c] + ec[(i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1))];
				xn[c] = v;
				if (k == 0 && plo) plo[c] = v;
				if (k == d.nzl - 1 && phi) phi[c - (long long)k * d.plane] = v;
			}
	}
	if (sp.cm) signal_neighbours(sp.cm, sp.seq);
}

// ---- the shared-memory tail of the V-cycle ----------------------------------------------------------------
constexpr int TAIL_MAX_LEVELS = 12;
constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_ARRAYS = 6; // wx wy wz dd x b
constexpr int TAIL_SLOTS = 4;  // pairs of x-adjacent cells a thread may own on one tail level (setup_tail keeps larger levels out of the tail)

struct TailLevel {
	Dims d;
	const float *wx, *wy, *wz, *dd; // global, pointing at plane 0 (ghost planes allocated)
	int offset;                     // first float of this level's block in shared memory
	int stride;                     // ncell + 2*plane: floats per array
};
struct TailArgs {
	int nlev;
	int pre, post, coarse;
	const float *b_in; // right-hand side of the first tail level (global)
	float *x_out;      // its solution (global)
	TailLevel L[TAIL_MAX_LEVELS];
};

// (one CTA of TAIL_THREADS threads; sm = its dynamic shared memory)
__device__ __forceinline__ void tail_body(const TailArgs &A, float *sm) {
	const int tid = threadIdx.x;
	auto arr = [&](int l, int which) -> float * { return sm + A.L[l].offset + which * A.L[l].stride + (int)A.L[l].d.plane; };
	// stage the operators; x and the ghost planes of b start at zero
	for (int l = 0; l < A.nlev; ++l) {
		const TailLevel &L = A.L[l];
		const int plane = (int)L.d.plane, n = L.stride;
		float *wx = arr(l, 0), *wy = arr(l, 1), *wz = arr(l, 2), *dd = arr(l, 3), *x = arr(l, 4), *b = arr(l, 5);
		for (int c = tid - plane; c < n - plane; c += TAIL_THREADS) {
			wx[c] = L.wx[c]; wy[c] = L.wy[c]; wz[c] = L.wz[c]; dd[c] = L.dd[c];
			x[c] = 0.f;
			b[c] = (l == 0 && c >= 0 && c < (int)L.d.ncell) ? __ldcg(A.b_in + c) : 0.f; // (L2: inside k_vcycle_mid other SMs have just written it)
		}
	}
	__syncthreads();
	// Relaxation of one colour. The cells of a level are walked as PAIRS along x (hx = ceil(nx / 2) pairs per row): pair slot q = tid + TAIL_THREADS * m belongs
	// to this thread, the colour picks which cell of the pair — every thread of a warp works, on every pass (a walk over all cells that skips the other colour
	// keeps half of each warp idle). load_slots() resolves the thread's slots of a level once per visit (the index divisions are as expensive as the relaxation
	// itself); only the threads that own a slot take part in the level's barriers: a named barrier over those warps, or none at all when one warp does the
	// whole level (the 4^3 bottom of the hierarchy, 16 sweeps per V-cycle).
	__shared__ int s_cells[TAIL_SLOTS][TAIL_THREADS]; // (shared, not registers: 1024 threads leave 64 registers each) per slot: flat index of the pair's first cell | bit 29: the pair has no second cell (odd nx, last pair) | bit 30: (j + k + k0) & 1; -1: no such slot
	int n_active = 0;       // threads with a slot, rounded up to whole warps
	auto load_slots = [&](int l) {
		const Dims &d = A.L[l].d;
		const int hx = (d.nx + 1) >> 1, slots = hx * d.ny * d.nzl;
		n_active = min(TAIL_THREADS, (slots + 31) & ~31);
#pragma unroll
		for (int m = 0; m < TAIL_SLOTS; ++m) {
			const int q = tid + TAIL_THREADS * m;
			int v = -1;
			if (q < slots) {
				const int rowi = q / hx, p2 = q - rowi * hx, k = rowi / d.ny, j = rowi - k * d.ny;
				v = (d.nx * rowi + 2 * p2) | ((2 * p2 + 1 >= d.nx) ? (1 << 29) : 0) | (((j + k + d.k0) & 1) << 30);
			}
			s_cells[m][tid] = v;
		}
	};
	auto level_sync = [&]() {
		if (n_active <= 32) __syncwarp();
		else if (n_active < TAIL_THREADS) asm volatile("bar.sync 1, %0;" ::"r"(n_active) : "memory");
		else __syncthreads();
	};
	auto half = [&](int l, int color, bool zero_x) { // (threads without a slot never get here)
		const Dims &d = A.L[l].d;
		const int nx = d.nx, plane = (int)d.plane;
		const float *wx = arr(l, 0), *wy = arr(l, 1), *wz = arr(l, 2), *dd = arr(l, 3), *b = arr(l, 5);
		float *x = arr(l, 4);
#pragma unroll 1
		for (int m = 0; m < TAIL_SLOTS; ++m) {
			const int sc = s_cells[m][tid];
			if (sc < 0) break;
			const int second = (color + (sc >> 30)) & 1; // which cell of the pair has this colour
			if (second && (sc & (1 << 29))) continue;
			const int c = (sc & 0xffffff) + second;
			const float w0 = wx[c], w1 = wx[c + 1], w2 = wy[c], w3 = wy[c + nx], w4 = wz[c], w5 = wz[c + plane];
			// (a neighbour across a zero coefficient — a wall, whose flat index wraps to a cell of the SAME colour — is not read: the product is an
			// exact zero either way, and the in-place update then never reads a value another thread may be writing)
			x[c] = zero_x ? gs_relax0(w0, w1, w2, w3, w4, w5, dd[c], b[c])
			              : gs_relax(w0, w1, w2, w3, w4, w5, dd[c], b[c], w0 != 0.f ? x[c - 1] : 0.f, w1 != 0.f ? x[c + 1] : 0.f, w2 != 0.f ? x[c - nx] : 0.f,
			                         w3 != 0.f ? x[c + nx] : 0.f, w4 != 0.f ? x[c - plane] : 0.f, w5 != 0.f ? x[c + plane] : 0.f, x[c]);
		}
		level_sync();
	};
	// `sweeps` red-black sweeps of level l, first colour `c0`; zero: x starts from zero. Ends with a block barrier.
	auto smooth = [&](int l, int sweeps, int c0, bool zero) {
		load_slots(l);
		if (tid < n_active)
			for (int sw = 0; sw < sweeps; ++sw) {
				half(l, c0, zero && sw == 0);
				half(l, c0 ^ 1, false);
			}
		__syncthreads();
	};
	for (int l = 0; l < A.nlev; ++l) { // descend
		const bool last = l + 1 == A.nlev;
		smooth(l, last ? A.coarse : A.pre, 0, true);
		if (last) break;
		const Dims &d = A.L[l].d, &dc = A.L[l + 1].d;
		const int nx = d.nx, ny = d.ny, plane = (int)d.plane;
		const float *wx = arr(l, 0), *wy = arr(l, 1), *wz = arr(l, 2), *dd = arr(l, 3), *x = arr(l, 4), *b = arr(l, 5);
		float *bc = arr(l + 1, 5);
		const int cplane = (int)dc.plane;
		for (int C = tid; C < (int)dc.ncell; C += TAIL_THREADS) {
			const int K = C / cplane, rem = C - K * cplane, J = rem / dc.nx, I = rem - J * dc.nx;
			float acc = 0.f;
			for (int dk = 0; dk < 2; ++dk)
				for (int dj = 0; dj < 2; ++dj)
					for (int di = 0; di < 2; ++di) {
						const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
						if (i < nx && j < ny && k < d.nzl) {
							const int c = i + nx * (j + ny * k);
							acc += residual7(wx[c], wx[c + 1], wy[c], wy[c + nx], wz[c], wz[c + plane], dd[c], b[c], x[c], x[c - 1], x[c + 1], x[c - nx],
							                 x[c + nx], x[c - plane], x[c + plane]);
						}
					}
			bc[C] = acc;
		}
		__syncthreads();
	}
	for (int l = A.nlev - 1; l >= 0; --l) { // ascend
		const bool last = l + 1 == A.nlev;
		if (!last) {
			const Dims &d = A.L[l].d, &dc = A.L[l + 1].d;
			const int nx = d.nx, plane = (int)d.plane;
			float *x = arr(l, 4);
			const float *ec = arr(l + 1, 4);
			for (int c = tid; c < (int)d.ncell; c += TAIL_THREADS) {
				const int k = c / plane, rem = c - k * plane, j = rem / nx, i = rem - j * nx;
				x[c] += ec[(i >> 1) + dc.nx * ((j >> 1) + dc.ny * (k >> 1))];
			}
			__syncthreads();
		}
		smooth(l, last ? A.coarse : A.post, 1, false);
	}
	const float *x = arr(0, 4);
	for (int c = tid; c < (int)A.L[0].d.ncell; c += TAIL_THREADS) A.x_out[c] = x[c];
}

__global__ void __launch_bounds__(TAIL_THREADS) k_vcycle_tail(TailArgs A, const CGState *__restrict__ st) {
	if (st && st->done) return;
	extern __shared__ float sm[];
	tail_body(A, sm);
}

// ---- the middle of the V-cycle in ONE cooperative launch -----------------------------------------------------------------
// Levels that are too large for one CTA's shared memory but far too small to fill the GPU (a few 10^5 unknowns and less: from the third level down on a
// 512^3 liquid scene, the whole hierarchy of a 64^3 one) used to cost five launches each per V-cycle, every one of them ~12 us of launch latency, pipeline
// fill and tail for microseconds of work — a quarter of the solve time for a tenth of its bytes. Here ONE launch of one CTA per SM walks all of them down
// and up again, in place, separated by grid-wide barriers (cooperative launch: every CTA is resident); the data lives in L2. The shared-memory tail runs
// inside the same launch on CTA 0. Arithmetic is that of the tiled kernels (same atoms, same order): in-place red-black relaxation is what the out-of-place
// fused sweep computes, so results agree bit for bit with every other path (tests/test_gpu_parity.py).
//   work item = one row of one plane of one active tile, taken by a warp: lane l relaxes cell i0 + 2 l + (colour offset of the row)
constexpr int MID_MAX_LEVELS = 8;
constexpr int MID_THREADS = 1024;
struct MidLevel {
	Dims d;
	Tiles tiles;
	const float *wx, *wy, *wz, *dd;
	float *b, *x; // right-hand side, solution (in place)
};
struct MidArgs {
	int nlev;              // levels walked here, finest first
	int pre, post, coarse; // sweeps (coarse: on the last level when no tail follows)
	int has_tail;          // the shared-memory tail (TailArgs) continues below the last level
	int dot;               // finest level is level 0 of the solve: reduce z.b into the CG state at the end
	unsigned *barrier;     // [0] arrivals, [1] generation
	MidLevel L[MID_MAX_LEVELS];
};

// Grid-wide barrier of a cooperative launch: arrivals counted with a release atomic (the CTA's writes, ordered before it by the block barrier, become
// visible with it), the last arrival resets the counter and publishes the next generation, everybody else polls the generation with acquire loads.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &gen) {
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned target = ++gen;
		unsigned prev;
		asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
		if (prev == gridDim.x - 1) {
			asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
			asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(bar + 1), "r"(target) : "memory");
		} else {
			unsigned seen;
			do {
				asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory");
			} while (seen != target);
		}
	}
	__syncthreads();
}

enum MidMode { MID_ZERO_A, MID_ZERO_B, MID_PLAIN, MID_PROLONG_A, MID_PROLONG_B };

__global__ void __launch_bounds__(MID_THREADS, 1) k_vcycle_mid(const __grid_constant__ MidArgs A, const __grid_constant__ TailArgs TA, RedBuf rb, CGState *st) {
	if (st && st->done) return;
	extern __shared__ float sm[];
	unsigned gen = *reinterpret_cast<volatile unsigned *>(&A.barrier[1]);
	const int lane = threadIdx.x & 31;
	__shared__ int s_ntiles[MID_MAX_LEVELS], s_bz[MID_MAX_LEVELS]; // the levels' active-tile counts and tile depths, fetched once
	if (threadIdx.x < MID_MAX_LEVELS) {
		const bool have = (int)threadIdx.x < A.nlev;
		s_ntiles[threadIdx.x] = have ? A.L[threadIdx.x].tiles.count[0] : 0;
		s_bz[threadIdx.x] = have ? (A.L[threadIdx.x].tiles.bz ? A.L[threadIdx.x].tiles.bz : A.L[threadIdx.x].tiles.count[1]) : 2;
	}
	__syncthreads();
	const long long gwarp = ((long long)blockIdx.x * MID_THREADS + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * MID_THREADS) >> 5;

	// one colour of one red-black sweep on level l; ec / dc: the coarse correction of the PROLONG modes
	auto relax = [&](int l, int colour, int mode, const float *ec, const Dims &dc) {
		const MidLevel &L = A.L[l];
		const Dims &d = L.d;
		Tiles T = L.tiles;
		T.bz = s_bz[l];
		const long long nx = d.nx, ny = d.ny, plane = d.plane;
		const int ntiles = s_ntiles[l];
		const long long items = (long long)ntiles * T.bz * TY;
		auto E = [&](int i, int j, int k) -> float { return __ldcg(ec + ((i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1)))); };
		for (long long it = gwarp; it < items; it += nwarps) {
			const int t = (int)(it / (T.bz * TY)), rem = (int)(it - (long long)t * (T.bz * TY));
			int i0, j0, kb;
			tile_origin(T, T.ids[t], i0, j0, kb);
			const int k = kb + rem / TY, j = j0 + rem % TY;
			if (k >= d.nzl || j >= d.ny) continue;
			const int i = i0 + 2 * lane + ((colour + j + k + d.k0) & 1);
			if (i >= d.nx) continue;
			const long long c = i + nx * (j + ny * k);
			const float w0 = L.wx[c], w1 = L.wx[c + 1], w2 = L.wy[c], w3 = L.wy[c + nx], w4 = L.wz[c], w5 = L.wz[c + plane], dg = L.dd[c];
			const float bb = __ldcg(L.b + c);
			float xn;
			if (mode == MID_ZERO_A) xn = gs_relax0(w0, w1, w2, w3, w4, w5, dg, bb);
			else {
				// every operand in one round trip: a neighbour across a zero coefficient (a wall, whose flat index wraps to another cell of the allocation,
				// or a cell without an equation) is read like any other — whatever finite value it holds meets an exact zero, as in the tiled kernels
				float x0 = __ldcg(L.x + c - 1), x1 = __ldcg(L.x + c + 1), x2 = __ldcg(L.x + c - nx), x3 = __ldcg(L.x + c + nx), x4 = __ldcg(L.x + c - plane),
				      x5 = __ldcg(L.x + c + plane);
				float xc = mode == MID_ZERO_B ? 0.f : __ldcg(L.x + c);
				if (mode == MID_PROLONG_A) { // neighbours (the other colour) and the cell itself still lack the coarse correction
					if (i > 0) x0 += E(i - 1, j, k);
					if (i + 1 < d.nx) x1 += E(i + 1, j, k);
					if (j > 0) x2 += E(i, j - 1, k);
					if (j + 1 < d.ny) x3 += E(i, j + 1, k);
					if (k > 0) x4 += E(i, j, k - 1);
					if (k + 1 < d.nzl) x5 += E(i, j, k + 1);
				}
				if (mode == MID_PROLONG_A || mode == MID_PROLONG_B) xc += E(i, j, k);
				xn = gs_relax(w0, w1, w2, w3, w4, w5, dg, bb, x0, x1, x2, x3, x4, x5, xc);
			}
			__stcg(L.x + c, xn);
		}
	};
	// coarse b = P^T (b - A x), level l -> l + 1 (bc: the coarse right-hand side, which may be the tail's)
	auto restrict_to = [&](int l, float *bc, const Dims &dc) {
		const MidLevel &L = A.L[l];
		const Dims &d = L.d;
		Tiles T = L.tiles;
		T.bz = s_bz[l];
		const long long nx = d.nx, ny = d.ny, plane = d.plane;
		const int ntiles = s_ntiles[l];
		const int per_tile = (T.bz >> 1) * (TY >> 1);
		const long long items = (long long)ntiles * per_tile;
		for (long long it = gwarp; it < items; it += nwarps) {
			const int t = (int)(it / per_tile), rem = (int)(it - (long long)t * per_tile);
			int i0, j0, kb;
			tile_origin(T, T.ids[t], i0, j0, kb);
			const int k = kb + 2 * (rem / (TY >> 1)), J = (j0 >> 1) + rem % (TY >> 1), I = (i0 >> 1) + lane;
			if (k >= d.nzl || I >= dc.nx || J >= dc.ny) continue;
			float acc = 0.f;
#pragma unroll
			for (int dk = 0; dk < 2; ++dk)
#pragma unroll
				for (int dj = 0; dj < 2; ++dj)
#pragma unroll
					for (int di = 0; di < 2; ++di) {
						const int i = 2 * I + di, j = 2 * J + dj, kk = k + dk;
						if (i < d.nx && j < d.ny && kk < d.nzl) {
							const long long c = i + nx * (j + ny * kk);
							acc += residual7(L.wx[c], L.wx[c + 1], L.wy[c], L.wy[c + nx], L.wz[c], L.wz[c + plane], L.dd[c], __ldcg(L.b + c), __ldcg(L.x + c), __ldcg(L.x + c - 1),
							                 __ldcg(L.x + c + 1), __ldcg(L.x + c - nx), __ldcg(L.x + c + nx), __ldcg(L.x + c - plane), __ldcg(L.x + c + plane));
						}
					}
			__stcg(bc + (I + (long long)dc.nx * (J + (long long)dc.ny * (k >> 1))), acc);
		}
	};
	// x += P e on every cell of the active tiles (MGPostSweeps = 0 only: otherwise the first post-sweep folds it in)
	auto prolong = [&](int l, const float *ec, const Dims &dc) {
		const MidLevel &L = A.L[l];
		const Dims &d = L.d;
		Tiles T = L.tiles;
		T.bz = s_bz[l];
		const int ntiles = s_ntiles[l];
		const long long items = (long long)ntiles * T.bz * TY;
		for (long long it = gwarp; it < items; it += nwarps) {
			const int t = (int)(it / (T.bz * TY)), rem = (int)(it - (long long)t * (T.bz * TY));
			int i0, j0, kb;
			tile_origin(T, T.ids[t], i0, j0, kb);
			const int k = kb + rem / TY, j = j0 + rem % TY;
			if (k >= d.nzl || j >= d.ny) continue;
			for (int e = 0; e < 2; ++e) {
				const int i = i0 + 2 * lane + e;
				if (i >= d.nx) continue;
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				__stcg(L.x + c, __ldcg(L.x + c) + __ldcg(ec + ((i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1)))));
			}
		}
	};

	// ---- descend
	for (int l = 0; l < A.nlev; ++l) {
		const bool last = l + 1 == A.nlev && !A.has_tail;
		const int sweeps = last ? A.coarse : A.pre;
		for (int sw = 0; sw < sweeps; ++sw) {
			relax(l, 0, sw == 0 ? MID_ZERO_A : MID_PLAIN, nullptr, A.L[l].d);
			grid_barrier(A.barrier, gen);
			relax(l, 1, sw == 0 ? MID_ZERO_B : MID_PLAIN, nullptr, A.L[l].d);
			grid_barrier(A.barrier, gen);
		}
		if (last) break;
		if (l + 1 < A.nlev) restrict_to(l, A.L[l + 1].b, A.L[l + 1].d);
		else restrict_to(l, const_cast<float *>(TA.b_in), TA.L[0].d);
		grid_barrier(A.barrier, gen);
	}
	// ---- the shared-memory tail on CTA 0 (the others wait at the barrier)
	if (A.has_tail) {
		if (blockIdx.x == 0) tail_body(TA, sm);
		grid_barrier(A.barrier, gen);
	}
	// ---- ascend
	for (int l = A.nlev - 1; l >= 0; --l) {
		const bool last = l + 1 == A.nlev && !A.has_tail;
		const int sweeps = last ? A.coarse : A.post;
		const float *ec = nullptr;
		Dims dc = A.L[l].d;
		if (!last) {
			if (l + 1 < A.nlev) { ec = A.L[l + 1].x; dc = A.L[l + 1].d; }
			else { ec = TA.x_out; dc = TA.L[0].d; }
			if (sweeps == 0) {
				prolong(l, ec, dc);
				grid_barrier(A.barrier, gen);
			}
		}
		for (int sw = 0; sw < sweeps; ++sw) {
			const bool pro = !last && sw == 0;
			relax(l, 1, pro ? MID_PROLONG_A : MID_PLAIN, ec, dc);
			grid_barrier(A.barrier, gen);
			relax(l, 0, pro ? MID_PROLONG_B : MID_PLAIN, ec, dc);
			grid_barrier(A.barrier, gen);
		}
	}
	if (A.dot) { // level 0 of the solve: rho' = z.b0, beta (pcg_solver.h:286-288)
		const MidLevel &L = A.L[0];
		const Dims &d = L.d;
		Tiles T = L.tiles;
		T.bz = s_bz[0];
		const int ntiles = s_ntiles[0];
		const long long items = (long long)ntiles * T.bz * TY;
		double red[1] = {0.0};
		for (long long it = gwarp; it < items; it += nwarps) {
			const int t = (int)(it / (T.bz * TY)), rem = (int)(it - (long long)t * (T.bz * TY));
			int i0, j0, kb;
			tile_origin(T, T.ids[t], i0, j0, kb);
			const int k = kb + rem / TY, j = j0 + rem % TY;
			if (k >= d.nzl || j >= d.ny) continue;
			for (int e = 0; e < 2; ++e) {
				const int i = i0 + 2 * lane + e;
				if (i >= d.nx) continue;
				const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
				red[0] += (double)__ldcg(L.x + c) * (double)__ldcg(L.b + c);
			}
		}
		grid_reduce<1, 0u>(red, rb, [&](double (&tot)[1]) {
			const double zr = tot[0];
			st->beta = st->iter == 0 ? 0.0 : zr / st->rho;
			st->rho = zr;
			if (zr == 0.0 || zr != zr) st->done = 1;
		});
	}
}

// ---- hierarchy setup -----------------------------------------------------------------------------------------
// Coarse operator = scale * P^T A P (once per projection): coarse face coupling = sum of the four fine
// couplings crossing the coarse face, coarse dd = sum of the children's dd. Flags the coarse tile of every
// coarse cell that carries an equation.
// Persistent over the UNION list of the fine level (tiles that hold an unknown now or held one in the previous projection, `U`): the coarse cells under
// every other fine tile were zero before and stay zero. Block (TX/2, TY/2): one coarse column of the tile's footprint per thread.
__global__ void __launch_bounds__((TX / 2) * (TY / 2)) k_coarsen_operator(Dims df, Dims dc, Tiles U, Tiles Tc, float scale, const float *__restrict__ wx, const float *__restrict__ wy,
                                                                         const float *__restrict__ wz, const float *__restrict__ dd, float *__restrict__ cwx,
                                                                         float *__restrict__ cwy, float *__restrict__ cwz, float *__restrict__ cdd,
                                                                         unsigned char *__restrict__ tile_flags) {
	resolve_tiles(U);
	const int ntiles = *U.count;
	int i0, j0, kb, ke;
	for (TileWalk w(U, ntiles); w.next(U, df.nzl, i0, j0, kb, ke);) {
		const int I = (i0 >> 1) + threadIdx.x, J = (j0 >> 1) + threadIdx.y;
		if (I >= dc.nx || J >= dc.ny) continue;
		for (int k0 = kb; k0 < ke; k0 += 2) {
			const int K = k0 >> 1;
			float sx = 0.f, sy = 0.f, sz = 0.f, sd = 0.f;
			bool live = false;
#pragma unroll
			for (int dk = 0; dk < 2; ++dk)
#pragma unroll
				for (int dj = 0; dj < 2; ++dj)
#pragma unroll
					for (int di = 0; di < 2; ++di) {
						const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
						if (i >= df.nx || j >= df.ny || k >= df.nzl) continue;
						const long long c = i + (long long)df.nx * (j + (long long)df.ny * k);
						const float a = wx[c], bq = wy[c], cq = wz[c], e = dd[c];
						sd += e;
						if (!di) sx += a;
						if (!dj) sy += bq;
						if (!dk) sz += cq;
						// a child with an equation has a positive diagonal: its own lower faces, its upper faces, or dd
						live = live || a > 0.f || bq > 0.f || cq > 0.f || e > 0.f || wx[c + 1] > 0.f || wy[c + df.nx] > 0.f || wz[c + df.plane] > 0.f;
					}
			const long long C = I + (long long)dc.nx * (J + (long long)dc.ny * K);
			cwx[C] = scale * sx;
			cwy[C] = scale * sy;
			cwz[C] = scale * sz;
			cdd[C] = scale * sd;
			if (live) tile_flags[slice_of(Tc, I, J, K)] = 1;
		}
	}
}

// flag the tiles of a level that hold at least one cell with an equation (positive diagonal)
__global__ void __launch_bounds__(256) k_flag_live_tiles(Dims d, Tiles T, const float *__restrict__ wx, const float *__restrict__ wy, const float *__restrict__ wz,
                                                        const float *__restrict__ dd, unsigned char *__restrict__ tile_flags) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = blockIdx.z;
	if (i >= d.nx || j >= d.ny) return;
	const long long c = i + (long long)d.nx * (j + (long long)d.ny * k);
	if (gs_diag(wx[c], wx[c + 1], wy[c], wy[c + d.nx], wz[c], wz[c + d.plane], dd[c]) > 0.f) tile_flags[slice_of(T, i, j, k)] = 1;
}

// Slice flags -> the projection's tile lists (single CTA; a level has at most a few 10^4 slices).
//   * tile depth: `bz_fixed` != 0, or chosen here among bz_max, bz_max / 2, ... down to the slice depth. A persistent sweep CTA spends (bz + 2) plane
//     steps per tile and the grid of `sweep_grid` CTAs needs ceil(tiles / grid) rounds, so the depth that minimises (bz + 2) * rounds wins (ties: the
//     deeper one): deep tiles for a grid full of unknowns (fewest halo planes), shallower ones when a liquid scene leaves only a few tiles per CTA
//     — they fill the waves better and hug the free surface more tightly. A level whose every slice is active keeps bz_max.
//   * ids / count: active tiles of that depth, ascending; count[1] = the depth (what Tiles::bz == 0 reads).
//   * uids / ucount: the UNION list — tiles with a slice flagged now or in the previous projection (`dirty`), i.e. every tile whose arrays may hold
//     something other than zeros: what the kernels that WRITE a level's arrays walk over. Afterwards `dirty` becomes the current flags.
__global__ void __launch_bounds__(1024) k_compact_tiles(const unsigned char *__restrict__ flags, unsigned char *__restrict__ dirty, int ntx, int nty, int nslices_z, int slice,
                                                       int bz_max, int bz_fixed, int sweep_grid, int *__restrict__ ids, int *__restrict__ count, int *__restrict__ uids,
                                                       int *__restrict__ ucount) {
	__shared__ int warp_sums[2][32];
	__shared__ int base[2];
	__shared__ int s_cnt[4], s_bz;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int per_plane = ntx * nty;
	// ---- the depth
	if (tid == 0) s_bz = bz_fixed ? bz_fixed : bz_max;
	if (tid < 4) s_cnt[tid] = 0;
	__syncthreads();
	if (!bz_fixed && bz_max > slice) {
		int ncand = 0;
		for (int bz = bz_max; bz >= slice && ncand < 4; bz >>= 1) ++ncand;
		for (int cnd = 0; cnd < ncand; ++cnd) {
			const int m = (bz_max >> cnd) / slice, ntz = (nslices_z + m - 1) / m;
			int mine = 0;
			for (int t = tid; t < per_plane * ntz; t += 1024) {
				const int tz = t / per_plane, xy = t - tz * per_plane;
				bool any = false;
				for (int q = 0; q < m && tz * m + q < nslices_z; ++q) any = any || flags[xy + per_plane * (tz * m + q)];
				mine += any ? 1 : 0;
			}
			for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
			if (lane == 0 && mine) atomicAdd(&s_cnt[cnd], mine);
		}
		__syncthreads();
		if (tid == 0) {
			const int finest = (bz_max >> (ncand - 1)) / slice; // slices per tile of the shallowest candidate
			const int total_finest = per_plane * ((nslices_z + finest - 1) / finest);
			int best = 0;
			if (s_cnt[ncand - 1] < total_finest) { // (a level that is active everywhere keeps the deepest tiles)
				long long best_cost = -1;
				for (int cnd = 0; cnd < ncand; ++cnd) {
					const int bz = bz_max >> cnd, g = s_cnt[cnd] < sweep_grid ? (s_cnt[cnd] > 0 ? s_cnt[cnd] : 1) : sweep_grid;
					const long long cost = (long long)(bz + 2) * ((s_cnt[cnd] + g - 1) / g);
					if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cnd; }
				}
			}
			s_bz = bz_max >> best;
		}
		__syncthreads();
	}
	const int bz = s_bz, m = bz / slice > 0 ? bz / slice : 1, ntz = (nslices_z + m - 1) / m, n = per_plane * ntz;
	if (tid < 2) base[tid] = 0;
	__syncthreads();
	// ---- the lists
	for (int start = 0; start < n; start += 1024) {
		const int idx = start + tid;
		int f = 0, u = 0;
		if (idx < n) {
			const int tz = idx / per_plane, xy = idx - tz * per_plane;
			for (int q = 0; q < m && tz * m + q < nslices_z; ++q) {
				const int sl = xy + per_plane * (tz * m + q);
				f |= flags[sl] ? 1 : 0;
				u |= (flags[sl] || dirty[sl]) ? 1 : 0;
			}
		}
		const unsigned bf = __ballot_sync(0xffffffffu, f), bu = __ballot_sync(0xffffffffu, u);
		const unsigned below = (1u << lane) - 1u;
		if (lane == 0) { warp_sums[0][wid] = __popc(bf); warp_sums[1][wid] = __popc(bu); }
		__syncthreads();
		int wf = 0, wu = 0;
		for (int w = 0; w < wid; ++w) { wf += warp_sums[0][w]; wu += warp_sums[1][w]; }
		if (f) ids[base[0] + wf + __popc(bf & below)] = idx;
		if (u) uids[base[1] + wu + __popc(bu & below)] = idx;
		__syncthreads();
		if (tid < 2) {
			int tot = 0;
			for (int w = 0; w < 32; ++w) tot += warp_sums[tid][w];
			base[tid] += tot;
		}
		__syncthreads();
	}
	for (int sl = tid; sl < per_plane * nslices_z; sl += 1024) dirty[sl] = flags[sl];
	if (tid == 0) { count[0] = base[0]; count[1] = bz; ucount[0] = base[1]; ucount[1] = bz; }
}

} // namespace shkz
