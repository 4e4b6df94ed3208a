// Shared device-side definitions of libshkz_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slab_comm.cuh"

namespace shkz {

// Local z-slab of a global nx*ny*nzg cell grid: planes [k0, k0+nzl). Every internal cell array is
// allocated with one ghost plane below and above and addressed through a pointer to plane 0, so
// indices -plane .. ncell+plane-1 are valid. Walls and non-row cells are encoded as ZERO coefficients,
// which is why no kernel in the solve needs (i,j,k) bounds logic for its neighbour reads.
struct Dims {
	int nx, ny, nzl; // local extent
	int k0, nzg;     // first global plane, global z extent
	long long plane; // nx*ny
	long long ncell; // plane*nzl
};

// Device-resident CG control block: the host never needs a value from here inside the loop.
struct CGState {
	double rho;     // z.r of the current iteration
	double sz;      // s.(A s)
	double alpha, beta;
	double rnorm;   // |r|_inf
	double rr;      // r.r (plain CG) or z.r (MG) for the next rho
	double bnorm;   // |b|_inf
	double tol;     // residual * |b|_inf
	double sum_x;   // sum of x over rows (singular systems: mean removal)
	unsigned long long n_rows;
	int iter;       // completed iterations, counted like pcg_solver.h:282
	int max_iter;
	int done;       // 1: converged, max_iter reached, or trivial rhs — later kernels become no-ops
	int converged;
	int has_dirichlet;
	int pad;
};

// Work decomposition of every solve kernel: the cell grid of a level is cut into TX x TY x bz tiles and
// only tiles that hold at least one unknown are visited (compacted, ascending list built at assembly
// time on the device). Kernels are persistent: a fixed grid strides over the list, whose length is read
// from device memory, so launch geometry (and a captured CUDA graph) never depends on the scene.
constexpr int TX = 64, TY = 16;

// Tiles are flagged in SLICES of `slice` planes (8, or the level's tile depth where that is smaller); how many slices make a tile — the tile depth bz — is
// decided per projection ON THE DEVICE by the compaction kernel (k_compact_tiles: deep tiles when the scene fills the grid, shallower ones when a liquid
// scene leaves a persistent grid with a handful of tiles per CTA), which stores it next to the count. bz != 0 here fixes the depth (z-slab levels, small levels).
struct Tiles {
	const int *ids;    // active tile ids, ascending:  id = tx + ntx * (ty + nty * tz)
	const int *count;  // [0] number of active tiles  [1] planes per tile chosen for this projection
	int ntx, nty;      // tile grid of the level in x and y
	int slice;         // planes per flag slice
	int bz;            // planes per tile (even); 0: read it from count[1]
	int balanced;      // how a persistent grid divides the list (TileWalk): bit 0 element-wise kernels, bit 1 stencil kernels may take the balanced walk; bit 2: hybrid walk
};
// every tile kernel starts with this: after it T.bz is the depth in force
__device__ __forceinline__ void resolve_tiles(Tiles &T) { if (T.bz == 0) T.bz = T.count[1]; }

__device__ __forceinline__ void tile_origin(const Tiles &T, int id, int &i0, int &j0, int &kb) {
	const int tx = id % T.ntx;
	const int r = id / T.ntx;
	i0 = tx * TX;
	j0 = (r % T.nty) * TY;
	kb = (r / T.nty) * T.bz;
}

// The work of one CTA of a persistent tile kernel, as a sequence of pieces (tile footprint i0, j0 and planes [kb, ke)).
//   strided    CTA b takes the tiles b, b+G, b+2G, ... of the list, each whole. The CTAs of a wave then work on NEIGHBOURING tiles at the same time: they share
//              halo lines in L2 and open the same DRAM pages together. Every stencil kernel (sweeps, residual, SpMV) walks this way — the even split below
//              made the 512^3 sweep 2x slower (measured): 296 far-apart streams per array defeat the DRAM row buffers.
//   balanced   the active tiles, cut into PAIRS of planes, form one long run that is divided evenly: CTA b takes the pairs [b*tot/G, (b+1)*tot/G), i.e. a few
//              pieces of up to a whole tile — for the element-wise kernels (xpay, axpy2, init, dots), which gain 10-20 % from it on liquid scenes whose tile
//              count is a small, odd multiple of the grid (Tiles::balanced; z-slab levels stay strided).
// Pieces start on even planes. Every thread of the CTA walks the same sequence.
struct TileWalk {
	int u, end; // balanced: next / last plane-pair unit of this CTA; strided: next tile index / number of tiles
	bool bal;
	int tail_u, tail_end; // hybrid: the plane-pair units of the last, partial round that follow the strided part
	// (with four or more tiles per CTA the strided walk loses little to the last partial round and keeps its locality: measured, all-fluid 512^3)
	//   hybrid     (whole-grid levels) the strided walk covers the WHOLE rounds only — the first floor(n / G) * G tiles —, and the tiles of the last, partial round
	//              are cut into plane pairs and divided evenly like the balanced walk does: a liquid scene at 512^3 has 1384 level-0 tiles for 296 CTAs, i.e. 4.68
	//              rounds that used to cost 5 (level 1: 173 tiles, a single round with 123 idle CTAs). Measured (B200, 512^3): k_residual_restrict -17 % (dam-break) / -18 % (FLIP)
	//              on level 0 and -56 % on level 1, k_xpay_spmv_tma -1..-7 %; k_sweep_tma is the exception (+3..5 % on liquid scenes: each piece pays its two halo planes
	//              and the fill of the three-stage pipeline) and passes hybrid = false.
	__device__ __forceinline__ TileWalk(const Tiles &T, int ntiles, bool elementwise = false, bool hybrid = true)
	    : bal((T.balanced & (elementwise ? 1 : 2)) && ntiles < 4 * (int)gridDim.x), tail_u(0), tail_end(0) {
		const int G = (int)gridDim.x, b = (int)blockIdx.x;
		if (bal) {
			const long long tot = (long long)ntiles * (T.bz >> 1);
			u = (int)(tot * b / G);
			end = (int)(tot * (b + 1) / G);
		} else {
			u = b;
			end = ntiles;
			if (hybrid && (T.balanced & 4)) {
				const int full = (ntiles / G) * G, U = T.bz >> 1;
				const long long tot = (long long)(ntiles - full) * U, first = (long long)full * U;
				if (tot > 0) {
					end = full;
					tail_u = (int)(first + tot * b / G);
					tail_end = (int)(first + tot * (b + 1) / G);
				}
			}
		}
	}
	__device__ __forceinline__ bool next(const Tiles &T, int nzl, int &i0, int &j0, int &kb, int &ke) {
		if (!bal && u >= end && tail_u < tail_end) { // the strided rounds are done: on to this CTA's share of the partial round
			bal = true;
			u = tail_u;
			end = tail_end;
		}
		while (u < end) {
			if (!bal) {
				tile_origin(T, T.ids[u], i0, j0, kb);
				ke = min(kb + T.bz, nzl);
				u += (int)gridDim.x;
				return true;
			}
			const int U = T.bz >> 1, t = u / U, o = u - t * U, len = min(U - o, end - u);
			tile_origin(T, T.ids[t], i0, j0, kb);
			const int kt = min(kb + T.bz, nzl);
			kb += 2 * o;
			ke = min(kb + 2 * len, kt);
			u += len;
			if (kb < ke) return true;
		}
		return false;
	}
};

// flag slot (slice) of a cell
__device__ __forceinline__ int slice_of(const Tiles &T, int i, int j, int k) { return (i / TX) + T.ntx * ((j / TY) + T.nty * (k / T.slice)); }

struct RedBuf {
	double *partials;      // [blocks][N]
	unsigned int *counter; // zero between kernels
	const CommDev *comm;   // z-slab solvers: the result is folded over all ranks (nullptr on a whole grid)
};

__device__ __forceinline__ unsigned linear_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
__device__ __forceinline__ unsigned linear_bid() { return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); }

template <int N, unsigned MAXMASK>
__device__ __forceinline__ void warp_combine(double (&v)[N]) {
#pragma unroll
	for (int n = 0; n < N; ++n) {
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			double o = __shfl_xor_sync(0xffffffffu, v[n], off);
			v[n] = ((MAXMASK >> n) & 1u) ? fmax(v[n], o) : v[n] + o;
		}
	}
}

template <int N, unsigned MAXMASK>
__device__ __forceinline__ void block_combine(double (&v)[N], double (*sm)[32]) {
	const unsigned tid = linear_tid(), lane = tid & 31u, wid = tid >> 5;
	const unsigned nwarps = (blockDim.x * blockDim.y * blockDim.z + 31u) >> 5;
	warp_combine<N, MAXMASK>(v);
	__syncthreads(); // sm may still be read from a previous use
	if (lane == 0) {
#pragma unroll
		for (int n = 0; n < N; ++n) sm[n][wid] = v[n];
	}
	__syncthreads();
	if (wid == 0) {
#pragma unroll
		for (int n = 0; n < N; ++n) v[n] = lane < nwarps ? sm[n][lane] : 0.0;
		warp_combine<N, MAXMASK>(v);
	}
}

// Grid-wide reduction with a deterministic combine order: every block stores its partial, the block
// that arrives last folds all partials in a fixed order and runs `fin(total)` on one thread. On a z-slab
// solver that thread first exchanges the rank totals with the other GPUs through peer memory
// (slab_comm.cuh), so the reduction kernel IS the all-reduce: no separate collective is launched.
// All values reduced here are sums, or maxima of non-negative numbers (identity 0 for both).
template <int N, unsigned MAXMASK, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[N], RedBuf rb, Fin fin) {
	__shared__ double sm[N][32];
	__shared__ int is_last;
	const unsigned tid = linear_tid();
	const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
	const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
	block_combine<N, MAXMASK>(v, sm);
	if (tid == 0) {
		const unsigned bid = linear_bid();
#pragma unroll
		for (int n = 0; n < N; ++n) __stcg(&rb.partials[(size_t)bid * N + n], v[n]);
		__threadfence();
		const unsigned ticket = atomicAdd(rb.counter, 1u);
		is_last = (ticket == nblocks - 1);
	}
	__syncthreads();
	if (!is_last) return;
	__threadfence();
	double acc[N];
#pragma unroll
	for (int n = 0; n < N; ++n) acc[n] = 0.0;
	for (unsigned b = tid; b < nblocks; b += nthreads) {
#pragma unroll
		for (int n = 0; n < N; ++n) {
			const double p = __ldcg(&rb.partials[(size_t)b * N + n]);
			acc[n] = ((MAXMASK >> n) & 1u) ? fmax(acc[n], p) : acc[n] + p;
		}
	}
	block_combine<N, MAXMASK>(acc, sm);
	if (tid == 0) {
		if (rb.comm) cross_rank_combine<N, MAXMASK>(acc, rb.comm); // peer-memory all-reduce, same bits on every rank
		fin(acc);
		*rb.counter = 0u;
	}
}

} // namespace shkz
