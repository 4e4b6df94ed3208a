// The conjugate-gradient product with its operands staged through shared memory by the Tensor Memory Accelerator (sm_100a), and with the
// direction update folded into the stage:
//
//      s_new = z + beta s_old          (pcg_solver.h:289)        formerly k_xpay:      reads z, s        writes s          20 B / unknown
//      q     = A s_new,  s_new . q     (pcg_solver.h:276-277)    formerly k_spmv_dot4: reads s, w, dd    writes q          32 B / unknown
//                                                                 here:                 reads z, s, w, dd writes s_new, q   44 B / unknown, one launch
//
// A CTA marches a TX x TY tile through its planes. A three-deep ring of plane stages is filled by six 3-D tiled bulk-tensor loads per plane
// (cp.async.bulk.tensor -> UTMALDG, completion on an mbarrier, issued two planes ahead by one elected thread): the z and s_old boxes carry the
// one-cell halo of the tile, so every thread forms s_new for its own quad AND the halo ring is formed redundantly by a ninth warp — no thread ever
// reads a direction value that another CTA is writing, because s_new goes to the OTHER buffer of a ping-pong pair (the host swaps the two after
// the launch). x / y neighbours of the product come from a shared plane of s_new, z neighbours ride in registers (the own quad of plane p+1 is
// formed one step ahead). Boxes are zero-filled outside the grid, which is what a wall needs (zero coefficient).
//
// Arithmetic per cell is that of k_xpay + k_spmv_dot4 (spmv_cell, same order); the reduction order differs, like between any two launch geometries.
#pragma once
#include <cuda.h>
#include "kernels_cg.cuh"
#include "kernels_mg_tma.cuh"

namespace shkz {

template <class VecT> struct SpmvStage {
	static constexpr int SX0 = sizeof(VecT) == 8 ? 2 : 4;     // box column of tile column 0 (box rows start on a 16-byte boundary)
	static constexpr int SW = TX + 2 * SX0;                    // s_old box: columns i0-SX0 .. i0+TX+SX0-1
	static constexpr int SROWS = TY + 2;                       // rows j0-1 .. j0+TY
	static constexpr int ZX0 = 4, ZW = TX + 8;                 // z box (float): columns i0-4 .. i0+TX+3
	static constexpr int WXW = TX + 4;                         // wx box: columns i0 .. i0+TX+3, rows j0 .. j0+TY-1
	static constexpr int WYROWS = TY + 1;                      // wy box: rows j0 .. j0+TY
	static constexpr int pad128(int bytes) { return (bytes + 127) / 128 * 128; }
	static constexpr int OFF_S = 0;
	static constexpr int OFF_Z = OFF_S + pad128(SW * SROWS * (int)sizeof(VecT));
	static constexpr int OFF_WX = OFF_Z + pad128(ZW * SROWS * 4);
	static constexpr int OFF_WY = OFF_WX + pad128(WXW * TY * 4);
	static constexpr int OFF_WZ = OFF_WY + pad128(TX * WYROWS * 4);
	static constexpr int OFF_DD = OFF_WZ + pad128(TX * TY * 4);
	static constexpr int BYTES = OFF_DD + pad128(TX * TY * 4);
	static constexpr unsigned TX_BYTES = (unsigned)(SW * SROWS * sizeof(VecT) + ZW * SROWS * 4 + WXW * TY * 4 + TX * WYROWS * 4 + 2 * TX * TY * 4);
	static constexpr int PLANE_W = TX + 2;                     // shared plane of s_new: columns i0-1 .. i0+TX, rows j0-1 .. j0+TY
	static constexpr int PLANE_BYTES = pad128(PLANE_W * SROWS * (int)sizeof(VecT));
	static constexpr size_t SMEM = 3 * (size_t)BYTES + PLANE_BYTES + 128;
};

struct SpmvMaps { // tensor maps of the operands, box shapes as in SpmvStage
	CUtensorMap s[2], z[2], wx, wy, wz, dd; // s: the two direction buffers; z: the two level-0 multigrid buffers (whichever holds the V-cycle's result)
};

constexpr int SPMV_TMA_THREADS = 32 * 9; // 8 warps own the tile's 16 x 16 quads, the ninth forms the halo ring of s_new

// SLAB (z-slab solvers): nothing of s ever crosses NVLink. The ghost planes of z arrive with the V-cycle's last sweep (its boundary planes of x_new go straight into
// the neighbours' ghost planes; this kernel waits for that exchange, `wait_in`, in its prologue), and every rank keeps the ghost planes of s itself:
// s_new(ghost) = z(ghost) + beta s_old(ghost) is the arithmetic the neighbour applies to the same bits on its own boundary plane, so the copies agree bit for bit —
// the tiles at the bottom / top of the slab store the value they form for plane -1 / nzl anyway. One exchange per iteration less than k_xpay (push) + k_spmv_dot4 (wait).
template <class VecT, bool SLAB>
__global__ void __launch_bounds__(SPMV_TMA_THREADS, 2) k_xpay_spmv_tma(Dims d, Tiles T, const __grid_constant__ SpmvMaps M, int s_in, int z_in, const VecT *__restrict__ s_old,
                                                                      const float *__restrict__ z, VecT *__restrict__ s_new, VecT *__restrict__ q, RedBuf rb, CGState *st,
                                                                      unsigned long long wait_in) {
	if (st->done) return;
	if (SLAB && wait_in) block_wait_neighbours(rb.comm, wait_in);
	using SS = SpmvStage<VecT>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	unsigned char *stage_base = smem_raw;
	VecT *plane_s = reinterpret_cast<VecT *>(smem_raw + 3 * SS::BYTES);
	unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + 3 * SS::BYTES + SS::PLANE_BYTES);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool ringwarp = warp == 8;
	const bool producer = tid == SPMV_TMA_THREADS - 1;
	const int tx = tid & 15, ty = (tid >> 4) & 15; // quad column / row of an owner thread
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx, plane = d.plane;
	const VecT beta = (VecT)st->beta;
	double red[1] = {0.0};
	if (tid == 0) {
		for (int n = 0; n < 3; ++n) mbar_init(&full[n], 1);
		fence_barrier_init();
	}
	__syncthreads();
	unsigned loads_done = 0;
	const CUtensorMap *map_s = &M.s[s_in], *map_z = &M.z[z_in];

	// The walk looks one tile ahead and the loads of all the CTA's tiles form one stream (load n -> stage n % 3, requested as soon as load n-3 has been
	// consumed), so the first planes of the next tile arrive while the last steps of this one run (k_sweep_tma, kernels_mg_tma.cuh, does the same).
	TileWalk w(T, ntiles);
	int i0, j0, kb, ke, ni0 = 0, nj0 = 0, nkb = 0, nke = 0;
	bool have = w.next(T, d.nzl, i0, j0, kb, ke);
	unsigned issued = 0; // (producer) loads requested so far
	while (have) {
		const bool nhave = w.next(T, d.nzl, ni0, nj0, nkb, nke);
		const unsigned base = loads_done; // load index of plane kb
		const unsigned cur_n = (unsigned)(ke - kb + 1), next_n = nhave ? (unsigned)(nke - nkb + 1) : 0u;
		auto stage_of = [&](int p) -> unsigned char * { return stage_base + ((base + (unsigned)(p - kb)) % 3) * SS::BYTES; };
		auto issue_at = [&](int ti0, int tj0, int p, unsigned n) {
			unsigned char *sp = stage_base + (n % 3) * SS::BYTES;
			unsigned long long *bar = &full[n % 3];
			fence_proxy_async();
			mbar_expect_tx(bar, SS::TX_BYTES);
			tma_load_3d(sp + SS::OFF_S, map_s, bar, ti0 - SS::SX0, tj0 - 1, p + 1);
			tma_load_3d(sp + SS::OFF_Z, map_z, bar, ti0 - SS::ZX0, tj0 - 1, p + 1);
			tma_load_3d(sp + SS::OFF_WX, &M.wx, bar, ti0, tj0, p + 1);
			tma_load_3d(sp + SS::OFF_WY, &M.wy, bar, ti0, tj0, p + 1);
			tma_load_3d(sp + SS::OFF_WZ, &M.wz, bar, ti0, tj0, p + 1);
			tma_load_3d(sp + SS::OFF_DD, &M.dd, bar, ti0, tj0, p + 1);
		};
		auto pump = [&](unsigned consumed) { // (producer) request every load whose stage is free
			while (issued < consumed + 3 && issued < base + cur_n + next_n) {
				const unsigned idx = issued - base;
				if (idx < cur_n) issue_at(i0, j0, kb + (int)idx, issued);
				else issue_at(ni0, nj0, nkb + (int)(idx - cur_n), issued);
				++issued;
			}
		};
		auto wait_plane = [&](int p) {
			const unsigned n = base + (unsigned)(p - kb);
			mbar_wait(&full[n % 3], (n / 3) & 1u);
		};
		// s_new of one cell of plane p from its stage: box row r (grid row j0-1+r), tile column c (grid column i0+c, -1 <= c <= TX)
		auto snew_at = [&](const unsigned char *sp, int r, int c) -> VecT {
			const VecT so = reinterpret_cast<const VecT *>(sp + SS::OFF_S)[r * SS::SW + SS::SX0 + c];
			const float zz = reinterpret_cast<const float *>(sp + SS::OFF_Z)[r * SS::ZW + SS::ZX0 + c];
			return fma(beta, so, (VecT)zz);
		};
		if (producer) pump(base); // planes kb .. ke (the last one is only read for the z neighbour above the tile)

		const int i = i0 + 4 * tx, j = j0 + ty;
		const bool valid = !ringwarp && i < d.nx && j < d.ny;
		const long long row = i + nx * j;
		// own quad of s_new in planes p-1 (sm), p (sc), p+1 (sp): registers
		V4<VecT> sm{0, 0, 0, 0}, sc{0, 0, 0, 0};
		if (valid) { // plane kb-1: straight from global memory, once per tile (z-slab ghost plane / wall: meets a zero coefficient where it is not a value)
			const long long c = row + plane * (kb - 1);
			const V4<VecT> so = ldv4(s_old + c);
			const float4 zz = *reinterpret_cast<const float4 *>(z + c);
			sm = V4<VecT>{fma(beta, so.a, (VecT)zz.x), fma(beta, so.b, (VecT)zz.y), fma(beta, so.c, (VecT)zz.z), fma(beta, so.d, (VecT)zz.w)};
			if (SLAB && kb == 0) stv4(s_new + c, sm); // the lower ghost plane of s, kept by this rank itself
		}
		wait_plane(kb);
		if (!ringwarp) {
			const unsigned char *sp = stage_of(kb);
			sc = V4<VecT>{snew_at(sp, ty + 1, 4 * tx), snew_at(sp, ty + 1, 4 * tx + 1), snew_at(sp, ty + 1, 4 * tx + 2), snew_at(sp, ty + 1, 4 * tx + 3)};
		}
		for (int p = kb; p < ke; ++p) {
			if (p + 1 == ke && nhave && !ringwarp) { // the next tile starts with direct loads of its plane kb-1: have them in L2 by then
				const int pi = ni0 + 4 * tx, pj = nj0 + ty;
				if (pi < d.nx && pj < d.ny) {
					const long long c = pi + nx * pj + plane * (nkb - 1);
					asm volatile("prefetch.global.L2 [%0];" ::"l"(s_old + c));
					asm volatile("prefetch.global.L2 [%0];" ::"l"(z + c));
				}
			}
			wait_plane(p + 1);
			const unsigned char *P = stage_of(p), *N = stage_of(p + 1);
			V4<VecT> sp4{0, 0, 0, 0};
			if (ringwarp) {
				// the halo ring of plane p: rows j0-1 and j0+TY (one quad per lane: 2 x 16 quads), columns i0-1 and i0+TX (one cell per lane: 2 x 16 cells)
				const int rr = lane < 16 ? 0 : SS::SROWS - 1, qc = 4 * (lane & 15);
#pragma unroll
				for (int e = 0; e < 4; ++e) plane_s[rr * SS::PLANE_W + 1 + qc + e] = snew_at(P, rr, qc + e);
				const int cr = 1 + (lane & 15), cc = lane < 16 ? -1 : TX;
				plane_s[cr * SS::PLANE_W + 1 + cc] = snew_at(P, cr, cc);
			} else {
				sp4 = V4<VecT>{snew_at(N, ty + 1, 4 * tx), snew_at(N, ty + 1, 4 * tx + 1), snew_at(N, ty + 1, 4 * tx + 2), snew_at(N, ty + 1, 4 * tx + 3)};
				VecT *dst = plane_s + (ty + 1) * SS::PLANE_W + 1 + 4 * tx;
				dst[0] = sc.a; dst[1] = sc.b; dst[2] = sc.c; dst[3] = sc.d;
			}
			__syncthreads();
			if (valid) {
				const long long c = row + plane * p;
				const float *WX = reinterpret_cast<const float *>(P + SS::OFF_WX) + ty * SS::WXW + 4 * tx;
				const float *WY = reinterpret_cast<const float *>(P + SS::OFF_WY) + ty * TX + 4 * tx;
				const float4 wzc = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(P + SS::OFF_WZ) + ty * TX + 4 * tx);
				const float4 wzp = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(N + SS::OFF_WZ) + ty * TX + 4 * tx);
				const float4 ddq = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(P + SS::OFF_DD) + ty * TX + 4 * tx);
				const float4 wxq = *reinterpret_cast<const float4 *>(WX), wyq = *reinterpret_cast<const float4 *>(WY), wyu = *reinterpret_cast<const float4 *>(WY + TX);
				const float wx4 = WX[4];
				const VecT *S0 = plane_s + (ty + 1) * SS::PLANE_W + 1 + 4 * tx;
				const VecT sl = S0[-1], sr = S0[4];
				const VecT *SD = S0 - SS::PLANE_W, *SU = S0 + SS::PLANE_W;
				V4<VecT> v;
				v.a = spmv_cell<VecT, float>(ddq.x, wxq.x, wxq.y, wyq.x, wyu.x, wzc.x, wzp.x, sc.a, sl, sc.b, SD[0], SU[0], sm.a, sp4.a);
				v.b = spmv_cell<VecT, float>(ddq.y, wxq.y, wxq.z, wyq.y, wyu.y, wzc.y, wzp.y, sc.b, sc.a, sc.c, SD[1], SU[1], sm.b, sp4.b);
				v.c = spmv_cell<VecT, float>(ddq.z, wxq.z, wxq.w, wyq.z, wyu.z, wzc.z, wzp.z, sc.c, sc.b, sc.d, SD[2], SU[2], sm.c, sp4.c);
				v.d = spmv_cell<VecT, float>(ddq.w, wxq.w, wx4, wyq.w, wyu.w, wzc.w, wzp.w, sc.d, sc.c, sr, SD[3], SU[3], sm.d, sp4.d);
				stv4(q + c, v);
				stv4(s_new + c, sc);
				red[0] += (double)sc.a * (double)v.a + (double)sc.b * (double)v.b + (double)sc.c * (double)v.c + (double)sc.d * (double)v.d;
			}
			sm = sc; sc = sp4;
			__syncthreads(); // stage p and the shared plane are free (after the last step: the stage of plane ke too)
			if (producer) pump(p + 1 == ke ? base + cur_n : base + (unsigned)(p - kb) + 1u);
		}
		if (SLAB && ke == d.nzl && valid) stv4(s_new + row + plane * ke, sc); // (sc now holds s_new of plane ke: the upper ghost plane)
		loads_done = base + cur_n;
		have = nhave;
		i0 = ni0; j0 = nj0; kb = nkb; ke = nke;
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) {
		st->sz = t[0];
		st->alpha = st->rho / t[0];
	});
}

} // namespace shkz
