// The red-black sweep of kernels_mg.cuh with its operands staged through shared memory by the Tensor Memory
// Accelerator (sm_100a): a three-deep ring of plane stages, each filled by six 3-D tiled bulk-tensor loads
// (cp.async.bulk.tensor -> UTMALDG) that one elected thread issues two planes ahead and that complete on an mbarrier.
// The compute threads never wait on a global load: all their operands come from shared memory, the only global
// traffic they issue themselves is the float4 store of x_new (and the coarse correction of the PROLONG variant).
// Out-of-range box parts (tile halo beyond the grid, planes beyond the ghost planes) are zero-filled by the TMA
// unit, which is exactly what a wall needs (zero coefficient).
//
// Arithmetic, thread mapping and the half-updated-plane ring H are those of k_sweep4: results are bit-identical.
#pragma once
#include <cuda.h>
#include "kernels_mg.cuh"

namespace shkz {

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
	unsigned ok;
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
	    "selp.u32 %0, 1, 0, p;\n"
	    "}\n"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
	while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
	             "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
	             : "memory");
}

// ---- stage layout (floats); every array starts on a 128-byte boundary ------------------------------------------------
constexpr int ST_W = S4_PITCH;                 // 72 columns: grid columns i0-4 .. i0+67
constexpr int ST_XO_ROWS = TY + 4;             // rows j0-2 .. j0+17
constexpr int ST_WY_ROWS = TY + 3;             // rows j0-1 .. j0+17
constexpr int ST_ROWS = TY + 2;                // rows j0-1 .. j0+16
constexpr int pad32(int n) { return (n + 31) / 32 * 32; }
constexpr int ST_XO = 0;
constexpr int ST_WY = ST_XO + pad32(ST_XO_ROWS * ST_W);
constexpr int ST_WX = ST_WY + pad32(ST_WY_ROWS * ST_W);
constexpr int ST_WZ = ST_WX + pad32(ST_ROWS * ST_W);
constexpr int ST_DD = ST_WZ + pad32(ST_ROWS * ST_W);
constexpr int ST_B = ST_DD + pad32(ST_ROWS * ST_W);
constexpr int ST_FLOATS = ST_B + pad32(ST_ROWS * ST_W);
constexpr int ST_STAGES = 3;
constexpr unsigned ST_TX_BYTES = (unsigned)((ST_XO_ROWS + ST_WY_ROWS + 4 * ST_ROWS) * ST_W * sizeof(float)); // bytes one stage's six boxes deliver
constexpr unsigned ST_TX_BYTES_NOX = (unsigned)((ST_WY_ROWS + 4 * ST_ROWS) * ST_W * sizeof(float));          // ... without x_old (ZERO_X)
constexpr int H_FLOATS = 3 * S4_ROWS * S4_PITCH;
constexpr size_t SWEEP_TMA_SMEM = (size_t)(ST_STAGES * ST_FLOATS + H_FLOATS) * sizeof(float) + 128;

struct SweepMaps { // tensor maps of one level's arrays, box widths as documented above
	CUtensorMap wx, wy, wz, dd, b, xo;
};

template <int FIRST, bool ZERO_X, bool PROLONG, bool DOT, bool SLAB>
__global__ void __launch_bounds__(S4_THREADS, 2) k_sweep_tma(Dims d, Tiles T, const __grid_constant__ SweepMaps M, const float *__restrict__ xo, float *__restrict__ xn,
                                                            const float *__restrict__ ec, Dims dc, const SlabSweep sl, RedBuf rb, CGState *st) {
	if (st && st->done) return;
	constexpr bool slab_ghosts = SLAB; // z-slab solver: this launch also carries the sweep's halo traffic (SlabSweep, kernels_mg.cuh)
	extern __shared__ __align__(128) float smem[];
	float *stage_base = smem;
	float(*H)[S4_ROWS][S4_PITCH] = reinterpret_cast<float(*)[S4_ROWS][S4_PITCH]>(smem + ST_STAGES * ST_FLOATS);
	unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + ST_STAGES * ST_FLOATS + H_FLOATS);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool colwarp = warp == S4_ROW_WARPS;
	const bool producer = tid == S4_THREADS - 1; // last lane of the halo-column warp
	const int tx = lane & 15, half = lane >> 4;
	const int r = warp < S4_ROW_WARPS - 1 ? ((warp >> 1) * 4 + (warp & 1) + 2 * half) : TY + half;
	resolve_tiles(T);
	const int ntiles = *T.count;
	const long long nx = d.nx, ny = d.ny, plane = d.plane;
	const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
	double red[1] = {0.0};

	if (tid == 0) {
		for (int sidx = 0; sidx < ST_STAGES; ++sidx) mbar_init(&full[sidx], 1);
		fence_barrier_init();
	}
	__syncthreads();

	auto EC = [&](int i, int j, int k) -> long long { return (i >> 1) + (long long)dc.nx * ((j >> 1) + (long long)dc.ny * (k >> 1)); };
	unsigned loads_done = 0; // stage loads this CTA has consumed so far (same value in every thread)

	float *push_lo = nullptr, *push_hi = nullptr; // neighbours' ghost planes of x_new that take the own boundary planes
	if constexpr (SLAB) {
		if (sl.wait_in) block_wait_neighbours(sl.cm, sl.wait_in);
		slab_push_half_planes<FIRST, ZERO_X, PROLONG>(d, sl, xo, ec, dc);
		if (sl.seq_out) {
			if (sl.cm->lo) push_lo = reinterpret_cast<float *>(sl.cm->lo + sl.off_xn) + (long long)(d.nzl + 1) * plane;
			if (sl.cm->hi) push_hi = reinterpret_cast<float *>(sl.cm->hi + sl.off_xn);
		}
	}

	for (int pass = 0; pass < (slab_ghosts ? 2 : 1); ++pass) {
	if (SLAB && pass == 1) block_wait_neighbours(sl.cm, sl.seq_half); // the neighbours' half-updated planes are in the ghost planes of x_old
	// The loads of a CTA form ONE stream over all its tiles (planes kb-1 .. ke+1 of the first, then of the next, ...): load n lands in stage n % 3, and the producer
	// requests a load as soon as the one three before it has been consumed. The halo-column warp, which holds the producer, looks one tile ahead, so while the
	// last two steps of a tile run, the first planes of the NEXT tile are already on their way (a tile of a liquid scene is 8-16 planes deep: the pipeline
	// fill was ~8 % of it). The row warps know nothing of this — they wait on the same mbarriers as before, which have simply completed earlier.
	TileWalk w(T, ntiles, false, false); // (whole tiles only: see TileWalk)
	auto fetch = [&](TileWalk &tw, int &ti0, int &tj0, int &tkb, int &tke) -> bool {
		while (tw.next(T, d.nzl, ti0, tj0, tkb, tke))
			if (!(slab_ghosts && (tkb < 2 || tke + 1 >= d.nzl) != (pass == 1))) return true; // pass 0: tiles that read no ghost plane; pass 1: the others
		return false;
	};
	int i0, j0, kb, ke;
	unsigned issued = loads_done; // (producer) loads requested so far
	while (fetch(w, i0, j0, kb, ke)) {
		const unsigned base = loads_done;           // load index of plane kb-1
		const unsigned cur_n = (unsigned)(ke - kb + 3);
		auto stage_of = [&](int p) -> const float * { return stage_base + ((base + (unsigned)(p - (kb - 1))) % ST_STAGES) * ST_FLOATS; };
		auto wait_plane = [&](int p) {
			const unsigned n = base + (unsigned)(p - (kb - 1));
			mbar_wait(&full[n % ST_STAGES], (n / ST_STAGES) & 1u);
		};
		// PROLONG: x_old + P e_coarse is formed IN the stage. As soon as the box of plane p+2 has landed (one step before anybody
		// needs it) every row-warp thread adds the coarse correction to its OWN quad of it — the only part of that plane it reads
		// before the next block barrier — and the halo-column warp does the rim (rows 0 and 19, quads 0 and 17 of the other rows),
		// which others read one barrier later. No extra barrier, no extra registers downstream: after this, the kernel reads
		// corrected values from shared memory exactly like the plain sweep. The float2 of e is requested a whole step ahead.
		// Cells outside the grid keep the TMA's zero fill (they only ever meet zero coefficients).
		auto ec_quad = [&](int rowi, int c4, int pz) -> float2 { // correction of x_old box quad (rowi, c4) in plane pz, 0 where there is none
			const int gi = i0 - 4 + 4 * c4, gj = j0 - 2 + rowi;
			return (gi >= 0 && gi < d.nx && gj >= 0 && gj < d.ny && pz >= -1 && pz <= d.nzl) ? *reinterpret_cast<const float2 *>(ec + EC(gi, gj, pz)) : make_float2(0.f, 0.f);
		};
		auto add_quad = [&](int pz, int rowi, int c4, float2 e) {
			float4 *cell = reinterpret_cast<float4 *>(stage_base + ((base + (unsigned)(pz - (kb - 1))) % ST_STAGES) * ST_FLOATS + ST_XO + rowi * ST_W + 4 * c4);
			float4 v = *cell;
			if (SLAB && (pz < 0 || pz >= d.nzl)) {
				// a ghost plane of a z-slab holds the neighbour's half-updated plane by the time a tile reads it: relaxed values on the first colour, the plain
				// x_old on the second (slab_push_half_planes) — only the latter still lacks the correction
				const int par0 = (j0 - 2 + rowi + pz + d.k0) & 1; // colour of the quad's cells 0 and 2 (its first column is a multiple of 4)
				if (par0 != FIRST) { v.x += e.x; v.z += e.y; }
				else { v.y += e.x; v.w += e.y; }
			} else {
				v.x += e.x; v.y += e.x; v.z += e.y; v.w += e.y;
			}
			*cell = v;
		};

		if (colwarp) {
			// ---- the producer's view of the load stream: this tile and the next
			TileWalk peek = w;
			int ni0 = 0, nj0 = 0, nkb = 0, nke = 0;
			const bool nhave = fetch(peek, ni0, nj0, nkb, nke);
			const unsigned next_n = nhave ? (unsigned)(nke - nkb + 3) : 0u;
			auto issue_at = [&](int ti0, int tj0, int p, unsigned n) { // fill stage n % 3 with plane p of the tile at (ti0, tj0)
				float *sp = stage_base + (n % ST_STAGES) * ST_FLOATS;
				unsigned long long *bar = &full[n % ST_STAGES];
				fence_proxy_async(); // the stage's previous contents were read through the generic proxy
				mbar_expect_tx(bar, ZERO_X ? ST_TX_BYTES_NOX : ST_TX_BYTES);
				if (!ZERO_X) tma_load_3d(sp + ST_XO, &M.xo, bar, ti0 - 4, tj0 - 2, p + 1);
				tma_load_3d(sp + ST_WY, &M.wy, bar, ti0 - 4, tj0 - 1, p + 1);
				tma_load_3d(sp + ST_WX, &M.wx, bar, ti0 - 4, tj0 - 1, p + 1);
				tma_load_3d(sp + ST_WZ, &M.wz, bar, ti0 - 4, tj0 - 1, p + 1);
				tma_load_3d(sp + ST_DD, &M.dd, bar, ti0 - 4, tj0 - 1, p + 1);
				tma_load_3d(sp + ST_B, &M.b, bar, ti0 - 4, tj0 - 1, p + 1);
			};
			// request every load whose stage is free: `consumed` loads of the stream have been read for the last time
			auto pump = [&](unsigned consumed) {
				while (issued < consumed + ST_STAGES && issued < base + cur_n + next_n) {
					const unsigned idx = issued - base;
					if (idx < cur_n) issue_at(i0, j0, kb - 1 + (int)idx, issued);
					else issue_at(ni0, nj0, nkb - 1 + (int)(idx - cur_n), issued);
					++issued;
				}
			};
			// after the block barrier of step p the stage of plane p is free (the last step also reads plane ke+1 for the last time)
			auto consumed_after = [&](int p) -> unsigned { return p == ke ? base + cur_n : base + (unsigned)(p - (kb - 1)) + 1u; };
			if (producer) pump(base);
			// ---- halo columns (stage columns 3 and TX+4) of the TY tile rows: one cell per lane and plane
			const int side = lane >> 4, cc = side ? TX + 4 : 3, rr = (lane & 15) + 1; // stage column, row slot
			const int ci = i0 - 4 + cc, cj = j0 - 1 + rr;
			const bool cv = ci >= 0 && ci < d.nx && cj < d.ny;
			float xm = 0.f, xc = 0.f;
			// ZERO_X on a z-slab: x_old is zero inside the slab, the ghost planes hold the neighbours' half-updated planes
			auto ghost1 = [&](int k) -> float { return (ZERO_X && slab_ghosts && cv && (k < 0 || k >= d.nzl) && k >= -1 && k <= d.nzl) ? __ldcg(xo + ci + nx * (cj + ny * k)) : 0.f; };
			if (!ZERO_X && cv && kb - 2 >= -1) {
				xm = xo[ci + nx * (cj + ny * (kb - 2))];
				if (PROLONG) xm += ec[EC(ci, cj, kb - 2)];
			}
			if (ZERO_X) xm = ghost1(kb - 2);
			// PROLONG: the rim of the x_old box, 72 quads over 32 lanes — rows 0 and 19 whole, quads 0 and 17 of rows 1..18
			auto rim = [&](int u, int &rowi, int &c4) -> bool {
				const int qn = lane + 32 * u;
				if (qn < 36) { rowi = qn < 18 ? 0 : ST_XO_ROWS - 1; c4 = qn % 18; }
				else { rowi = 1 + ((qn - 36) >> 1); c4 = (qn & 1) ? 17 : 0; }
				return qn < 72;
			};
			auto rim_load = [&](int pz, float2 (&e)[3]) {
#pragma unroll
				for (int u = 0; u < 3; ++u) { int rowi, c4; e[u] = rim(u, rowi, c4) ? ec_quad(rowi, c4, pz) : make_float2(0.f, 0.f); }
			};
			auto rim_add = [&](int pz, const float2 (&e)[3]) {
				wait_plane(pz);
#pragma unroll
				for (int u = 0; u < 3; ++u) { int rowi, c4; if (rim(u, rowi, c4)) add_quad(pz, rowi, c4, e[u]); }
				__syncwarp();
			};
			float2 e_rim[3];
			if (PROLONG) {
				rim_load(kb - 1, e_rim); rim_add(kb - 1, e_rim);
				__syncthreads(); // plane kb-1 is read by everybody right away
				rim_load(kb, e_rim); rim_add(kb, e_rim);
			} else wait_plane(kb - 1);
			if (!ZERO_X) xc = stage_of(kb - 1)[ST_XO + (rr + 1) * ST_W + cc];
			else xc = ghost1(kb - 1);
			for (int p = kb - 1; p <= ke; ++p) {
				const bool fix = PROLONG && p + 2 <= ke + 1;
				if (fix) rim_load(p + 2, e_rim);
				if (!PROLONG) wait_plane(p + 1);
				const float *P = stage_of(p), *N = stage_of(p + 1);
				const bool in_slab = p >= 0 && p < d.nzl;
				float xp = 0.f;
				if (!ZERO_X) xp = N[ST_XO + (rr + 1) * ST_W + cc];
				else xp = ghost1(p + 1);
				float h = xc;
				if (in_slab && ((ci + cj + p + d.k0) & 1) == FIRST) {
					const float w0 = P[ST_WX + rr * ST_W + cc], w1 = P[ST_WX + rr * ST_W + cc + 1], w2 = P[ST_WY + rr * ST_W + cc], w3 = P[ST_WY + (rr + 1) * ST_W + cc];
					const float w4 = P[ST_WZ + rr * ST_W + cc], w5 = N[ST_WZ + rr * ST_W + cc], dg = P[ST_DD + rr * ST_W + cc], bb = P[ST_B + rr * ST_W + cc];
					if (ZERO_X) h = gs_relax0(w0, w1, w2, w3, w4, w5, dg, bb);
					else {
						const float x0 = P[ST_XO + (rr + 1) * ST_W + cc - 1], x1 = P[ST_XO + (rr + 1) * ST_W + cc + 1], x2 = P[ST_XO + rr * ST_W + cc], x3 = P[ST_XO + (rr + 2) * ST_W + cc];
						h = gs_relax(w0, w1, w2, w3, w4, w5, dg, bb, x0, x1, x2, x3, xm, xp, xc);
					}
				}
				H[(p + 3) % 3][rr][cc] = cv ? h : 0.f;
				__syncthreads();
				if (producer) pump(consumed_after(p)); // the stage of plane p is free: every phase-1 read of it is behind the barrier
				if (fix) rim_add(p + 2, e_rim);
				xm = xc; xc = xp;
			}
		} else {

		const int i = i0 + 4 * tx, j = j0 - 1 + r;
		const bool valid = i < d.nx && j >= 0 && j < d.ny;
		const bool finish = valid && r >= 1 && r <= TY;
		const long long row = i + nx * j;
		const int q = 4 + 4 * tx; // stage / H column of the own quad
		auto LDQ = [&](const float *sp, int arr, int rowi) -> float4 { return *reinterpret_cast<const float4 *>(sp + arr + rowi * ST_W + q); };
		auto ECQ = [&](float4 v, int ii, int jj, int kk) -> float4 { // + coarse correction of an aligned quad
			const float2 e = *reinterpret_cast<const float2 *>(ec + EC(ii, jj, kk));
			v.x += e.x; v.y += e.x; v.z += e.y; v.w += e.y;
			return v;
		};
		float4 xm = zero4, xc = zero4, wz_cur = zero4;
		auto ghost4 = [&](int k) -> float4 { // ZERO_X on a z-slab: the ghost planes are the only non-zero part of x_old
			return (ZERO_X && slab_ghosts && valid && (k < 0 || k >= d.nzl) && k >= -1 && k <= d.nzl) ? ld4cg(xo + row + plane * k) : zero4;
		};
		if (!ZERO_X && valid && kb - 2 >= -1) {
			xm = ld4(xo + row + plane * (kb - 2));
			if (PROLONG) xm = ECQ(xm, i, j, kb - 2);
		}
		if (ZERO_X) xm = ghost4(kb - 2);
		float2 e_own = make_float2(0.f, 0.f); // PROLONG: correction of the own quad, box row r+1, quad tx+1
		if (PROLONG) {
			e_own = ec_quad(r + 1, tx + 1, kb - 1);
			wait_plane(kb - 1);
			add_quad(kb - 1, r + 1, tx + 1, e_own);
			__syncthreads(); // plane kb-1 is read by everybody right away
			e_own = ec_quad(r + 1, tx + 1, kb);
			wait_plane(kb);
			add_quad(kb, r + 1, tx + 1, e_own);
		} else wait_plane(kb - 1);
		{
			const float *P = stage_of(kb - 1);
			if (!ZERO_X) xc = LDQ(P, ST_XO, r + 1);
			else xc = ghost4(kb - 1);
			wz_cur = LDQ(P, ST_WZ, r);
		}
		float4 hm = zero4, hc = zero4;
		Quad prv;
		prv.wx = prv.wy = prv.wyu = prv.wz = prv.dd = prv.b = zero4;
		prv.wx4 = 0.f;
		for (int p = kb - 1; p <= ke; ++p) {
			const int slot = (p + 3) % 3;
			const bool in_slab = p >= 0 && p < d.nzl;
			const bool fix = PROLONG && p + 2 <= ke + 1;
			if (fix) e_own = ec_quad(r + 1, tx + 1, p + 2);
			if (!PROLONG) wait_plane(p + 1);
			const float *P = stage_of(p), *N = stage_of(p + 1);
			Quad cur;
			cur.wz = wz_cur;
			cur.wx = LDQ(P, ST_WX, r); cur.wx4 = P[ST_WX + r * ST_W + q + 4];
			cur.wy = LDQ(P, ST_WY, r); cur.wyu = LDQ(P, ST_WY, r + 1);
			cur.dd = LDQ(P, ST_DD, r); cur.b = LDQ(P, ST_B, r);
			const float4 wz_next = LDQ(N, ST_WZ, r);
			float4 xp = zero4, xd = zero4, xu = zero4;
			float xl = 0.f, xr = 0.f;
			if (!ZERO_X) {
				xp = LDQ(N, ST_XO, r + 1);
				xl = P[ST_XO + (r + 1) * ST_W + q - 1]; xr = P[ST_XO + (r + 1) * ST_W + q + 4];
				xd = LDQ(P, ST_XO, r); xu = LDQ(P, ST_XO, r + 2);
			} else xp = ghost4(p + 1);
			// ---- phase 1: half-updated plane p
			float4 hp = xc;
			if (in_slab) {
				const int a1 = (FIRST + j + p + d.k0) & 1;
				if (a1 == 0) hp = relax_quad<0, ZERO_X>(cur, wz_next, xc, xl, xr, xd, xu, xm, xp);
				else hp = relax_quad<1, ZERO_X>(cur, wz_next, xc, xl, xr, xd, xu, xm, xp);
			}
			if (!valid) hp = zero4;
			*reinterpret_cast<float4 *>(&H[slot][r][q]) = hp;
			__syncthreads();
			// ---- phase 2: finish plane k = p - 1
			const int k = p - 1;
			if (finish && k >= kb) {
				const int ks = (k + 3) % 3;
				const float hl = H[ks][r][q - 1], hr = H[ks][r][q + 4];
				const float4 hd = *reinterpret_cast<const float4 *>(&H[ks][r - 1][q]);
				const float4 hu = *reinterpret_cast<const float4 *>(&H[ks][r + 1][q]);
				const int a2 = (FIRST + 1 + j + k + d.k0) & 1;
				float4 xnew;
				if (a2 == 0) xnew = relax_quad<0, false>(prv, cur.wz, hc, hl, hr, hd, hu, hm, hp);
				else xnew = relax_quad<1, false>(prv, cur.wz, hc, hl, hr, hd, hu, hm, hp);
				*reinterpret_cast<float4 *>(xn + row + plane * k) = xnew;
				if (SLAB && k == 0 && push_lo) *reinterpret_cast<float4 *>(push_lo + row) = xnew;
				if (SLAB && k == d.nzl - 1 && push_hi) *reinterpret_cast<float4 *>(push_hi + row) = xnew;
				if (DOT) red[0] += (double)xnew.x * (double)prv.b.x + (double)xnew.y * (double)prv.b.y + (double)xnew.z * (double)prv.b.z + (double)xnew.w * (double)prv.b.w;
			}
			if (fix) { // the box of plane p+2 has had a whole step to land
				wait_plane(p + 2);
				add_quad(p + 2, r + 1, tx + 1, e_own);
			}
			hm = hc; hc = hp;
			xm = xc; xc = xp;
			wz_cur = wz_next;
			prv = cur;
		}
		}
		loads_done = base + cur_n;
		__syncthreads();
	}
	}
	if (slab_ghosts && sl.seq_out) signal_neighbours(sl.cm, sl.seq_out, HDR_PUSH_TICKET2);
	if (DOT) {
		grid_reduce<1, 0u>(red, rb, [&](double (&tot)[1]) {
			const double zr = tot[0];
			st->beta = st->iter == 0 ? 0.0 : zr / st->rho; // pcg_solver.h:286-288
			st->rho = zr;
			if (zr == 0.0 || zr != zr) st->done = 1;        // pcg_solver.h:263-271
		});
	}
}

} // namespace shkz
