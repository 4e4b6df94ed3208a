// Host <-> device traffic of shkz_b200_project_host on LIQUID scenes, done by kernels instead of copy engines.
//
// A projection only ever looks at the velocity and the solid level set next to a wet cell (fluid < 0): the assembly reads the six faces and the
// eight nodes of wet cells only (assemble_cell), the update changes input-ACTIVE faces only (macpressuresolver3.cpp:252-268), and the pressure is
// zero off the row set. A dam-break at 512^3 has 13 % of its cells wet — yet whole-array copies move 3.1 GB in and 2.7 GB out over PCIe (103 ms
// for 16 ms of GPU work). With the caller's buffers page-locked (shkz_b200_host_alloc, what Array=b200array3 hands over) the GPU can address
// them directly, so after the liquid level set has arrived (one DMA) the kernels below
//   k_flag_wet_slices   mark the 64 x 16 x `slice` cell blocks that hold a wet cell (the assembly's own test, on its own unit grid),
//   k_pull_slices       read the faces / nodes of exactly those blocks out of host memory (coalesced loads over PCIe),
//   k_push_faces        store back every face that was active on input (value; and the mask where the projection switched it off),
// and k_store_pressure writes the caller's pressure grids itself, over the union tile list (tiles that hold or held unknowns). The activity
// masks travel whole on a second stream, behind the solve. Results on the host are byte-for-byte those of the whole-array path
// (tests/test_gpu_host_sparse.py).
#pragma once
#include "common.cuh"
#include "kernels_assemble.cuh"

namespace shkz {

struct XferGeom {
	int ntx, nty;  // blocks per plane in x and y (TX x TY footprints)
	int slice;     // planes per block
	int nslices_z;
};

// unit = one plane of one footprint (block (TX, 4): four cells per thread), the unit grid of k_build_system
template <class RealT>
__global__ void __launch_bounds__(TX * 4) k_flag_wet_slices(Dims d, XferGeom g, const RealT *__restrict__ fluid, int *__restrict__ flags, int *__restrict__ list,
                                                            int *__restrict__ count) {
	const int per_plane = g.ntx * g.nty;
	const long long units = (long long)per_plane * d.nzl;
	for (long long u = blockIdx.x; u < units; u += gridDim.x) {
		const int k = (int)(u / per_plane), r = (int)(u - (long long)k * per_plane);
		const int i = (r % g.ntx) * TX + threadIdx.x, j0 = (r / g.ntx) * TY;
		bool wet = false;
#pragma unroll
		for (int m = 0; m < TY / 4; ++m) {
			const int j = j0 + threadIdx.y + 4 * m;
			if (i < d.nx && j < d.ny) wet = wet || fluid[i + (long long)d.nx * (j + (long long)d.ny * k)] < (RealT)0;
		}
		if (__syncthreads_or(wet) && threadIdx.x == 0 && threadIdx.y == 0) {
			const int s = r + per_plane * (k / g.slice);
			if (atomicExch(&flags[s], 1) == 0) list[atomicAdd(count, 1)] = s;
		}
	}
}

// rows [j0, j0+nj) x planes [k0, k0+nk) x columns [i0, i0+ni) of a W x H x * array: one row per warp at a time, lanes along x.
// mask != nullptr (params.velocity_masked: the entries of inactive faces are unspecified): a face whose mask byte is 0 arrives as 0, the value an
// inactive face of the simulators' velocity grids reads as.
template <class T>
__device__ __forceinline__ void copy_box(const T *__restrict__ src, T *__restrict__ dst, long long W, long long H, int i0, int j0, int k0, int ni, int nj, int nk,
                                         const uint8_t *__restrict__ mask = nullptr) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
	for (int row = warp; row < nj * nk; row += nwarps) {
		const int jj = row % nj, kk = row / nj;
		const long long base = i0 + W * ((j0 + jj) + H * (long long)(k0 + kk));
		T v[3];
		uint8_t m[3] = {1, 1, 1};
#pragma unroll
		for (int u = 0; u < 3; ++u) {
			const int c = lane + 32 * u;
			if (c < ni) {
				v[u] = src[base + c];
				if (mask) m[u] = mask[base + c];
			}
		}
#pragma unroll
		for (int u = 0; u < 3; ++u) {
			const int c = lane + 32 * u;
			if (c < ni) dst[base + c] = m[u] ? v[u] : (T)0;
		}
	}
}

// v = 0 where the mask byte is 0 (params.velocity_masked on whole arrays already on the device; `act` may also be a mapped host pointer)
template <class RealT>
__global__ void __launch_bounds__(256) k_zero_inactive(long long n, RealT *__restrict__ v, const uint8_t *__restrict__ act) {
	const bool vec_ok = (reinterpret_cast<unsigned long long>(act) & 3ull) == 0ull;
	for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < (n + 3) >> 2; q += (long long)gridDim.x * blockDim.x) {
		const long long f = q << 2;
		if (vec_ok && f + 3 < n) {
			const uchar4 m = *reinterpret_cast<const uchar4 *>(act + f);
			if (!m.x) v[f] = (RealT)0;
			if (!m.y) v[f + 1] = (RealT)0;
			if (!m.z) v[f + 2] = (RealT)0;
			if (!m.w) v[f + 3] = (RealT)0;
		} else
			for (long long e = f; e < n && e < f + 4; ++e)
				if (!act[e]) v[e] = (RealT)0;
	}
}

// the six faces and the eight nodes of every cell of the flagged blocks, host -> device staging (same dense layouts on both sides)
template <class RealT>
__global__ void __launch_bounds__(256) k_pull_slices(Dims d, XferGeom g, const int *__restrict__ list, const int *__restrict__ count, ConstFaceGrids<RealT> hvel,
                                                      FaceGrids<RealT> dvel, const RealT *__restrict__ hsolid, RealT *__restrict__ dsolid, FaceMasks hmask) {
	static_assert(TX + 1 <= 96, "copy_box moves at most 96 columns");
	const int n = *count;
	for (int u = blockIdx.x; u < n; u += gridDim.x) {
		const int s = list[u], tx = s % g.ntx, r = s / g.ntx;
		const int i0 = tx * TX, j0 = (r % g.nty) * TY, k0 = (r / g.nty) * g.slice;
		const int ni = min(TX, d.nx - i0), nj = min(TY, d.ny - j0), nk = min(g.slice, d.nzl - k0);
		copy_box(hvel.p[0], dvel.p[0], d.nx + 1, d.ny, i0, j0, k0, ni + 1, nj, nk, hmask.p[0]); // (hmask.p: the caller's masks, or null: values as they are)
		copy_box(hvel.p[1], dvel.p[1], d.nx, d.ny + 1, i0, j0, k0, ni, nj + 1, nk, hmask.p[1]);
		copy_box(hvel.p[2], dvel.p[2], d.nx, d.ny, i0, j0, k0, ni, nj, nk + 1, hmask.p[2]);
		if (hsolid) copy_box(hsolid, dsolid, d.nx + 1, d.ny + 1, i0, j0, k0, ni + 1, nj + 1, nk + 1);
	}
}

// Results back: the projection changes a face only if it was active on input (a face it switches off carries XFER_OFF_MARK in the staged mask
// instead of 0, see finish_face), so that is what travels — value for every such face, mask byte for the switched-off ones. Four faces per thread.
constexpr uint8_t XFER_OFF_MARK = 2;
template <class RealT>
__global__ void __launch_bounds__(256) k_push_faces(long long nf, const RealT *__restrict__ dvel, uint8_t *__restrict__ dact, RealT *__restrict__ hvel, uint8_t *__restrict__ hact,
                                                     unsigned long long *__restrict__ pushed) {
	unsigned n_val = 0, n_off = 0;
	const long long quads = (nf + 3) >> 2;
	const bool vec_ok = (reinterpret_cast<unsigned long long>(hvel) & 15ull) == 0ull, mask_vec_ok = (reinterpret_cast<unsigned long long>(hact) & 3ull) == 0ull; // (the caller's buffers may start anywhere)
	for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
		const long long f = q << 2;
		uint8_t m[4];
		if (f + 3 < nf) {
			const uchar4 mm = *reinterpret_cast<const uchar4 *>(dact + f);
			m[0] = mm.x; m[1] = mm.y; m[2] = mm.z; m[3] = mm.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) m[e] = f + e < nf ? dact[f + e] : (uint8_t)0;
		}
		if (!(m[0] | m[1] | m[2] | m[3])) continue;
		if (sizeof(RealT) == 4 && vec_ok && m[0] && m[1] && m[2] && m[3] && f + 3 < nf) {
			*reinterpret_cast<float4 *>(hvel + f) = *reinterpret_cast<const float4 *>(dvel + f);
			n_val += 4;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e)
				if (m[e]) { hvel[f + e] = dvel[f + e]; ++n_val; }
		}
		// masks: the four bytes in one store when any of them changed (single bytes over PCIe cost a transaction each: a scene whose active band
		// reaches far into dry or solid regions switches off 10^8 faces)
		if (m[0] == XFER_OFF_MARK || m[1] == XFER_OFF_MARK || m[2] == XFER_OFF_MARK || m[3] == XFER_OFF_MARK) {
			if (f + 3 < nf && mask_vec_ok) {
				const uchar4 now = make_uchar4(m[0] == 1, m[1] == 1, m[2] == 1, m[3] == 1);
				*reinterpret_cast<uchar4 *>(hact + f) = now;
				*reinterpret_cast<uchar4 *>(dact + f) = now;
				n_off += 4;
			} else {
#pragma unroll
				for (int e = 0; e < 4; ++e)
					if (f + e < nf && m[e] == XFER_OFF_MARK) { hact[f + e] = 0; dact[f + e] = 0; ++n_off; }
			}
		}
	}
	// bytes stored into host memory, for the stats (one atomic per warp)
	unsigned long long bytes = (unsigned long long)n_val * sizeof(RealT) + n_off;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, off);
	if ((threadIdx.x & 31) == 0 && bytes) atomicAdd(pushed, bytes);
}

template <class T>
__global__ void __launch_bounds__(256) k_fill_zero(T *__restrict__ p, long long n) {
	for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) p[c] = (T)0;
}

} // namespace shkz
