// Device side of the z-slab communicator: halo planes and CG scalars move between the GPUs of one NVLink domain
// by plain loads / stores on peer-mapped memory inside our own kernels (no NCCL call anywhere in the solve).
//
// Every slab solver carves all of its ghosted cell arrays from ONE device allocation (the "arena"), exported
// once through CUDA IPC (or used directly when the slabs live in one process). All ranks run the same allocation
// sequence on slabs of equal size, so an array sits at the same offset in every arena and a neighbour's ghost plane
// is just (mapped arena base + offset). The first ARENA_HEADER bytes hold the synchronisation words:
//   flag_from_lo / flag_from_hi   number of the last halo exchange whose planes the lower / upper neighbour has
//                                 stored here (written remotely, release order: planes, system fence, flag)
//   red_seq                       count of cross-rank reductions this rank has taken part in
//   mail[4 slots][8 ranks][8]     all-to-all mailbox of the reductions: every rank stores its partial sums into
//                                 the slot of every rank (itself included), then each rank folds the world's
//                                 partials in rank order, so all ranks obtain bit-identical results
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slab_comm.h"

namespace shkz {

// ---- bounded spin: every device-side wait of the communicator goes through here --------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
// Spin until *f >= seq. Returns false — after storing the abort word into EVERY rank's arena — when that takes longer than cm->timeout_ns, or at once
// when another rank (or a host) has already aborted. A failed rank therefore costs the others one timeout, not a hung GPU: the kernels run to completion
// on whatever data is there, the host finds the abort word after its stream synchronise and the call ends with SHKZ_B200_ERR_COMM.
static __device__ __noinline__ bool spin_until(volatile unsigned long long *f, unsigned long long seq, const CommDev *cm, int what) {
	volatile unsigned long long *abort_word = reinterpret_cast<volatile unsigned long long *>(cm->self + HDR_ABORT);
	unsigned long long t0 = 0;
	unsigned polls = 0;
	while (*f < seq) {
		if ((++polls & 255u) != 0) continue;
		if (*abort_word) return false;
		const unsigned long long now = global_ns();
		if (!t0) t0 = now;
		else if (now - t0 > cm->timeout_ns) {
			const unsigned long long w = (unsigned long long)(cm->rank + 1) | ((unsigned long long)what << 8) | (seq << 16);
			for (int p = 0; p < cm->world; ++p) *reinterpret_cast<volatile unsigned long long *>(cm->peer[p] + HDR_ABORT) = w;
			__threadfence_system();
			return false;
		}
	}
	return true;
}

// Fold the world's partial results (sums, or maxima where MAXMASK has the bit set) — called by ONE thread per rank.
template <int N, unsigned MAXMASK>
__device__ __forceinline__ void cross_rank_combine(double (&v)[N], const CommDev *cm) {
	static_assert(N <= 7, "mailbox entries hold seven values and a sequence word");
	unsigned long long *seqp = reinterpret_cast<unsigned long long *>(cm->self + HDR_RED_SEQ);
	const unsigned long long seq = *seqp + 1ull;
	*seqp = seq;
	const size_t entry = ((size_t)(seq & 3ull) * COMM_MAX_WORLD + (size_t)cm->rank) * 8;
	for (int p = 0; p < cm->world; ++p) {
		volatile double *m = reinterpret_cast<volatile double *>(cm->peer[p] + HDR_MAIL) + entry;
#pragma unroll
		for (int n = 0; n < N; ++n) m[n] = v[n];
	}
	__threadfence_system();
	for (int p = 0; p < cm->world; ++p)
		*(reinterpret_cast<volatile unsigned long long *>(cm->peer[p] + HDR_MAIL) + entry + 7) = seq;
	double tot[N];
#pragma unroll
	for (int n = 0; n < N; ++n) tot[n] = 0.0;
	for (int q = 0; q < cm->world; ++q) {
		const size_t e = ((size_t)(seq & 3ull) * COMM_MAX_WORLD + (size_t)q) * 8;
		volatile unsigned long long *f = reinterpret_cast<volatile unsigned long long *>(cm->self + HDR_MAIL) + e + 7;
		if (!spin_until(f, seq, cm, ABORT_WAIT_MAIL)) break; // (aborted: the totals are garbage, the host call fails)
		__threadfence_system();
		const volatile double *m = reinterpret_cast<const volatile double *>(cm->self + HDR_MAIL) + e;
#pragma unroll
		for (int n = 0; n < N; ++n) {
			const double p = m[n];
			tot[n] = ((MAXMASK >> n) & 1u) ? fmax(tot[n], p) : tot[n] + p;
		}
	}
#pragma unroll
	for (int n = 0; n < N; ++n) v[n] = tot[n];
}

// Called by every thread of every block of a kernel that has stored planes into the neighbours' arenas: the block that
// finishes last publishes exchange number `seq` in the neighbours' flag words (release order: data, system fence, flag).
// A kernel that publishes two exchanges (the fused slab sweep) counts the second one on its own ticket word.
__device__ __forceinline__ void signal_neighbours(const CommDev *cm, unsigned long long seq, size_t ticket_word = HDR_PUSH_TICKET) {
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
		unsigned int *ticket = reinterpret_cast<unsigned int *>(cm->self + ticket_word);
		const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
		if (atomicAdd(ticket, 1u) == nblocks - 1) {
			*ticket = 0u;
			__threadfence_system();
			if (cm->lo) *reinterpret_cast<volatile unsigned long long *>(cm->lo + HDR_FLAG_FROM_HI) = seq;
			if (cm->hi) *reinterpret_cast<volatile unsigned long long *>(cm->hi + HDR_FLAG_FROM_LO) = seq;
		}
	}
}

// Block until both neighbours have published exchange number `seq` here (ONE thread of a kernel calls this last; the
// kernel then cannot complete, and the stream cannot move on, before the ghost planes are in place).
__device__ __forceinline__ void wait_neighbours(const CommDev *cm, unsigned long long seq) {
	if (cm->lo) spin_until(reinterpret_cast<volatile unsigned long long *>(cm->self + HDR_FLAG_FROM_LO), seq, cm, ABORT_WAIT_LO);
	if (cm->hi) spin_until(reinterpret_cast<volatile unsigned long long *>(cm->self + HDR_FLAG_FROM_HI), seq, cm, ABORT_WAIT_HI);
	__threadfence_system();
}

// A kernel that writes a cell array and sends its boundary planes along: `off` = arena offset of the array's lower ghost plane
// (cm == nullptr: whole grid, or nobody reads the ghost planes); the block that finishes last publishes exchange `seq`.
struct SlabPush {
	const CommDev *cm;
	size_t off;
	unsigned long long seq;
};
template <class T> __device__ __forceinline__ T *push_target_lo(const SlabPush &sp, long long plane, int nzl) { // lower neighbour's upper ghost plane
	return (sp.cm && sp.cm->lo) ? reinterpret_cast<T *>(sp.cm->lo + sp.off) + (long long)(nzl + 1) * plane : nullptr;
}
template <class T> __device__ __forceinline__ T *push_target_hi(const SlabPush &sp) { // upper neighbour's lower ghost plane
	return (sp.cm && sp.cm->hi) ? reinterpret_cast<T *>(sp.cm->hi + sp.off) : nullptr;
}

// A block of a kernel that reads ghost planes whose exchange `seq` was published by a kernel that did not wait for it
// (the fused slab sweep): one thread acquires the neighbours' flags, the barrier hands the ordering to the block.
__device__ __forceinline__ void block_wait_neighbours(const CommDev *cm, unsigned long long seq) {
	if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) wait_neighbours(cm, seq);
	__syncthreads();
}

// ... and the stand-alone form for consumers that have no such prologue
static __global__ void k_comm_wait(const CommDev *cm, unsigned long long seq) {
	if (threadIdx.x == 0) wait_neighbours(cm, seq);
}

// Halo exchange number `seq` of the cell array at arena offset `off` (offset of its lower ghost plane): store the own
// boundary planes into the neighbours' ghost planes, publish `seq` in their flag words, wait for their planes.
static __global__ void __launch_bounds__(256) k_halo_push(const CommDev *cm, size_t off, size_t plane_bytes, int nzl, unsigned long long seq) {
	const char *src_lo = cm->self + off + plane_bytes;                 // own plane 0
	const char *src_hi = cm->self + off + plane_bytes * (size_t)nzl;   // own plane nzl-1
	char *dst_lo = cm->lo ? cm->lo + off + plane_bytes * (size_t)(nzl + 1) : nullptr; // lower neighbour's upper ghost plane
	char *dst_hi = cm->hi ? cm->hi + off : nullptr;                                    // upper neighbour's lower ghost plane
	const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if ((plane_bytes & 15) == 0) {
		const size_t n16 = plane_bytes >> 4;
		for (size_t e = tid; e < n16; e += nth) {
			if (dst_lo) reinterpret_cast<uint4 *>(dst_lo)[e] = reinterpret_cast<const uint4 *>(src_lo)[e];
			if (dst_hi) reinterpret_cast<uint4 *>(dst_hi)[e] = reinterpret_cast<const uint4 *>(src_hi)[e];
		}
	} else {
		for (size_t e = tid; e < plane_bytes; e += nth) {
			if (dst_lo) dst_lo[e] = src_lo[e];
			if (dst_hi) dst_hi[e] = src_hi[e];
		}
	}
	signal_neighbours(cm, seq);
	if (blockIdx.x == 0 && threadIdx.x == 0) wait_neighbours(cm, seq); // ... and the kernel ends when theirs have arrived
}

// All-gather by stores: `bytes` at arena offset src_off of this rank go to arena offset dst_off of EVERY rank (the offsets
// already include this rank's position in the gathered array). The last block to finish runs a barrier over all ranks
// (a mailbox reduction of nothing), so when the kernel ends on a rank, every rank's part has landed there.
static __global__ void __launch_bounds__(256) k_gather_push(const CommDev *cm, size_t src_off, size_t dst_off, size_t bytes) {
	const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	const char *src = cm->self + src_off;
	if ((bytes & 15) == 0) {
		const size_t n16 = bytes >> 4;
		for (size_t e = tid; e < n16; e += nth) {
			const uint4 v = reinterpret_cast<const uint4 *>(src)[e];
			for (int p = 0; p < cm->world; ++p) reinterpret_cast<uint4 *>(cm->peer[p] + dst_off)[e] = v;
		}
	} else {
		for (size_t e = tid; e < bytes; e += nth)
			for (int p = 0; p < cm->world; ++p) cm->peer[p][dst_off + e] = src[e];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int *ticket = reinterpret_cast<unsigned int *>(cm->self + HDR_PUSH_TICKET);
		if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
			*ticket = 0u;
			__threadfence_system();
			double nothing[1] = {0.0};
			cross_rank_combine<1, 0u>(nothing, cm);
		}
	}
}

} // namespace shkz
