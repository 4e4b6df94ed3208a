/*
 * b200dense.h — the message by which a Shiokaze module asks an array for direct access to its storage.
 *
 * array3<T>::const_send_message(B200_DENSE_MESSAGE, &descriptor) forwards to the array's core (include/shiokaze/array/array3.h:125-150); the
 * b200array3 core (`Array=b200array3`, plugin/b200array3.cpp) fills the descriptor and returns true, every stock core returns false.
 *   values   nx*ny*nz elements of element_bytes bytes, index i + nx*(j + ny*k) — the dense layout of include/shkz_b200.h. Entries of inactive cells
 *            are unspecified: a reader that wants array3::operator() semantics writes the background / fill value there first (b200pressure3 does)
 *   active   one byte per cell, 1 = active (array3::active()): the mask format of the C-ABI
 *   filled   one byte per cell, 1 = flood-filled; NULL when the array was never flood-filled
 *   pinned   1: values and active are page-locked (shkz_b200_host_alloc) — the C-ABI copies run at PCIe speed straight from / into them
 */
#ifndef SHKZ_B200_DENSE_H
#define SHKZ_B200_DENSE_H
#include <stdint.h>
#define B200_DENSE_MESSAGE "b200:dense"
typedef struct b200_dense_descriptor {
	unsigned nx, ny, nz, element_bytes;
	void *values;
	uint8_t *active;
	const uint8_t *filled;
	int pinned;
} b200_dense_descriptor;

#ifdef __cplusplus
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>
// std::fill over a large page-locked buffer, split over the host threads (one thread fills ~8 GB/s: five whole grids at 256^3 took tens of milliseconds)
template <class T> static void b200_parallel_fill( T *p, size_t count, T value ) {
	const unsigned nthreads = (unsigned)std::max<size_t>(1,std::min<size_t>((size_t)NUM_THREAD,count/(1u<<20)));
	if( nthreads <= 1 ) { std::fill(p,p+count,value); return; }
	std::vector<std::thread> pool;
	for( unsigned t=0; t<nthreads; ++t ) pool.emplace_back([=]() { std::fill(p+count*t/nthreads,p+count*(t+1)/nthreads,value); });
	for( auto &t : pool ) t.join();
}

// Every Shiokaze module of this repository carries one of these. CUDA loads a kernel's code lazily, on its first launch, and that load takes a process-wide
// driver lock and may wait for the launching context to drain. With `GPUs=N` the module drives N devices from N host threads of ONE process, and its kernels wait
// for each other across devices: a thread that holds the loader lock while its device spins on a neighbour, whose host thread in turn needs that lock to launch
// the kernel the spin is waiting for, is a deadlock (seen as a rare 20 s communicator time-out in the first projection of a run). Loading everything when the
// context is created removes the lock from the solve. Runs when the module is dlopen'ed, i.e. before the first CUDA call of a Shiokaze host; a value the
// user has set is kept.
namespace { struct b200_eager_cuda_modules { b200_eager_cuda_modules() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); } } b200_eager_cuda_modules_instance; }
#endif
#endif
