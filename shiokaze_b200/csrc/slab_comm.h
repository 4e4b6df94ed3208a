// z-slab communicator: halo planes between neighbouring ranks and the CG scalar all-reduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace shkz {

struct CGState;

class SlabComm {
public:
	SlabComm(long long plane_cells, int device);
	~SlabComm();
	static int unique_id(uint8_t *id, std::string &err);
	int export_window(uint8_t *ipc);
	int connect(int rank, int world, const uint8_t *id, const uint8_t *lower_ipc, const uint8_t *upper_ipc);
	// fill the ghost planes of a cell array (pointer at plane 0) from the neighbouring slabs
	int exchange(void *p, long long plane_cells, int nzl, size_t elem, cudaStream_t stream);
	int allreduce_begin_state(CGState *st, cudaStream_t stream);
	int allreduce_sum_x(CGState *st, cudaStream_t stream);
	const char *error() const { return m_error.c_str(); }
private:
	std::string m_error;
	long long m_plane;
	int m_device;
};

} // namespace shkz
