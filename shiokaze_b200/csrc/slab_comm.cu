#include "slab_comm.h"

namespace shkz {

SlabComm::SlabComm(long long plane_cells, int device) : m_plane(plane_cells), m_device(device) {}
SlabComm::~SlabComm() {}
int SlabComm::unique_id(uint8_t *, std::string &err) { err = "slab communicator not built yet"; return 1; }
int SlabComm::export_window(uint8_t *) { m_error = "slab communicator not built yet"; return 1; }
int SlabComm::connect(int, int, const uint8_t *, const uint8_t *, const uint8_t *) { m_error = "slab communicator not built yet"; return 1; }
int SlabComm::exchange(void *, long long, int, size_t, cudaStream_t) { m_error = "slab communicator not built yet"; return 1; }
int SlabComm::allreduce_begin_state(CGState *, cudaStream_t) { m_error = "slab communicator not built yet"; return 1; }
int SlabComm::allreduce_sum_x(CGState *, cudaStream_t) { m_error = "slab communicator not built yet"; return 1; }

} // namespace shkz
