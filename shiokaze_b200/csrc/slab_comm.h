// Host side of the z-slab communicator (device side: slab_comm.cuh): the arena every ghosted array of a slab solver
// is carved from, its export / import through CUDA IPC (one process per GPU) or direct peer access (one process,
// several GPUs), and the numbering of halo exchanges.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace shkz {

constexpr int COMM_MAX_WORLD = 8;
constexpr size_t HDR_FLAG_FROM_LO = 0, HDR_FLAG_FROM_HI = 128, HDR_RED_SEQ = 256, HDR_PUSH_TICKET = 384, HDR_PUSH_TICKET2 = 512, HDR_ABORT = 640, HDR_MAIL = 1024;
// HDR_ABORT: 0 while the communicator is healthy. A rank whose device-side wait ran out of time (CommDev::timeout_ns), or whose host gave up
// before launching (SlabComm::raise_abort), stores a non-zero word here IN EVERY RANK'S ARENA: every wait of every rank then returns at once,
// the streams drain, and the host call ends with SHKZ_B200_ERR_COMM instead of hanging the NVLink domain.
//   bits 0..7  = 1 + rank that raised it    bits 8..15 = what it waited for (1 lower neighbour, 2 upper neighbour, 3 reduction mailbox, 4 host)
//   bits 16..  = the exchange / reduction number it waited for
constexpr int ABORT_WAIT_LO = 1, ABORT_WAIT_HI = 2, ABORT_WAIT_MAIL = 3, ABORT_HOST = 4;
constexpr size_t ARENA_HEADER = 65536;

struct CommDev {
	int rank, world;
	char *self;                 // own arena
	char *lo, *hi;              // arenas of the z-neighbours as mapped into this process (nullptr at the domain ends)
	char *peer[COMM_MAX_WORLD]; // every rank's arena (peer[rank] == self)
	unsigned long long timeout_ns; // longest a device-side wait may spin before it aborts the communicator
};


class SlabComm {
public:
	SlabComm(int device) : m_device(device) {}
	~SlabComm();
	// arena
	int create_arena(size_t bytes);
	void *carve(size_t bytes); // 256-byte aligned, zero-filled; nullptr when the arena is exhausted
	bool owns(const void *p) const { return m_base && p >= (const void *)m_base && p < (const void *)(m_base + m_size); }
	size_t offset_of(const void *p) const { return (size_t)((const char *)p - m_base); }
	size_t mark() const { return m_used; }
	void rewind(size_t mark) { m_used = mark; } // drop everything carved after mark()
	// wiring
	int export_handle(uint8_t *blob, size_t blob_bytes);                             // this rank's arena, for the other processes
	int connect_ipc(int rank, int world, const uint8_t *blobs, size_t blob_bytes);   // blobs: world entries, rank-major
	int connect_local(int rank, int world, SlabComm *const *all);                    // same process: peer access
	bool connected() const { return m_connected; }
	int rank() const { return m_rank; }
	int world() const { return m_world; }
	const CommDev *device_view() const { return m_dev; }
	unsigned long long next_exchange() { return ++m_exchange; }
	// health: the abort word of the own arena (0 = fine), read after a stream synchronise; raise_abort() releases every rank's device-side waits
	// when this rank's host cannot take part in a call the others may already be running
	unsigned long long read_abort();
	void raise_abort();
	static std::string describe_abort(unsigned long long word);
	const char *error() const { return m_error.c_str(); }

private:
	int finish_connect();
	std::string m_error;
	int m_device;
	char *m_base = nullptr;
	size_t m_size = 0, m_used = 0;
	bool m_connected = false, m_ipc = false;
	int m_rank = 0, m_world = 1;
	char *m_peer[COMM_MAX_WORLD] = {};
	CommDev *m_dev = nullptr;
	unsigned long long m_exchange = 0;
};

} // namespace shkz
