#include "slab_comm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace shkz {

namespace {
struct Blob { // what shkz_b200_slab_export hands to the other ranks
	uint32_t magic;
	int32_t device;
	uint64_t arena_bytes;
	cudaIpcMemHandle_t handle;
};
constexpr uint32_t kMagic = 0x5a4b4853u; // "SHKZ"
std::string cuda_err(const char *what, cudaError_t e) { return std::string(what) + ": " + cudaGetErrorString(e); }
} // namespace

SlabComm::~SlabComm() {
	cudaSetDevice(m_device);
	if (m_ipc)
		for (int r = 0; r < m_world; ++r)
			if (r != m_rank && m_peer[r]) cudaIpcCloseMemHandle(m_peer[r]);
	if (m_dev) cudaFree(m_dev);
	if (m_base) cudaFree(m_base);
}

int SlabComm::create_arena(size_t bytes) {
	if (m_base) { m_error = "arena already created"; return 1; }
	cudaError_t e = cudaMalloc((void **)&m_base, bytes);
	if (e != cudaSuccess) { m_error = cuda_err("cudaMalloc(arena)", e); m_base = nullptr; return 1; }
	e = cudaMemset(m_base, 0, bytes);
	if (e != cudaSuccess) { m_error = cuda_err("cudaMemset(arena)", e); return 1; }
	m_size = bytes;
	m_used = ARENA_HEADER;
	return 0;
}

void *SlabComm::carve(size_t bytes) {
	const size_t need = (bytes + 255) / 256 * 256;
	if (!m_base || m_used + need > m_size) return nullptr;
	void *p = m_base + m_used;
	m_used += need;
	cudaMemset(p, 0, need);
	return p;
}

int SlabComm::export_handle(uint8_t *blob, size_t blob_bytes) {
	if (!m_base) { m_error = "no arena"; return 1; }
	if (blob_bytes < sizeof(Blob)) { m_error = "export blob too small"; return 1; }
	Blob b{};
	b.magic = kMagic;
	b.device = m_device;
	b.arena_bytes = m_size;
	cudaError_t e = cudaIpcGetMemHandle(&b.handle, m_base);
	if (e != cudaSuccess) { m_error = cuda_err("cudaIpcGetMemHandle", e); return 1; }
	memset(blob, 0, blob_bytes);
	memcpy(blob, &b, sizeof b);
	return 0;
}

int SlabComm::connect_ipc(int rank, int world, const uint8_t *blobs, size_t blob_bytes) {
	if (m_connected) { m_error = "already connected"; return 1; }
	if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world) { m_error = "bad rank / world"; return 1; }
	m_rank = rank; m_world = world; m_ipc = true;
	for (int r = 0; r < world; ++r) {
		if (r == rank) { m_peer[r] = m_base; continue; }
		Blob b;
		memcpy(&b, blobs + (size_t)r * blob_bytes, sizeof b);
		if (b.magic != kMagic) { m_error = "export blob of rank " + std::to_string(r) + " is not valid"; return 1; }
		if (b.arena_bytes != m_size) { m_error = "rank " + std::to_string(r) + " has a different arena size (slabs must be equal)"; return 1; }
		void *p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { m_error = cuda_err("cudaIpcOpenMemHandle", e); return 1; }
		m_peer[r] = static_cast<char *>(p);
	}
	return finish_connect();
}

int SlabComm::connect_local(int rank, int world, SlabComm *const *all) {
	if (m_connected) { m_error = "already connected"; return 1; }
	if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world) { m_error = "bad rank / world"; return 1; }
	m_rank = rank; m_world = world; m_ipc = false;
	cudaSetDevice(m_device);
	for (int r = 0; r < world; ++r) {
		if (all[r]->m_size != m_size) { m_error = "slabs must be equal (arena sizes differ)"; return 1; }
		m_peer[r] = all[r]->m_base;
		if (r != rank && all[r]->m_device != m_device) {
			cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->m_device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { m_error = cuda_err("cudaDeviceEnablePeerAccess", e); return 1; }
			cudaGetLastError();
		}
	}
	return finish_connect();
}

unsigned long long SlabComm::read_abort() {
	unsigned long long w = 0;
	if (!m_base) return 0;
	cudaSetDevice(m_device);
	if (cudaMemcpy(&w, m_base + HDR_ABORT, sizeof w, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 0; }
	return w;
}

void SlabComm::raise_abort() {
	if (!m_connected) return;
	cudaSetDevice(m_device);
	const unsigned long long w = (unsigned long long)(m_rank + 1) | ((unsigned long long)ABORT_HOST << 8);
	cudaStream_t st = nullptr;
	if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; } // (the solver's own stream may be the one that is stuck)
	for (int r = 0; r < m_world; ++r)
		if (m_peer[r]) cudaMemcpyAsync(m_peer[r] + HDR_ABORT, &w, sizeof w, cudaMemcpyHostToDevice, st);
	cudaStreamSynchronize(st);
	cudaStreamDestroy(st);
	cudaGetLastError();
}

std::string SlabComm::describe_abort(unsigned long long w) {
	const int who = (int)(w & 0xff) - 1, what = (int)((w >> 8) & 0xff);
	const unsigned long long seq = w >> 16;
	const char *names[] = {"?", "the halo planes of its lower neighbour", "the halo planes of its upper neighbour", "the reduction mailbox", "its host gave up before launching"};
	char buf[256];
	if (what == ABORT_HOST) snprintf(buf, sizeof buf, "slab communicator aborted: rank %d failed on the host before it could take part", who);
	else snprintf(buf, sizeof buf, "slab communicator aborted: rank %d timed out waiting for %s (exchange %llu)", who, names[what >= 0 && what <= 3 ? what : 0], seq);
	return buf;
}

int SlabComm::finish_connect() {
	CommDev h{};
	{
		// device-side waits give up after this long (default 20 s; a projection takes milliseconds): SHKZ_B200_COMM_TIMEOUT_MS
		double ms = 20000.0;
		if (const char *e = getenv("SHKZ_B200_COMM_TIMEOUT_MS")) { const double v = atof(e); if (v > 0.0) ms = v; }
		h.timeout_ns = (unsigned long long)(ms * 1e6);
	}
	h.rank = m_rank; h.world = m_world; h.self = m_base;
	h.lo = m_rank > 0 ? m_peer[m_rank - 1] : nullptr;
	h.hi = m_rank + 1 < m_world ? m_peer[m_rank + 1] : nullptr;
	for (int r = 0; r < m_world; ++r) h.peer[r] = m_peer[r];
	cudaError_t e = cudaMalloc((void **)&m_dev, sizeof(CommDev));
	if (e == cudaSuccess) e = cudaMemcpy(m_dev, &h, sizeof h, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { m_error = cuda_err("CommDev upload", e); return 1; }
	m_connected = true;
	return 0;
}

} // namespace shkz
