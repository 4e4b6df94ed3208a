// Host side of libshkz_b200: device memory, launch sequences, and the C-ABI of include/shkz_b200.h.
// No CPU compute path exists in this file: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/shkz_b200.h"
#include "common.cuh"
#include "kernels_assemble.cuh"
#include "kernels_cg.cuh"
#include "kernels_mg.cuh"
#ifdef SHKZ_B200_TEST_HOOKS
#include "kernels_legacy.cuh" // one-launch-per-colour validation kernels: only in the test-hook build of the library
#endif
#include "kernels_mg_tma.cuh"
#include "kernels_cg_tma.cuh"
#include "kernels_post.cuh"
#include "kernels_xfer.cuh"
#include "slab_comm.h"

using namespace shkz;

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_error = buf;
	return code;
}

#define CK(call)                                                                                              \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if (e_ != cudaSuccess) return fail(SHKZ_B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)
#define CKR(call)                     \
	do {                              \
		int r_ = (call);              \
		if (r_ != SHKZ_B200_OK) return r_; \
	} while (0)

// A cell-shaped device array with one ghost plane on each side; `p` points at plane 0. Slab solvers carve these
// from their communicator's arena (slab_comm.h) so that neighbours can address the ghost planes.
struct CellArray {
	void *base = nullptr;
	size_t bytes = 0;
	bool in_arena = false;
	int alloc(const Dims &d, size_t elem, SlabComm *arena = nullptr) {
		bytes = (size_t)(d.nzl + 2) * (size_t)d.plane * elem;
		if (arena) {
			base = arena->carve(bytes);
			in_arena = true;
			if (!base) return fail(SHKZ_B200_ERR_CUDA, "slab arena exhausted (%zu more bytes needed)", bytes);
			return SHKZ_B200_OK;
		}
		CK(cudaMalloc(&base, bytes));
		CK(cudaMemset(base, 0, bytes));
		return SHKZ_B200_OK;
	}
	void release() {
		if (base && !in_arena) cudaFree(base);
		base = nullptr;
		bytes = 0;
		in_arena = false;
	}
	template <class T> T *ptr(const Dims &d) const { return base ? static_cast<T *>(base) + d.plane : nullptr; }
};

struct PlainArray {
	void *base = nullptr;
	size_t bytes = 0;
	int alloc(size_t n) {
		bytes = n ? n : 1;
		CK(cudaMalloc(&base, bytes));
		CK(cudaMemset(base, 0, bytes));
		return SHKZ_B200_OK;
	}
	void release() {
		if (base) cudaFree(base);
		base = nullptr;
		bytes = 0;
	}
};

Dims make_dims(int nx, int ny, int nzl, int k0, int nzg) {
	Dims d;
	d.nx = nx; d.ny = ny; d.nzl = nzl; d.k0 = k0; d.nzg = nzg;
	d.plane = (long long)nx * ny;
	d.ncell = d.plane * nzl;
	return d;
}

size_t face_count(const Dims &d, int dim) {
	return (size_t)(d.nx + (dim == 0)) * (size_t)(d.ny + (dim == 1)) * (size_t)(d.nzl + (dim == 2));
}

struct Profiler {
	bool on = false;
	std::vector<cudaEvent_t> ev; // pairs
	std::vector<int> id;         // entry index per pair
	size_t used = 0;
	std::vector<std::string> names;
	std::vector<uint64_t> count;
	std::vector<double> ms;
	std::map<std::string, int> index;
	static constexpr size_t kMaxPairs = 16384;
	int entry(const char *tag) {
		auto it = index.find(tag);
		if (it != index.end()) return it->second;
		const int n = (int)names.size();
		names.push_back(tag); count.push_back(0); ms.push_back(0.0);
		index[tag] = n;
		return n;
	}
	int begin(const char *tag, cudaStream_t stream) {
		if (!on || used >= kMaxPairs) return -1;
		if (ev.size() < 2 * (used + 1)) {
			cudaEvent_t a, b;
			if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
			ev.push_back(a); ev.push_back(b);
			id.push_back(0);
		}
		id[used] = entry(tag);
		cudaEventRecord(ev[2 * used], stream);
		return (int)used;
	}
	void end(int slot, cudaStream_t stream) {
		if (slot < 0) return;
		cudaEventRecord(ev[2 * slot + 1], stream);
		used = slot + 1;
	}
	void collect() { // after the stream has been synchronised
		for (size_t n = 0; n < used; ++n) {
			float t = 0.f;
			if (cudaEventElapsedTime(&t, ev[2 * n], ev[2 * n + 1]) == cudaSuccess) { ms[id[n]] += t; count[id[n]] += 1; }
		}
		used = 0;
	}
	void reset() { used = 0; names.clear(); count.clear(); ms.clear(); index.clear(); }
	void destroy() { for (auto e : ev) cudaEventDestroy(e); ev.clear(); id.clear(); reset(); }
};

struct HostLevel {
	Dims d;
	int bz = 4, tiles_total = 0; // deepest tile depth of the level; number of flag slices (= the largest number of tiles any depth gives)
	int slice = 4, nslices_z = 0, ntx = 0, nty = 0;
	bool adaptive = false;       // tile depth chosen per projection by k_compact_tiles (levels with deep tiles)
	std::string tag_sweep, tag_restrict, tag_prolong, tag_coarsen, tag_compact;
	CellArray wx, wy, wz, dd, xa, xb, b;
	CellArray legacy_r; // residual array of the unfused validation path (allocated on first use)
	bool own_coef = true, own_b = true; // level 0 may alias CG arrays
	PlainArray tile_flags, tile_ids, tile_count;
	// tiles whose arrays may hold something other than zeros: flagged now or in the previous projection (`tile_dirty` = last call's flags). The kernels
	// that WRITE a level's arrays (assembly, coarsening, pressure scatter) walk this union list; everything outside it is zero and stays zero.
	PlainArray tile_dirty, tile_uids, tile_ucount;
	Tiles utiles{};
	MGLevel view;
	bool tma = false; // tensor maps built: the level's sweeps run the TMA-staged kernel
	CUtensorMap map_wx, map_wy, map_wz, map_dd, map_b, map_xa, map_xb;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
		else cudaGetLastError();
	}
	return fn;
}

// 3-D map of a cell array WITH its ghost planes (z coordinate = plane + 1), box = cols x rows x 1 plane of float / double elements, zero fill outside
bool make_box_map(CUtensorMap *map, void *base, const Dims &d, size_t elem, int cols, int rows) {
	EncodeTiledFn enc = encode_tiled();
	if (!enc || !base) return false;
	const cuuint64_t dims[3] = {(cuuint64_t)d.nx, (cuuint64_t)d.ny, (cuuint64_t)d.nzl + 2};
	const cuuint64_t strides[2] = {(cuuint64_t)d.nx * elem, (cuuint64_t)d.plane * elem};
	const cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)rows, 1u};
	const cuuint32_t estr[3] = {1u, 1u, 1u};
	return enc(map, elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_plane_map(CUtensorMap *map, void *base, const Dims &d, int rows) { return make_box_map(map, base, d, sizeof(float), ST_W, rows); }

} // namespace

struct shkz_b200_solver {
	Dims d;
	double dx = 0;
	int real = SHKZ_B200_REAL_F32;
	int device = 0;
	int num_sms = 148;
	size_t real_bytes = 4;
	bool whole_grid = true;
	// geometry / assembly products (RealT)
	CellArray phi, pressure, curv;
	CellArray in_rows;
	PlainArray areas[3], rhos[3];
	// operator + CG vectors; element sizes depend on the precision mode they were allocated for
	int alloc_precision = -1;
	CellArray wx, wy, wz, dd; // CoefT
	CellArray b, x, r, s, q;    // VecT
	CellArray s2;               // VecT: the other direction buffer of the fused xpay + product kernel (k_xpay_spmv_tma writes s_new out of place)
	bool spmv_tma = false;      // tensor maps of that kernel built (whole grid, float coefficients, nx % 4 == 0)
	SpmvMaps spmv_maps{};
	CellArray p_prev;           // VecT: WarmStart=Yes — the previous call's pressure per cell (allocated on first use)
	bool poisoned = false;      // the last solve ended on a non-finite scalar: the solver's persistent vectors are re-zeroed before the next one
	std::vector<HostLevel> levels;
	int mg_min_size_built = -1;
	int tail_first = -1;        // first level of the shared-memory tail of the V-cycle (-1: none)
	size_t tail_smem = 0;
	TailArgs tail_args{};
	// z-slab solvers: from level `agg_level` down the hierarchy is GLOBAL — every rank gathers the whole coarse
	// problem over NVLink and solves it redundantly with the whole-grid kernels (glevels[0] is level agg_level)
	std::vector<HostLevel> glevels;
	int agg_level = -1;
	int gtail_first = -1;
	size_t gtail_smem = 0;
	TailArgs gtail_args{};
	// levels [mid_first, tail_first) run inside ONE cooperative launch together with the tail (k_vcycle_mid, kernels_mg.cuh); -1: no such levels
	int mid_first = -1, gmid_first = -1;
	PlainArray mid_barrier;
	std::map<const void *, int> occupancy; // resident CTAs per SM of each persistent kernel
	// reductions / control
	PlainArray partials, counter, state;
	CGState *h_state = nullptr; // pinned
	size_t max_blocks = 0;
	unsigned last_iterations = 0; // of the previous solve: sizes the first batch of iterations before the host looks
	// host-call staging (device)
	PlainArray st_vel[3], st_act[3], st_solid, st_fluid, st_pressure, st_pact;
	// sparse host copies (kernels_xfer.cuh): flags / list / count of the wet transfer blocks, byte counter of the face push, their pinned mirror
	PlainArray xf_flags, xf_list, xf_count, xf_pushed;
	unsigned long long *h_xf = nullptr; // pinned: [0] wet blocks [1] bytes pushed
	cudaStream_t copy_stream = nullptr; // non-blocking: the activity masks travel behind the solve
	bool xfer_sparse = false;           // set around the project_impl of a sparse host call: the caller's pressure grids are host memory written by k_store_pressure
	                                    // over the union tile list (no wholesale clear), the masks arrive on copy_stream (ev[11]), switched-off faces are marked
	const void *last_host_pressure = nullptr, *last_host_pact = nullptr; // the grids the previous sparse call wrote: everything off its tiles is known to be zero there
	uint64_t project_serial = 0, host_serial = ~0ull; // ... provided no other projection ran on this solver in between
	// extrapolation + solid constraint after the projection (kernels_post.cuh): the other activity mask of the rounds, the velocity before the constraint
	PlainArray post_act[3], post_vel[3], post_mask[3];
	// slab communicator (nullptr on a whole grid)
	SlabComm *comm = nullptr;
	size_t arena_mark = 0; // arena fill level after the creation-time arrays
	unsigned long long pending_wait = 0; // exchange published by the last fused slab sweep that no kernel has waited for yet
	// bookkeeping
	uint64_t launches = 0;
	Profiler prof;
	bool have_system = false;
	bool have_hierarchy = false;
	bool fractions_stale = false; // the face arrays of the fractions were not written by the last project() (they never are: debug_fetch materialises them)
	const void *last_solid = nullptr; // the caller's device solid grid of the last project() (debug_fetch of the area fractions reads it again)
	const float *debug_vcycle_result = nullptr;
	int sweep_mode = 0; // 0: best kernel per level (TMA-staged > quad > scalar); 1: no TMA; 2: scalar only (debug / A-B timing)
	AsmParams last_asm{};
	cudaEvent_t ev[12]{};
	bool events = false;

	SlabComm *arena() const { return whole_grid ? nullptr : comm; }
	RedBuf redbuf() const {
		return RedBuf{static_cast<double *>(partials.base), static_cast<unsigned int *>(counter.base), (comm && comm->connected()) ? comm->device_view() : nullptr};
	}
	CGState *dstate() const { return static_cast<CGState *>(state.base); }
};

namespace {

void release_precision_arrays(shkz_b200_solver *S) {
	for (CellArray *a : {&S->wx, &S->wy, &S->wz, &S->dd, &S->b, &S->x, &S->r, &S->s, &S->q, &S->p_prev, &S->s2}) a->release();
	S->spmv_tma = false;
	for (HostLevel &L : S->levels) {
		if (L.own_coef) { L.wx.release(); L.wy.release(); L.wz.release(); L.dd.release(); }
		if (L.own_b) L.b.release();
		L.xa.release(); L.xb.release(); L.legacy_r.release();
		L.tile_flags.release(); L.tile_ids.release(); L.tile_count.release(); L.tile_dirty.release(); L.tile_uids.release(); L.tile_ucount.release();
	}
	for (HostLevel &L : S->glevels) {
		L.wx.release(); L.wy.release(); L.wz.release(); L.dd.release(); L.b.release();
		L.xa.release(); L.xb.release(); L.legacy_r.release();
		L.tile_flags.release(); L.tile_ids.release(); L.tile_count.release(); L.tile_dirty.release(); L.tile_uids.release(); L.tile_ucount.release();
	}
	S->glevels.clear();
	S->agg_level = -1;
	S->gtail_first = -1;
	S->mid_first = S->gmid_first = -1;
	S->levels.clear();
	S->last_host_pressure = S->last_host_pact = nullptr; // (the union tile lists start over)
	S->alloc_precision = -1;
	S->have_system = false;
	S->have_hierarchy = false;
	S->tail_first = -1;
	if (S->arena()) S->arena()->rewind(S->arena_mark); // every rank re-carves the same sequence, offsets stay in step
}

// planes per tile: as deep as possible (each tile relaxes two extra halo planes) while the level still
// yields about one full wave of tiles
int pick_bz(const Dims &d) {
	const long long xy = (long long)((d.nx + TX - 1) / TX) * ((d.ny + TY - 1) / TY);
	int bz = 32;
	while (bz > 4 && xy * ((d.nzl + bz - 1) / bz) < 592) bz >>= 1;
	if (const char *e = getenv("SHKZ_B200_BZ")) { const int v = atoi(e); if (v >= 2 && v <= 64 && !(v & 1) && v < bz) bz = v; } // (experiments)
	return bz;
}

// arrays, tile list and tensor maps of one multigrid level (coefficient / rhs arrays may have been aliased by the caller)
// balanced: the level's kernels divide the active tiles evenly over the persistent grid (TileWalk, common.cuh) — every whole-grid level; the levels of a
// z-slab keep whole tiles per CTA, the fused slab sweep orders them by whether they touch a ghost plane
int finish_level(HostLevel &L, const Dims &cur, const std::string &n, SlabComm *arena, bool balanced) {
	L.d = cur;
	if (!L.wx.base) for (CellArray *a : {&L.wx, &L.wy, &L.wz, &L.dd}) CKR(a->alloc(cur, sizeof(float), arena));
	if (!L.b.base) CKR(L.b.alloc(cur, sizeof(float), arena));
	CKR(L.xa.alloc(cur, sizeof(float), arena));
	CKR(L.xb.alloc(cur, sizeof(float), arena));
	L.bz = pick_bz(cur);
	L.slice = L.bz < 8 ? L.bz : 8;
	L.adaptive = L.bz > L.slice && !getenv("SHKZ_B200_BZ"); // (z-slab levels too: every rank picks the depth that suits its own slab)
	const int ntx = (cur.nx + TX - 1) / TX, nty = (cur.ny + TY - 1) / TY;
	L.ntx = ntx; L.nty = nty;
	L.nslices_z = (cur.nzl + L.slice - 1) / L.slice;
	L.tiles_total = ntx * nty * L.nslices_z;
	CKR(L.tile_flags.alloc((size_t)L.tiles_total));
	CKR(L.tile_ids.alloc((size_t)L.tiles_total * sizeof(int)));
	CKR(L.tile_count.alloc(2 * sizeof(int)));
	CKR(L.tile_dirty.alloc((size_t)L.tiles_total));
	CKR(L.tile_uids.alloc((size_t)L.tiles_total * sizeof(int)));
	CKR(L.tile_ucount.alloc(2 * sizeof(int)));
	L.tag_sweep = "sweep@" + n; L.tag_restrict = "residual_restrict@" + n; L.tag_prolong = "prolong_add@" + n;
	L.tag_coarsen = "coarsen_operator@" + n; L.tag_compact = "compact_tiles@" + n;
	L.view.d = cur;
	L.view.tiles = Tiles{static_cast<const int *>(L.tile_ids.base), static_cast<const int *>(L.tile_count.base), ntx, nty, L.slice, L.adaptive ? 0 : L.bz, balanced ? (getenv("SHKZ_B200_NO_HYBRID") ? 1 : 5) : 0}; // (the variable: A-B timing)
	L.utiles = L.view.tiles;
	L.utiles.ids = static_cast<const int *>(L.tile_uids.base);
	L.utiles.count = static_cast<const int *>(L.tile_ucount.base);
	L.view.wx = L.wx.ptr<float>(cur); L.view.wy = L.wy.ptr<float>(cur); L.view.wz = L.wz.ptr<float>(cur); L.view.dd = L.dd.ptr<float>(cur);
	L.view.b = L.b.ptr<float>(cur); L.view.xa = L.xa.ptr<float>(cur); L.view.xb = L.xb.ptr<float>(cur);
	L.tma = (cur.nx & 3) == 0 && make_plane_map(&L.map_wx, L.wx.base, cur, ST_ROWS) && make_plane_map(&L.map_wy, L.wy.base, cur, ST_WY_ROWS) &&
	        make_plane_map(&L.map_wz, L.wz.base, cur, ST_ROWS) && make_plane_map(&L.map_dd, L.dd.base, cur, ST_ROWS) &&
	        make_plane_map(&L.map_b, L.b.base, cur, ST_ROWS) && make_plane_map(&L.map_xa, L.xa.base, cur, ST_XO_ROWS) &&
	        make_plane_map(&L.map_xb, L.xb.base, cur, ST_XO_ROWS);
	return SHKZ_B200_OK;
}

int big_extent(const Dims &d) { return d.nx > d.ny ? (d.nx > d.nzg ? d.nx : d.nzg) : (d.ny > d.nzg ? d.ny : d.nzg); }

// the shared-memory tail of a whole-grid hierarchy: the longest run of coarsest levels that fits one CTA's shared memory
int setup_tail(const std::vector<HostLevel> &lv, int &tail_first, size_t &tail_smem, TailArgs &A) {
	tail_first = -1;
	const size_t limit = 200 * 1024;
	size_t bytes = 0;
	int first = (int)lv.size();
	while (first > 0 && (int)lv.size() - (first - 1) <= TAIL_MAX_LEVELS) {
		const Dims &dl = lv[first - 1].d;
		const size_t add = (size_t)TAIL_ARRAYS * (size_t)(dl.ncell + 2 * dl.plane) * sizeof(float);
		if (bytes + add > limit) break;
		if ((long long)((dl.nx + 1) / 2) * dl.ny * dl.nzl > (long long)TAIL_SLOTS * TAIL_THREADS) break; // tail_body: at most TAIL_SLOTS cell pairs per thread
		bytes += add;
		--first;
	}
	if (first >= (int)lv.size()) return SHKZ_B200_OK;
	tail_first = first;
	tail_smem = bytes;
	A = TailArgs{};
	A.nlev = (int)lv.size() - first;
	int off = 0;
	for (int m = 0; m < A.nlev; ++m) {
		const HostLevel &L = lv[first + m];
		A.L[m].d = L.d;
		A.L[m].wx = L.view.wx; A.L[m].wy = L.view.wy; A.L[m].wz = L.view.wz; A.L[m].dd = L.view.dd;
		A.L[m].stride = (int)(L.d.ncell + 2 * L.d.plane);
		A.L[m].offset = off;
		off += TAIL_ARRAYS * A.L[m].stride;
	}
	A.b_in = lv[first].view.b;
	A.x_out = lv[first].view.xa;
	return SHKZ_B200_OK;
}

// z-slab solvers: levels this small (largest global extent, or global cell count: a long thin stack of slabs) are gathered and solved
// redundantly on every rank — a whole-grid kernel on 2 M cells takes ~10 us, a slab sweep never less than two NVLink round trips
// levels of at most this many cells (dense) leave the tiled kernels for the cooperative mid-V-cycle kernel
long long mid_max_cells() {
	static long long v = -1;
	if (v < 0) {
		v = 1ll << 21;
		if (const char *e = getenv("SHKZ_B200_MID_CELLS")) v = atoll(e); // 0 switches the kernel off (A-B timing)
	}
	return v;
}
int pick_mid_first(const std::vector<HostLevel> &lv, int tail_first, bool allow_level0) {
	const int end = tail_first >= 0 ? tail_first : (int)lv.size();
	int first = -1;
	for (int l = end - 1; l >= (allow_level0 ? 0 : 1); --l) {
		if (lv[l].d.ncell > mid_max_cells() || end - l > MID_MAX_LEVELS) break;
		first = l;
	}
	return first;
}

constexpr int AGG_MAX_EXTENT = 64;
constexpr long long AGG_MAX_CELLS = 1ll << 23;

template <class VecT, class CoefT>
int ensure_precision_arrays(shkz_b200_solver *S, int precision, const shkz_b200_params &P) {
	const int min_size = P.mg_min_size < 2 ? 2 : P.mg_min_size;
	if (S->alloc_precision == precision && S->mg_min_size_built == min_size) return SHKZ_B200_OK;
	release_precision_arrays(S);
	const Dims &d = S->d;
	SlabComm *arena = S->arena();
	for (CellArray *a : {&S->wx, &S->wy, &S->wz, &S->dd}) CKR(a->alloc(d, sizeof(CoefT), arena));
	for (CellArray *a : {&S->b, &S->x, &S->r, &S->s, &S->q}) CKR(a->alloc(d, sizeof(VecT), arena));
	// multigrid hierarchy (always allocated: switching the preconditioner must not reallocate; level 0 also
	// carries the tile list every CG kernel runs over)
	Dims cur = d;
	for (int l = 0;; ++l) {
		S->levels.emplace_back();
		HostLevel &L = S->levels.back();
		if (l == 0 && sizeof(CoefT) == sizeof(float)) {
			L.own_coef = false;
			L.wx = S->wx; L.wy = S->wy; L.wz = S->wz; L.dd = S->dd;
		}
		if (l == 0 && sizeof(VecT) == sizeof(float)) {
			L.own_b = false;
			L.b = S->r; // an all-float CG smooths against r directly
		}
		CKR(finish_level(L, cur, std::to_string(l), arena, S->whole_grid));
		if (big_extent(cur) <= min_size) break;
		if (!S->whole_grid) {
			// slabs: aggregates stay inside one rank (even local extent and even first plane); small levels go global
			if (l >= 1 && (big_extent(cur) <= AGG_MAX_EXTENT || (long long)cur.plane * cur.nzg <= AGG_MAX_CELLS)) { S->agg_level = l; break; }
			if ((cur.nzl & 1) || (cur.k0 & 1) || cur.nzl < 2) { if (l >= 1) S->agg_level = l; break; }
		}
		cur = make_dims((cur.nx + 1) / 2, (cur.ny + 1) / 2, (cur.nzl + 1) / 2, cur.k0 / 2, (cur.nzg + 1) / 2);
	}
	S->tail_first = -1;
	if (S->whole_grid) {
		CKR(setup_tail(S->levels, S->tail_first, S->tail_smem, S->tail_args));
	} else if (S->agg_level >= 0) {
		// the global continuation of the hierarchy: whole-grid levels, in the arena because every rank stores its part of
		// the operator and of each right-hand side straight into every other rank's copy
		const Dims &ds = S->levels[S->agg_level].d;
		Dims g = make_dims(ds.nx, ds.ny, ds.nzg, 0, ds.nzg);
		for (int l = 0;; ++l) {
			S->glevels.emplace_back();
			CKR(finish_level(S->glevels.back(), g, "g" + std::to_string(l), arena, true));
			if (big_extent(g) <= min_size) break;
			g = make_dims((g.nx + 1) / 2, (g.ny + 1) / 2, (g.nzl + 1) / 2, 0, (g.nzg + 1) / 2);
		}
		CKR(setup_tail(S->glevels, S->gtail_first, S->gtail_smem, S->gtail_args));
	}
	{
		const size_t smem = S->tail_smem > S->gtail_smem ? S->tail_smem : S->gtail_smem;
		if (smem) CK(cudaFuncSetAttribute(k_vcycle_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		if (smem) CK(cudaFuncSetAttribute(k_vcycle_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		S->mid_first = S->whole_grid ? pick_mid_first(S->levels, S->tail_first, true) : -1; // (slab levels exchange halos between sweeps: tiled kernels)
		S->gmid_first = S->glevels.empty() ? -1 : pick_mid_first(S->glevels, S->gtail_first, false);
		int coop = 0;
		if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, S->device) != cudaSuccess || !coop) S->mid_first = S->gmid_first = -1;
		if (!S->mid_barrier.base) CKR(S->mid_barrier.alloc(64));
	}
	// the fused direction update + product (kernels_cg_tma.cuh): whole grids with float coefficients whose rows are whole quads
	S->spmv_tma = false;
	// (z-slabs too — k_xpay_spmv_tma<SLAB> — unless the whole problem is gathered from level 0 on: its V-cycle result then lives in the global arrays)
	if ((S->whole_grid || (S->agg_level != 0 && !getenv("SHKZ_B200_NO_SLAB_SPMV_TMA"))) && sizeof(CoefT) == sizeof(float) && (d.nx & 3) == 0 && !getenv("SHKZ_B200_NO_SPMV_TMA")) {
		using SS = SpmvStage<VecT>;
		CKR(S->s2.alloc(d, sizeof(VecT), arena));
		HostLevel &L0 = S->levels[0];
		SpmvMaps &M = S->spmv_maps;
		S->spmv_tma = make_box_map(&M.s[0], S->s.base, d, sizeof(VecT), SS::SW, SS::SROWS) && make_box_map(&M.s[1], S->s2.base, d, sizeof(VecT), SS::SW, SS::SROWS) &&
		              make_box_map(&M.z[0], L0.xa.base, d, sizeof(float), SS::ZW, SS::SROWS) && make_box_map(&M.z[1], L0.xb.base, d, sizeof(float), SS::ZW, SS::SROWS) &&
		              make_box_map(&M.wx, S->wx.base, d, sizeof(float), SS::WXW, TY) && make_box_map(&M.wy, S->wy.base, d, sizeof(float), TX, SS::WYROWS) &&
		              make_box_map(&M.wz, S->wz.base, d, sizeof(float), TX, TY) && make_box_map(&M.dd, S->dd.base, d, sizeof(float), TX, TY);
	}
	S->alloc_precision = precision;
	S->mg_min_size_built = min_size;
	return SHKZ_B200_OK;
}

dim3 cell_block() { return dim3(32, 8, 1); }
dim3 cell_grid(const Dims &d, int ex, int ey, int ez) { return dim3((d.nx + ex + 31) / 32, (d.ny + ey + 7) / 8, d.nzl + ez); }
int flat_blocks(long long n) {
	long long b = (n + 255) / 256;
	const long long cap = 148 * 16;
	return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define LAUNCH(S, tag, kernel, grid, block, stream, ...)  \
	do {                                                 \
		const int slot_ = (S)->prof.begin(tag, stream);  \
		kernel<<<grid, block, 0, stream>>>(__VA_ARGS__); \
		(S)->prof.end(slot_, stream);                    \
		(S)->launches++;                                 \
	} while (0)

// persistent grid of a tile kernel: one resident wave, never more CTAs than the level has tiles
template <class K>
int tile_grid(shkz_b200_solver *S, K kernel, dim3 block, int tiles_total, size_t smem = 0) {
	const void *key = reinterpret_cast<const void *>(kernel);
	auto it = S->occupancy.find(key);
	int per_sm;
	if (it == S->occupancy.end()) {
		per_sm = 1;
		if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(block.x * block.y * block.z), smem) != cudaSuccess || per_sm < 1) {
			cudaGetLastError();
			per_sm = 1;
		}
		S->occupancy[key] = per_sm;
	} else per_sm = it->second;
	const long long wave = (long long)per_sm * S->num_sms;
	const long long g = tiles_total < wave ? tiles_total : wave;
	return (int)(g < 1 ? 1 : g);
}
#define LAUNCH_TILES(S, tag, kernel, block, tiles_total, stream, ...) \
	LAUNCH(S, tag, kernel, tile_grid(S, kernel, block, tiles_total), block, stream, __VA_ARGS__)
#define LAUNCH_TILES_SMEM(S, tag, kernel, block, tiles_total, smem, stream, ...)                            \
	do {                                                                                                    \
		const int grid_ = tile_grid(S, kernel, block, tiles_total, smem);                                   \
		const int slot_ = (S)->prof.begin(tag, stream);                                                     \
		kernel<<<grid_, block, smem, stream>>>(__VA_ARGS__);                                                \
		(S)->prof.end(slot_, stream);                                                                       \
		(S)->launches++;                                                                                    \
	} while (0)

// ---- halo exchange of one cell array (ghost planes), a no-op on a whole grid: the own boundary planes are stored
// ---- into the neighbours' ghost planes over NVLink by our own kernel, then the stream waits for theirs
template <class T>
int halo(shkz_b200_solver *S, const Dims &d, T *p, cudaStream_t st) {
	if (S->whole_grid) return SHKZ_B200_OK;
	if (!S->comm || !S->comm->connected()) return fail(SHKZ_B200_ERR_STATE, "slab solver is not connected");
	const void *base = reinterpret_cast<const char *>(p) - (size_t)d.plane * sizeof(T);
	if (!S->comm->owns(base)) return fail(SHKZ_B200_ERR_STATE, "halo exchange of an array outside the slab arena");
	const size_t plane_bytes = (size_t)d.plane * sizeof(T);
	const unsigned long long seq = S->comm->next_exchange();
	const size_t chunks = (plane_bytes + 15) / 16;
	const int blocks = (int)((chunks + 255) / 256 > 296 ? 296 : ((chunks + 255) / 256 < 1 ? 1 : (chunks + 255) / 256));
	LAUNCH(S, "halo", k_halo_push, blocks, 256, st, S->comm->device_view(), S->comm->offset_of(base), plane_bytes, d.nzl, seq);
	return SHKZ_B200_OK;
}

int compact_tiles(shkz_b200_solver *S, HostLevel &H, cudaStream_t stream) {
	LAUNCH(S, H.tag_compact.c_str(), k_compact_tiles, 1, 1024, stream, static_cast<const unsigned char *>(H.tile_flags.base), static_cast<unsigned char *>(H.tile_dirty.base),
	       H.ntx, H.nty, H.nslices_z, H.slice, H.bz, H.adaptive ? 0 : H.bz, 2 * S->num_sms, static_cast<int *>(H.tile_ids.base), static_cast<int *>(H.tile_count.base),
	       static_cast<int *>(H.tile_uids.base), static_cast<int *>(H.tile_ucount.base));
	return SHKZ_B200_OK;
}

// ---- multigrid ----
template <int FIRST, bool ZERO_X, bool PROLONG, bool DOT>
void launch_sweep(shkz_b200_solver *S, const HostLevel &H, const float *xo, float *xn, const float *ec, const Dims &dc, CGState *st, cudaStream_t stream,
                  const SlabSweep sl = SlabSweep{}) {
	const int slab_ghosts = sl.cm != nullptr;
	const MGLevel &L = H.view;
	// profiler entry per variant: "sweep@<level>" + z (first pre-sweep, x_old = 0) / p (prolongation folded in) / d (z.r reduction folded in)
	const std::string tag = H.tag_sweep + (ZERO_X ? "z" : "") + (PROLONG ? "p" : "") + (DOT ? "d" : "");
	if (H.tma && S->sweep_mode == 0) { // operands staged through shared memory by TMA
		SweepMaps maps;
		maps.wx = H.map_wx; maps.wy = H.map_wy; maps.wz = H.map_wz; maps.dd = H.map_dd; maps.b = H.map_b;
		maps.xo = xo == L.xb ? H.map_xb : H.map_xa;
		if (slab_ghosts) { // z-slab: the same kernel also moves the sweep's halo planes
			LAUNCH_TILES_SMEM(S, tag.c_str(), (k_sweep_tma<FIRST, ZERO_X, PROLONG, DOT, true>), dim3(S4_THREADS, 1, 1), H.tiles_total, SWEEP_TMA_SMEM, stream, L.d,
			                  L.tiles, maps, xo, xn, ec, dc, sl, S->redbuf(), st);
			return;
		}
		LAUNCH_TILES_SMEM(S, tag.c_str(), (k_sweep_tma<FIRST, ZERO_X, PROLONG, DOT, false>), dim3(S4_THREADS, 1, 1), H.tiles_total, SWEEP_TMA_SMEM, stream, L.d,
		                  L.tiles, maps, xo, xn, ec, dc, sl, S->redbuf(), st);
	} else if ((L.d.nx & 3) == 0 && S->sweep_mode <= 1) // aligned quads, direct global loads
		LAUNCH_TILES(S, tag.c_str(), (k_sweep4<FIRST, ZERO_X, PROLONG, DOT>), dim3(S4_THREADS, 1, 1), H.tiles_total, stream, L.d, L.tiles, (const float *)L.wx,
		             (const float *)L.wy, (const float *)L.wz, (const float *)L.dd, (const float *)L.b, xo, xn, ec, dc, slab_ghosts, S->redbuf(), st);
	else
		LAUNCH_TILES(S, tag.c_str(), (k_sweep<FIRST, ZERO_X, PROLONG, DOT>), sweep_block(), H.tiles_total, stream, L.d, L.tiles, (const float *)L.wx,
		             (const float *)L.wy, (const float *)L.wz, (const float *)L.dd, (const float *)L.b, xo, xn, ec, dc, slab_ghosts, S->redbuf(), st);
}

// One V-cycle on level l and below, right-hand side in the level's b. *result = buffer holding the solution.
// dot: also reduce (solution . b) into the CG state (level 0 only).
int vcycle_slab(shkz_b200_solver *S, size_t l, const shkz_b200_params &P, CGState *st, cudaStream_t stream, bool dot, const float **result, bool push_result = false);

// global = true: the gathered coarse hierarchy of a z-slab solver (whole-grid semantics on every rank)
int vcycle(shkz_b200_solver *S, size_t l, const shkz_b200_params &P, CGState *st, cudaStream_t stream, bool dot, const float **result, bool global = false) {
	if (!S->whole_grid && !global) return vcycle_slab(S, l, P, st, stream, dot, result);
	std::vector<HostLevel> &lv = global ? S->glevels : S->levels;
	const int tail_first = global ? S->gtail_first : S->tail_first;
	HostLevel &H = lv[l];
	const MGLevel &L = H.view;
	const int coarse = P.mg_coarse_sweeps < 1 ? 1 : P.mg_coarse_sweeps;
	const int mid_first = global ? S->gmid_first : S->mid_first;
	if ((int)l == mid_first && P.mg_gamma <= 1 && S->sweep_mode == 0) {
		// levels [l, tail_first) and the tail: one cooperative launch, one CTA per SM
		MidArgs A{};
		const int end = tail_first >= 0 ? tail_first : (int)lv.size();
		A.nlev = end - (int)l;
		A.pre = P.mg_pre_sweeps < 1 ? 1 : P.mg_pre_sweeps;
		A.post = P.mg_post_sweeps < 0 ? 0 : P.mg_post_sweeps;
		A.coarse = coarse;
		A.has_tail = tail_first >= 0 ? 1 : 0;
		A.dot = (dot && l == 0 && !global) ? 1 : 0;
		A.barrier = static_cast<unsigned *>(S->mid_barrier.base);
		for (int m = 0; m < A.nlev; ++m) {
			const MGLevel &V = lv[l + m].view;
			A.L[m].d = V.d; A.L[m].tiles = V.tiles;
			A.L[m].wx = V.wx; A.L[m].wy = V.wy; A.L[m].wz = V.wz; A.L[m].dd = V.dd;
			A.L[m].b = V.b; A.L[m].x = V.xa;
		}
		TailArgs T = global ? S->gtail_args : S->tail_args;
		T.pre = A.pre; T.post = A.post; T.coarse = coarse;
		RedBuf rb = S->redbuf();
		void *args[] = {&A, &T, &rb, &st};
		const size_t smem = A.has_tail ? (global ? S->gtail_smem : S->tail_smem) : 0;
		const int slot_ = S->prof.begin("vcycle_mid", stream);
		CK(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_vcycle_mid), dim3(S->num_sms), dim3(MID_THREADS), args, smem, stream));
		S->prof.end(slot_, stream);
		S->launches++;
		*result = L.xa;
		return SHKZ_B200_OK;
	}
	if ((int)l == tail_first) {
		TailArgs A = global ? S->gtail_args : S->tail_args;
		A.pre = P.mg_pre_sweeps < 1 ? 1 : P.mg_pre_sweeps;
		A.post = P.mg_post_sweeps < 0 ? 0 : P.mg_post_sweeps;
		A.coarse = coarse;
		const int slot_ = S->prof.begin("vcycle_tail", stream);
		k_vcycle_tail<<<1, TAIL_THREADS, global ? S->gtail_smem : S->tail_smem, stream>>>(A, st);
		S->prof.end(slot_, stream);
		S->launches++;
		*result = L.xa;
		if (dot) LAUNCH_TILES(S, "dot_zb", k_dot_zb, cg_block(), H.tiles_total, stream, L.d, L.tiles, (const float *)L.xa, (const float *)L.b, S->redbuf(), st);
		return SHKZ_B200_OK;
	}
	const bool last = (l + 1 == lv.size());
	int pre = last ? coarse : (P.mg_pre_sweeps < 1 ? 1 : P.mg_pre_sweeps);
	int post = last ? coarse : (P.mg_post_sweeps < 0 ? 0 : P.mg_post_sweeps);
	if (!last && l >= 1) if (const char *e = getenv("SHKZ_B200_COARSE_SWEEPS")) pre = post = atoi(e); // (experiments)
	float *bufs[2] = {L.xa, L.xb};
	const float *cur = nullptr;
	int w = 0;
	const Dims none{};
	for (int sw = 0; sw < pre; ++sw) {
		if (sw == 0) launch_sweep<0, true, false, false>(S, H, nullptr, bufs[w], nullptr, none, st, stream);
		else launch_sweep<0, false, false, false>(S, H, cur, bufs[w], nullptr, none, st, stream);
		cur = bufs[w];
		w ^= 1;
	}
	const float *ec = nullptr;
	Dims dc = none;
	if (!last) {
		const MGLevel &C = lv[l + 1].view;
		dc = C.d;
		// gamma coarse-grid visits (W-cycle: 2) on every level but the finest; the last correction is folded into the first post-sweep
		const int gamma = (l >= 1 && P.mg_gamma > 1) ? P.mg_gamma : 1;
		for (int g = 0; g < gamma; ++g) {
			LAUNCH_TILES(S, H.tag_restrict.c_str(), k_residual_restrict, restrict_block(), H.tiles_total, stream, L.d, L.tiles, (const float *)L.wx, (const float *)L.wy,
			             (const float *)L.wz, (const float *)L.dd, (const float *)L.b, cur, C.d, C.b, (const CGState *)st, (const CommDev *)nullptr, 0ull);
			CKR(vcycle(S, l + 1, P, st, stream, false, &ec, global));
			if (g + 1 < gamma || post == 0) {
				LAUNCH_TILES(S, H.tag_prolong.c_str(), k_prolong_add, dim3(TX, 8, 1), H.tiles_total, stream, L.d, L.tiles, dc, ec, cur, bufs[w], (const CGState *)st, SlabPush{});
				cur = bufs[w];
				w ^= 1;
			}
		}
	}
	for (int sw = 0; sw < post; ++sw) {
		const bool prolong = !last && sw == 0, dotnow = dot && sw + 1 == post;
		if (prolong && dotnow) launch_sweep<1, false, true, true>(S, H, cur, bufs[w], ec, dc, st, stream);
		else if (prolong) launch_sweep<1, false, true, false>(S, H, cur, bufs[w], ec, dc, st, stream);
		else if (dotnow) launch_sweep<1, false, false, true>(S, H, cur, bufs[w], nullptr, none, st, stream);
		else launch_sweep<1, false, false, false>(S, H, cur, bufs[w], nullptr, none, st, stream);
		cur = bufs[w];
		w ^= 1;
	}
	if (dot && post == 0) LAUNCH_TILES(S, "dot_zb", k_dot_zb, cg_block(), H.tiles_total, stream, L.d, L.tiles, cur, (const float *)L.b, S->redbuf(), st);
	*result = cur;
	return SHKZ_B200_OK;
}

// The V-cycle on a z-slab: same kernels, same arithmetic, same result bits as on the whole grid. Before every sweep
// each rank relaxes the first colour of its two boundary planes and stores them into the neighbours' ghost planes
// (k_boundary_half_push); after it, the finished boundary planes follow (halo). The shared-memory tail and the folded
// prolongation are whole-grid features: here every level is swept in place and the correction is added by its own kernel.
// z-slab solvers: the own nzl planes of a slab array -> planes [k0, k0+nzl) of the same-shaped GLOBAL array in every rank's arena;
// the kernel ends with a barrier over all ranks, so afterwards every copy is complete
int gather_planes(shkz_b200_solver *S, const Dims &d, const float *src_plane0, float *dst_plane0, cudaStream_t stream) {
	const size_t plane_bytes = (size_t)d.plane * sizeof(float);
	const size_t src_off = S->comm->offset_of(src_plane0), dst_off = S->comm->offset_of(dst_plane0) + (size_t)d.k0 * plane_bytes;
	const size_t chunks = (plane_bytes * (size_t)d.nzl + 15) / 16;
	const int blocks = (int)((chunks + 255) / 256 > 296 ? 296 : ((chunks + 255) / 256 < 1 ? 1 : (chunks + 255) / 256));
	LAUNCH(S, "gather_planes", k_gather_push, blocks, 256, stream, S->comm->device_view(), src_off, dst_off, plane_bytes * (size_t)d.nzl);
	return SHKZ_B200_OK;
}

// the SlabPush of a kernel that writes cell array `p` of a slab solver (exchange number taken here)
template <class T>
SlabPush slab_push(shkz_b200_solver *S, const Dims &d, T *p) {
	SlabPush sp{};
	if (S->whole_grid) return sp;
	sp.cm = S->comm->device_view();
	sp.off = S->comm->offset_of(reinterpret_cast<const char *>(p) - (size_t)d.plane * sizeof(T));
	sp.seq = S->comm->next_exchange();
	return sp;
}

// a consumer of ghost planes without a wait of its own: settle what the last fused sweep left pending
void settle_pending(shkz_b200_solver *S, cudaStream_t stream) {
	if (!S->pending_wait) return;
	LAUNCH(S, "comm_wait", k_comm_wait, 1, 32, stream, S->comm->device_view(), S->pending_wait);
	S->pending_wait = 0;
}

// One red-black sweep of a slab level. push_out: a later kernel reads the ghost planes of x_new (another sweep or the
// residual), so the boundary planes of x_new travel too. Levels swept by the TMA kernel do all of it in ONE launch
// (SlabSweep, kernels_mg.cuh); the others keep the three-launch form (half-plane push + wait, sweep, halo push + wait).
template <int FIRST>
int slab_sweep(shkz_b200_solver *S, HostLevel &H, bool zero_x, const float *xo, float *xn, bool dot, CGState *st, cudaStream_t stream, bool push_out,
               const float *ec = nullptr, const Dims *dcp = nullptr) { // ec != nullptr (TMA levels only): x_old + P e_coarse folded into this sweep
	const MGLevel &L = H.view;
	const Dims none{};
	const size_t off = S->comm->offset_of(reinterpret_cast<const char *>(xo) - (size_t)L.d.plane * sizeof(float));
	if (H.tma && S->sweep_mode == 0) {
		SlabSweep sl{};
		sl.cm = S->comm->device_view();
		sl.off_xo = off;
		sl.off_xn = S->comm->offset_of(reinterpret_cast<const char *>(xn) - (size_t)L.d.plane * sizeof(float));
		sl.wait_in = zero_x ? 0ull : S->pending_wait; // (x_old = 0: nothing of the ghost planes is read before the half planes arrive)
		sl.seq_half = S->comm->next_exchange();
		sl.seq_out = push_out ? S->comm->next_exchange() : 0ull;
		sl.wx = L.wx; sl.wy = L.wy; sl.wz = L.wz; sl.dd = L.dd; sl.b = L.b;
		if (zero_x) launch_sweep<FIRST, true, false, false>(S, H, xo, xn, nullptr, none, st, stream, sl);
		else if (ec && dot) launch_sweep<FIRST, false, true, true>(S, H, xo, xn, ec, *dcp, st, stream, sl);
		else if (ec) launch_sweep<FIRST, false, true, false>(S, H, xo, xn, ec, *dcp, st, stream, sl);
		else if (dot) launch_sweep<FIRST, false, false, true>(S, H, xo, xn, nullptr, none, st, stream, sl);
		else launch_sweep<FIRST, false, false, false>(S, H, xo, xn, nullptr, none, st, stream, sl);
		S->pending_wait = sl.seq_out;
		return SHKZ_B200_OK;
	}
	settle_pending(S, stream);
	const unsigned long long seq = S->comm->next_exchange();
	const long long pb = (L.d.plane + 255) / 256;
	const dim3 grid((unsigned)(pb > 148 ? 148 : (pb < 1 ? 1 : pb)), 2, 1);
	if (zero_x) LAUNCH(S, "boundary_half", (k_boundary_half_push<FIRST, true>), grid, 256, stream, L.d, (const float *)L.wx, (const float *)L.wy, (const float *)L.wz,
	                   (const float *)L.dd, (const float *)L.b, xo, S->comm->device_view(), off, seq);
	else LAUNCH(S, "boundary_half", (k_boundary_half_push<FIRST, false>), grid, 256, stream, L.d, (const float *)L.wx, (const float *)L.wy, (const float *)L.wz,
	            (const float *)L.dd, (const float *)L.b, xo, S->comm->device_view(), off, seq);
	SlabSweep ghosts{};
	ghosts.cm = S->comm->device_view(); // (k_sweep4 / k_sweep only look at cm != nullptr: the ghost planes of x_old are live)
	if (zero_x) launch_sweep<FIRST, true, false, false>(S, H, xo, xn, nullptr, none, st, stream, ghosts);
	else if (dot) launch_sweep<FIRST, false, false, true>(S, H, xo, xn, nullptr, none, st, stream, ghosts);
	else launch_sweep<FIRST, false, false, false>(S, H, xo, xn, nullptr, none, st, stream, ghosts);
	return push_out ? halo(S, L.d, xn, stream) : SHKZ_B200_OK;
}

// push_result: the caller folds this level's result into its first post-sweep and reads its ghost planes: the last sweep pushes its boundary planes too
int vcycle_slab(shkz_b200_solver *S, size_t l, const shkz_b200_params &P, CGState *st, cudaStream_t stream, bool dot, const float **result, bool push_result) {
	HostLevel &H = S->levels[l];
	const MGLevel &L = H.view;
	if ((int)l == S->agg_level) {
		// every rank stores its planes of this level's right-hand side into every rank's copy of the global level, then all
		// of them run the same whole-grid V-cycle on it; the slab's view of the result starts k0 planes into the global array
		HostLevel &G = S->glevels[0];
		CKR(gather_planes(S, L.d, L.b, G.view.b, stream));
		const float *xg = nullptr;
		CKR(vcycle(S, 0, P, st, stream, false, &xg, true));
		*result = xg + (long long)L.d.k0 * L.d.plane;
		return SHKZ_B200_OK;
	}
	const bool last = (l + 1 == S->levels.size());
	const int coarse = P.mg_coarse_sweeps < 1 ? 1 : P.mg_coarse_sweeps;
	const int pre = last ? coarse : (P.mg_pre_sweeps < 1 ? 1 : P.mg_pre_sweeps);
	const int post = last ? coarse : (P.mg_post_sweeps < 0 ? 0 : P.mg_post_sweeps);
	float *bufs[2] = {L.xa, L.xb};
	const float *cur = nullptr;
	int w = 0;
	bool fold = false;
	const float *fold_ec = nullptr;
	Dims fold_dc{};
	for (int sw = 0; sw < pre; ++sw) {
		// the very first sweep starts from x = 0: the "old" buffer is only a landing place for the neighbours' boundary planes
		CKR(slab_sweep<0>(S, H, sw == 0, sw == 0 ? bufs[w ^ 1] : cur, bufs[w], false, st, stream, true));
		cur = bufs[w];
		w ^= 1;
	}
	if (!last) {
		const MGLevel &C = S->levels[l + 1].view;
		LAUNCH_TILES(S, H.tag_restrict.c_str(), k_residual_restrict, restrict_block(), H.tiles_total, stream, L.d, L.tiles, (const float *)L.wx, (const float *)L.wy,
		             (const float *)L.wz, (const float *)L.dd, (const float *)L.b, cur, C.d, C.b, (const CGState *)st, S->comm->device_view(), S->pending_wait);
		S->pending_wait = 0;
		const float *ec = nullptr;
		// levels swept by the TMA kernel fold "x + P e" into the first post-sweep, ghost planes included (k_sweep_tma<PROLONG, SLAB>): no k_prolong_add launch
		fold = post > 0 && H.tma && S->sweep_mode == 0 && !getenv("SHKZ_B200_NO_SLAB_PROLONG");
		CKR(vcycle_slab(S, l + 1, P, st, stream, false, &ec, fold));
		fold_ec = ec; fold_dc = C.d;
		if (!fold) {
			// (the correction's boundary planes travel with the kernel that writes them; whoever reads the ghost planes next waits)
			const SlabPush sp = post > 0 ? slab_push(S, L.d, bufs[w]) : SlabPush{};
			LAUNCH_TILES(S, H.tag_prolong.c_str(), k_prolong_add, dim3(TX, 8, 1), H.tiles_total, stream, L.d, L.tiles, C.d, ec, cur, bufs[w], (const CGState *)st, sp);
			if (sp.cm) S->pending_wait = sp.seq;
			cur = bufs[w];
			w ^= 1;
		}
	}
	for (int sw = 0; sw < post; ++sw) {
		// (the last sweep of the whole V-cycle pushes its boundary planes too when the fused direction update + product reads the ghost planes of z)
		const bool pro = fold && sw == 0;
		CKR(slab_sweep<1>(S, H, false, cur, bufs[w], dot && sw + 1 == post, st, stream, sw + 1 < post || (dot && l == 0 && S->spmv_tma) || push_result,
		                  pro ? fold_ec : nullptr, pro ? &fold_dc : nullptr));
		cur = bufs[w];
		w ^= 1;
	}
	if (dot && post == 0) LAUNCH_TILES(S, "dot_zb", k_dot_zb, cg_block(), H.tiles_total, stream, L.d, L.tiles, cur, (const float *)L.b, S->redbuf(), st);
	*result = cur;
	return SHKZ_B200_OK;
}

#ifdef SHKZ_B200_TEST_HOOKS
// The same V-cycle with one launch per colour / transfer step on dense grids (validation only).
int legacy_vcycle(shkz_b200_solver *S, size_t l, const shkz_b200_params &P, cudaStream_t stream) {
	HostLevel &H = S->levels[l];
	const MGLevel &L = H.view;
	const bool last = (l + 1 == S->levels.size());
	const int coarse = P.mg_coarse_sweeps < 1 ? 1 : P.mg_coarse_sweeps;
	const int pre = last ? coarse : (P.mg_pre_sweeps < 1 ? 1 : P.mg_pre_sweeps);
	const int post = last ? coarse : (P.mg_post_sweeps < 0 ? 0 : P.mg_post_sweeps);
	const dim3 block(32, 8, 1), grid(((L.d.nx + 1) / 2 + 31) / 32, (L.d.ny + 7) / 8, L.d.nzl);
	const CGState *st = nullptr;
	float *x = L.xa;
	CK(cudaMemsetAsync(H.xa.base, 0, H.xa.bytes, stream)); // x = 0: with omega != 1 the second colour of the first sweep reads its own old value
	for (int sw = 0; sw < pre; ++sw) {
		if (sw == 0) LAUNCH(S, "legacy_rbgs", k_legacy_rbgs<true>, grid, block, stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, 0, st);
		else LAUNCH(S, "legacy_rbgs", k_legacy_rbgs<false>, grid, block, stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, 0, st);
		LAUNCH(S, "legacy_rbgs", k_legacy_rbgs<false>, grid, block, stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, 1, st);
	}
	if (!last) {
		HostLevel &HC = S->levels[l + 1];
		const MGLevel &C = HC.view;
		if (!H.legacy_r.base) CKR(H.legacy_r.alloc(L.d, sizeof(float)));
		float *r = H.legacy_r.ptr<float>(L.d);
		LAUNCH(S, "legacy_residual", k_legacy_residual, legacy_stencil_grid(L.d), legacy_stencil_block(), stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, r, st);
		LAUNCH(S, "legacy_restrict", k_legacy_restrict, cell_grid(C.d, 0, 0, 0), cell_block(), stream, L.d, C.d, r, C.b, st);
		CKR(legacy_vcycle(S, l + 1, P, stream));
		LAUNCH(S, "legacy_prolong", k_legacy_prolong_add, cell_grid(L.d, 0, 0, 0), cell_block(), stream, L.d, C.d, C.xa, x, st);
	}
	for (int sw = 0; sw < post; ++sw) {
		LAUNCH(S, "legacy_rbgs", k_legacy_rbgs<false>, grid, block, stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, 1, st);
		LAUNCH(S, "legacy_rbgs", k_legacy_rbgs<false>, grid, block, stream, L.d, L.wx, L.wy, L.wz, L.dd, L.b, x, 0, st);
	}
	return SHKZ_B200_OK;
}

#endif // SHKZ_B200_TEST_HOOKS

int coarsen_levels(shkz_b200_solver *S, std::vector<HostLevel> &lv, bool slab, const shkz_b200_params &P, cudaStream_t stream) {
	for (size_t l = 0; l + 1 < lv.size(); ++l) {
		const MGLevel &F = lv[l].view;
		HostLevel &HC = lv[l + 1];
		const MGLevel &C = HC.view;
		CK(cudaMemsetAsync(HC.tile_flags.base, 0, (size_t)HC.tiles_total, stream));
		LAUNCH_TILES(S, lv[l].tag_coarsen.c_str(), k_coarsen_operator, restrict_block(), lv[l].tiles_total, stream, F.d, C.d, lv[l].utiles, C.tiles, (float)P.mg_coarse_scale,
		             (const float *)F.wx, (const float *)F.wy, (const float *)F.wz, (const float *)F.dd, C.wx, C.wy, C.wz, C.dd, static_cast<unsigned char *>(HC.tile_flags.base));
		if (slab) CKR(halo(S, C.d, C.wz, stream));
		CKR(compact_tiles(S, HC, stream));
	}
	return SHKZ_B200_OK;
}

int build_hierarchy(shkz_b200_solver *S, const shkz_b200_params &P, cudaStream_t stream) {
	CKR(coarsen_levels(S, S->levels, !S->whole_grid, P, stream));
	if (S->agg_level >= 0) {
		// gather the operator of the first global level from the slabs, then coarsen it on every rank
		const MGLevel &L = S->levels[S->agg_level].view;
		HostLevel &HG = S->glevels[0];
		const MGLevel &G = HG.view;
		CKR(gather_planes(S, L.d, L.wx, G.wx, stream));
		CKR(gather_planes(S, L.d, L.wy, G.wy, stream));
		CKR(gather_planes(S, L.d, L.wz, G.wz, stream));
		CKR(gather_planes(S, L.d, L.dd, G.dd, stream));
		CK(cudaMemsetAsync(HG.tile_flags.base, 0, (size_t)HG.tiles_total, stream));
		LAUNCH(S, "flag_live_tiles", k_flag_live_tiles, cell_grid(G.d, 0, 0, 0), cell_block(), stream, G.d, G.tiles, (const float *)G.wx, (const float *)G.wy,
		       (const float *)G.wz, (const float *)G.dd, static_cast<unsigned char *>(HG.tile_flags.base));
		CKR(compact_tiles(S, HG, stream));
		CKR(coarsen_levels(S, S->glevels, false, P, stream));
	}
	CK(cudaGetLastError());
	S->have_hierarchy = true;
	return SHKZ_B200_OK;
}

// relaxation factor of the sweeps: a per-device constant, set before anything that smooths
int set_omega(const shkz_b200_params &P, cudaStream_t stream) {
	const float w = (float)P.mg_omega;
	CK(cudaMemcpyToSymbolAsync(c_mg_omega, &w, sizeof w, 0, cudaMemcpyHostToDevice, stream));
	return SHKZ_B200_OK;
}

// z-slab solvers, after a stream synchronise: a device-side wait that ran out of time (or a rank that gave up) leaves an abort word in the arena
int comm_health(shkz_b200_solver *S) {
	if (S->whole_grid || !S->comm) return SHKZ_B200_OK;
	const unsigned long long w = S->comm->read_abort();
	if (w) return fail(SHKZ_B200_ERR_COMM, "%s", SlabComm::describe_abort(w).c_str());
	return SHKZ_B200_OK;
}

// The solve kernels only ever write the ACTIVE tiles of x, s, q and of the multigrid buffers and rely on what lies outside them being finite (it meets
// zero coefficients). A solve that ended on NaN / Inf (a non-finite input velocity, say) would leave those values behind in tiles that are dry next time:
// wipe everything such a solve may have touched before the next one.
int scrub_after_nonfinite(shkz_b200_solver *S, cudaStream_t stream) {
	if (!S->poisoned) return SHKZ_B200_OK;
	for (CellArray *a : {&S->x, &S->r, &S->s, &S->s2, &S->q, &S->p_prev})
		if (a->base) CK(cudaMemsetAsync(a->base, 0, a->bytes, stream));
	for (std::vector<HostLevel> *lv : {&S->levels, &S->glevels})
		for (HostLevel &L : *lv) {
			if (L.xa.base) CK(cudaMemsetAsync(L.xa.base, 0, L.xa.bytes, stream));
			if (L.xb.base) CK(cudaMemsetAsync(L.xb.base, 0, L.xb.bytes, stream));
			if (L.own_b && L.b.base) CK(cudaMemsetAsync(L.b.base, 0, L.b.bytes, stream));
		}
	S->poisoned = false;
	return SHKZ_B200_OK;
}

// ---- the CG driver (pcg_solver.h:246-295 with the loop control on the device) ----
template <class VecT, class CoefT>
int solve(shkz_b200_solver *S, const shkz_b200_params &P, cudaStream_t stream) {
	const Dims &d = S->d;
	const RedBuf rb = S->redbuf();
	CGState *st = S->dstate();
	HostLevel &H0 = S->levels[0];
	const Tiles T = H0.view.tiles;
	const int tt = H0.tiles_total;
	VecT *b = S->b.ptr<VecT>(d), *x = S->x.ptr<VecT>(d), *r = S->r.ptr<VecT>(d), *s = S->s.ptr<VecT>(d), *q = S->q.ptr<VecT>(d);
	VecT *sbuf[2] = {s, S->s2.ptr<VecT>(d)};
	int scur = 0; // which of the two direction buffers holds s
	const CoefT *wx = S->wx.ptr<CoefT>(d), *wy = S->wy.ptr<CoefT>(d), *wz = S->wz.ptr<CoefT>(d), *dd = S->dd.ptr<CoefT>(d);
	const bool mg = P.precond == SHKZ_B200_PRECOND_MG;
	constexpr bool kFloatVec = sizeof(VecT) == sizeof(float);
	float *b0 = kFloatVec ? nullptr : H0.view.b; // an all-float CG hands r itself to multigrid
	CKR(scrub_after_nonfinite(S, stream));

	LAUNCH(S, "cg_begin", k_cg_begin, 1, 32, stream, P.residual, (int)P.max_iterations, st);
	if (mg && !kFloatVec) LAUNCH_TILES(S, "cg_init", (k_cg_init<VecT, true>), cg_block(), tt, stream, d, T, (const VecT *)b, x, r, s, b0);
	else LAUNCH_TILES(S, "cg_init", (k_cg_init<VecT, false>), cg_block(), tt, stream, d, T, (const VecT *)b, x, r, s, b0);

	const float *z = nullptr;
	if (mg) CKR(vcycle(S, 0, P, st, stream, true, &z));
	else LAUNCH_TILES(S, "dot_rr", k_dot_rr<VecT>, cg_block(), tt, stream, d, T, (const VecT *)r, rb, st);

	const unsigned check = P.check_every < 1 ? 1 : (unsigned)P.check_every;
	unsigned it = 0;
	// the host looks at the device's convergence flag after a first batch sized by the previous solve, then every `check`
	unsigned batch = S->last_iterations ? S->last_iterations : check;
	while (it < P.max_iterations) {
		for (unsigned c = 0; c < batch && it < P.max_iterations; ++c, ++it) {
			// z-slabs: k_xpay stores the boundary planes of s into the neighbours' ghost planes, the product waits for theirs
			if (mg && S->spmv_tma && S->sweep_mode == 0 && sizeof(CoefT) == sizeof(float) && (S->whole_grid || (P.mg_post_sweeps > 0 && (z == H0.view.xa || z == H0.view.xb)))) {
				// s_new = z + beta s and q = A s_new in ONE TMA-staged launch, s_new into the other direction buffer
				const int zin = z == H0.view.xb ? 1 : 0;
				if (S->whole_grid) {
					LAUNCH_TILES_SMEM(S, "xpay_spmv_dot", (k_xpay_spmv_tma<VecT, false>), dim3(SPMV_TMA_THREADS, 1, 1), tt, SpmvStage<VecT>::SMEM, stream, d, T, S->spmv_maps, scur, zin,
					                  (const VecT *)sbuf[scur], z, sbuf[scur ^ 1], q, rb, st, 0ull);
				} else { // z-slab: the ghost planes of z come with the V-cycle's last sweep (vcycle_slab pushes them for this kernel)
					LAUNCH_TILES_SMEM(S, "xpay_spmv_dot", (k_xpay_spmv_tma<VecT, true>), dim3(SPMV_TMA_THREADS, 1, 1), tt, SpmvStage<VecT>::SMEM, stream, d, T, S->spmv_maps, scur, zin,
					                  (const VecT *)sbuf[scur], z, sbuf[scur ^ 1], q, rb, st, S->pending_wait);
					S->pending_wait = 0;
				}
				scur ^= 1;
				s = sbuf[scur];
				if (kFloatVec) LAUNCH_TILES(S, "axpy2_norm", (k_axpy2_norm<VecT, false, false>), cg_block(), tt, stream, d, T, (const VecT *)s, (const VecT *)q, x, r, b0, rb, st);
				else LAUNCH_TILES(S, "axpy2_norm", (k_axpy2_norm<VecT, false, true>), cg_block(), tt, stream, d, T, (const VecT *)s, (const VecT *)q, x, r, b0, rb, st);
				CKR(vcycle(S, 0, P, st, stream, true, &z));
				continue;
			}
			const SlabPush sp = S->whole_grid ? SlabPush{} : slab_push(S, d, s);
			if (mg) LAUNCH_TILES(S, "xpay", (k_xpay<VecT, float>), cg_block(), tt, stream, d, T, z, s, (const CGState *)st, sp);
			else LAUNCH_TILES(S, "xpay", (k_xpay<VecT, VecT>), cg_block(), tt, stream, d, T, (const VecT *)r, s, (const CGState *)st, sp);
			if ((d.nx & 3) == 0) LAUNCH_TILES(S, "spmv_dot", (k_spmv_dot4<VecT, CoefT>), cg_block4(), tt, stream, d, T, wx, wy, wz, dd, (const VecT *)s, q, rb, st, sp.seq);
			else LAUNCH_TILES(S, "spmv_dot", (k_spmv_dot<VecT, CoefT>), cg_block(), tt, stream, d, T, wx, wy, wz, dd, (const VecT *)s, q, rb, st, sp.seq);
			if (mg) {
				if (kFloatVec) LAUNCH_TILES(S, "axpy2_norm", (k_axpy2_norm<VecT, false, false>), cg_block(), tt, stream, d, T, (const VecT *)s, (const VecT *)q, x, r, b0, rb, st);
				else LAUNCH_TILES(S, "axpy2_norm", (k_axpy2_norm<VecT, false, true>), cg_block(), tt, stream, d, T, (const VecT *)s, (const VecT *)q, x, r, b0, rb, st);
				CKR(vcycle(S, 0, P, st, stream, true, &z));
			} else {
				LAUNCH_TILES(S, "axpy2_norm", (k_axpy2_norm<VecT, true, false>), cg_block(), tt, stream, d, T, (const VecT *)s, (const VecT *)q, x, r, b0, rb, st);
			}
		}
		CK(cudaMemcpyAsync(S->h_state, st, sizeof(CGState), cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		S->prof.collect();
		if (S->h_state->done) break;
		CKR(comm_health(S)); // (a slab whose neighbour never arrived has given up: do not iterate on garbage)
		batch = check;
	}
	CK(cudaMemcpyAsync(S->h_state, st, sizeof(CGState), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	CK(cudaGetLastError());
	S->last_iterations = S->h_state->converged ? (unsigned)S->h_state->iter : 0u;
	{
		const CGState &h = *S->h_state;
		S->poisoned = !(std::isfinite(h.rnorm) && std::isfinite(h.rho) && std::isfinite(h.alpha) && std::isfinite(h.beta) && std::isfinite(h.bnorm));
	}
	return comm_health(S);
}

void fill_stats(const shkz_b200_solver *S, shkz_b200_stats *out) {
	if (!out) return;
	const CGState &h = *S->h_state;
	out->n_rows = h.n_rows; // (z-slab solvers: the count is reduced over all ranks, like every CG scalar)
	out->n_rows_global = h.n_rows;
	out->iterations = (uint32_t)h.iter;
	out->converged = h.converged;
	out->rhs_absmax = h.bnorm;
	out->reresid = h.bnorm > 0 ? h.rnorm / h.bnorm : 0.0;
	out->has_dirichlet = h.has_dirichlet;
	out->mg_levels = (int)(S->levels.size() + (S->glevels.empty() ? 0 : S->glevels.size() - 1));
	out->kernel_launches = S->launches;
	out->mg_mid_level = S->mid_first;
	out->mg_tail_level = S->tail_first;
}

template <class RealT>
int post_impl(shkz_b200_solver *S, void *const vel_v[3], uint8_t *const act[3], const void *solid_v, int width, cudaStream_t stream);

template <class RealT, class VecT, class CoefT>
int project_impl(shkz_b200_solver *S, double dt, void *const vel_v[3], uint8_t *const act[3], const void *solid_v, const void *fluid_v, int fluid_levelset,
                 const shkz_b200_params &P, void *pressure_v, uint8_t *pressure_active, shkz_b200_stats *stats, cudaStream_t stream) {
	const Dims &d = S->d;
	CKR((ensure_precision_arrays<VecT, CoefT>(S, P.precision, P)));
	RealT *phi = S->phi.ptr<RealT>(d);
	RealT *pres = S->pressure.ptr<RealT>(d);
	uint8_t *in_rows = S->in_rows.ptr<uint8_t>(d);
	const RealT *solid = static_cast<const RealT *>(solid_v);
	FaceGrids<RealT> vel;
	ConstFaceGrids<RealT> cvel;
	FaceMasks masks;
	for (int dim = 0; dim < 3; ++dim) {
		vel.p[dim] = static_cast<RealT *>(vel_v[dim]); cvel.p[dim] = vel.p[dim];
		masks.p[dim] = act[dim];
	}
	AsmParams A{};
	A.dt = dt; A.dx = S->dx;
	A.eps_fluid = P.eps_fluid; A.eps_solid = P.eps_solid;
	A.surface_tension = P.surface_tension; A.rhs_correct = P.rhs_correct;
	A.second_order_fluid = P.second_order_fluid; A.second_order_solid = P.second_order_solid;
	A.have_solid = solid != nullptr; A.fluid_levelset = fluid_levelset;
	A.apply_rhs_correct = P.apply_rhs_correct;
	A.off_mark = S->xfer_sparse ? XFER_OFF_MARK : 0;
	S->project_serial++;
	{
		int e = 0;
		A.dx_pow2 = frexp(S->dx, &e) == 0.5 ? 1 : 0;
		A.inv_dx = 1.0 / S->dx;
	}
	S->last_asm = A;
	const RedBuf rb = S->redbuf();
	CGState *st = S->dstate();

	CK(cudaEventRecord(S->ev[0], stream));
	// fluid -> internal array with ghost planes (neighbour slabs fill them)
	CK(cudaMemcpyAsync(phi, fluid_v, sizeof(RealT) * (size_t)d.ncell, cudaMemcpyDeviceToDevice, stream));
	CKR(halo(S, d, phi, stream));
	// a2-a4: the area / liquid fractions are recomputed from the level sets by every kernel that needs them (kernels_assemble.cuh: face_area / face_rho,
	// closed form without a level set); the six face arrays are only materialised when somebody asks for them (shkz_b200_debug_fetch)
	S->fractions_stale = true;
	S->last_solid = solid_v;
	const bool tension = P.surface_tension != 0.0 && A.fluid_levelset; // (without a level set every rho is 1: no face takes the increment, macpressuresolver3.cpp:104-113)
	if (tension) {
		CK(cudaEventRecord(S->ev[8], stream));
		RealT *curv = S->curv.ptr<RealT>(d);
		if (!curv) { CKR(S->curv.alloc(d, sizeof(RealT))); curv = S->curv.ptr<RealT>(d); }
		LAUNCH(S, "curvature", k_curvature<RealT>, cell_grid(d, 0, 0, 0), cell_block(), stream, d, A, (const RealT *)phi, curv);
		CKR(halo(S, d, curv, stream));
		LAUNCH(S, "surface_tension", k_surface_tension<RealT>, cell_grid(d, 1, 1, 1), cell_block(), stream, d, A, (const RealT *)phi, (const RealT *)curv, vel, masks);
		CK(cudaEventRecord(S->ev[9], stream));
	}
	{
		const bool share = sizeof(CoefT) == sizeof(float);
		HostLevel &H0 = S->levels[0];
		const MGLevel &L0 = H0.view;
		CK(cudaMemsetAsync(H0.tile_flags.base, 0, (size_t)H0.tiles_total, stream));
		{
			// persistent over (tile footprint, plane) units
			const long long units = (long long)L0.tiles.ntx * L0.tiles.nty * d.nzl;
#define BUILD_SYSTEM(HS, LS)                                                                                                                                          \
	do {                                                                                                                                                          \
		const int bgrid = tile_grid(S, k_build_system<RealT, CoefT, VecT, HS, LS>, dim3(TX, BS_ROWS, 1), (int)(units > (1ll << 30) ? (1ll << 30) : units));       \
		LAUNCH(S, "build_system", (k_build_system<RealT, CoefT, VecT, HS, LS>), bgrid, dim3(TX, BS_ROWS, 1), stream, d, A, (const RealT *)phi, solid, in_rows,     \
		       cvel, S->wx.ptr<CoefT>(d), S->wy.ptr<CoefT>(d), S->wz.ptr<CoefT>(d), S->dd.ptr<CoefT>(d), share ? nullptr : L0.wx, share ? nullptr : L0.wy,         \
		       share ? nullptr : L0.wz, share ? nullptr : L0.dd, S->b.ptr<VecT>(d), L0.tiles, static_cast<unsigned char *>(H0.tile_flags.base),                   \
		       static_cast<const unsigned char *>(H0.tile_dirty.base), rb, st);                                                                                   \
	} while (0)
			if (A.have_solid && A.fluid_levelset) BUILD_SYSTEM(true, true);
			else if (A.have_solid) BUILD_SYSTEM(true, false);
			else if (A.fluid_levelset) BUILD_SYSTEM(false, true);
			else BUILD_SYSTEM(false, false);
#undef BUILD_SYSTEM
		}
		CKR(compact_tiles(S, H0, stream));
		CKR(halo(S, d, S->wz.ptr<CoefT>(d), stream));
		if (!share) CKR(halo(S, d, L0.wz, stream));
		if (P.warm_start) { // a10: b -= A p_prev (macpressuresolver3.cpp:221-230); the first call finds p_prev = 0
			if (!S->p_prev.base) CKR(S->p_prev.alloc(d, sizeof(VecT), S->arena()));
			VecT *pp = S->p_prev.ptr<VecT>(d);
			CKR(halo(S, d, pp, stream));
			LAUNCH_TILES(S, "warm_rhs", (k_warm_rhs<VecT, CoefT>), cg_block(), H0.tiles_total, stream, d, L0.tiles, (const CoefT *)S->wx.ptr<CoefT>(d),
			             (const CoefT *)S->wy.ptr<CoefT>(d), (const CoefT *)S->wz.ptr<CoefT>(d), (const CoefT *)S->dd.ptr<CoefT>(d), (const VecT *)pp, S->b.ptr<VecT>(d), rb, st);
		}
	}
	CK(cudaGetLastError());
	CK(cudaEventRecord(S->ev[1], stream));
	S->have_hierarchy = false;
	if (P.precond == SHKZ_B200_PRECOND_MG) CKR(build_hierarchy(S, P, stream));
	CK(cudaEventRecord(S->ev[2], stream));
	S->have_system = true;
	S->have_hierarchy = S->have_hierarchy && P.precond == SHKZ_B200_PRECOND_MG;
	CKR(set_omega(P, stream));
	CKR((solve<VecT, CoefT>(S, P, stream)));
	CK(cudaEventRecord(S->ev[3], stream));
	// pressure scatter + velocity update
	if (!S->h_state->has_dirichlet && S->h_state->n_rows && P.precond == SHKZ_B200_PRECOND_MG && P.mg_post_sweeps <= 0) {
		LAUNCH(S, "sum_rows", k_sum_rows<VecT>, flat_blocks(d.ncell), 256, stream, d, (const VecT *)S->x.ptr<VecT>(d), (const uint8_t *)in_rows, rb, st);
	}
	// the caller's grids are cleared wholesale, then written where a tile holds (or held) unknowns
	// (a sparse host call passes the caller's page-locked grids here: off the union list they already hold zeros, shkz_b200.h)
	if (pressure_v && !S->xfer_sparse) CK(cudaMemsetAsync(pressure_v, 0, sizeof(RealT) * (size_t)d.ncell, stream));
	if (pressure_active && !S->xfer_sparse) CK(cudaMemsetAsync(pressure_active, 0, (size_t)d.ncell, stream));
	LAUNCH_TILES(S, "store_pressure", (k_store_pressure<RealT, VecT>), dim3(TX, 8, 1), S->levels[0].tiles_total, stream, d, S->levels[0].utiles, (const VecT *)S->x.ptr<VecT>(d),
	             (const uint8_t *)in_rows, (const CGState *)st, pres, static_cast<RealT *>(pressure_v), pressure_active, P.warm_start ? S->p_prev.ptr<VecT>(d) : (VecT *)nullptr);
	CKR(halo(S, d, pres, stream));
	if (S->xfer_sparse) CK(cudaStreamWaitEvent(stream, S->ev[11], 0)); // the activity masks, uploaded on copy_stream behind the solve
	{
		const dim3 ugrid((d.nx + 1 + 31) / 32, (d.ny + 1 + 8 * UV_ROWS - 1) / (8 * UV_ROWS), d.nzl + 1), ublock(32, 8, 1);
		if (A.have_solid && A.fluid_levelset) LAUNCH(S, "update_velocity", (k_update_velocity<RealT, true, true>), ugrid, ublock, stream, d, A, (const RealT *)phi, solid, (const RealT *)pres, vel, masks);
		else if (A.have_solid) LAUNCH(S, "update_velocity", (k_update_velocity<RealT, true, false>), ugrid, ublock, stream, d, A, (const RealT *)phi, solid, (const RealT *)pres, vel, masks);
		else if (A.fluid_levelset) LAUNCH(S, "update_velocity", (k_update_velocity<RealT, false, true>), ugrid, ublock, stream, d, A, (const RealT *)phi, solid, (const RealT *)pres, vel, masks);
		else LAUNCH(S, "update_velocity", (k_update_velocity<RealT, false, false>), ugrid, ublock, stream, d, A, (const RealT *)phi, solid, (const RealT *)pres, vel, masks);
	}
	if (P.extrapolate_width > 0) CKR((post_impl<RealT>(S, vel_v, act, solid_v, P.extrapolate_width, stream))); // (part of ms_update)
	CK(cudaEventRecord(S->ev[4], stream));
	CK(cudaStreamSynchronize(stream));
	CK(cudaGetLastError());
	S->prof.collect();
	if (stats) {
		fill_stats(S, stats);
		cudaEventElapsedTime(&stats->ms_assemble, S->ev[0], S->ev[1]);
		cudaEventElapsedTime(&stats->ms_setup, S->ev[1], S->ev[2]);
		cudaEventElapsedTime(&stats->ms_solve, S->ev[2], S->ev[3]);
		cudaEventElapsedTime(&stats->ms_update, S->ev[3], S->ev[4]);
		cudaEventElapsedTime(&stats->ms_total, S->ev[0], S->ev[4]);
		if (tension) cudaEventElapsedTime(&stats->ms_surftension, S->ev[8], S->ev[9]);
		int nt[2] = {0, 0};
		const HostLevel &H0 = S->levels[0];
		if (cudaMemcpy(nt, H0.tile_count.base, sizeof nt, cudaMemcpyDeviceToHost) == cudaSuccess && nt[1] > 0) {
			stats->active_tiles = (uint32_t)nt[0];
			stats->tile_depth = (uint32_t)nt[1];
			stats->total_tiles = (uint32_t)(H0.ntx * H0.nty * ((d.nzl + nt[1] - 1) / nt[1]));
		}
	}
	return comm_health(S);
}

// macutility3::extrapolate_and_constrain_velocity (src/utility/macutility3.cpp:89-93) on device arrays, in place
template <class RealT>
int post_impl(shkz_b200_solver *S, void *const vel_v[3], uint8_t *const act[3], const void *solid_v, int width, cudaStream_t stream) {
	const Dims &d = S->d;
	if (!S->whole_grid) return fail(SHKZ_B200_ERR_STATE, "extrapolate_constrain: whole-grid solvers only");
	const dim3 block(32, 8, 1);
	for (int dim = 0; dim < 3; ++dim) {
		const int w = d.nx + (dim == 0), h = d.ny + (dim == 1), dz = d.nzl + (dim == 2);
		const size_t nf = face_count(d, dim);
		if (width > 0 && !S->post_act[dim].base) CKR(S->post_act[dim].alloc(nf));
		uint8_t *cur = act[dim], *other = static_cast<uint8_t *>(S->post_act[dim].base);
		const dim3 grid((w + 31) / 32, (h + 7) / 8, dz);
		for (int round = 0; round < width; ++round) {
			LAUNCH(S, "extrapolate", k_extrapolate_round<RealT>, grid, block, stream, w, h, dz, static_cast<RealT *>(vel_v[dim]), (const uint8_t *)cur, other);
			uint8_t *t = cur; cur = other; other = t;
		}
		if (cur != act[dim]) CK(cudaMemcpyAsync(act[dim], cur, nf, cudaMemcpyDeviceToDevice, stream));
	}
	if (solid_v) { // levelset_exist(solid): the caller passes NULL otherwise, and then not even the wall rule applies (macutility3.cpp:66)
		ConstFaceGrids<RealT> vs;
		FaceMasks as;
		for (int dim = 0; dim < 3; ++dim) {
			const size_t nf = face_count(d, dim);
			if (!S->post_vel[dim].base) { CKR(S->post_vel[dim].alloc(nf * sizeof(RealT))); CKR(S->post_mask[dim].alloc(nf)); }
			CK(cudaMemcpyAsync(S->post_vel[dim].base, vel_v[dim], nf * sizeof(RealT), cudaMemcpyDeviceToDevice, stream));
			CK(cudaMemcpyAsync(S->post_mask[dim].base, act[dim], nf, cudaMemcpyDeviceToDevice, stream));
			vs.p[dim] = static_cast<const RealT *>(S->post_vel[dim].base);
			as.p[dim] = static_cast<uint8_t *>(S->post_mask[dim].base);
		}
		for (int dim = 0; dim < 3; ++dim) {
			const int w = d.nx + (dim == 0), h = d.ny + (dim == 1), dz = d.nzl + (dim == 2);
			LAUNCH(S, "constrain_velocity", k_constrain_velocity<RealT>, dim3((w + 31) / 32, (h + 7) / 8, dz), block, stream, d, S->dx, dim, static_cast<const RealT *>(solid_v), vs, as,
			       static_cast<RealT *>(vel_v[dim]), (const uint8_t *)act[dim]);
		}
	}
	CK(cudaGetLastError());
	return SHKZ_B200_OK;
}

typedef int (*project_fn)(shkz_b200_solver *, double, void *const[3], uint8_t *const[3], const void *, const void *, int, const shkz_b200_params &, void *,
                          uint8_t *, shkz_b200_stats *, cudaStream_t);

project_fn pick_project(int real, int precision) {
	if (real == SHKZ_B200_REAL_F32) {
		if (precision == SHKZ_B200_PREC_FP64) return project_impl<float, double, double>;
		if (precision == SHKZ_B200_PREC_MIXED) return project_impl<float, double, float>;
		if (precision == SHKZ_B200_PREC_FP32) return project_impl<float, float, float>;
	} else if (real == SHKZ_B200_REAL_F64) {
		if (precision == SHKZ_B200_PREC_FP64) return project_impl<double, double, double>;
		if (precision == SHKZ_B200_PREC_MIXED) return project_impl<double, double, float>;
		if (precision == SHKZ_B200_PREC_FP32) return project_impl<double, float, float>;
	}
	return nullptr;
}

int check_params(const shkz_b200_params *p, shkz_b200_params &out) {
	if (!p) {
		shkz_b200_default_params(&out);
		return SHKZ_B200_OK;
	}
	if (p->struct_size != sizeof(shkz_b200_params)) return fail(SHKZ_B200_ERR_ARG, "params.struct_size %u != %zu", p->struct_size, sizeof(shkz_b200_params));
	out = *p;
	if (out.precond != SHKZ_B200_PRECOND_NONE && out.precond != SHKZ_B200_PRECOND_MG) return fail(SHKZ_B200_ERR_ARG, "unknown precond %d", out.precond);
	if (out.precision < 0 || out.precision > 2) return fail(SHKZ_B200_ERR_ARG, "unknown precision %d", out.precision);
	if (!(out.mg_coarse_scale > 0.0)) return fail(SHKZ_B200_ERR_ARG, "mg_coarse_scale must be > 0");
	if (!(out.mg_omega > 0.0 && out.mg_omega < 2.0)) return fail(SHKZ_B200_ERR_ARG, "mg_omega must be in (0, 2)");
	return SHKZ_B200_OK;
}

// device staging arrays of the host-buffer entry points (allocated on first use, or ahead of time by shkz_b200_prepare)
static int ensure_host_staging(shkz_b200_solver *S, bool have_solid) {
	const Dims &d = S->d;
	const size_t rb = S->real_bytes, ncell = (size_t)d.ncell, nodal = (size_t)(d.nx + 1) * (d.ny + 1) * (d.nzl + 1);
	for (int dim = 0; dim < 3; ++dim) {
		const size_t nf = face_count(d, dim);
		if (!S->st_vel[dim].base) { CKR(S->st_vel[dim].alloc(nf * rb)); CKR(S->st_act[dim].alloc(nf)); }
	}
	if (!S->st_fluid.base) { CKR(S->st_fluid.alloc(ncell * rb)); CKR(S->st_pressure.alloc(ncell * rb)); CKR(S->st_pact.alloc(ncell)); }
	if (have_solid && !S->st_solid.base) CKR(S->st_solid.alloc(nodal * rb));
	return SHKZ_B200_OK;
}

// device-visible address of a page-locked host buffer; nullptr for pageable memory (and for NULL)
static void *mapped_host(const void *p) {
	if (!p) return nullptr;
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// The sparse upload of a host call (kernels_xfer.cuh), after the liquid level set has been copied into its staging array: flag the wet transfer blocks, and — unless
// they are most of the grid — pull velocity and solid nodes around them straight out of the caller's page-locked buffers and send the masks after them on
// copy_stream. *sparse = false leaves everything but the level set to the whole-array copies. pulled = bytes read from host memory by k_pull_slices.
template <class RealT>
static int host_sparse_upload(shkz_b200_solver *S, void *const hvel[3], uint8_t *const vel_active[3], const void *hsolid, cudaStream_t stream, bool *sparse, uint64_t *pulled,
                              void *const masked[3]) { // masked: device-visible addresses of the caller's masks when params.velocity_masked, else null
	const Dims &d = S->d;
	XferGeom g{};
	g.ntx = (d.nx + TX - 1) / TX; g.nty = (d.ny + TY - 1) / TY;
	g.slice = d.nzl < 8 ? d.nzl : 8;
	g.nslices_z = (d.nzl + g.slice - 1) / g.slice;
	const size_t nslices = (size_t)g.ntx * g.nty * g.nslices_z;
	if (nslices * sizeof(int) > S->xf_flags.bytes || !S->h_xf || !S->copy_stream) return fail(SHKZ_B200_ERR_STATE, "sparse host copies: bookkeeping arrays missing");
	CK(cudaMemsetAsync(S->xf_flags.base, 0, nslices * sizeof(int), stream));
	CK(cudaMemsetAsync(S->xf_count.base, 0, sizeof(int), stream));
	const long long units = (long long)g.ntx * g.nty * d.nzl;
	const int fgrid = (int)(units < 8ll * S->num_sms ? units : 8ll * S->num_sms);
	LAUNCH(S, "flag_wet_slices", k_flag_wet_slices<RealT>, fgrid, dim3(TX, 4, 1), stream, d, g, static_cast<const RealT *>(S->st_fluid.base), static_cast<int *>(S->xf_flags.base),
	       static_cast<int *>(S->xf_list.base), static_cast<int *>(S->xf_count.base));
	S->h_xf[0] = 0;
	CK(cudaMemcpyAsync(&S->h_xf[0], S->xf_count.base, sizeof(int), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	const size_t wet = (size_t)(S->h_xf[0] & 0xffffffffull);
	*sparse = 2 * wet <= nslices; // (beyond that the copy engines' whole-array streams are the faster way)
	if (!*sparse) return SHKZ_B200_OK;
	if (wet) {
		ConstFaceGrids<RealT> hv;
		FaceGrids<RealT> dv;
		for (int dim = 0; dim < 3; ++dim) {
			hv.p[dim] = static_cast<const RealT *>(hvel[dim]);
			dv.p[dim] = static_cast<RealT *>(S->st_vel[dim].base);
		}
		const int pgrid = (int)(wet < 8ull * S->num_sms ? wet : 8ull * S->num_sms);
		FaceMasks hm;
		for (int dim = 0; dim < 3; ++dim) hm.p[dim] = masked ? static_cast<uint8_t *>(masked[dim]) : nullptr;
		LAUNCH(S, "pull_slices", k_pull_slices<RealT>, pgrid, 256, stream, d, g, static_cast<const int *>(S->xf_list.base), static_cast<const int *>(S->xf_count.base), hv, dv,
		       static_cast<const RealT *>(hsolid), static_cast<RealT *>(S->st_solid.base), hm);
	}
	if (!S->whole_grid) {
		// z-slab: the z faces of plane 0 and plane nzl are shared with the neighbouring slabs and finished by BOTH sides (the host keeps the upper slab's copy):
		// one of them may lie between a dry cell of this slab and a wet cell of the neighbour, which no block of this slab's list covers. Those two face planes
		// and the node planes their area fractions read travel whole (a few MB).
		const size_t rbs = sizeof(RealT), fplane = (size_t)d.plane, nplane = (size_t)(d.nx + 1) * (d.ny + 1);
		const RealT *hw = static_cast<const RealT *>(hvel[2]);
		RealT *dw = static_cast<RealT *>(S->st_vel[2].base);
		CK(cudaMemcpyAsync(dw, hw, fplane * rbs, cudaMemcpyHostToDevice, stream));
		CK(cudaMemcpyAsync(dw + fplane * d.nzl, hw + fplane * d.nzl, fplane * rbs, cudaMemcpyHostToDevice, stream));
		*pulled += 2 * fplane * rbs;
		if (masked) {
			const uint8_t *hm = static_cast<const uint8_t *>(masked[2]);
			LAUNCH(S, "zero_inactive", k_zero_inactive<RealT>, flat_blocks((long long)(fplane + 3) / 4), 256, stream, (long long)fplane, dw, hm);
			LAUNCH(S, "zero_inactive", k_zero_inactive<RealT>, flat_blocks((long long)(fplane + 3) / 4), 256, stream, (long long)fplane, dw + fplane * d.nzl, hm + fplane * d.nzl);
		}
		if (hsolid) {
			const RealT *hs = static_cast<const RealT *>(hsolid);
			RealT *ds = static_cast<RealT *>(S->st_solid.base);
			CK(cudaMemcpyAsync(ds, hs, nplane * rbs, cudaMemcpyHostToDevice, stream));
			CK(cudaMemcpyAsync(ds + nplane * d.nzl, hs + nplane * d.nzl, nplane * rbs, cudaMemcpyHostToDevice, stream));
			*pulled += 2 * nplane * rbs;
		}
	}
	const uint64_t per_block = (uint64_t)((TX + 1) * TY * g.slice + TX * (TY + 1) * g.slice + TX * TY * (g.slice + 1) + (hsolid ? (TX + 1) * (TY + 1) * (g.slice + 1) : 0));
	*pulled += (uint64_t)wet * per_block * sizeof(RealT); // (blocks on the grid's upper edges are smaller: an upper bound by a few percent)
	if (masked) *pulled += (uint64_t)wet * (uint64_t)((TX + 1) * TY * g.slice + TX * (TY + 1) * g.slice + TX * TY * (g.slice + 1)); // the mask bytes of the same faces
	// the masks are first needed by the velocity update: they cross PCIe after the pull and behind assembly and solve
	CK(cudaEventRecord(S->ev[10], stream));
	CK(cudaStreamWaitEvent(S->copy_stream, S->ev[10], 0));
	for (int dim = 0; dim < 3; ++dim) CK(cudaMemcpyAsync(S->st_act[dim].base, vel_active[dim], face_count(d, dim), cudaMemcpyHostToDevice, S->copy_stream));
	CK(cudaEventRecord(S->ev[11], S->copy_stream));
	return SHKZ_B200_OK;
}

template <class RealT>
static int host_sparse_download(shkz_b200_solver *S, void *const hvel[3], uint8_t *const hact[3], cudaStream_t stream) {
	const Dims &d = S->d;
	CK(cudaMemsetAsync(S->xf_pushed.base, 0, sizeof(unsigned long long), stream));
	for (int dim = 0; dim < 3; ++dim) {
		const long long nf = (long long)face_count(d, dim);
		LAUNCH(S, "push_faces", k_push_faces<RealT>, flat_blocks((nf + 3) / 4), 256, stream, nf, static_cast<const RealT *>(S->st_vel[dim].base),
		       static_cast<uint8_t *>(S->st_act[dim].base), static_cast<RealT *>(hvel[dim]), hact[dim], static_cast<unsigned long long *>(S->xf_pushed.base));
	}
	CK(cudaMemcpyAsync(&S->h_xf[1], S->xf_pushed.base, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
	return SHKZ_B200_OK;
}

int device_ready(int device) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return fail(SHKZ_B200_ERR_NO_DEVICE, "no CUDA device available (%s); libshkz_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
	}
	if (device < 0 || device >= n) return fail(SHKZ_B200_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
	return SHKZ_B200_OK;
}

} // namespace

// =============================================== C-ABI ===============================================
extern "C" {

int shkz_b200_abi_version(void) { return SHKZ_B200_ABI_VERSION; }

const char *shkz_b200_last_error(void) { return g_error.c_str(); }

void shkz_b200_default_params(shkz_b200_params *p) {
	if (!p) return;
	memset(p, 0, sizeof *p);
	p->struct_size = sizeof *p;
	p->second_order_fluid = 1;
	p->second_order_solid = 1;
	p->eps_fluid = 1e-2;
	p->eps_solid = 1e-2;
	p->residual = 1e-4;
	p->max_iterations = 30000;
	p->precond = SHKZ_B200_PRECOND_MG;
	p->precision = SHKZ_B200_PREC_MIXED;
	p->mg_pre_sweeps = 2;
	p->mg_post_sweeps = 2;
	p->mg_coarse_sweeps = 8;
	p->mg_min_size = 4;
	p->check_every = 4;
	p->mg_coarse_scale = 0.5;
	p->mg_gamma = 1;
	p->mg_omega = 1.15;
}

int shkz_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int shkz_b200_create_slab(int nx, int ny, int nz, int k0, int k1, double dx, int real, int device, shkz_b200_solver **out) {
	if (!out) return fail(SHKZ_B200_ERR_ARG, "out is NULL");
	*out = nullptr;
	if (nx < 1 || ny < 1 || nz < 1 || k0 < 0 || k1 > nz || k1 <= k0) return fail(SHKZ_B200_ERR_ARG, "bad grid %dx%dx%d slab [%d,%d)", nx, ny, nz, k0, k1);
	if (!(dx > 0.0)) return fail(SHKZ_B200_ERR_ARG, "dx must be positive");
	if (real != SHKZ_B200_REAL_F32 && real != SHKZ_B200_REAL_F64) return fail(SHKZ_B200_ERR_ARG, "unknown real type %d", real);
	if ((long long)nz + 2 > 65535) return fail(SHKZ_B200_ERR_ARG, "nz too large for the launch geometry");
	CKR(device_ready(device));
	CK(cudaSetDevice(device));
	int num_sms = 0;
	CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
	shkz_b200_solver *S = new shkz_b200_solver();
	S->num_sms = num_sms > 0 ? num_sms : 148;
	if (const char *e = getenv("SHKZ_B200_SWEEP_MODE")) S->sweep_mode = atoi(e); // A-B timing of the sweep kernels
	S->d = make_dims(nx, ny, k1 - k0, k0, nz);
	S->dx = dx;
	S->real = real;
	S->device = device;
	S->real_bytes = real == SHKZ_B200_REAL_F64 ? 8 : 4;
	S->whole_grid = (k0 == 0 && k1 == nz);
	const Dims &d = S->d;
	int rc = SHKZ_B200_OK;
	auto tryalloc = [&](int r) { if (rc == SHKZ_B200_OK) rc = r; };
	if (!S->whole_grid) {
		// one arena for every ghosted array (see slab_comm.h); generous bound on what any precision mode carves
		if (nz % (k1 - k0) != 0 || k0 % (k1 - k0) != 0) rc = fail(SHKZ_B200_ERR_ARG, "z-slabs must be equal: nz %d, slab [%d,%d)", nz, k0, k1);
		S->comm = new SlabComm(device);
		const size_t cells = (size_t)(d.nzl + 2) * (size_t)d.plane;
		if (rc == SHKZ_B200_OK && S->comm->create_arena(ARENA_HEADER + cells * 136 + (size_t)160 * 1024 * 1024)) rc = fail(SHKZ_B200_ERR_CUDA, "%s", S->comm->error());
	}
	tryalloc(S->phi.alloc(d, S->real_bytes, S->arena()));
	tryalloc(S->pressure.alloc(d, S->real_bytes, S->arena()));
	tryalloc(S->in_rows.alloc(d, 1, S->arena()));
	if (!S->whole_grid) {
		tryalloc(S->curv.alloc(d, S->real_bytes, S->arena())); // (a whole grid allocates it on first use)
		S->arena_mark = S->comm->mark();
	}
	// reduction scratch: the largest grid any reducing kernel uses
	{
		const dim3 g2 = cell_grid(d, 1, 1, 1);
		S->max_blocks = (size_t)g2.x * g2.y * g2.z + 148 * 16;
		tryalloc(S->partials.alloc(S->max_blocks * 4 * sizeof(double)));
		tryalloc(S->counter.alloc(64));
		tryalloc(S->state.alloc(sizeof(CGState)));
	}
	// bookkeeping of the sparse host copies (kernels_xfer.cuh). Allocated HERE, not on first use: a page-locked allocation synchronises every device of the
	// process, and a z-slab solver's first host call may run while a neighbouring slab's kernel already spins on this rank (one process, N host threads).
	{
		const size_t nslices = (size_t)((d.nx + TX - 1) / TX) * ((d.ny + TY - 1) / TY) * ((d.nzl + 7) / 8 + 1);
		tryalloc(S->xf_flags.alloc(nslices * sizeof(int)));
		tryalloc(S->xf_list.alloc(nslices * sizeof(int)));
		tryalloc(S->xf_count.alloc(sizeof(int)));
		tryalloc(S->xf_pushed.alloc(sizeof(unsigned long long)));
		if (rc == SHKZ_B200_OK && cudaHostAlloc(reinterpret_cast<void **>(&S->h_xf), 2 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess)
			rc = fail(SHKZ_B200_ERR_CUDA, "cudaHostAlloc failed");
		if (rc == SHKZ_B200_OK && cudaStreamCreateWithFlags(&S->copy_stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(SHKZ_B200_ERR_CUDA, "cudaStreamCreate failed");
	}
	if (rc == SHKZ_B200_OK && cudaMallocHost((void **)&S->h_state, sizeof(CGState)) != cudaSuccess) rc = fail(SHKZ_B200_ERR_CUDA, "cudaMallocHost failed");
	if (rc == SHKZ_B200_OK) {
		memset(S->h_state, 0, sizeof(CGState));
		for (auto &e : S->ev)
			if (cudaEventCreate(&e) != cudaSuccess) rc = fail(SHKZ_B200_ERR_CUDA, "cudaEventCreate failed");
		S->events = true;
	}
	if (rc != SHKZ_B200_OK) {
		shkz_b200_destroy(S);
		return rc;
	}
	*out = S;
	return SHKZ_B200_OK;
}

int shkz_b200_create(int nx, int ny, int nz, double dx, int real, int device, shkz_b200_solver **out) {
	return shkz_b200_create_slab(nx, ny, nz, 0, nz, dx, real, device, out);
}

void shkz_b200_destroy(shkz_b200_solver *S) {
	if (!S) return;
	cudaSetDevice(S->device);
	cudaDeviceSynchronize();
	if (S->comm) { delete S->comm; S->comm = nullptr; }
	release_precision_arrays(S);
	S->phi.release(); S->pressure.release(); S->curv.release(); S->in_rows.release();
	for (int dim = 0; dim < 3; ++dim) {
		S->areas[dim].release(); S->rhos[dim].release(); S->st_vel[dim].release(); S->st_act[dim].release();
		S->post_act[dim].release(); S->post_vel[dim].release(); S->post_mask[dim].release();
	}
	S->st_solid.release(); S->st_fluid.release(); S->st_pressure.release(); S->st_pact.release();
	S->xf_flags.release(); S->xf_list.release(); S->xf_count.release(); S->xf_pushed.release();
	S->partials.release(); S->counter.release(); S->state.release(); S->mid_barrier.release();
	if (S->h_state) cudaFreeHost(S->h_state);
	if (S->h_xf) cudaFreeHost(S->h_xf);
	if (S->copy_stream) cudaStreamDestroy(S->copy_stream);
	if (S->events) for (auto &e : S->ev) if (e) cudaEventDestroy(e);
	S->prof.destroy();
	delete S;
}

int shkz_b200_project_device(shkz_b200_solver *S, double dt, void *const vel[3], uint8_t *const vel_active[3], const void *solid, const void *fluid,
                             int fluid_levelset, const shkz_b200_params *params, void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats,
                             void *cuda_stream) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!vel || !vel_active || !fluid) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active / fluid must not be NULL");
	for (int dim = 0; dim < 3; ++dim)
		if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
	if (!S->whole_grid && !(S->comm && S->comm->connected())) return fail(SHKZ_B200_ERR_STATE, "slab solver is not connected (call shkz_b200_slab_connect)");
	shkz_b200_params P;
	CKR(check_params(params, P));
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	project_fn fn = pick_project(S->real, P.precision);
	if (!fn) return fail(SHKZ_B200_ERR_ARG, "unsupported real/precision combination");
	if (stats) memset(stats, 0, sizeof *stats);
	S->launches = 0;
	if (P.velocity_masked && !S->xfer_sparse) {
		// the entries of inactive faces are unspecified: make them what an inactive face reads as before anything looks at them (a sparse host call has done
		// this while pulling the values it needs, kernels_xfer.cuh)
		cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
		for (int dim = 0; dim < 3; ++dim) {
			const long long nf = (long long)face_count(S->d, dim);
			if (S->real == SHKZ_B200_REAL_F32) { LAUNCH(S, "zero_inactive", k_zero_inactive<float>, flat_blocks((nf + 3) / 4), 256, stream, nf, static_cast<float *>(vel[dim]), (const uint8_t *)vel_active[dim]); }
			else { LAUNCH(S, "zero_inactive", k_zero_inactive<double>, flat_blocks((nf + 3) / 4), 256, stream, nf, static_cast<double *>(vel[dim]), (const uint8_t *)vel_active[dim]); }
		}
	}
	const int rc = fn(S, dt, vel, vel_active, solid, fluid, fluid_levelset, P, pressure, pressure_active, stats, static_cast<cudaStream_t>(cuda_stream));
	// a z-slab call that fails on this rank's host must not leave the other ranks' kernels spinning on planes that will never arrive
	if (rc != SHKZ_B200_OK && rc != SHKZ_B200_ERR_COMM && !S->whole_grid && S->comm) S->comm->raise_abort();
	return rc;
}

static int project_host_impl(shkz_b200_solver *S, double dt, void *const vel[3], uint8_t *const vel_active[3], const void *solid, const void *fluid,
                             int fluid_levelset, const shkz_b200_params *params, void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats);

int shkz_b200_project_host(shkz_b200_solver *S, double dt, void *const vel[3], uint8_t *const vel_active[3], const void *solid, const void *fluid,
                           int fluid_levelset, const shkz_b200_params *params, void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	const int rc = project_host_impl(S, dt, vel, vel_active, solid, fluid, fluid_levelset, params, pressure, pressure_active, stats);
	if (rc != SHKZ_B200_OK && rc != SHKZ_B200_ERR_COMM && !S->whole_grid && S->comm) S->comm->raise_abort(); // (see shkz_b200_project_device)
	return rc;
}

static int project_host_impl(shkz_b200_solver *S, double dt, void *const vel[3], uint8_t *const vel_active[3], const void *solid, const void *fluid,
                             int fluid_levelset, const shkz_b200_params *params, void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats) {
	if (!vel || !vel_active || !fluid) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active / fluid must not be NULL");
	for (int dim = 0; dim < 3; ++dim)
		if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
	shkz_b200_params P;
	CKR(check_params(params, P));
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	const Dims &d = S->d;
	const size_t rb = S->real_bytes;
	const size_t nodal = (size_t)(d.nx + 1) * (d.ny + 1) * (d.nzl + 1);
	const size_t ncell = (size_t)d.ncell;
	cudaStream_t stream = nullptr;
	void *dvel[3];
	uint8_t *dact[3];
	CKR(ensure_host_staging(S, solid != nullptr));
	for (int dim = 0; dim < 3; ++dim) {
		dvel[dim] = S->st_vel[dim].base;
		dact[dim] = static_cast<uint8_t *>(S->st_act[dim].base);
	}

	// Sparse copies (kernels_xfer.cuh) need a liquid level set (without one every cell is wet), nothing that reads the velocity away from wet cells
	// (surface tension walks the masks early, the extrapolation fills the whole band), and every buffer addressable from the device.
	void *mvel[3] = {nullptr, nullptr, nullptr}, *mact[3] = {nullptr, nullptr, nullptr}, *msolid = nullptr, *mpres = nullptr, *mpact = nullptr;
	bool sparse = fluid_levelset && P.surface_tension == 0.0 && P.extrapolate_width <= 0;
	if (const char *e = getenv("SHKZ_B200_HOST_COPIES")) sparse = sparse && strcmp(e, "dense") != 0;
	if (sparse) {
		for (int dim = 0; dim < 3 && sparse; ++dim) sparse = (mvel[dim] = mapped_host(vel[dim])) && (mact[dim] = mapped_host(vel_active[dim]));
		if (sparse && solid) sparse = (msolid = mapped_host(solid)) != nullptr;
		if (sparse && pressure) sparse = (mpres = mapped_host(pressure)) != nullptr;
		if (sparse && pressure_active) sparse = (mpact = mapped_host(pressure_active)) != nullptr;
	}
	uint64_t h2d = 0, d2h = 0, pulled = 0;
	S->launches = 0;
	CK(cudaEventRecord(S->ev[5], stream));
	CK(cudaMemcpyAsync(S->st_fluid.base, fluid, ncell * rb, cudaMemcpyHostToDevice, stream));
	h2d += ncell * rb;
	if (sparse) {
		if (S->real == SHKZ_B200_REAL_F32) CKR((host_sparse_upload<float>(S, mvel, vel_active, msolid, stream, &sparse, &pulled, P.velocity_masked ? mact : nullptr)));
		else CKR((host_sparse_upload<double>(S, mvel, vel_active, msolid, stream, &sparse, &pulled, P.velocity_masked ? mact : nullptr)));
	}
	for (int dim = 0; dim < 3; ++dim) h2d += face_count(d, dim);
	if (sparse) {
		h2d += pulled;
		// everything off the tiles the previous sparse call wrote is zero in the caller's pressure grids — if they are the same grids and no other projection
		// has moved the solver's tile lists since; otherwise clear them once
		const bool known = S->host_serial == S->project_serial;
		if (pressure && !(known && S->last_host_pressure == mpres)) {
			if (rb == 4) LAUNCH(S, "fill_zero", k_fill_zero<float>, flat_blocks((long long)ncell), 256, stream, static_cast<float *>(mpres), (long long)ncell);
			else LAUNCH(S, "fill_zero", k_fill_zero<double>, flat_blocks((long long)ncell), 256, stream, static_cast<double *>(mpres), (long long)ncell);
			d2h += ncell * rb;
		}
		if (pressure_active && !(known && S->last_host_pact == mpact)) {
			LAUNCH(S, "fill_zero", k_fill_zero<uint8_t>, flat_blocks((long long)ncell), 256, stream, static_cast<uint8_t *>(mpact), (long long)ncell);
			d2h += ncell;
		}
	} else {
		for (int dim = 0; dim < 3; ++dim) {
			const size_t nf = face_count(d, dim);
			CK(cudaMemcpyAsync(S->st_vel[dim].base, vel[dim], nf * rb, cudaMemcpyHostToDevice, stream));
			CK(cudaMemcpyAsync(S->st_act[dim].base, vel_active[dim], nf, cudaMemcpyHostToDevice, stream));
			h2d += nf * rb;
		}
		if (solid) {
			CK(cudaMemcpyAsync(S->st_solid.base, solid, nodal * rb, cudaMemcpyHostToDevice, stream));
			h2d += nodal * rb;
		}
	}
	S->last_host_pressure = S->last_host_pact = nullptr;
	CK(cudaEventRecord(S->ev[6], stream));
	const uint64_t setup_launches = S->launches;
	S->xfer_sparse = sparse;
	int rc = shkz_b200_project_device(S, dt, dvel, dact, solid ? S->st_solid.base : nullptr, S->st_fluid.base, fluid_levelset, params, sparse ? mpres : S->st_pressure.base,
	                                  sparse ? static_cast<uint8_t *>(mpact) : static_cast<uint8_t *>(S->st_pact.base), stats, stream);
	S->xfer_sparse = false;
	if (rc != SHKZ_B200_OK) {
		if (sparse) cudaStreamSynchronize(S->copy_stream); // (nothing of this call may still be reading the caller's buffers)
		return rc;
	}
	CK(cudaEventRecord(S->ev[6 + 1], stream));
	const uint64_t before_push = S->launches;
	if (sparse) {
		if (S->real == SHKZ_B200_REAL_F32) CKR((host_sparse_download<float>(S, mvel, reinterpret_cast<uint8_t *const *>(mact), stream)));
		else CKR((host_sparse_download<double>(S, mvel, reinterpret_cast<uint8_t *const *>(mact), stream)));
	} else {
		for (int dim = 0; dim < 3; ++dim) {
			const size_t nf = face_count(d, dim);
			CK(cudaMemcpyAsync(vel[dim], S->st_vel[dim].base, nf * rb, cudaMemcpyDeviceToHost, stream));
			CK(cudaMemcpyAsync(vel_active[dim], S->st_act[dim].base, nf, cudaMemcpyDeviceToHost, stream));
			d2h += nf * (rb + 1);
		}
		if (pressure) { CK(cudaMemcpyAsync(pressure, S->st_pressure.base, ncell * rb, cudaMemcpyDeviceToHost, stream)); d2h += ncell * rb; }
		if (pressure_active) { CK(cudaMemcpyAsync(pressure_active, S->st_pact.base, ncell, cudaMemcpyDeviceToHost, stream)); d2h += ncell; }
	}
	CK(cudaEventRecord(S->ev[0], stream));
	CK(cudaStreamSynchronize(stream));
	if (sparse) {
		d2h += S->h_xf[1];
		int nu[2] = {0, 0};
		const HostLevel &H0 = S->levels[0];
		CK(cudaMemcpy(nu, H0.tile_ucount.base, sizeof nu, cudaMemcpyDeviceToHost)); // union tiles k_store_pressure wrote, and their depth
		uint64_t cells = (uint64_t)nu[0] * TX * TY * (uint64_t)(nu[1] > 0 ? nu[1] : H0.bz);
		if (cells > ncell) cells = ncell;
		d2h += cells * ((pressure ? rb : 0) + (pressure_active ? 1 : 0));
		S->last_host_pressure = mpres;
		S->last_host_pact = mpact;
		S->host_serial = S->project_serial;
	}
	if (stats) {
		cudaEventElapsedTime(&stats->ms_h2d, S->ev[5], S->ev[6]);
		cudaEventElapsedTime(&stats->ms_d2h, S->ev[7], S->ev[0]);
		stats->ms_total += stats->ms_h2d + stats->ms_d2h;
		stats->host_copies = sparse ? 1 : 0;
		stats->h2d_bytes = h2d;
		stats->d2h_bytes = d2h;
		stats->kernel_launches += setup_launches + (S->launches - before_push);
	}
	return SHKZ_B200_OK;
}

int shkz_b200_prepare(shkz_b200_solver *S, const shkz_b200_params *params, int host_buffers, int have_solid) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	shkz_b200_params P;
	CKR(check_params(params, P));
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	if (P.precision == SHKZ_B200_PREC_FP64) CKR((ensure_precision_arrays<double, double>(S, P.precision, P)));
	else if (P.precision == SHKZ_B200_PREC_MIXED) CKR((ensure_precision_arrays<double, float>(S, P.precision, P)));
	else CKR((ensure_precision_arrays<float, float>(S, P.precision, P)));
	if (P.warm_start && !S->p_prev.base) CKR(S->p_prev.alloc(S->d, P.precision == SHKZ_B200_PREC_FP32 ? sizeof(float) : sizeof(double), S->arena()));
	if (host_buffers) CKR(ensure_host_staging(S, have_solid != 0));
	CK(cudaDeviceSynchronize());
	return SHKZ_B200_OK;
}

int shkz_b200_extrapolate_constrain_device(shkz_b200_solver *S, void *const vel[3], uint8_t *const vel_active[3], const void *solid, int width, void *cuda_stream) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!vel || !vel_active) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active must not be NULL");
	for (int dim = 0; dim < 3; ++dim)
		if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
	if (width < 0) return fail(SHKZ_B200_ERR_ARG, "width must be >= 0");
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	S->launches = 0;
	const int rc = S->real == SHKZ_B200_REAL_F64 ? post_impl<double>(S, vel, vel_active, solid, width, stream) : post_impl<float>(S, vel, vel_active, solid, width, stream);
	if (rc != SHKZ_B200_OK) return rc;
	CK(cudaStreamSynchronize(stream));
	CK(cudaGetLastError());
	return SHKZ_B200_OK;
}

int shkz_b200_extrapolate_constrain_host(shkz_b200_solver *S, void *const vel[3], uint8_t *const vel_active[3], const void *solid, int width) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!vel || !vel_active) return fail(SHKZ_B200_ERR_ARG, "vel / vel_active must not be NULL");
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	const Dims &d = S->d;
	const size_t rb = S->real_bytes, nodal = (size_t)(d.nx + 1) * (d.ny + 1) * (d.nzl + 1);
	cudaStream_t stream = nullptr;
	void *dvel[3];
	uint8_t *dact[3];
	for (int dim = 0; dim < 3; ++dim) {
		if (!vel[dim] || !vel_active[dim]) return fail(SHKZ_B200_ERR_ARG, "vel[%d] / vel_active[%d] is NULL", dim, dim);
		const size_t nf = face_count(d, dim);
		if (!S->st_vel[dim].base) { CKR(S->st_vel[dim].alloc(nf * rb)); CKR(S->st_act[dim].alloc(nf)); }
		CK(cudaMemcpyAsync(S->st_vel[dim].base, vel[dim], nf * rb, cudaMemcpyHostToDevice, stream));
		CK(cudaMemcpyAsync(S->st_act[dim].base, vel_active[dim], nf, cudaMemcpyHostToDevice, stream));
		dvel[dim] = S->st_vel[dim].base;
		dact[dim] = static_cast<uint8_t *>(S->st_act[dim].base);
	}
	if (solid) {
		if (!S->st_solid.base) CKR(S->st_solid.alloc(nodal * rb));
		CK(cudaMemcpyAsync(S->st_solid.base, solid, nodal * rb, cudaMemcpyHostToDevice, stream));
	}
	CKR(shkz_b200_extrapolate_constrain_device(S, dvel, dact, solid ? S->st_solid.base : nullptr, width, stream));
	for (int dim = 0; dim < 3; ++dim) {
		const size_t nf = face_count(d, dim);
		CK(cudaMemcpyAsync(vel[dim], S->st_vel[dim].base, nf * rb, cudaMemcpyDeviceToHost, stream));
		CK(cudaMemcpyAsync(vel_active[dim], S->st_act[dim].base, nf, cudaMemcpyDeviceToHost, stream));
	}
	CK(cudaStreamSynchronize(stream));
	return SHKZ_B200_OK;
}

int shkz_b200_host_alloc(size_t bytes, void **out) {
	if (!out) return fail(SHKZ_B200_ERR_ARG, "out is NULL");
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return fail(SHKZ_B200_ERR_NO_DEVICE, "no CUDA device available; libshkz_b200 has no CPU fallback");
	}
	CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
	return SHKZ_B200_OK;
}

void shkz_b200_host_free(void *ptr) {
	if (ptr) cudaFreeHost(ptr);
}

int shkz_b200_resolve(shkz_b200_solver *S, const shkz_b200_params *params, shkz_b200_stats *stats, void *cuda_stream) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!S->have_system) return fail(SHKZ_B200_ERR_STATE, "resolve() needs a prior project()");
	shkz_b200_params P;
	CKR(check_params(params, P));
	if (P.precision != S->alloc_precision) return fail(SHKZ_B200_ERR_STATE, "resolve() precision differs from the assembled system");
	CKR(device_ready(S->device));
	CK(cudaSetDevice(S->device));
	cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
	if (stats) memset(stats, 0, sizeof *stats);
	S->launches = 0;
	CK(cudaEventRecord(S->ev[1], stream));
	if (P.precond == SHKZ_B200_PRECOND_MG) CKR(build_hierarchy(S, P, stream));
	CK(cudaEventRecord(S->ev[2], stream));
	CKR(set_omega(P, stream));
	int rc;
	if (P.precision == SHKZ_B200_PREC_FP64) rc = solve<double, double>(S, P, stream);
	else if (P.precision == SHKZ_B200_PREC_MIXED) rc = solve<double, float>(S, P, stream);
	else rc = solve<float, float>(S, P, stream);
	if (rc != SHKZ_B200_OK) return rc;
	CK(cudaEventRecord(S->ev[3], stream));
	CK(cudaStreamSynchronize(stream));
	S->prof.collect();
	if (stats) {
		fill_stats(S, stats);
		cudaEventElapsedTime(&stats->ms_setup, S->ev[1], S->ev[2]);
		cudaEventElapsedTime(&stats->ms_solve, S->ev[2], S->ev[3]);
		stats->ms_total = stats->ms_setup + stats->ms_solve;
	}
	return SHKZ_B200_OK;
}

int shkz_b200_profile_enable(shkz_b200_solver *S, int on) {
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	S->prof.reset();
	S->prof.on = on != 0;
	return SHKZ_B200_OK;
}

int shkz_b200_profile_count(shkz_b200_solver *S) { return S ? (int)S->prof.names.size() : 0; }

int shkz_b200_profile_get(shkz_b200_solver *S, int index, char *name, size_t name_bytes, uint64_t *launches, double *total_ms) {
	if (!S || index < 0 || (size_t)index >= S->prof.names.size()) return fail(SHKZ_B200_ERR_ARG, "profile index out of range");
	if (name && name_bytes) { strncpy(name, S->prof.names[index].c_str(), name_bytes - 1); name[name_bytes - 1] = 0; }
	if (launches) *launches = S->prof.count[index];
	if (total_ms) *total_ms = S->prof.ms[index];
	return SHKZ_B200_OK;
}

int shkz_b200_debug_fetch(shkz_b200_solver *S, const char *name, void *dst, size_t dst_bytes, size_t *needed) {
	if (!S || !name) return fail(SHKZ_B200_ERR_ARG, "solver / name is NULL");
	CK(cudaSetDevice(S->device));
	const Dims &d = S->d;
	const std::string n(name);
	const void *src = nullptr;
	size_t bytes = 0;
	if (S->fractions_stale && (n.compare(0, 5, "areas") == 0 || n.compare(0, 4, "rhos") == 0)) {
		const dim3 grid((d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8, d.nzl + 1), block(32, 8, 1);
		for (int dim = 0; dim < 3; ++dim) {
			if (!S->areas[dim].base) CKR(S->areas[dim].alloc(face_count(d, dim) * S->real_bytes));
			if (!S->rhos[dim].base) CKR(S->rhos[dim].alloc(face_count(d, dim) * S->real_bytes));
		}
		if (S->real == SHKZ_B200_REAL_F32) {
			FaceGrids<float> a, r;
			for (int dim = 0; dim < 3; ++dim) { a.p[dim] = static_cast<float *>(S->areas[dim].base); r.p[dim] = static_cast<float *>(S->rhos[dim].base); }
			k_face_fractions<float><<<grid, block>>>(d, S->last_asm, static_cast<const float *>(S->last_solid), S->phi.ptr<float>(d), a, r);
		} else {
			FaceGrids<double> a, r;
			for (int dim = 0; dim < 3; ++dim) { a.p[dim] = static_cast<double *>(S->areas[dim].base); r.p[dim] = static_cast<double *>(S->rhos[dim].base); }
			k_face_fractions<double><<<grid, block>>>(d, S->last_asm, static_cast<const double *>(S->last_solid), S->phi.ptr<double>(d), a, r);
		}
		CK(cudaDeviceSynchronize());
		S->fractions_stale = false;
	}
	const size_t vec = S->alloc_precision == SHKZ_B200_PREC_FP32 ? 4 : 8;
	const size_t coef = S->alloc_precision == SHKZ_B200_PREC_FP64 ? 8 : 4;
	auto cell = [&](const CellArray &a, size_t elem) { src = a.base ? static_cast<const char *>(a.base) + (size_t)d.plane * elem : nullptr; bytes = (size_t)d.ncell * elem; };
	if (n == "dd") cell(S->dd, coef);
	else if (n == "wx") cell(S->wx, coef);
	else if (n == "wy") cell(S->wy, coef);
	else if (n == "wz") cell(S->wz, coef);
	else if (n == "rhs") cell(S->b, vec);
	else if (n == "x") cell(S->x, vec);
	else if (n == "vcycle" && S->debug_vcycle_result) { src = S->debug_vcycle_result; bytes = (size_t)d.ncell * 4; }
	else if (n == "tile_count" && !S->levels.empty()) { src = S->levels[0].tile_count.base; bytes = sizeof(int); }
	else if (n == "in_rows") cell(S->in_rows, 1);
	else if (n == "phi") cell(S->phi, S->real_bytes);
	else if (n.size() == 6 && n.compare(0, 5, "areas") == 0 && n[5] >= '0' && n[5] <= '2') { src = S->areas[n[5] - '0'].base; bytes = face_count(d, n[5] - '0') * S->real_bytes; }
	else if (n.size() == 5 && n.compare(0, 4, "rhos") == 0 && n[4] >= '0' && n[4] <= '2') { src = S->rhos[n[4] - '0'].base; bytes = face_count(d, n[4] - '0') * S->real_bytes; }
	else if (n.compare(0, 3, "mg_") == 0) { // mg_<level>_<dd|wx|wy|wz>
		int l = -1;
		char what[16] = {0};
		if (sscanf(name, "mg_%d_%15s", &l, what) == 2 && l >= 0 && (size_t)l < S->levels.size()) {
			const HostLevel &L = S->levels[l];
			const std::string w(what);
			const CellArray *a = w == "dd" ? &L.dd : w == "wx" ? &L.wx : w == "wy" ? &L.wy : w == "wz" ? &L.wz : nullptr;
			if (a) { src = a->base ? static_cast<const char *>(a->base) + (size_t)L.d.plane * 4 : nullptr; bytes = (size_t)L.d.ncell * 4; }
		}
	}
	if (needed) *needed = bytes;
	if (!src) return fail(SHKZ_B200_ERR_ARG, "unknown or unallocated array '%s'", name);
	if (!dst) return SHKZ_B200_OK;
	if (dst_bytes < bytes) return fail(SHKZ_B200_ERR_ARG, "buffer too small for '%s': %zu < %zu", name, dst_bytes, bytes);
	CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
	return SHKZ_B200_OK;
}

int shkz_b200_debug_vcycle(shkz_b200_solver *S, const shkz_b200_params *params, int legacy) {
#ifndef SHKZ_B200_TEST_HOOKS
	(void)params; (void)legacy;
	return fail(SHKZ_B200_ERR_STATE, "shkz_b200_debug_vcycle is only built into libshkz_b200_testhooks.so (-DSHKZ_B200_TEST_HOOKS)%s", S ? "" : "");
#else
	if (!S) return fail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!S->have_system || !S->have_hierarchy) return fail(SHKZ_B200_ERR_STATE, "debug_vcycle() needs a prior project() with the multigrid preconditioner");
	shkz_b200_params P;
	CKR(check_params(params, P));
	if (P.precision != S->alloc_precision) return fail(SHKZ_B200_ERR_STATE, "debug_vcycle() precision differs from the assembled system");
	CK(cudaSetDevice(S->device));
	cudaStream_t stream = nullptr;
	const Dims &d = S->d;
	HostLevel &H0 = S->levels[0];
	const Tiles T = H0.view.tiles;
	CKR(set_omega(P, stream));
	// level-0 right-hand side := float(rhs of the last project())
	if (P.precision == SHKZ_B200_PREC_FP32)
		LAUNCH_TILES(S, "cg_init", (k_cg_init<float, false>), cg_block(), H0.tiles_total, stream, d, T, (const float *)S->b.ptr<float>(d), S->x.ptr<float>(d), S->r.ptr<float>(d), S->s.ptr<float>(d), (float *)nullptr);
	else
		LAUNCH_TILES(S, "cg_init", (k_cg_init<double, true>), cg_block(), H0.tiles_total, stream, d, T, (const double *)S->b.ptr<double>(d), S->x.ptr<double>(d), S->r.ptr<double>(d), S->s.ptr<double>(d), H0.view.b);
	const float *z = nullptr;
	if (legacy == 1) {
		CKR(legacy_vcycle(S, 0, P, stream));
		z = H0.view.xa;
	} else {
		const int keep = S->sweep_mode;
		S->sweep_mode = legacy == 2 ? 2 : (legacy == 3 ? 1 : 0);
		const int rc = vcycle(S, 0, P, nullptr, stream, false, &z);
		S->sweep_mode = keep;
		CKR(rc);
	}
	CK(cudaStreamSynchronize(stream));
	CK(cudaGetLastError());
	S->debug_vcycle_result = z;
	return SHKZ_B200_OK;
#endif
}

int shkz_b200_slab_export(shkz_b200_solver *S, uint8_t ipc[SHKZ_B200_IPC_BYTES]) {
	if (!S || !ipc) return fail(SHKZ_B200_ERR_ARG, "solver / ipc is NULL");
	if (S->whole_grid || !S->comm) return fail(SHKZ_B200_ERR_STATE, "not a z-slab solver");
	CK(cudaSetDevice(S->device));
	if (S->comm->export_handle(ipc, SHKZ_B200_IPC_BYTES)) return fail(SHKZ_B200_ERR_COMM, "%s", S->comm->error());
	return SHKZ_B200_OK;
}

int shkz_b200_slab_connect(shkz_b200_solver *S, int rank, int world, const uint8_t *all_ipc) {
	if (!S || !all_ipc) return fail(SHKZ_B200_ERR_ARG, "solver / all_ipc is NULL");
	if (S->whole_grid || !S->comm) return fail(SHKZ_B200_ERR_STATE, "not a z-slab solver");
	if (world < 1 || world > COMM_MAX_WORLD) return fail(SHKZ_B200_ERR_ARG, "world must be 1..%d", COMM_MAX_WORLD);
	if (S->d.nzl * world != S->d.nzg || S->d.k0 != rank * S->d.nzl) return fail(SHKZ_B200_ERR_ARG, "rank %d of %d does not own slab [%d,%d) of %d planes", rank, world, S->d.k0, S->d.k0 + S->d.nzl, S->d.nzg);
	CK(cudaSetDevice(S->device));
	if (S->comm->connect_ipc(rank, world, all_ipc, SHKZ_B200_IPC_BYTES)) return fail(SHKZ_B200_ERR_COMM, "%s", S->comm->error());
	return SHKZ_B200_OK;
}

int shkz_b200_slab_connect_local(shkz_b200_solver *const *solvers, int world) {
	if (!solvers || world < 1 || world > COMM_MAX_WORLD) return fail(SHKZ_B200_ERR_ARG, "solvers is NULL or world not in 1..%d", COMM_MAX_WORLD);
	SlabComm *all[COMM_MAX_WORLD] = {};
	for (int r = 0; r < world; ++r) {
		shkz_b200_solver *S = solvers[r];
		if (!S || S->whole_grid || !S->comm) return fail(SHKZ_B200_ERR_STATE, "solver %d is not a z-slab solver", r);
		if (S->d.nzl * world != S->d.nzg || S->d.k0 != r * S->d.nzl) return fail(SHKZ_B200_ERR_ARG, "solver %d does not own slab %d of %d", r, r, world);
		all[r] = S->comm;
	}
	for (int r = 0; r < world; ++r) {
		CK(cudaSetDevice(solvers[r]->device));
		if (solvers[r]->comm->connect_local(r, world, all)) return fail(SHKZ_B200_ERR_COMM, "%s", solvers[r]->comm->error());
	}
	return SHKZ_B200_OK;
}

} // extern "C"
