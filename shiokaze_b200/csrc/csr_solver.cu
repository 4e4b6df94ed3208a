// CG on an assembled sparse matrix (CSR in, fp64) — the GPU stand-in for Shiokaze's `LinSolver` modules
// (RCMatrix_solver_interface::solve, include/shiokaze/linsolver/RCMatrix_solver.h:77; reference modules
// src/linsolver/pcg.cpp:45-73 and src/linsolver/cg.cpp). SURVEY.md 8f rank 2: the callers of the pressure path that
// assemble an RCMatrix themselves (the stock macpressuresolver3, macstreamfuncsolver3, the 2-D solvers) keep their
// assembly and hand the system to the GPU through the reference's own solver plug-in point.
//
// Algorithm = pcg_solver.h:246-295 as it EFFECTIVELY runs in the reference (its MIC(0) result is discarded, :383):
// plain CG with the infinity-norm stopping rule tol = Residual * |b|_inf, count = it + 1, reresid = |r|_inf / |b|_inf;
// optional Jacobi scaling (Precond=jacobi) as an additive flag. Three kernels per iteration, loop control on the device
// (CGState, common.cuh), deterministic two-stage reductions — the same skeleton as the matrix-free solver.
//
// Matrix layout: rows of at most 32 entries with little padding waste go to ELL (entry k of every row contiguous:
// coalesced, one thread per row); anything else stays CSR with one warp per row.
// No CPU compute path exists in this file.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/shkz_b200.h"
#include "common.cuh"

using namespace shkz;

namespace {

thread_local std::string g_csr_error;
int cfail(int code, const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_csr_error = buf;
	return code;
}
#define CCK(call)                                                                                                       \
	do {                                                                                                                \
		cudaError_t e_ = (call);                                                                                        \
		if (e_ != cudaSuccess) return cfail(SHKZ_B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

constexpr int FLAT_THREADS = 256;

// ---- layout ------------------------------------------------------------------------------------------------------
// CSR -> ELL (padding: column = own row, value 0), and the inverse diagonal for Jacobi (1 where the diagonal is not positive)
__global__ void __launch_bounds__(FLAT_THREADS) k_csr_to_ell(long long n, int width, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                                            const double *__restrict__ val, int *__restrict__ ecol, double *__restrict__ eval, double *__restrict__ invd) {
	for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
		const long long a = rowptr[row], b = rowptr[row + 1];
		double dg = 0.0;
		for (int k = 0; k < width; ++k) {
			const bool have = a + k < b;
			const int c = have ? col[a + k] : (int)row;
			const double v = have ? val[a + k] : 0.0;
			ecol[(long long)k * n + row] = c;
			eval[(long long)k * n + row] = v;
			if (have && c == row) dg += v;
		}
		invd[row] = dg > 0.0 ? 1.0 / dg : 1.0;
	}
}

__global__ void __launch_bounds__(FLAT_THREADS) k_csr_diag(long long n, const long long *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                          double *__restrict__ invd) {
	for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
		double dg = 0.0;
		for (long long e = rowptr[row]; e < rowptr[row + 1]; ++e)
			if (col[e] == row) dg += val[e];
		invd[row] = dg > 0.0 ? 1.0 / dg : 1.0;
	}
}

// ---- the CG kernels ------------------------------------------------------------------------------------------------
// x = 0, r = b, s = 0; |b|_inf, rho = r . M^-1 r ; last block: tol, trivial-rhs exit          (pcg_solver.h:249-271)
template <bool JACOBI>
__global__ void __launch_bounds__(FLAT_THREADS) k_flat_init(long long n, const double *__restrict__ b, const double *__restrict__ invd, double *__restrict__ x,
                                                           double *__restrict__ r, double *__restrict__ s, double residual, int max_iter, RedBuf rb, CGState *st) {
	double red[2] = {0.0, 0.0};
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
		const double v = b[e];
		x[e] = 0.0; s[e] = 0.0; r[e] = v;
		red[0] = fmax(red[0], fabs(v));
		red[1] += JACOBI ? v * v * invd[e] : v * v;
	}
	grid_reduce<2, 0x1u>(red, rb, [&](double (&t)[2]) {
		const double factor = residual < 1e-30 ? 1e-30 : residual; // pcg_solver.h:239
		st->bnorm = t[0]; st->rnorm = t[0];
		st->tol = factor * t[0];
		st->rho = t[1]; st->beta = 0.0; st->alpha = 0.0; st->sz = 0.0;
		st->iter = 0; st->max_iter = max_iter;
		st->n_rows = (unsigned long long)n;
		st->converged = t[0] == 0.0 ? 1 : 0;
		st->done = (t[0] == 0.0 || max_iter <= 0 || t[1] == 0.0 || t[1] != t[1]) ? 1 : 0; // pcg_solver.h:254-271
	});
}

// s = M^-1 r + beta s                                                                        (pcg_solver.h:272,289)
template <bool JACOBI>
__global__ void __launch_bounds__(FLAT_THREADS) k_flat_xpay(long long n, const double *__restrict__ r, const double *__restrict__ invd, double *__restrict__ s,
                                                           const CGState *__restrict__ st) {
	if (st->done) return;
	const double beta = st->beta;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
		s[e] = (JACOBI ? r[e] * invd[e] : r[e]) + beta * s[e];
}

// q = A s, s.q ; last block: alpha = rho / s.q                                                (pcg_solver.h:276-277)
__global__ void __launch_bounds__(FLAT_THREADS) k_ell_spmv_dot(long long n, int width, const int *__restrict__ ecol, const double *__restrict__ eval,
                                                              const double *__restrict__ s, double *__restrict__ q, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
		double acc = 0.0;
		for (int k = 0; k < width; ++k) acc += eval[(long long)k * n + row] * s[ecol[(long long)k * n + row]];
		q[row] = acc;
		red[0] += s[row] * acc;
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) { st->sz = t[0]; st->alpha = st->rho / t[0]; });
}

__global__ void __launch_bounds__(FLAT_THREADS) k_csr_spmv_dot(long long n, const long long *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                              const double *__restrict__ s, double *__restrict__ q, RedBuf rb, CGState *st) {
	if (st->done) return;
	double red[1] = {0.0};
	const int lane = threadIdx.x & 31;
	const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long row = warp; row < n; row += nwarps) {
		double acc = 0.0;
		for (long long e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) acc += val[e] * s[col[e]];
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
		if (lane == 0) {
			q[row] = acc;
			red[0] += s[row] * acc;
		}
	}
	grid_reduce<1, 0u>(red, rb, [&](double (&t)[1]) { st->sz = t[0]; st->alpha = st->rho / t[0]; });
}

// x += alpha s ; r -= alpha q ; |r|_inf ; rho' = r . M^-1 r ; last block: stop test, beta, rho   (pcg_solver.h:278-289)
template <bool JACOBI>
__global__ void __launch_bounds__(FLAT_THREADS) k_flat_axpy2_norm(long long n, const double *__restrict__ s, const double *__restrict__ q, const double *__restrict__ invd,
                                                                 double *__restrict__ x, double *__restrict__ r, RedBuf rb, CGState *st) {
	if (st->done) return;
	const double alpha = st->alpha;
	double red[2] = {0.0, 0.0};
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
		x[e] += alpha * s[e];
		const double rv = r[e] - alpha * q[e];
		r[e] = rv;
		red[0] = fmax(red[0], fabs(rv));
		red[1] += JACOBI ? rv * rv * invd[e] : rv * rv;
	}
	grid_reduce<2, 0x1u>(red, rb, [&](double (&t)[2]) {
		st->rnorm = t[0];
		st->iter += 1;
		if (t[0] <= st->tol) { st->done = 1; st->converged = 1; }
		else if (st->iter >= st->max_iter) st->done = 1;
		st->beta = t[1] / st->rho;
		st->rho = t[1];
	});
}

struct Buf {
	void *p = nullptr;
	size_t bytes = 0;
	int ensure(size_t need) {
		if (need <= bytes) return SHKZ_B200_OK;
		if (p) cudaFree(p);
		p = nullptr; bytes = 0;
		const size_t want = need + need / 8;
		CCK(cudaMalloc(&p, want));
		bytes = want;
		return SHKZ_B200_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

} // namespace

struct shkz_b200_csr {
	int device = 0;
	int num_sms = 148;
	Buf rowptr, col, val, ecol, eval, invd, b, x, r, s, q, partials, counter, state;
	CGState *h_state = nullptr;
	cudaEvent_t ev[4]{};
	bool events = false;
};

extern "C" {

const char *shkz_b200_csr_last_error(void) { return g_csr_error.c_str(); }

void shkz_b200_csr_default_params(shkz_b200_csr_params *p) {
	if (!p) return;
	memset(p, 0, sizeof *p);
	p->struct_size = sizeof *p;
	p->residual = 1e-4;          /* pcg.cpp:76 */
	p->max_iterations = 30000;   /* pcg.cpp:77 */
	p->precond = SHKZ_B200_CSR_PRECOND_NONE;
	p->check_every = 16;
}

int shkz_b200_csr_create(int device, shkz_b200_csr **out) {
	if (!out) return cfail(SHKZ_B200_ERR_ARG, "out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		return cfail(SHKZ_B200_ERR_NO_DEVICE, "no CUDA device available (%s); libshkz_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
	}
	if (device < 0 || device >= n) return cfail(SHKZ_B200_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
	CCK(cudaSetDevice(device));
	shkz_b200_csr *S = new shkz_b200_csr();
	S->device = device;
	int sms = 0;
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) S->num_sms = sms;
	int rc = SHKZ_B200_OK;
	if (cudaMallocHost((void **)&S->h_state, sizeof(CGState)) != cudaSuccess) rc = cfail(SHKZ_B200_ERR_CUDA, "cudaMallocHost failed");
	for (auto &ev : S->ev)
		if (rc == SHKZ_B200_OK && cudaEventCreate(&ev) != cudaSuccess) rc = cfail(SHKZ_B200_ERR_CUDA, "cudaEventCreate failed");
	S->events = rc == SHKZ_B200_OK;
	if (rc != SHKZ_B200_OK) { shkz_b200_csr_destroy(S); return rc; }
	*out = S;
	return SHKZ_B200_OK;
}

void shkz_b200_csr_destroy(shkz_b200_csr *S) {
	if (!S) return;
	cudaSetDevice(S->device);
	cudaDeviceSynchronize();
	for (Buf *b : {&S->rowptr, &S->col, &S->val, &S->ecol, &S->eval, &S->invd, &S->b, &S->x, &S->r, &S->s, &S->q, &S->partials, &S->counter, &S->state}) b->release();
	if (S->h_state) cudaFreeHost(S->h_state);
	for (auto &ev : S->ev) if (ev) cudaEventDestroy(ev);
	delete S;
}

int shkz_b200_csr_solve_host(shkz_b200_csr *S, uint64_t n64, const int64_t *rowptr, const int32_t *col, const double *val, const double *rhs, double *x,
                             const shkz_b200_csr_params *params, shkz_b200_csr_stats *stats) {
	if (!S) return cfail(SHKZ_B200_ERR_ARG, "solver is NULL");
	if (!rowptr || !rhs || !x) return cfail(SHKZ_B200_ERR_ARG, "rowptr / rhs / x must not be NULL");
	shkz_b200_csr_params P;
	if (params) {
		if (params->struct_size != sizeof P) return cfail(SHKZ_B200_ERR_ARG, "params.struct_size %u != %zu", params->struct_size, sizeof P);
		P = *params;
	} else shkz_b200_csr_default_params(&P);
	if (P.precond != SHKZ_B200_CSR_PRECOND_NONE && P.precond != SHKZ_B200_CSR_PRECOND_JACOBI) return cfail(SHKZ_B200_ERR_ARG, "unknown precond %d", P.precond);
	if (n64 >= (1ull << 31)) return cfail(SHKZ_B200_ERR_ARG, "n = %llu does not fit the 32-bit column indices", (unsigned long long)n64);
	if (stats) memset(stats, 0, sizeof *stats);
	const long long n = (long long)n64;
	if (n == 0) return SHKZ_B200_OK;
	if (rowptr[0] != 0 || rowptr[n] < 0) return cfail(SHKZ_B200_ERR_ARG, "rowptr must start at 0 and be non-decreasing");
	const long long nnz = rowptr[n];
	if (nnz > 0 && (!col || !val)) return cfail(SHKZ_B200_ERR_ARG, "col / val must not be NULL");
	long long width = 0;
	for (long long r = 0; r < n; ++r) {
		const long long w = rowptr[r + 1] - rowptr[r];
		if (w < 0) return cfail(SHKZ_B200_ERR_ARG, "rowptr decreases at row %lld", r);
		if (w > width) width = w;
	}
	int dn = 0;
	if (cudaGetDeviceCount(&dn) != cudaSuccess || dn <= 0) { cudaGetLastError(); return cfail(SHKZ_B200_ERR_NO_DEVICE, "no CUDA device available; libshkz_b200 has no CPU fallback"); }
	CCK(cudaSetDevice(S->device));
	cudaStream_t stream = nullptr;
	const bool ell = width <= 32 && width * n <= nnz + nnz / 2 + n;
	const bool jac = P.precond == SHKZ_B200_CSR_PRECOND_JACOBI;
	const int blocks = (int)((n + FLAT_THREADS - 1) / FLAT_THREADS < (long long)S->num_sms * 8 ? (n + FLAT_THREADS - 1) / FLAT_THREADS : (long long)S->num_sms * 8);
	const int vblocks = (int)((n * 32 + FLAT_THREADS - 1) / FLAT_THREADS < (long long)S->num_sms * 8 ? (n * 32 + FLAT_THREADS - 1) / FLAT_THREADS : (long long)S->num_sms * 8);
#define ENS(buf, bytes) do { int r_ = (buf).ensure(bytes); if (r_ != SHKZ_B200_OK) return r_; } while (0)
	ENS(S->rowptr, (size_t)(n + 1) * 8); ENS(S->col, (size_t)(nnz ? nnz : 1) * 4); ENS(S->val, (size_t)(nnz ? nnz : 1) * 8);
	ENS(S->invd, (size_t)n * 8);
	for (Buf *b : {&S->b, &S->x, &S->r, &S->s, &S->q}) ENS(*b, (size_t)n * 8);
	ENS(S->partials, (size_t)(S->num_sms * 8 + 8) * 4 * sizeof(double)); ENS(S->counter, 64); ENS(S->state, sizeof(CGState));
	if (ell) { ENS(S->ecol, (size_t)width * n * 4); ENS(S->eval, (size_t)width * n * 8); }
#undef ENS
	uint64_t launches = 0;
	CCK(cudaEventRecord(S->ev[0], stream));
	CCK(cudaMemsetAsync(S->counter.p, 0, 64, stream));
	CCK(cudaMemcpyAsync(S->rowptr.p, rowptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
	if (nnz) {
		CCK(cudaMemcpyAsync(S->col.p, col, (size_t)nnz * 4, cudaMemcpyHostToDevice, stream));
		CCK(cudaMemcpyAsync(S->val.p, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, stream));
	}
	CCK(cudaMemcpyAsync(S->b.p, rhs, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
	CCK(cudaEventRecord(S->ev[1], stream));
	const long long *d_rowptr = static_cast<const long long *>(S->rowptr.p);
	const int *d_col = static_cast<const int *>(S->col.p);
	const double *d_val = static_cast<const double *>(S->val.p);
	double *d_invd = static_cast<double *>(S->invd.p), *d_b = static_cast<double *>(S->b.p), *d_x = static_cast<double *>(S->x.p);
	double *d_r = static_cast<double *>(S->r.p), *d_s = static_cast<double *>(S->s.p), *d_q = static_cast<double *>(S->q.p);
	const RedBuf rb{static_cast<double *>(S->partials.p), static_cast<unsigned int *>(S->counter.p), nullptr};
	CGState *st = static_cast<CGState *>(S->state.p);
	if (ell) k_csr_to_ell<<<blocks, FLAT_THREADS, 0, stream>>>(n, (int)width, d_rowptr, d_col, d_val, static_cast<int *>(S->ecol.p), static_cast<double *>(S->eval.p), d_invd);
	else k_csr_diag<<<blocks, FLAT_THREADS, 0, stream>>>(n, d_rowptr, d_col, d_val, d_invd);
	++launches;
	if (jac) k_flat_init<true><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_b, d_invd, d_x, d_r, d_s, P.residual, (int)P.max_iterations, rb, st);
	else k_flat_init<false><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_b, d_invd, d_x, d_r, d_s, P.residual, (int)P.max_iterations, rb, st);
	++launches;
	const unsigned check = P.check_every < 1 ? 1u : (unsigned)P.check_every;
	unsigned it = 0;
	while (it < P.max_iterations) {
		for (unsigned c = 0; c < check && it < P.max_iterations; ++c, ++it) {
			if (jac) k_flat_xpay<true><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_r, d_invd, d_s, st);
			else k_flat_xpay<false><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_r, d_invd, d_s, st);
			if (ell) k_ell_spmv_dot<<<blocks, FLAT_THREADS, 0, stream>>>(n, (int)width, static_cast<const int *>(S->ecol.p), static_cast<const double *>(S->eval.p), d_s, d_q, rb, st);
			else k_csr_spmv_dot<<<vblocks, FLAT_THREADS, 0, stream>>>(n, d_rowptr, d_col, d_val, d_s, d_q, rb, st);
			if (jac) k_flat_axpy2_norm<true><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_s, d_q, d_invd, d_x, d_r, rb, st);
			else k_flat_axpy2_norm<false><<<blocks, FLAT_THREADS, 0, stream>>>(n, d_s, d_q, d_invd, d_x, d_r, rb, st);
			launches += 3;
		}
		CCK(cudaMemcpyAsync(S->h_state, st, sizeof(CGState), cudaMemcpyDeviceToHost, stream));
		CCK(cudaStreamSynchronize(stream));
		if (S->h_state->done) break;
	}
	CCK(cudaMemcpyAsync(S->h_state, st, sizeof(CGState), cudaMemcpyDeviceToHost, stream));
	CCK(cudaEventRecord(S->ev[2], stream));
	CCK(cudaMemcpyAsync(x, d_x, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
	CCK(cudaEventRecord(S->ev[3], stream));
	CCK(cudaStreamSynchronize(stream));
	CCK(cudaGetLastError());
	if (stats) {
		const CGState &h = *S->h_state;
		stats->iterations = (uint32_t)h.iter;
		stats->converged = h.converged;
		stats->reresid = h.bnorm > 0 ? h.rnorm / h.bnorm : 0.0;
		stats->rhs_absmax = h.bnorm;
		stats->ell_width = ell ? (int32_t)width : 0;
		stats->kernel_launches = launches;
		cudaEventElapsedTime(&stats->ms_h2d, S->ev[0], S->ev[1]);
		cudaEventElapsedTime(&stats->ms_solve, S->ev[1], S->ev[2]);
		cudaEventElapsedTime(&stats->ms_d2h, S->ev[2], S->ev[3]);
	}
	return SHKZ_B200_OK;
}

} // extern "C"
