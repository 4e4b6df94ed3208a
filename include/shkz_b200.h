/*
 * shkz_b200.h — thin C-ABI of the B200 pressure-projection library (libshkz_b200.so).
 *
 * This is the drop-in boundary for ONE path of ryichando/shiokaze: the 3D MAC-grid pressure
 * projection `macproject3_interface::project()` as implemented by
 *     src/projection/macpressuresolver3.cpp:50-272   (orchestration, assembly, velocity update)
 *     src/utility/macutility3.cpp:94-194             (solid area / liquid fractions)
 *     src/math/RCMatrix.cpp + src/linsolver/pcg.cpp:45-73 + local/include/pcgsolver/pcg_solver.h:246-295
 *                                                     (matrix + conjugate-gradient solve)
 * The C++ module that a Shiokaze host loads (shiokaze_b200/plugin/b200pressure3.cpp, selected with
 * `Projection=b200pressure3`) gathers dense buffers from the host's array3 / macarray3 grids and calls
 * the functions below; nothing else crosses the boundary. Plain pointers and sizes only, no C++ or
 * torch types, no exceptions; every function returns an error code and records a message retrievable
 * with shkz_b200_last_error(). There is no CPU fallback: without a usable CUDA device every compute
 * entry point fails with SHKZ_B200_ERR_NO_DEVICE.
 *
 * Dense layouts (the reference's own index order, include/shiokaze/math/shape.h:883-888: x fastest):
 *     cell grid   nx * ny * nzl               index i + nx*(j + ny*k)
 *     x faces     (nx+1) * ny * nzl           y faces  nx * (ny+1) * nzl        z faces  nx * ny * (nzl+1)
 *     nodal grid  (nx+1) * (ny+1) * (nzl+1)
 * nzl = number of z planes held by this solver (= nz for a whole grid, = k1-k0 for a z-slab).
 * Values are what array3::operator() returns (active value / flood-fill value / background,
 * include/shiokaze/array/array3.h:796-801); face activity is array3::active().
 */
#ifndef SHKZ_B200_H
#define SHKZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHKZ_B200_ABI_VERSION 4

enum shkz_b200_status {
	SHKZ_B200_OK = 0,
	SHKZ_B200_ERR_ARG = 1,       /* bad argument / unsupported combination */
	SHKZ_B200_ERR_NO_DEVICE = 2, /* no CUDA device: the library never falls back to the CPU */
	SHKZ_B200_ERR_CUDA = 3,      /* a CUDA runtime call failed */
	SHKZ_B200_ERR_COMM = 4,      /* slab communicator (CUDA IPC / peer access) failure */
	SHKZ_B200_ERR_STATE = 5      /* call sequence error */
};

enum shkz_b200_precond {
	SHKZ_B200_PRECOND_NONE = 0, /* plain CG == what the reference's "pcg" computes (pcg_solver.h:383 discards MIC(0)) */
	SHKZ_B200_PRECOND_MG = 1    /* aggregation-multigrid V-cycle, red-black Gauss-Seidel smoothing */
};

enum shkz_b200_precision {
	SHKZ_B200_PREC_FP64 = 0,  /* CG vectors and operator coefficients double (the reference's FLOAT_TYPE=double) */
	SHKZ_B200_PREC_MIXED = 1, /* CG vectors double, operator coefficients float */
	SHKZ_B200_PREC_FP32 = 2   /* everything float */
};

enum shkz_b200_real {
	SHKZ_B200_REAL_F32 = 0, /* host built with Real=float (include/shiokaze/core/config.h:34, the default) */
	SHKZ_B200_REAL_F64 = 1  /* host built with Real=double */
};

/* Flags of the reference module and its children, same meaning and defaults
 * (macpressuresolver3.cpp:274-280,296-305; macutility3.cpp:408-412,417-421; pcg.cpp:39-44,75-80). */
typedef struct shkz_b200_params {
	uint32_t struct_size;         /* = sizeof(shkz_b200_params), for forward compatibility */
	int32_t second_order_fluid;   /* SecondOrderAccurateFluid (Yes) */
	int32_t second_order_solid;   /* SecondOrderAccurateSolid (Yes) */
	int32_t apply_rhs_correct;    /* nonzero: add rhs_correct to every row (Gain && target volume set) */
	double eps_fluid;             /* MacUtility.EpsFluid (1e-2) */
	double eps_solid;             /* MacUtility.EpsSolid (1e-2) */
	double surface_tension;       /* project()'s surface_tension argument */
	double rhs_correct;           /* volume-correction constant, macpressuresolver3.cpp:204-214 (host computes the PI controller) */
	double residual;              /* LinSolver.Residual (1e-4): stop when |r|_inf <= residual * |b|_inf */
	uint32_t max_iterations;      /* LinSolver.MaxIterations (30000) */
	/* additive flags of this implementation */
	int32_t precond;              /* shkz_b200_precond (default MG) */
	int32_t precision;            /* shkz_b200_precision (default MIXED) */
	int32_t mg_pre_sweeps;        /* red-black sweeps before coarse correction (default 2) */
	int32_t mg_post_sweeps;       /* and after, reversed colour order (default 2) */
	int32_t mg_coarse_sweeps;     /* sweeps on the coarsest level, each direction (default 8) */
	int32_t mg_min_size;          /* stop coarsening when the largest extent is <= this (default 4) */
	int32_t check_every;          /* host reads the convergence flag every this many iterations (default 4) */
	double mg_coarse_scale;       /* coarse operator = scale * (P^T A P), piecewise-constant P (default 0.5) */
	int32_t mg_gamma;             /* coarse-grid visits per level below level 0: 1 = V-cycle (default), 2 = W-cycle */
	int32_t warm_start;           /* WarmStart (No): solve for the correction to the previous call's pressure, macpressuresolver3.cpp:221-242
	                                 (kept per CELL by the solver; the reference keeps it per row number) */
	double mg_omega;              /* relaxation factor of the red-black sweeps, 0 < omega < 2 (1 = Gauss-Seidel; default 1.15) */
	int32_t extrapolate_width;    /* > 0: project() ends with shkz_b200_extrapolate_constrain on the velocity it still holds on the device (default 0: the
	                                 host's own macutility3::extrapolate_and_constrain_velocity call does it, as with the reference module) */
	int32_t velocity_masked;      /* nonzero: the ENTRIES of inactive faces in vel[] are unspecified on input and on output — the library reads an inactive face as 0, the
	                                 background value of the simulators' velocity grids (what a dense array core hands over in place holds stale values there;
	                                 default 0: every entry is what array3::operator() returns) */
} shkz_b200_params;

typedef struct shkz_b200_stats {
	uint64_t n_rows;        /* number of unknowns (cells in the row set) on this solver's slab */
	uint64_t n_rows_global; /* ... over all slabs */
	uint32_t iterations;    /* as the reference counts them (pcg_solver.h:282) */
	int32_t converged;
	double reresid;         /* |r|_inf / |b|_inf at exit */
	double rhs_absmax;      /* |b|_inf */
	int32_t has_dirichlet;  /* 0: pure Neumann (singular) system, pressure mean removed */
	int32_t mg_levels;
	uint64_t kernel_launches; /* kernels of this library launched by the call */
	float ms_h2d, ms_assemble, ms_setup, ms_solve, ms_update, ms_d2h, ms_total; /* CUDA-event times */
	float ms_surftension;   /* part of ms_assemble spent on the surface-tension force (macpressuresolver3.cpp:85-115); 0 without it */
	uint32_t active_tiles;  /* level-0 tiles (64 x 16 x tile_depth cells) that hold an unknown: what every solve kernel walks over */
	uint32_t total_tiles;
	int32_t mg_mid_level;   /* first multigrid level of the cooperative mid-V-cycle launch (levels of <= 2^21 cells), -1: none */
	int32_t mg_tail_level;  /* first level of the shared-memory tail, which runs inside that launch (or alone when mg_mid_level is -1) */
	uint32_t tile_depth;    /* planes per level-0 tile this projection (chosen on the device: deep tiles for full grids, shallower for liquid scenes) */
	uint32_t host_copies;   /* shkz_b200_project_host: 0 = whole arrays through the copy engines, 1 = sparse (only what the projection touches, moved by kernels) */
	uint64_t h2d_bytes;     /* shkz_b200_project_host: bytes that crossed PCIe towards the device ... */
	uint64_t d2h_bytes;     /* ... and back to the host, this call */
} shkz_b200_stats;

typedef struct shkz_b200_solver shkz_b200_solver; /* opaque */

int shkz_b200_abi_version(void);
const char *shkz_b200_last_error(void);
void shkz_b200_default_params(shkz_b200_params *params);

/* Number of CUDA devices visible (0 when there is none; never an error). */
int shkz_b200_device_count(void);

/*
 * Create a solver for a whole nx*ny*nz grid on CUDA device `device`.
 * Replaces macpressuresolver3::initialize(shape,dx) + post_initialize() (macpressuresolver3.cpp:281-291).
 * real: shkz_b200_real — the element type of every grid buffer passed to project().
 */
int shkz_b200_create(int nx, int ny, int nz, double dx, int real, int device, shkz_b200_solver **out);

/*
 * Create a solver for the z-slab [k0,k1) of an nx*ny*nz grid (one solver per GPU / rank).
 * All grid buffers then hold nzl = k1-k0 cell planes (z faces / nodes: nzl+1, the shared plane duplicated).
 * Until shkz_b200_slab_connect() succeeds the solver refuses to project when (k0,k1) != (0,nz).
 */
int shkz_b200_create_slab(int nx, int ny, int nz, int k0, int k1, double dx, int real, int device, shkz_b200_solver **out);

void shkz_b200_destroy(shkz_b200_solver *solver);

/*
 * project(): replaces macproject3_interface::project (macproject3_interface.h:67-72).
 *   dt               time step
 *   vel[3]           in/out face velocities; only ACTIVE faces are modified (macpressuresolver3.cpp:252-268)
 *   vel_active[3]    in/out face activity (1/0); faces the reference would set_off() become 0
 *   solid            nodal solid level set, or NULL when the host's levelset_exist(solid) is false
 *                    (include/shiokaze/array/array_utility3.h:112-122)
 *   fluid            cell liquid level set (dense read, smoke passes its constant -1 grid)
 *   fluid_levelset   host's levelset_exist(fluid): 0 => every liquid fraction is 1 (macutility3.cpp:191-193)
 *   pressure         out, cell grid, 0 outside the row set        (get_pressure(), macproject3_interface.h:79)
 *   pressure_active  out, 1 on the row set (the reference activates exactly these cells, :245-248); may be NULL
 * Grid element type is float or double as chosen at creation. `_host` takes host pointers and does
 * the H2D / D2H copies itself; `_device` takes device pointers on the solver's GPU and runs on
 * `cuda_stream` (a cudaStream_t, NULL = default stream), returning after the stream has been synchronised.
 *
 * `_host` on a liquid scene (fluid_levelset != 0) whose buffers are ALL page-locked (shkz_b200_host_alloc, cudaHostAlloc, cudaHostRegister) moves only
 * what the projection touches: the liquid level set and the activity masks travel whole, velocity and solid nodes only around wet cells (fluid < 0),
 * and only the faces that were active on input and the pressure tiles that hold (or held, in the previous call) unknowns are written back — by
 * the library's own kernels addressing the host buffers directly (stats.host_copies = 1, stats.h2d_bytes / d2h_bytes say what moved). The buffers
 * end up byte for byte as with whole-array copies, under one assumption: between two calls that pass the SAME pressure / pressure_active pointers
 * the caller does not write into them (cells off the current and the previous row set are not rewritten; the first call with a new pointer clears
 * the whole grid). Environment SHKZ_B200_HOST_COPIES=dense forces whole-array copies.
 */
int shkz_b200_project_host(shkz_b200_solver *solver, double dt, void *const vel[3], uint8_t *const vel_active[3],
                           const void *solid, const void *fluid, int fluid_levelset, const shkz_b200_params *params,
                           void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats);

int shkz_b200_project_device(shkz_b200_solver *solver, double dt, void *const vel[3], uint8_t *const vel_active[3],
                             const void *solid, const void *fluid, int fluid_levelset, const shkz_b200_params *params,
                             void *pressure, uint8_t *pressure_active, shkz_b200_stats *stats, void *cuda_stream);

/*
 * Optional: allocate, now, everything the next project() with these parameters would allocate on first use (precision-dependent arrays, multigrid hierarchy,
 * and — host_buffers != 0 — the device staging arrays of shkz_b200_project_host; have_solid: a solid level set will be passed). One process that drives several
 * z-slab solvers from several threads calls this for every solver BEFORE starting the threads (plugin/b200pressure3.cpp does): a device allocation made while peer
 * access is enabled waits for the peer devices, and a peer whose kernel is already spinning on this rank's planes never becomes idle.
 */
int shkz_b200_prepare(shkz_b200_solver *solver, const shkz_b200_params *params, int host_buffers, int have_solid);

/*
 * The step the simulators run right after every projection (src/liquid/macliquid3.cpp:309-319, src/smoke/macsmoke3.cpp:294), SURVEY.md 8f rank 3:
 *     macutility3::extrapolate_and_constrain_velocity(solid, velocity, width)        src/utility/macutility3.cpp:89-93
 *       = macarray_extrapolator3::extrapolate(velocity, width)                       include/shiokaze/array/macarray_extrapolator3.h:49-53
 *         (width rounds: every inactive face next to an active one becomes active with the mean of its active neighbours, array_extrapolator3.h:51-82)
 *       + macutility3::constrain_velocity(solid, velocity)                           src/utility/macutility3.cpp:61-88
 *         (faces inside the solid lose the velocity component that points into it; wall faces may not point outwards; nothing at all without a solid level set)
 * vel / vel_active in place, same dense layouts as project(); solid = nodal level set or NULL (the host's levelset_exist(solid) is false). Results are
 * the reference's, bit for bit (tests/test_gpu_post.py). Whole-grid solvers only. params.extrapolate_width > 0 makes project() end with this step.
 */
int shkz_b200_extrapolate_constrain_device(shkz_b200_solver *solver, void *const vel[3], uint8_t *const vel_active[3], const void *solid, int width,
                                           void *cuda_stream);
int shkz_b200_extrapolate_constrain_host(shkz_b200_solver *solver, void *const vel[3], uint8_t *const vel_active[3], const void *solid, int width);

/*
 * Page-locked host memory for the buffers handed to shkz_b200_project_host: the H2D / D2H copies then run at PCIe speed (pageable
 * std::vector storage costs 3-4x the time: measured through the Shiokaze module). Free with shkz_b200_host_free.
 */
int shkz_b200_host_alloc(size_t bytes, void **out);
void shkz_b200_host_free(void *ptr);

/*
 * Re-run only the linear solve of the last project() (same matrix, same right-hand side, x = 0):
 * the timed unit of the solver benchmark. Velocity and pressure outputs are not touched.
 */
int shkz_b200_resolve(shkz_b200_solver *solver, const shkz_b200_params *params, shkz_b200_stats *stats, void *cuda_stream);

/* ---- z-slab communicator: one solver per GPU of an NVLink domain. Halo planes and the CG scalars travel by direct
 * peer-memory stores / loads inside the library's own kernels; no collective library is involved. Slabs must be equal
 * (nz divisible by the number of ranks, rank r owns planes [r*nz/world, (r+1)*nz/world)).
 * One process per GPU: every rank calls slab_export, the host exchanges the blobs by any means it has (the Python
 * mirror uses torch.distributed), then every rank calls slab_connect with all blobs in rank order.
 * One process driving several GPUs (what a Shiokaze host would do): create the solvers, then slab_connect_local. ---- */
#define SHKZ_B200_IPC_BYTES 128
int shkz_b200_slab_export(shkz_b200_solver *solver, uint8_t ipc[SHKZ_B200_IPC_BYTES]);
int shkz_b200_slab_connect(shkz_b200_solver *solver, int rank, int world, const uint8_t *all_ipc /* world * SHKZ_B200_IPC_BYTES */);
int shkz_b200_slab_connect_local(shkz_b200_solver *const *solvers, int world);

/* ---- assembled systems: CG on a CSR matrix ------------------------------------------------------
 * Replaces RCMatrix_solver_interface<size_t,double>::solve (include/shiokaze/linsolver/RCMatrix_solver.h:77) as implemented by
 * the reference's `pcg` module (src/linsolver/pcg.cpp:45-73 -> local/include/pcgsolver/pcg_solver.h:246-295), for callers that
 * assemble an RCMatrix themselves (the stock macpressuresolver3, macstreamfuncsolver3, the 2-D solvers). The Shiokaze module
 * on top is shiokaze_b200/plugin/b200cg.cpp (`LinSolver=b200cg`).
 * Algorithm: the reference's effective one — plain CG (its MIC(0) output is overwritten, pcg_solver.h:383), stop when
 * |r|_inf <= residual * |b|_inf, iterations counted as it+1, reresid = |r|_inf / |b|_inf; |b|_inf == 0 => 0 iterations, x = 0.
 * The matrix must be symmetric positive (semi-)definite; rows in CSR with 64-bit row pointers and 32-bit column indices. */
enum shkz_b200_csr_precond { SHKZ_B200_CSR_PRECOND_NONE = 0, SHKZ_B200_CSR_PRECOND_JACOBI = 1 };

typedef struct shkz_b200_csr_params {
	uint32_t struct_size;      /* = sizeof(shkz_b200_csr_params) */
	uint32_t max_iterations;   /* LinSolver.MaxIterations (30000, pcg.cpp:77) */
	double residual;           /* LinSolver.Residual (1e-4, pcg.cpp:76) */
	int32_t precond;           /* shkz_b200_csr_precond (default NONE = what the reference computes) */
	int32_t check_every;       /* host reads the convergence flag every this many iterations (default 16) */
} shkz_b200_csr_params;

typedef struct shkz_b200_csr_stats {
	uint32_t iterations;       /* as the reference counts them (pcg_solver.h:282) */
	int32_t converged;
	double reresid;            /* |r|_inf / |b|_inf at exit */
	double rhs_absmax;
	int32_t ell_width;         /* > 0: rows were laid out as ELL of this width; 0: CSR, one warp per row */
	int32_t reserved;
	uint64_t kernel_launches;
	float ms_h2d, ms_solve, ms_d2h;
	float reserved2;
} shkz_b200_csr_stats;

typedef struct shkz_b200_csr shkz_b200_csr; /* opaque: device buffers reused across solves */

const char *shkz_b200_csr_last_error(void);
void shkz_b200_csr_default_params(shkz_b200_csr_params *params);
int shkz_b200_csr_create(int device, shkz_b200_csr **out);
void shkz_b200_csr_destroy(shkz_b200_csr *solver);
/* Solve A x = rhs (host pointers; x is overwritten, the start vector is 0 as in pcg_solver.h:249). */
int shkz_b200_csr_solve_host(shkz_b200_csr *solver, uint64_t n, const int64_t *rowptr /* n+1 */, const int32_t *col, const double *val,
                             const double *rhs, double *x, const shkz_b200_csr_params *params, shkz_b200_csr_stats *stats);

/* ---- the step BEFORE the projection: advection on the MAC grid (first part) ------------------------------------
 * Replaces macadvection3_interface::advect_vector / advect_scalar (include/shiokaze/advection/macadvection3_interface.h:52-71) as implemented by the
 * reference's `macadvection3` module (src/advection/macadvection3.cpp): semi-Lagrangian back-tracing with trilinear (array_interpolator3.h:50-105) or
 * sixth-order WENO (WENO3.h, WENO.h) interpolation, and the MacCormack scheme on top of it (forward, backward with -dt, correction limited to the
 * min / max of the eight corner values; first-order inside `trim_narrowband` cells of the liquid surface). The simulators call both right before
 * project(): src/liquid/macliquid3.cpp:343-346 (the level set through maclevelsetsurfacetracker3.cpp:51, then the velocity), src/smoke/macsmoke3.cpp:274,281
 * (density, velocity). The Shiokaze module on top is shiokaze_b200/plugin/b200advection3.cpp (`Advection=b200advection3`). Results are the reference's, bit for bit
 * (tests/test_gpu_advect.py). Same dense layouts as project(); whole grids only.
 *   - a velocity grid is read through its activity mask: inactive faces read as 0, the background value of the simulators' velocity grids;
 *   - cell grids (q, fluid) are read densely: every entry holds what array3::operator() returns (active value / flood-fill value / background);
 *   - only ACTIVE entries of u / q are written; the activity itself never changes (macadvection3.cpp:71-72, :195-196).
 * advect_vector: the reference traces the field with ITSELF — its `velocity` argument is not used (macadvection3.cpp:79) — hence no such argument here.
 * fluid: the liquid level set (a smoke solver passes its constant grid); may be NULL when maccormack == 0. */
typedef struct shkz_b200_advect_params {
	uint32_t struct_size;       /* = sizeof(shkz_b200_advect_params) */
	int32_t maccormack;         /* MacCormack (Yes, macadvection3.cpp:286) */
	int32_t weno;               /* WENO (No, :287): WENO3::interpolate, order 6, instead of trilinear interpolation */
	uint32_t trim_narrowband;   /* TrimNarrowBand (1, :288) */
	double scalar_background;   /* advect_scalar with MacCormack: array3::get_background_value() of q — what the forward result, a freshly borrowed grid of
	                               q's type (:245), reads off the active set when the backward pass interpolates in it (a level set: +halfwidth; density: 0) */
} shkz_b200_advect_params;

typedef struct shkz_b200_advect_stats {
	uint64_t kernel_launches;
	uint64_t h2d_bytes, d2h_bytes;       /* `_host` entry points */
	float ms_h2d, ms_advect, ms_d2h;     /* CUDA-event times */
	uint32_t host_copies;                /* advect_vector_host: 0 = whole arrays through the copy engines, 1 = sparse (masks whole, values of active faces by kernels) */
} shkz_b200_advect_stats;

typedef struct shkz_b200_advect shkz_b200_advect; /* opaque: work arrays (forward result, limiter record) and host staging, reused across calls */

const char *shkz_b200_advect_last_error(void);
void shkz_b200_advect_default_params(shkz_b200_advect_params *params);
/* Replaces macadvection3::initialize(shape,dx) (macadvection3.cpp:63-67). real: shkz_b200_real. */
int shkz_b200_advect_create(int nx, int ny, int nz, double dx, int real, int device, shkz_b200_advect **out);
void shkz_b200_advect_destroy(shkz_b200_advect *advect);
/* u in/out (face grids), u_active their activity. `_device`: device pointers on the handle's GPU, runs on cuda_stream, returns after synchronising it.
 * `_host` with page-locked u[3] (shkz_b200_host_alloc, cudaHostAlloc: what Array=b200array3 hands over) on a scene where at most half of the faces are active
 * moves the masks whole and the VALUES OF ACTIVE FACES ONLY, both ways, by kernels that address the host buffers directly (an inactive face reads as 0 and is
 * never written, so nothing else matters); stats.host_copies / h2d_bytes / d2h_bytes say what moved. SHKZ_B200_HOST_COPIES=dense forces whole arrays. */
int shkz_b200_advect_vector_host(shkz_b200_advect *advect, double dt, void *const u[3], const uint8_t *const u_active[3], const void *fluid,
                                 const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats);
int shkz_b200_advect_vector_device(shkz_b200_advect *advect, double dt, void *const u[3], const uint8_t *const u_active[3], const void *fluid,
                                   const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats, void *cuda_stream);
/* q in/out (cell grid) carried by vel; `fluid` may be the grid q was copied from (a level set carried by itself, maclevelsetsurfacetracker3.cpp:49-51). */
int shkz_b200_advect_scalar_host(shkz_b200_advect *advect, double dt, void *q, const uint8_t *q_active, const void *const vel[3],
                                 const uint8_t *const vel_active[3], const void *fluid, const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats);
int shkz_b200_advect_scalar_device(shkz_b200_advect *advect, double dt, void *q, const uint8_t *q_active, const void *const vel[3],
                                   const uint8_t *const vel_active[3], const void *fluid, const shkz_b200_advect_params *params, shkz_b200_advect_stats *stats,
                                   void *cuda_stream);

/* ---- per-kernel timing (CUDA events around every launch; slows the call down, never on by default) ----
 * enable(1) resets the accumulators; entries are "<kernel>" or "<kernel>@<multigrid level>". */
int shkz_b200_profile_enable(shkz_b200_solver *solver, int on);
int shkz_b200_profile_count(shkz_b200_solver *solver);
int shkz_b200_profile_get(shkz_b200_solver *solver, int index, char *name, size_t name_bytes, uint64_t *launches, double *total_ms);

/* ---- test hook: copy an internal device array to the host (names: see DESIGN.md, e.g. "diag","rhs") ---- */
int shkz_b200_debug_fetch(shkz_b200_solver *solver, const char *name, void *dst, size_t dst_bytes, size_t *needed_bytes);


/* ---- test hook, ONLY in the library built with -DSHKZ_B200_TEST_HOOKS (shiokaze_b200/_build/libshkz_b200_testhooks.so; the product library does
 * not carry it, nor the one-launch-per-colour validation kernels behind it): apply ONE multigrid V-cycle to the right-hand side of the last project() and keep the result for
 * debug_fetch("vcycle"). legacy: 0 = the product kernels; 1 = the unfused one-launch-per-colour kernels on dense
 * grids; 2 = the product path with the scalar sweep kernel forced; 3 = with the quad kernel (no TMA) forced. All must agree bit for bit
 * (tests/test_gpu_parity.py). */
int shkz_b200_debug_vcycle(shkz_b200_solver *solver, const shkz_b200_params *params, int legacy);

#ifdef __cplusplus
}
#endif
#endif
