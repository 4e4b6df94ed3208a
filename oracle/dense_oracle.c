/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference's pressure-projection path on DENSE arrays.
 * Never linked into, imported by or called from the shipped library; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs use it, and only as a checker.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this file against (a) the reference's own
 * known-answer test (src/examples/accuracytest3-example.cpp, golden max_norm values in
 * tests/golden/) and (b) outputs of the unmodified reference build (oracle/_ref, made by
 * oracle/Makefile from /root/reference) on dam-break / smoke / solid-obstacle scenes.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Layout: x fastest, index = i + w*(j + h*k)        (include/shiokaze/math/shape.h:883-888)
 * Real:   the reference stores grids as `Real` (float by default, include/shiokaze/core/config.h:34)
 *         while the linear system is double. All grid I/O here is double arrays that hold
 *         Real-representable values; `real_is_double` chooses where roundings happen.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
	int32_t nx, ny, nz;
	int32_t second_order_fluid, second_order_solid; /* SecondOrderAccurateFluid / Solid */
	int32_t real_is_double;                         /* 0: Real=float (shipping), 1: Real=double */
	int32_t fluid_levelset_exist;                   /* array_utility3.h:112-122 evaluated by the caller */
	int32_t have_solid;                             /* levelset_exist(solid); solid is nodal (nx+1)(ny+1)(nz+1) */
	uint32_t max_iterations;                        /* LinSolver.MaxIterations (pcg.cpp:43) */
	int32_t pad;
	double dx, dt;
	double eps_fluid, eps_solid;                    /* MacUtility.EpsFluid / EpsSolid (macutility3.cpp:408-409) */
	double surface_tension;
	double rhs_correct;                             /* volume-correction constant added to every row (0: off) */
	double residual;                                /* LinSolver.Residual (pcg.cpp:40) */
} oracle_params;

typedef struct {
	uint64_t n_rows, nnz;
	uint32_t iterations;
	int32_t converged;
	double reresid;
	double rhs_absmax;
} oracle_stats;

static double RR(const oracle_params *P, double x) { return P->real_is_double ? x : (double)(float)x; }

static size_t face_count(const oracle_params *P, int dim) {
	return (size_t)(P->nx + (dim == 0)) * (size_t)(P->ny + (dim == 1)) * (size_t)(P->nz + (dim == 2));
}
static size_t face_index(const oracle_params *P, int dim, int i, int j, int k) {
	size_t w = (size_t)(P->nx + (dim == 0)), h = (size_t)(P->ny + (dim == 1));
	return (size_t)i + w * ((size_t)j + h * (size_t)k);
}
static size_t cell_index(const oracle_params *P, int i, int j, int k) {
	return (size_t)i + (size_t)P->nx * ((size_t)j + (size_t)P->ny * (size_t)k);
}
static int clampi(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }

/* include/shiokaze/utility/utility.h:162-170 */
static double fraction(double phi0, double phi1) {
	if (phi0 * phi1 >= 0.0) {
		if (phi0 < 0.0 || phi1 < 0.0) return 1.0;
		return 0.0;
	}
	double denom = fabs(phi1 - phi0);
	if (denom < DBL_MIN) denom = DBL_MIN;
	return -(phi0 < phi1 ? phi0 : phi1) / denom;
}

/* include/shiokaze/utility/utility.h:179-214 — marching-squares polygon of {phi<0}, shoelace area */
static double get_area(double q00, double q10, double q11, double q01) {
	static const int quads[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
	double v[4] = {q00, q10, q11, q01};
	double p[8][2];
	int pnum = 0;
	for (int n = 0; n < 4; ++n) {
		if (v[n] < 0.0) {
			p[pnum][0] = quads[n][0];
			p[pnum][1] = quads[n][1];
			pnum++;
		}
		if (v[n] * v[(n + 1) % 4] < 0) {
			double y0 = v[n], y1 = v[(n + 1) % 4];
			if (y0 - y1) {
				double a = y0 / (y0 - y1);
				p[pnum][0] = (1.0 - a) * quads[n][0] + a * quads[(n + 1) % 4][0];
				p[pnum][1] = (1.0 - a) * quads[n][1] + a * quads[(n + 1) % 4][1];
				pnum++;
			}
		}
	}
	double sum = 0.0;
	for (int m = 0; m < pnum; ++m) sum += p[m][0] * p[(m + 1) % pnum][1] - p[m][1] * p[(m + 1) % pnum][0];
	return 0.5 * sum;
}

/* src/utility/macutility3.cpp:94-165. solid: nodal (nx+1)(ny+1)(nz+1) dense, or have_solid==0. */
void oracle_area_fraction(const oracle_params *P, const double *solid, double *areas[3]) {
	const int nx = P->nx, ny = P->ny, nz = P->nz;
	const size_t sw = (size_t)nx + 1, sh = (size_t)ny + 1;
#define S(i, j, k) solid[(size_t)(i) + sw * ((size_t)(j) + sh * (size_t)(k))]
	for (int dim = 0; dim < 3; ++dim) {
		const int w = nx + (dim == 0), h = ny + (dim == 1), d = nz + (dim == 2);
		const int n_dim = dim == 0 ? nx : (dim == 1 ? ny : nz);
		for (int k = 0; k < d; ++k) for (int j = 0; j < h; ++j) for (int i = 0; i < w; ++i) {
			const int pd = dim == 0 ? i : (dim == 1 ? j : k);
			double area;
			if (!P->have_solid) {
				/* :149-164  areas=1, six domain walls 0 */
				area = (pd == 0 || pd == n_dim) ? 0.0 : 1.0;
			} else {
				/* :120  pi[dim]==0 || pi[dim]==solid.shape()[dim]; solid is nodal so the second never fires */
				if (pd == 0 || pd == n_dim + 1) area = 0.0;
				else {
					double q00, q10, q11, q01;
					if (dim == 0) { q00 = S(i, j, k); q10 = S(i, j + 1, k); q11 = S(i, j + 1, k + 1); q01 = S(i, j, k + 1); }
					else if (dim == 1) { q00 = S(i, j, k); q10 = S(i + 1, j, k); q11 = S(i + 1, j, k + 1); q01 = S(i, j, k + 1); }
					else { q00 = S(i, j, k); q10 = S(i + 1, j, k); q11 = S(i + 1, j + 1, k); q01 = S(i, j + 1, k); }
					area = 1.0 - get_area(q00, q10, q11, q01);
				}
				if (area && area < P->eps_solid) area = P->eps_solid; /* :141 */
			}
			area = RR(P, area);
			if (!P->second_order_solid && area) area = 1.0; /* macpressuresolver3.cpp:76-80 */
			areas[dim][face_index(P, dim, i, j, k)] = area;
		}
	}
#undef S
}

/* src/utility/macutility3.cpp:166-194 */
void oracle_fluid_fraction(const oracle_params *P, const double *fluid, double *rhos[3]) {
	const int nx = P->nx, ny = P->ny, nz = P->nz;
	for (int dim = 0; dim < 3; ++dim) {
		const int w = nx + (dim == 0), h = ny + (dim == 1), d = nz + (dim == 2);
		for (int k = 0; k < d; ++k) for (int j = 0; j < h; ++j) for (int i = 0; i < w; ++i) {
			double rho;
			if (!P->fluid_levelset_exist) rho = 1.0; /* :191-193 */
			else {
				/* :179-182, shape3::clamp = include/shiokaze/math/shape.h:790-798 */
				double a = fluid[cell_index(P, clampi(i, nx), clampi(j, ny), clampi(k, nz))];
				double b = fluid[cell_index(P, clampi(i - (dim == 0), nx), clampi(j - (dim == 1), ny), clampi(k - (dim == 2), nz))];
				rho = fraction(a, b);
				if (rho && rho < P->eps_fluid) rho = P->eps_fluid; /* :183 */
			}
			rho = RR(P, rho);
			if (!P->second_order_fluid && rho) rho = 1.0; /* macpressuresolver3.cpp:70-74 */
			rhos[dim][face_index(P, dim, i, j, k)] = rho;
		}
	}
}

/* src/projection/macpressuresolver3.cpp:85-115 — surface tension adds to ACTIVE velocity faces */
static void surface_tension_force(const oracle_params *P, const double *fluid, double *const rhos[3],
                                  double *vel[3], const uint8_t *const active[3]) {
	const int nx = P->nx, ny = P->ny, nz = P->nz;
	const double dx = P->dx, dt = P->dt, kappa = P->surface_tension;
	double *curv = (double *)malloc(sizeof(double) * (size_t)nx * ny * nz);
	for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
#define F(a, b, c) fluid[cell_index(P, clampi(a, nx), clampi(b, ny), clampi(c, nz))]
		/* :93-98 — the six neighbour reads are Real values and add up as Real (float on the shipping build); "- 6.0*fluid" promotes to double */
		double six;
		if (P->real_is_double) six = F(i - 1, j, k) + F(i + 1, j, k) + F(i, j - 1, k) + F(i, j + 1, k) + F(i, j, k - 1) + F(i, j, k + 1);
		else six = (double)((float)F(i - 1, j, k) + (float)F(i + 1, j, k) + (float)F(i, j - 1, k) + (float)F(i, j + 1, k) + (float)F(i, j, k - 1) + (float)F(i, j, k + 1));
		double value = (six - 6.0 * F(i, j, k)) / (dx * dx);
		curv[cell_index(P, i, j, k)] = RR(P, value);
	}
	for (int dim = 0; dim < 3; ++dim) {
		const int w = nx + (dim == 0), h = ny + (dim == 1), d = nz + (dim == 2);
		for (int k = 0; k < d; ++k) for (int j = 0; j < h; ++j) for (int i = 0; i < w; ++i) {
			size_t f = face_index(P, dim, i, j, k);
			if (!active[dim][f]) continue;
			double rho = rhos[dim][f];
			if (rho && rho < 1.0) {
				double sgn = F(i, j, k) < 0.0 ? -1.0 : 1.0;
				double theta = sgn < 0 ? 1.0 - rho : rho;
				double face_c = theta * curv[cell_index(P, clampi(i, nx), clampi(j, ny), clampi(k, nz))]
				              + (1.0 - theta) * curv[cell_index(P, clampi(i - (dim == 0), nx), clampi(j - (dim == 1), ny), clampi(k - (dim == 2), nz))];
				double inc = -sgn * dt / (dx * rho) * kappa * face_c;
				if (P->real_is_double) vel[dim][f] += inc;
				else vel[dim][f] = (double)((float)vel[dim][f] + (float)inc); /* array3::increment, array3.h:631-639 */
			}
		}
#undef F
	}
	free(curv);
}

/* CSR system exactly as RCMatrix rows end up (sorted columns; src/math/RCMatrix.cpp:249-272) */
typedef struct {
	size_t n;
	size_t *rowstart; /* n+1 */
	size_t *col;
	double *val;
	double *rhs;
	double *dirichlet; /* part of the diagonal that comes from air neighbours (ghost-fluid Dirichlet faces) */
} csr_system;

static void csr_free(csr_system *A) {
	free(A->rowstart); free(A->col); free(A->val); free(A->rhs); free(A->dirichlet);
	memset(A, 0, sizeof *A);
}

/* src/projection/macpressuresolver3.cpp:121-156 (row labelling) and :159-199 (assembly).
 * Rows are numbered in natural dense order (the reference uses its tile order; only the
 * summation order of the CG dot products depends on it). index_map[c] = row or SIZE_MAX. */
static void assemble(const oracle_params *P, const double *fluid, double *const areas[3], double *const rhos[3],
                     double *const vel[3], size_t *index_map, csr_system *A) {
	const int nx = P->nx, ny = P->ny, nz = P->nz;
	const double dx = P->dx, dt = P->dt;
	static const int qoff[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
	static const int foff[6][3] = {{1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 0}};
	static const int direction[6] = {0, 0, 1, 1, 2, 2};
	static const int sgn[6] = {1, -1, 1, -1, 1, -1};
	size_t index = 0;
	for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
		size_t c = cell_index(P, i, j, k);
		int inside = 0;
		if (fluid[c] < 0.0) {
			for (int nq = 0; nq < 6; ++nq) {
				int qi = i + qoff[nq][0], qj = j + qoff[nq][1], qk = k + qoff[nq][2];
				if (qi < 0 || qj < 0 || qk < 0 || qi >= nx || qj >= ny || qk >= nz) continue; /* out_of_bounds */
				if (fluid[cell_index(P, qi, qj, qk)] < 0.0) {
					int dim = direction[nq];
					size_t f = face_index(P, dim, i + foff[nq][0], j + foff[nq][1], k + foff[nq][2]);
					if (areas[dim][f] && rhos[dim][f]) { inside = 1; break; }
				}
			}
		}
		index_map[c] = inside ? index++ : SIZE_MAX;
	}
	A->n = index;
	A->rowstart = (size_t *)calloc(index + 1, sizeof(size_t));
	A->col = (size_t *)malloc(sizeof(size_t) * 7 * (index ? index : 1));
	A->val = (double *)malloc(sizeof(double) * 7 * (index ? index : 1));
	A->rhs = (double *)calloc(index ? index : 1, sizeof(double));
	A->dirichlet = (double *)calloc(index ? index : 1, sizeof(double));
	size_t nnz = 0;
	for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
		size_t c = cell_index(P, i, j, k);
		size_t n_index = index_map[c];
		if (n_index == SIZE_MAX) continue;
		size_t cols[7]; double vals[7]; int cnt = 0;
		double diagonal = 0.0, rhs = 0.0, dirichlet = 0.0;
		for (int nq = 0; nq < 6; ++nq) {
			int dim = direction[nq];
			int qi = i + qoff[nq][0], qj = j + qoff[nq][1], qk = k + qoff[nq][2];
			if (qi < 0 || qj < 0 || qk < 0 || qi >= nx || qj >= ny || qk >= nz) continue;
			size_t f = face_index(P, dim, i + foff[nq][0], j + foff[nq][1], k + foff[nq][2]);
			double area = areas[dim][f];
			if (area) {
				double rho = rhos[dim][f];
				if (rho) {
					double value = dt * area / (dx * dx * rho);
					size_t q = cell_index(P, qi, qj, qk);
					if (fluid[q] < 0.0) {
						cols[cnt] = index_map[q]; vals[cnt] = -value; cnt++;
					} else dirichlet += value;
					diagonal += value;
				}
				rhs += -sgn[nq] * area * vel[dim][f] / dx;
			}
		}
		cols[cnt] = n_index; vals[cnt] = diagonal; cnt++;
		/* sorted insert, like RCMatrix::add_to_element */
		for (int a = 1; a < cnt; ++a) {
			size_t cc = cols[a]; double vv = vals[a]; int b = a - 1;
			while (b >= 0 && cols[b] > cc) { cols[b + 1] = cols[b]; vals[b + 1] = vals[b]; --b; }
			cols[b + 1] = cc; vals[b + 1] = vv;
		}
		A->rowstart[n_index] = nnz;
		for (int a = 0; a < cnt; ++a) { A->col[nnz] = cols[a]; A->val[nnz] = vals[a]; nnz++; }
		A->rhs[n_index] = rhs + P->rhs_correct; /* :204-217 */
		A->dirichlet[n_index] = dirichlet;
	}
	A->rowstart[index] = nnz;
}

static double dot(size_t n, const double *x, const double *y) { double s = 0; for (size_t i = 0; i < n; ++i) s += x[i] * y[i]; return s; }
static double abs_max(size_t n, const double *x) { double m = 0; for (size_t i = 0; i < n; ++i) { double v = fabs(x[i]); if (v > m) m = v; } return m; }
static void spmv(const csr_system *A, const double *x, double *y) {
	for (size_t r = 0; r < A->n; ++r) {
		double s = 0;
		for (size_t e = A->rowstart[r]; e < A->rowstart[r + 1]; ++e) s += A->val[e] * x[A->col[e]];
		y[r] = s;
	}
}

/* local/include/pcgsolver/pcg_solver.h:246-295 with apply_preconditioner == identity (:379-384,
 * the MIC(0) result is overwritten by `result = x`), called through src/linsolver/pcg.cpp:45-73. */
static void cg_solve(const oracle_params *P, const csr_system *A, double *x, oracle_stats *st) {
	const size_t n = A->n;
	double *r = (double *)malloc(sizeof(double) * (n ? n : 1));
	double *s = (double *)malloc(sizeof(double) * (n ? n : 1));
	double *z = (double *)malloc(sizeof(double) * (n ? n : 1));
	double tolerance_factor = P->residual < 1e-30 ? 1e-30 : P->residual; /* :239 */
	memset(x, 0, sizeof(double) * n);
	memcpy(r, A->rhs, sizeof(double) * n);
	double residual_out = abs_max(n, r), residual0 = residual_out;
	st->rhs_absmax = residual0;
	st->iterations = 0; st->converged = 1; st->reresid = 0.0;
	if (residual_out == 0) goto done;
	{
		double tol = tolerance_factor * residual_out;
		memcpy(z, r, sizeof(double) * n);
		double rho = dot(n, z, r);
		if (rho == 0 || rho != rho) { st->converged = 0; goto done; }
		memcpy(s, z, sizeof(double) * n);
		uint32_t iteration;
		st->converged = 0;
		for (iteration = 0; iteration < P->max_iterations; ++iteration) {
			spmv(A, s, z);
			double alpha = rho / dot(n, s, z);
			for (size_t i = 0; i < n; ++i) x[i] += alpha * s[i];
			for (size_t i = 0; i < n; ++i) r[i] += -alpha * z[i];
			residual_out = abs_max(n, r);
			if (residual_out <= tol) { st->iterations = iteration + 1; st->converged = 1; break; }
			memcpy(z, r, sizeof(double) * n);
			double rho_new = dot(n, z, r);
			double beta = rho_new / rho;
			for (size_t i = 0; i < n; ++i) z[i] += beta * s[i];
			{ double *t = s; s = z; z = t; }
			rho = rho_new;
		}
		if (!st->converged) st->iterations = iteration;
		st->reresid = residual_out / residual0;
	}
done:
	free(r); free(s); free(z);
}

/*
 * Whole project() call (src/projection/macpressuresolver3.cpp:50-272).
 *   vel[3], vel_active[3]: in/out, face-shaped; pressure / in_rows: out, cell-shaped.
 *   areas_out / rhos_out / rhs_out / diag_out: optional dumps (may be NULL) for kernel-level tests:
 *   rhs_out, diag_out and dirichlet_out are cell-shaped (0 outside the row set); dirichlet_out is the
 *   part of the diagonal contributed by air neighbours (diag = sum of couplings + dirichlet).
 */
int oracle_project_warm(const oracle_params *P, double *vel[3], uint8_t *vel_active[3], const double *solid, const double *fluid,
                        double *pressure, uint8_t *in_rows, double *areas_out[3], double *rhos_out[3],
                        double *rhs_out, double *diag_out, double *dirichlet_out, oracle_stats *st,
                        double *prev_pressure, uint64_t *prev_rows);

int oracle_project(const oracle_params *P, double *vel[3], uint8_t *vel_active[3], const double *solid, const double *fluid,
                   double *pressure, uint8_t *in_rows, double *areas_out[3], double *rhos_out[3],
                   double *rhs_out, double *diag_out, double *dirichlet_out, oracle_stats *st) {
	return oracle_project_warm(P, vel, vel_active, solid, fluid, pressure, in_rows, areas_out, rhos_out, rhs_out, diag_out, dirichlet_out, st, NULL, NULL);
}

/* ... with WarmStart=Yes (macpressuresolver3.cpp:221-230, 239-242) when prev_pressure != NULL: the previous result, indexed by ROW NUMBER
 * as the reference keeps it (capacity nx*ny*nz doubles; *prev_rows = its current length, "resize(index)" pads with zeros / truncates),
 * is multiplied into the right-hand side before the solve, added back to the solution afterwards and replaced by the sum. */
int oracle_project_warm(const oracle_params *P, double *vel[3], uint8_t *vel_active[3], const double *solid, const double *fluid,
                        double *pressure, uint8_t *in_rows, double *areas_out[3], double *rhos_out[3],
                        double *rhs_out, double *diag_out, double *dirichlet_out, oracle_stats *st,
                        double *prev_pressure, uint64_t *prev_rows) {
	const int nx = P->nx, ny = P->ny, nz = P->nz;
	const size_t ncell = (size_t)nx * ny * nz;
	double *areas[3], *rhos[3];
	for (int dim = 0; dim < 3; ++dim) {
		areas[dim] = (double *)malloc(sizeof(double) * face_count(P, dim));
		rhos[dim] = (double *)malloc(sizeof(double) * face_count(P, dim));
	}
	oracle_area_fraction(P, solid, areas);
	oracle_fluid_fraction(P, fluid, rhos);
	if (P->surface_tension) surface_tension_force(P, fluid, rhos, vel, (const uint8_t *const *)vel_active);
	size_t *index_map = (size_t *)malloc(sizeof(size_t) * ncell);
	csr_system A;
	memset(&A, 0, sizeof A);
	assemble(P, fluid, areas, rhos, vel, index_map, &A);
	st->n_rows = A.n;
	st->nnz = A.rowstart[A.n];
	double *x = (double *)calloc(A.n ? A.n : 1, sizeof(double));
	if (prev_pressure) { /* :221-230 */
		for (size_t r = (size_t)*prev_rows; r < A.n; ++r) prev_pressure[r] = 0.0;
		*prev_rows = A.n;
		double *ap = (double *)malloc(sizeof(double) * (A.n ? A.n : 1));
		spmv(&A, prev_pressure, ap);
		for (size_t r = 0; r < A.n; ++r) A.rhs[r] -= ap[r];
		free(ap);
	}
	cg_solve(P, &A, x, st);
	if (prev_pressure) /* :239-242 */
		for (size_t r = 0; r < A.n; ++r) { x[r] += prev_pressure[r]; prev_pressure[r] = x[r]; }
	/* :245-248 scatter to the Real pressure grid; activity == row set */
	for (size_t c = 0; c < ncell; ++c) {
		int in = index_map[c] != SIZE_MAX;
		in_rows[c] = (uint8_t)in;
		pressure[c] = in ? RR(P, x[index_map[c]]) : 0.0;
		if (rhs_out) rhs_out[c] = in ? A.rhs[index_map[c]] : 0.0;
		if (dirichlet_out) dirichlet_out[c] = in ? A.dirichlet[index_map[c]] : 0.0;
		if (diag_out) {
			double dg = 0.0;
			if (in) for (size_t e = A.rowstart[index_map[c]]; e < A.rowstart[index_map[c] + 1]; ++e) if (A.col[e] == index_map[c]) dg = A.val[e];
			diag_out[c] = dg;
		}
	}
	/* :252-268 velocity update on ACTIVE faces */
	for (int dim = 0; dim < 3; ++dim) {
		const int w = nx + (dim == 0), h = ny + (dim == 1), d = nz + (dim == 2);
		const int n_dim = dim == 0 ? nx : (dim == 1 ? ny : nz);
		for (int k = 0; k < d; ++k) for (int j = 0; j < h; ++j) for (int i = 0; i < w; ++i) {
			size_t f = face_index(P, dim, i, j, k);
			if (!vel_active[dim][f]) continue;
			const int pd = dim == 0 ? i : (dim == 1 ? j : k);
			double rho = rhos[dim][f];
			if (areas[dim][f] && rho) {
				if (pd == 0 || pd == n_dim) vel[dim][f] = 0.0;
				else {
					double p1 = pressure[cell_index(P, i, j, k)];
					double p0 = pressure[cell_index(P, i - (dim == 0), j - (dim == 1), k - (dim == 2))];
					if (P->real_is_double) {
						vel[dim][f] -= P->dt * (p1 - p0) / (rho * P->dx);
					} else {
						float diff = (float)p1 - (float)p0;                        /* float - float */
						float delta = (float)(P->dt * diff / (rho * P->dx));       /* converted to T by subtract() */
						vel[dim][f] = (double)((float)vel[dim][f] - delta);         /* array3.h:663-671 */
					}
				}
			} else {
				if (pd == 0 && fluid[cell_index(P, i, j, k)] < 0.0) vel[dim][f] = 0.0;
				else if (pd == n_dim && fluid[cell_index(P, i - (dim == 0), j - (dim == 1), k - (dim == 2))] < 0.0) vel[dim][f] = 0.0;
				else { vel_active[dim][f] = 0; vel[dim][f] = 0.0; } /* set_off(): reads back as background 0 */
			}
		}
	}
	for (int dim = 0; dim < 3; ++dim) {
		if (areas_out && areas_out[dim]) memcpy(areas_out[dim], areas[dim], sizeof(double) * face_count(P, dim));
		if (rhos_out && rhos_out[dim]) memcpy(rhos_out[dim], rhos[dim], sizeof(double) * face_count(P, dim));
		free(areas[dim]); free(rhos[dim]);
	}
	free(index_map); free(x); csr_free(&A);
	return 0;
}

/* The PI controller of the volume correction (macpressuresolver3.cpp:204-214); y_prev is caller state. */
double oracle_volume_correction(double gain, double current_volume, double target_volume, double dt, double *y_prev) {
	if (!(gain && target_volume)) return 0.0;
	double x = (current_volume - target_volume) / target_volume;
	double y = *y_prev + x * dt; *y_prev = y;
	double kp = gain * 2.3 / (25.0 * 0.01);
	double ki = kp * kp / 16.0;
	return -(kp * x + ki * y) / (x + 1.0);
}
