/*
 * b200cg — Shiokaze `LinSolver` module that solves an assembled RCMatrix system on the GPU.
 *
 * Built as libshiokaze_b200cg.so against the reference's own headers and selected at run time with
 *     LinSolver=b200cg
 * wherever the reference loads a linear solver (macpressuresolver3.cpp:296-305, macstreamfuncsolver3, the 2-D
 * solvers). It implements RCMatrix_solver_interface<size_t,double>
 * (include/shiokaze/linsolver/RCMatrix_solver.h:41-110) like the reference's pcg module (src/linsolver/pcg.cpp:30-82):
 * same flags (Residual, MaxIterations; ModifiedIC / MinDiagRatio are read and ignored — the reference's result
 * ignores them too, pcg_solver.h:374-383), same Result{count, reresid}. The rows are flattened to CSR with the
 * interface's own iterators and handed to shkz_b200_csr_solve_host (include/shkz_b200.h); there is no numerical
 * code and no CPU fallback here: a CUDA error is fatal, the host's own convention (src/core/module.cpp:67,114,142).
 */
#include <shiokaze/core/console.h>
#include <shiokaze/linsolver/RCMatrix_solver.h>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
//
#include "../../include/shkz_b200.h"
//
SHKZ_USING_NAMESPACE
//
template <class N, class T> class b200cg_solver : public RCMatrix_solver_interface<N,T> {
protected:
	//
	LONG_NAME("B200 Conjugate Gradient Solver")
	ARGUMENT_NAME("B200CG")
	//
	virtual void configure( configuration &config ) override {
		shkz_b200_csr_default_params(&m_param);
		double modified_ic (0.97), min_diag_ratio (0.25);
		config.get_double("Residual",m_param.residual,"Tolerable residual");
		config.get_double("ModifiedIC",modified_ic,"Accepted for compatibility (the reference discards its MIC(0) result)");
		config.get_double("MinDiagRatio",min_diag_ratio,"Accepted for compatibility");
		config.get_unsigned("MaxIterations",m_param.max_iterations,"Maximal iteration count");
		std::string precond ("none");
		config.get_string("Precond",precond,"Preconditioner: none (the reference's effective algorithm) or jacobi");
		config.get_integer("GPU",m_device,"CUDA device index");
		m_param.precond = precond == "jacobi" ? SHKZ_B200_CSR_PRECOND_JACOBI : SHKZ_B200_CSR_PRECOND_NONE;
	}
	virtual void post_initialize() override {
		release();
		if( shkz_b200_csr_create(m_device,&m_solver) != SHKZ_B200_OK ) fatal("shkz_b200_csr_create");
	}
	static void fatal( const char *what ) {
		console::dump( "<Red>b200cg: %s failed: %s<Default>\n", what, shkz_b200_csr_last_error());
		exit(-1);
	}
	virtual typename RCMatrix_solver_interface<N,T>::Result solve( const RCMatrix_interface<N,T> *A, const RCMatrix_vector_interface<N,T> *b, RCMatrix_vector_interface<N,T> *x ) const override {
		//
		const N n = A->rows();
		if( ! m_solver && shkz_b200_csr_create(m_device,&m_solver) != SHKZ_B200_OK ) fatal("shkz_b200_csr_create");
		//
		// Row pointers first (non_zeros is O(1) per row), then every thread flattens its own band of rows
		std::vector<int64_t> rowptr(n+1,0);
		for( N row=0; row<n; ++row ) rowptr[row+1] = rowptr[row] + (int64_t)A->non_zeros(row);
		std::vector<int32_t> col(rowptr[n]);
		std::vector<double> val(rowptr[n]);
		const unsigned nthreads = std::max(1u,std::min(std::thread::hardware_concurrency(),(unsigned)(n/4096+1)));
		std::vector<std::thread> workers;
		for( unsigned t=0; t<nthreads; ++t ) workers.emplace_back([&,t]() {
			for( N row=n*t/nthreads; row<n*(t+1)/nthreads; ++row ) {
				int64_t at = rowptr[row];
				A->const_for_each(row,[&]( N column, T value ) {
					col[at] = (int32_t)column;
					val[at++] = (double)value;
				});
			}
		});
		for( auto &w : workers ) w.join();
		//
		std::vector<double> rhs, result(n);
		b->convert_to(rhs);
		shkz_b200_csr_stats stats;
		if( shkz_b200_csr_solve_host(m_solver,n,rowptr.data(),col.data(),val.data(),rhs.data(),result.data(),&m_param,&stats) != SHKZ_B200_OK ) fatal("shkz_b200_csr_solve_host");
		x->convert_from(result);
		return {(N)stats.iterations,(T)stats.reresid};
	}
	void release() {
		if( m_solver ) { shkz_b200_csr_destroy(m_solver); m_solver = nullptr; }
	}
	virtual ~b200cg_solver() {
		release();
	}
	//
	shkz_b200_csr_params m_param;
	mutable shkz_b200_csr *m_solver {nullptr};
	int m_device {0};
};
//
extern "C" module * create_instance() {
	return new b200cg_solver<INDEX_TYPE,FLOAT_TYPE>();
}
//
extern "C" const char *license() {
	return "MIT";
}
//
