/*
 * b200pressure3 — Shiokaze module that replaces `macpressuresolver3` with the B200 CUDA path.
 *
 * Built as libshiokaze_b200pressure3.so against the reference's own headers and selected at run time with
 *     Projection=b200pressure3
 * through the reference's module loader (src/core/module.cpp:103-170: "lib"+"shiokaze_"+name+".so" via dlopen,
 * then extern "C" create_instance()). It implements macproject3_interface
 * (include/shiokaze/projection/macproject3_interface.h:38-101) with the same flags, console records and
 * timer names as src/projection/macpressuresolver3.cpp, and delegates every computation to the C-ABI of
 * include/shkz_b200.h. It contains no numerical code of its own and no CPU fallback: if the CUDA library
 * reports an error the module prints it and exits, which is the reference's own fatal-error convention
 * (src/core/module.cpp:67,114,142).
 *
 * Bridge contract: dense reads are array3::linearize() semantics
 * (include/shiokaze/array/array3.h:198-208), velocity writes touch ACTIVE faces only (set / set_off),
 * the pressure grid is cleared and then set exactly on the row set.
 */
#include <shiokaze/array/array3.h>
#include <shiokaze/array/array_utility3.h>
#include <shiokaze/array/macarray3.h>
#include <shiokaze/core/console.h>
#include <shiokaze/core/timer.h>
#include <shiokaze/projection/macproject3_interface.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
//
#include "../../include/shkz_b200.h"
#include "b200dense.h"
//
SHKZ_USING_NAMESPACE
//
class b200pressure3 : public macproject3_interface {
protected:
	//
	LONG_NAME("B200 MAC Pressure Solver 3D")
	//
	virtual void set_target_volume( double current_volume, double target_volume ) override {
		m_current_volume = current_volume;
		m_target_volume = target_volume;
	}
	//
	static void fatal( const char *what ) {
		console::dump( "<Red>b200pressure3: %s failed: %s<Default>\n", what, shkz_b200_last_error());
		exit(-1);
	}
	// Page-locked staging buffers (shkz_b200_host_alloc): one per grid handed to the C-ABI, kept for the life of the module
	// (pageable std::vector storage made the H2D / D2H copies the largest part of the GPU call)
	template <class T> struct staging {
		T *p {nullptr};
		size_t n {0};
		T * ensure( size_t count ) {
			if( count > n ) {
				release();
				void *q (nullptr);
				if( shkz_b200_host_alloc(count*sizeof(T),&q) != SHKZ_B200_OK ) fatal("shkz_b200_host_alloc");
				p = static_cast<T *>(q); n = count;
			}
			return p;
		}
		void release() { if( p ) shkz_b200_host_free(p); p = nullptr; n = 0; }
	};
	// array3::linearize() semantics (array3.h:198-208) with the activity mask alongside. A grid that is mostly active (the velocity
	// of a smoke solver) is read in ONE pass over all cells — the iterator hands out the active value, else the fill / background
	// value, exactly what operator() returns (array3.h:796-801) —, a sparse one (a narrow-band level set) by its actives and its inside.
	static void gather( const array3<Real> &a, Real *values, uint8_t *active ) {
		const shape3 s = a.shape();
		const size_t total = s.count();
		if( 2 * a.count() >= total ) {
			a.const_parallel_all([&]( int i, int j, int k, const auto &it ) {
				const size_t n = i + s.w * (j + s.h * (size_t)k);
				values[n] = it();
				if( active ) active[n] = it.active() ? 1 : 0;
			});
			return;
		}
		b200_parallel_fill(values,total,a.get_background_value());
		if( active ) b200_parallel_fill(active,total,(uint8_t)0);
		a.const_parallel_actives([&]( int i, int j, int k, const auto &it ) {
			const size_t n = i + s.w * (j + s.h * (size_t)k);
			values[n] = it();
			if( active ) active[n] = 1;
		});
		a.const_parallel_inside([&]( int i, int j, int k, const auto &it ) {
			if( ! it.active()) values[i + s.w * (j + s.h * (size_t)k)] = it();
		});
	}
	//
	// Zero-copy view of a grid whose core is b200array3 (`Array=b200array3`, plugin/b200array3.cpp): its value buffer and its one-byte-per-cell activity
	// mask ARE the dense buffers the C-ABI takes, page-locked. Entries of inactive cells are first rewritten to what array3::operator() returns for them
	// (fill value inside, background outside: array3.h:796-801) — a plain threaded sweep over memory instead of one std::function call per cell.
	// as_they_are: hand the buffers over without that sweep (velocity grids with params.velocity_masked: the library reads an inactive face as 0 itself)
	static bool dense_view( const array3<Real> &a, Real *&values, uint8_t *&active, bool as_they_are=false ) {
		b200_dense_descriptor d;
		if( ! a.const_send_message(B200_DENSE_MESSAGE,&d) || ! d.pinned || d.element_bytes != sizeof(Real) || ! d.values ) return false;
		Real *v = static_cast<Real *>(d.values);
		if( as_they_are ) { values = v; active = d.active; return true; }
		const size_t plane = (size_t)d.nx*d.ny;
		const Real background = a.get_background_value();
		Real fill = background;
		if( d.filled ) { // (the fill value has no getter: read it where it shows, at the first filled inactive cell)
			for( size_t n=0; n<plane*d.nz; ++n ) if( d.filled[n] && ! d.active[n] ) {
				fill = a((int)(n%d.nx),(int)((n%plane)/d.nx),(int)(n/plane));
				break;
			}
		}
		const unsigned nthreads = std::max(1u,std::min((unsigned)NUM_THREAD,d.nz));
		std::vector<std::thread> pool;
		for( unsigned t=0; t<nthreads; ++t ) pool.emplace_back([&,t]() {
			for( size_t n=plane*(d.nz*(size_t)t/nthreads); n<plane*(d.nz*(size_t)(t+1)/nthreads); ++n ) {
				if( ! d.active[n] ) v[n] = (d.filled && d.filled[n]) ? fill : background;
			}
		});
		for( auto &t : pool ) t.join();
		values = v; active = d.active;
		return true;
	}
	//
	virtual void project( double dt,
				macarray3<Real> &velocity,
				const array3<Real> &solid,
				const array3<Real> &fluid,
				double surface_tension,
				const std::vector<signed_rigidbody3_interface *> *rigidbodies ) override {
		//
		scoped_timer timer(this);
		timer.tick(); console::dump( ">>> Pressure Projection (B200) started...\n" );
		//
		// Host -> dense buffers
		timer.tick(); console::dump( "Gathering dense buffers..." );
		Real *vel[DIM3], *solid_dense (nullptr), *fluid_dense;
		uint8_t *vel_active[DIM3], *unused (nullptr);
		bool vel_in_place[DIM3];
		// Velocity grids on the dense core whose background value is 0 (every simulator's) go over as they are: rewriting their inactive entries was a host
		// sweep over three whole face grids per call — two thirds of this phase at 256^3 — for values the library can just as well read as 0 through the masks
		// it gets anyway (params.velocity_masked). One device only: the z-slab path keeps the rewritten entries.
		bool velocity_masked = m_slabs.empty();
		for( int dim : DIMS3 ) {
			b200_dense_descriptor probe;
			velocity_masked = velocity_masked && velocity[dim].get_background_value() == Real(0) && velocity[dim].const_send_message(B200_DENSE_MESSAGE,&probe) &&
				probe.pinned && probe.element_bytes == sizeof(Real) && probe.values;
		}
		for( int dim : DIMS3 ) {
			vel_in_place[dim] = dense_view(velocity[dim],vel[dim],vel_active[dim],velocity_masked); // b200array3 grids are read and written where they live
			if( ! vel_in_place[dim] ) {
				const size_t nf = velocity[dim].shape().count();
				vel[dim] = m_hvel[dim].ensure(nf);
				vel_active[dim] = m_hact[dim].ensure(nf);
				gather(velocity[dim],vel[dim],vel_active[dim]);
			}
		}
		if( ! dense_view(fluid,fluid_dense,unused)) {
			fluid_dense = m_hfluid.ensure(m_shape.count());
			gather(fluid,fluid_dense,nullptr);
		}
		const bool fluid_levelset = array_utility3::levelset_exist(fluid);
		bool have_solid = array_utility3::levelset_exist(solid);
		if( have_solid && solid.shape() != m_shape.nodal()) {
			console::dump( "<Red>b200pressure3: a solid level set must be nodal (shape+1).<Default>\n" );
			exit(-1);
		}
		if( have_solid && ! dense_view(solid,solid_dense,unused)) {
			solid_dense = m_hsolid.ensure(solid.shape().count());
			gather(solid,solid_dense,nullptr);
		}
		console::dump( "Done. Took %s\n", timer.stock("gather").c_str());
		//
		// Volume correction: the PI controller is host state, as in macpressuresolver3.cpp:204-217
		shkz_b200_params params = m_cuda_param;
		params.surface_tension = surface_tension;
		params.apply_rhs_correct = 0;
		params.rhs_correct = 0.0;
		params.velocity_masked = velocity_masked ? 1 : 0;
		if( m_param.gain && m_target_volume ) {
			timer.tick(); console::dump( "Computing volume correction...");
			double x = (m_current_volume-m_target_volume)/m_target_volume;
			double y = m_y_prev + x*dt; m_y_prev = y;
			double kp = m_param.gain * 2.3/(25.0*0.01);
			double ki = kp*kp/16.0;
			params.rhs_correct = -(kp*x+ki*y)/(x+1.0);
			params.apply_rhs_correct = 1;
			console::dump( "Done. Took %s\n", timer.stock("volume_correction").c_str()); // (the constant is added to the rows by the assembly kernel)
			console::write(get_argument_name()+"_volume_correct_rhs", params.rhs_correct);
		}
		//
		// The CUDA path
		timer.tick(); console::dump( "Solving on the GPU...");
		Real *pressure (nullptr);
		uint8_t *pressure_active (nullptr);
		const bool pressure_in_place = dense_view(m_pressure,pressure,pressure_active); // (the call overwrites every value and every activity byte)
		if( ! pressure_in_place ) {
			pressure = m_hpressure.ensure(m_shape.count());
			pressure_active = m_hpact.ensure(m_shape.count());
		}
		void *vel_ptr[3] = { vel[0], vel[1], vel[2] };
		uint8_t *act_ptr[3] = { vel_active[0], vel_active[1], vel_active[2] };
		shkz_b200_stats stats;
		if( m_slabs.empty()) {
			if( shkz_b200_project_host(m_solver,dt,vel_ptr,act_ptr,have_solid ? solid_dense : nullptr,fluid_dense,
					fluid_levelset,&params,pressure,pressure_active,&stats) != SHKZ_B200_OK ) fatal("shkz_b200_project_host");
		} else {
			project_slabs(dt,vel,vel_active,have_solid ? solid_dense : nullptr,fluid_dense,fluid_levelset,params,pressure,pressure_active,stats);
		}
		// the reference's records (macpressuresolver3.cpp:82,114,201,215,236-237,269,271 through scoped_timer::stock -> "<Arg>_<name>", milliseconds), fed
		// from the CUDA-event times of the phases that replace them: fractions + labelling + assembly are ONE fused phase here (ms_assemble; the surface-tension
		// part of it is reported separately like the reference does), the multigrid hierarchy is this solver's "build" step, the MG-PCG loop its "linsolve"
		const std::string arg = get_argument_name();
		console::write(arg+"_number_projection_iteration", stats.iterations);
		console::write(arg+"_solid_fluid_fractions", stats.ms_assemble-stats.ms_surftension);
		if( surface_tension ) console::write(arg+"_surftension_force", stats.ms_surftension);
		console::write(arg+"_build_highres_linsystem", stats.ms_setup);
		console::write(arg+"_linsolve", stats.ms_solve);
		console::write(arg+"_update_velocity", stats.ms_update);
		console::write(arg+"_h2d", stats.ms_h2d);
		console::write(arg+"_d2h", stats.ms_d2h);
		console::dump( "Done. Took %d iterations, Reresid=%e. Took %s\n", stats.iterations, stats.reresid, timer.stock("gpu_call").c_str());
		//
		// Dense buffers -> host grids
		timer.tick(); console::dump( "Scattering results...");
		// (exactly the row set is activated, macpressuresolver3.cpp:245-248; parallel_all + it.set() is the reference's own way of
		// filling a grid in parallel, e.g. macutility3.cpp:341-349)
		if( ! pressure_in_place ) {
			m_pressure.clear();
			m_pressure.parallel_all([&]( int i, int j, int k, auto &it ) {
				const size_t n = i + m_shape.w * (j + m_shape.h * (size_t)k);
				if( pressure_active[n] ) it.set(pressure[n]);
			});
		}
		for( int dim : DIMS3 ) if( ! vel_in_place[dim] ) {
			const shape3 s = velocity[dim].shape();
			if( params.extrapolate_width > 0 ) {
				// the extrapolation ACTIVATES faces around the active set (array_extrapolator3.h:51-82): every face is visited
				velocity[dim].parallel_all([&]( int i, int j, int k, auto &it, int tn ) {
					const size_t n = i + s.w * (j + s.h * (size_t)k);
					if( vel_active[dim][n] ) it.set(vel[dim][n]);
					else if( it.active()) it.set_off();
				});
			} else {
				// the projection itself never activates a face (macpressuresolver3.cpp:252-268): the active ones are enough
				velocity[dim].parallel_actives([&]( int i, int j, int k, auto &it, int tn ) {
					const size_t n = i + s.w * (j + s.h * (size_t)k);
					if( vel_active[dim][n] ) it.set(vel[dim][n]);
					else it.set_off();
				});
			}
		}
		console::dump( "Done. Took %s\n", timer.stock("scatter").c_str());
		console::dump( "<<< Projection done. Took %s.\n", timer.stock("projection").c_str());
	}
	//
	// GPUs=N: the grid cut into N equal z-slabs, one CUDA device and one host thread per slab (every call blocks on its own device while the slabs talk
	// to each other through peer memory inside the kernels: they must run concurrently). z is the slowest index of the dense layout, so a slab of the cell /
	// x-face / y-face / nodal grids is a contiguous range of the gathered buffers and is handed over in place; only the z-face grids, whose plane between two
	// slabs belongs to both, go through a per-slab copy.
	void project_slabs( double dt, Real *vel[DIM3], uint8_t *vel_active[DIM3], const Real *solid, const Real *fluid, bool fluid_levelset,
				const shkz_b200_params &params, Real *pressure, uint8_t *pressure_active, shkz_b200_stats &stats ) {
		const int world = (int)m_slabs.size();
		const size_t nx = m_shape.w, ny = m_shape.h, nzl = m_shape.d / world;
		const size_t cell_plane = nx*ny, xf_plane = (nx+1)*ny, yf_plane = nx*(ny+1), node_plane = (nx+1)*(ny+1);
		std::vector<int> rc (world,SHKZ_B200_OK);
		std::vector<std::string> message (world);
		std::vector<shkz_b200_stats> slab_stats (world);
		// everything a first call allocates is allocated here, slab by slab, before any slab's kernels start waiting for its neighbours (shkz_b200.h)
		for( int r=0; r<world; ++r ) {
			if( shkz_b200_prepare(m_slabs[r],&params,1,solid != nullptr) != SHKZ_B200_OK ) fatal("shkz_b200_prepare");
		}
		std::vector<std::thread> threads;
		for( int r=0; r<world; ++r ) {
			Real *wz = m_slab_w[r].ensure(cell_plane*(nzl+1));
			uint8_t *az = m_slab_wact[r].ensure(cell_plane*(nzl+1));
			std::copy(vel[2]+r*nzl*cell_plane,vel[2]+(r*nzl+nzl+1)*cell_plane,wz);
			std::copy(vel_active[2]+r*nzl*cell_plane,vel_active[2]+(r*nzl+nzl+1)*cell_plane,az);
			threads.emplace_back([&,r,wz,az]() {
				const size_t k0 = r*nzl;
				void *v[3] = { vel[0]+k0*xf_plane, vel[1]+k0*yf_plane, wz };
				uint8_t *a[3] = { vel_active[0]+k0*xf_plane, vel_active[1]+k0*yf_plane, az };
				rc[r] = shkz_b200_project_host(m_slabs[r],dt,v,a,solid ? solid+k0*node_plane : nullptr,fluid+k0*cell_plane,fluid_levelset,&params,
					pressure+k0*cell_plane,pressure_active+k0*cell_plane,&slab_stats[r]);
				if( rc[r] != SHKZ_B200_OK ) message[r] = shkz_b200_last_error(); // (the message is thread-local)
			});
		}
		for( auto &t : threads ) t.join();
		for( int r=0; r<world; ++r ) if( rc[r] != SHKZ_B200_OK ) {
			console::dump( "<Red>b200pressure3: slab %d of %d failed: %s<Default>\n", r, world, message[r].c_str());
			exit(-1);
		}
		// z faces back: every slab owns its lower planes, the last one also the top plane (the shared planes agree between neighbours)
		for( int r=0; r<world; ++r ) {
			const size_t planes = nzl + (r == world-1 ? 1 : 0);
			std::copy(m_slab_w[r].p,m_slab_w[r].p+planes*cell_plane,vel[2]+r*nzl*cell_plane);
			std::copy(m_slab_wact[r].p,m_slab_wact[r].p+planes*cell_plane,vel_active[2]+r*nzl*cell_plane);
		}
		// every slab reports the same globally reduced solve; phase times are the slowest slab's
		stats = slab_stats[0];
		for( int r=1; r<world; ++r ) {
			stats.ms_h2d = std::max(stats.ms_h2d,slab_stats[r].ms_h2d); stats.ms_assemble = std::max(stats.ms_assemble,slab_stats[r].ms_assemble);
			stats.ms_setup = std::max(stats.ms_setup,slab_stats[r].ms_setup); stats.ms_solve = std::max(stats.ms_solve,slab_stats[r].ms_solve);
			stats.ms_update = std::max(stats.ms_update,slab_stats[r].ms_update); stats.ms_d2h = std::max(stats.ms_d2h,slab_stats[r].ms_d2h);
			stats.ms_total = std::max(stats.ms_total,slab_stats[r].ms_total); stats.kernel_launches += slab_stats[r].kernel_launches;
		}
	}
	//
	virtual void configure( configuration &config ) override {
		// the reference module's flags (macpressuresolver3.cpp:274-280)
		bool second_order_fluid (true), second_order_solid (true), warm_start (false);
		config.get_bool("SecondOrderAccurateFluid",second_order_fluid,"Whether to enforce second order accuracy");
		config.get_bool("SecondOrderAccurateSolid",second_order_solid,"Whether to enforce second order accuracy for solid surfaces");
		config.get_double("Gain",m_param.gain,"Rate for volume correction");
		config.get_bool("WarmStart",warm_start,"Start from the solution of previous pressure");
		config.set_default_bool("ReportProgress",false);
		shkz_b200_default_params(&m_cuda_param);
		m_cuda_param.warm_start = warm_start; // the previous pressure lives on the device, per cell (shkz_b200.h)
		m_cuda_param.second_order_fluid = second_order_fluid;
		m_cuda_param.second_order_solid = second_order_solid;
		// the children's flags, under the groups the reference uses (macutility3.cpp:408-409, pcg.cpp:39-44)
		{
			configuration::auto_group group(config,"MAC Utility 3D","MacUtility");
			config.get_double("EpsFluid",m_cuda_param.eps_fluid,"Minimal bound for fluid fraction");
			config.get_double("EpsSolid",m_cuda_param.eps_solid,"Minimal bound for solid fraction");
		}
		{
			configuration::auto_group group(config,"Linear System Solver","LinSolver");
			double modified_ic (0.97), min_diag_ratio (0.25);
			config.get_double("Residual",m_cuda_param.residual,"Tolerable residual");
			config.get_unsigned("MaxIterations",m_cuda_param.max_iterations,"Maximal iteration count");
			config.get_double("ModifiedIC",modified_ic,"Accepted for compatibility (the reference discards its MIC(0) result)");
			config.get_double("MinDiagRatio",min_diag_ratio,"Accepted for compatibility");
		}
		// additive flags
		std::string precision ("mixed"), precond ("mg");
		config.get_string("Precision",precision,"CG arithmetic: mixed, fp64 or fp32");
		config.get_string("Precond",precond,"Preconditioner: mg (multigrid V-cycle) or none (the reference's plain CG)");
		config.get_integer("MGPreSweeps",m_cuda_param.mg_pre_sweeps,"Red-black sweeps before the coarse correction");
		config.get_integer("MGPostSweeps",m_cuda_param.mg_post_sweeps,"Red-black sweeps after the coarse correction");
		config.get_double("MGOmega",m_cuda_param.mg_omega,"Relaxation factor of the red-black sweeps (1 = Gauss-Seidel)");
		config.get_integer("ExtrapolateWidth",m_cuda_param.extrapolate_width,"> 0: finish project() with the velocity extrapolation + solid constraint on the device (macutility3::extrapolate_and_constrain_velocity)");
		config.get_integer("GPU",m_device,"CUDA device index (of the first slab when GPUs > 1)");
		config.get_integer("GPUs",m_gpus,"Number of CUDA devices: the grid is cut into this many equal z-slabs (1, 2, 4 or 8)");
		m_cuda_param.precision = precision == "fp64" ? SHKZ_B200_PREC_FP64 : (precision == "fp32" ? SHKZ_B200_PREC_FP32 : SHKZ_B200_PREC_MIXED);
		m_cuda_param.precond = precond == "none" ? SHKZ_B200_PRECOND_NONE : SHKZ_B200_PRECOND_MG;
	}
	virtual void initialize( const shape3 &shape, double dx ) override {
		m_shape = shape;
		m_dx = dx;
	}
	virtual void post_initialize() override {
		m_pressure.initialize(m_shape);
		m_target_volume = m_current_volume = m_y_prev = 0.0;
		release_solvers();
		const int real = sizeof(Real) == sizeof(double) ? SHKZ_B200_REAL_F64 : SHKZ_B200_REAL_F32;
		if( m_gpus <= 1 ) {
			if( shkz_b200_create(m_shape.w,m_shape.h,m_shape.d,m_dx,real,m_device,&m_solver) != SHKZ_B200_OK ) fatal("shkz_b200_create");
		} else {
			if( m_gpus > 8 || m_shape.d % m_gpus ) {
				console::dump( "<Red>b200pressure3: GPUs=%d needs a z extent (%d) divisible by it, and at most 8 devices.<Default>\n", m_gpus, (int)m_shape.d );
				exit(-1);
			}
			if( m_device+m_gpus > shkz_b200_device_count()) {
				console::dump( "<Red>b200pressure3: GPUs=%d from device %d, but %d CUDA device(s) visible.<Default>\n", m_gpus, m_device, shkz_b200_device_count());
				exit(-1);
			}
			const int nzl = m_shape.d / m_gpus;
			m_slabs.assign(m_gpus,nullptr);
			for( int r=0; r<m_gpus; ++r ) {
				if( shkz_b200_create_slab(m_shape.w,m_shape.h,m_shape.d,r*nzl,(r+1)*nzl,m_dx,real,m_device+r,&m_slabs[r]) != SHKZ_B200_OK ) fatal("shkz_b200_create_slab");
			}
			if( shkz_b200_slab_connect_local(m_slabs.data(),m_gpus) != SHKZ_B200_OK ) fatal("shkz_b200_slab_connect_local");
			m_slab_w.resize(m_gpus); m_slab_wact.resize(m_gpus);
		}
	}
	virtual const array3<Real> * get_pressure() const override {
		return &m_pressure;
	}
	virtual ~b200pressure3() {
		for( int dim : DIMS3 ) { m_hvel[dim].release(); m_hact[dim].release(); }
		m_hsolid.release(); m_hfluid.release(); m_hpressure.release(); m_hpact.release();
		release_solvers();
	}
	void release_solvers() {
		if( m_solver ) { shkz_b200_destroy(m_solver); m_solver = nullptr; }
		for( auto *s : m_slabs ) if( s ) shkz_b200_destroy(s);
		m_slabs.clear();
		for( auto &b : m_slab_w ) b.release();
		for( auto &b : m_slab_wact ) b.release();
	}
	//
	struct Parameters {
		double gain {1.0};
	};
	Parameters m_param;
	shkz_b200_params m_cuda_param;
	shkz_b200_solver *m_solver {nullptr};
	int m_device {0};
	int m_gpus {1};
	std::vector<shkz_b200_solver *> m_slabs;   // GPUs > 1: one z-slab solver per device, wired through peer memory
	std::vector<staging<Real>> m_slab_w;       // ... and the slabs' copies of the z-face grid
	std::vector<staging<uint8_t>> m_slab_wact;
	staging<Real> m_hvel[DIM3], m_hsolid, m_hfluid, m_hpressure;
	staging<uint8_t> m_hact[DIM3], m_hpact;
	//
	shape3 m_shape;
	double m_dx {0.0};
	array3<Real> m_pressure{this};
	//
	double m_target_volume {0.0};
	double m_current_volume {0.0};
	double m_y_prev {0.0};
};
//
extern "C" module * create_instance() {
	return new b200pressure3();
}
//
extern "C" const char *license() {
	return "MIT";
}
//
