/*
 * b200advection3 — Shiokaze module `Advection=b200advection3` (also `LevelsetAdvection.Advection=...` for the surface tracker's own instance):
 * drop-in for the reference's `macadvection3` (src/advection/macadvection3.cpp) behind macadvection3_interface
 * (include/shiokaze/advection/macadvection3_interface.h:39-89). SURVEY.md section 8(f) rank 4, first part: the step the simulators run right before every
 * projection (src/liquid/macliquid3.cpp:343-346, src/smoke/macsmoke3.cpp:274,281, src/surfacetracker/maclevelsetsurfacetracker3.cpp:49-51).
 *
 * Same flags, names, defaults and descriptions as the reference module (macadvection3.cpp:57-62, :285-289): MacCormack (Yes), WENO (No), TrimNarrowBand (1);
 * additive: GPU (device index). The module only bridges the host's grids to the C-ABI of libshkz_b200 (include/shkz_b200.h: shkz_b200_advect_*): grids on the
 * dense page-locked core (`Array=b200array3`) are handed over in place, grids on the stock cores are gathered with array3::linearize() semantics and the
 * active entries written back. Results are the reference's, bit for bit (tests/test_gpu_advect.py).
 * Records: the totals the reference writes — <Arg>_maccormack_u_<name>, _semilagrangian_u_<name>, _maccormack_cell_<name>, _semilagrangian_cell_<name>
 * (macadvection3.cpp:180,185,273,282); its per-phase records (forward / backward / final) have no counterpart: the backward pass and the limiter are one kernel.
 * Fatal errors follow the reference's convention: console::dump + exit (src/core/module.cpp:67,114,142).
 */
#include <shiokaze/advection/macadvection3_interface.h>
#include <shiokaze/array/array3.h>
#include <shiokaze/array/macarray3.h>
#include <shiokaze/core/console.h>
#include <shiokaze/core/timer.h>
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
class b200advection3 : public macadvection3_interface {
protected:
	//
	LONG_NAME("B200 MAC Advection 3D")
	//
	static void fatal( const char *what ) {
		console::dump( "<Red>b200advection3: %s failed: %s<Default>\n", what, shkz_b200_advect_last_error());
		exit(-1);
	}
	// One dense grid of a call: either the storage of a b200array3 core, in place, or a page-locked staging copy
	struct view {
		Real *values {nullptr};
		uint8_t *active {nullptr};
		bool in_place {false};
	};
	template <class T> struct staging {
		T *p {nullptr};
		size_t n {0};
		T * ensure( size_t count ) {
			if( count > n ) {
				release();
				void *q (nullptr);
				if( shkz_b200_host_alloc(count*sizeof(T),&q) != SHKZ_B200_OK ) {
					console::dump( "<Red>b200advection3: shkz_b200_host_alloc failed: %s<Default>\n", shkz_b200_last_error());
					exit(-1);
				}
				p = static_cast<T *>(q); n = count;
			}
			return p;
		}
		void release() { if( p ) shkz_b200_host_free(p); p = nullptr; n = 0; }
	};
	struct slot { staging<Real> values; staging<uint8_t> active; };
	//
	// What array3::operator() returns for every cell (active value / flood-fill value / background, array3.h:796-801), plus the activity mask
	// dense_read = false (velocity grids): the kernels read a velocity through its mask, what the inactive entries hold is never looked at — no rewrite of them
	view dense( const array3<Real> &a, slot &s, bool want_active, bool dense_read=true ) const {
		view v;
		const shape3 sh = a.shape();
		const size_t total = sh.count();
		b200_dense_descriptor d;
		if( a.const_send_message(B200_DENSE_MESSAGE,&d) && d.pinned && d.element_bytes == sizeof(Real) && d.values ) {
			Real *p = static_cast<Real *>(d.values);
			const size_t plane = (size_t)d.nx*d.ny;
			const Real background = a.get_background_value();
			Real fill = background;
			if( d.filled && dense_read ) {
				for( size_t n=0; n<total; ++n ) if( d.filled[n] && ! d.active[n] ) {
					fill = a((int)(n%d.nx),(int)((n%plane)/d.nx),(int)(n/plane));
					break;
				}
			}
			const unsigned nthreads = std::max(1u,std::min((unsigned)NUM_THREAD,d.nz));
			std::vector<std::thread> pool;
			if( dense_read ) for( unsigned t=0; t<nthreads; ++t ) pool.emplace_back([&,t]() {
				for( size_t n=plane*(d.nz*(size_t)t/nthreads); n<plane*(d.nz*(size_t)(t+1)/nthreads); ++n ) {
					if( ! d.active[n] ) p[n] = (d.filled && d.filled[n]) ? fill : background;
				}
			});
			for( auto &t : pool ) t.join();
			v.values = p; v.active = d.active; v.in_place = true;
			return v;
		}
		v.values = s.values.ensure(total);
		v.active = want_active ? s.active.ensure(total) : nullptr;
		if( 2 * a.count() >= total ) {
			a.const_parallel_all([&]( int i, int j, int k, const auto &it ) {
				const size_t n = i + sh.w * (j + sh.h * (size_t)k);
				v.values[n] = it();
				if( v.active ) v.active[n] = it.active() ? 1 : 0;
			});
			return v;
		}
		b200_parallel_fill(v.values,total,a.get_background_value());
		if( v.active ) b200_parallel_fill(v.active,total,(uint8_t)0);
		a.const_parallel_actives([&]( int i, int j, int k, const auto &it ) {
			const size_t n = i + sh.w * (j + sh.h * (size_t)k);
			v.values[n] = it();
			if( v.active ) v.active[n] = 1;
		});
		a.const_parallel_inside([&]( int i, int j, int k, const auto &it ) {
			if( ! it.active()) v.values[i + sh.w * (j + sh.h * (size_t)k)] = it();
		});
		return v;
	}
	// the advected values go back onto the active entries; the activity never changes (macadvection3.cpp:71-72, :195-196).
	// reset: do what the reference does to its output — clear(), then activate and set (:193-196). On the tiled core clear() also drops the flood-fill
	// state of a level set (cells inside the liquid read the background value afterwards, until the tracker's redistancing fills again,
	// maclevelsetsurfacetracker3.cpp:55): a dense read of the grid after the call then agrees with the reference everywhere, not only on the active set.
	static void write_back( array3<Real> &a, const view &v, bool reset=false ) {
		if( v.in_place ) return;
		const shape3 sh = a.shape();
		if( reset ) {
			a.clear();
			a.parallel_all([&]( int i, int j, int k, auto &it ) {
				const size_t n = i + sh.w * (j + sh.h * (size_t)k);
				if( v.active[n] ) it.set(v.values[n]);
			});
			return;
		}
		a.parallel_actives([&]( int i, int j, int k, auto &it ) {
			it.set(v.values[i + sh.w * (j + sh.h * (size_t)k)]);
		});
	}
	void ensure_handle() {
		if( m_advect ) return;
		if( shkz_b200_advect_create(m_shape.w,m_shape.h,m_shape.d,m_dx,sizeof(Real)==8 ? SHKZ_B200_REAL_F64 : SHKZ_B200_REAL_F32,m_device,&m_advect) != SHKZ_B200_OK ) fatal("shkz_b200_advect_create");
	}
	//
	virtual void advect_scalar( array3<Real> &scalar, const macarray3<Real> &velocity, const array3<Real> &fluid, double dt, std::string name="scalar" ) override {
		//
		scoped_timer timer(this);
		timer.tick(); console::dump( "Advecting %s on the GPU (%s, %s)...", name.c_str(), m_param.maccormack ? "MacCormack" : "semi-lagrangian", m_param.weno ? "WENO" : "Bilinear");
		ensure_handle();
		view q = dense(scalar,m_slot[3],true), f = dense(fluid,m_slot[4],false), vel[DIM3];
		void *vel_ptr[DIM3];
		const uint8_t *act_ptr[DIM3];
		for( int dim : DIMS3 ) {
			vel[dim] = dense(velocity[dim],m_slot[dim],true,false);
			vel_ptr[dim] = vel[dim].values; act_ptr[dim] = vel[dim].active;
		}
		shkz_b200_advect_params P = m_param;
		P.scalar_background = scalar.get_background_value();
		shkz_b200_advect_stats st;
		if( shkz_b200_advect_scalar_host(m_advect,dt,q.values,q.active,vel_ptr,act_ptr,f.values,&P,&st) != SHKZ_B200_OK ) fatal("shkz_b200_advect_scalar_host");
		write_back(scalar,q,scalar.is_levelset());
		console::dump( "Done. Took %s (h2d %.2f, kernels %.2f, d2h %.2f msec)\n", timer.stock((m_param.maccormack ? "maccormack_cell_" : "semilagrangian_cell_")+name).c_str(), st.ms_h2d, st.ms_advect, st.ms_d2h);
	}
	//
	virtual void advect_vector( macarray3<Real> &u, const macarray3<Real> &velocity, const array3<Real> &fluid, double dt, std::string name="vector" ) override {
		//
		// (the reference traces u with itself: its `velocity` argument is not used, macadvection3.cpp:79)
		scoped_timer timer(this);
		timer.tick(); console::dump( "Advecting %s on the GPU (%s, %s)...", name.c_str(), m_param.maccormack ? "MacCormack" : "semi-lagrangian", m_param.weno ? "WENO" : "Bilinear");
		ensure_handle();
		view f = dense(fluid,m_slot[4],false), vel[DIM3];
		void *vel_ptr[DIM3];
		const uint8_t *act_ptr[DIM3];
		for( int dim : DIMS3 ) {
			vel[dim] = dense(u[dim],m_slot[dim],true,false);
			vel_ptr[dim] = vel[dim].values; act_ptr[dim] = vel[dim].active;
		}
		shkz_b200_advect_stats st;
		if( shkz_b200_advect_vector_host(m_advect,dt,vel_ptr,act_ptr,f.values,&m_param,&st) != SHKZ_B200_OK ) fatal("shkz_b200_advect_vector_host");
		for( int dim : DIMS3 ) write_back(u[dim],vel[dim]);
		console::dump( "Done. Took %s (h2d %.2f, kernels %.2f, d2h %.2f msec)\n", timer.stock((m_param.maccormack ? "maccormack_u_" : "semilagrangian_u_")+name).c_str(), st.ms_h2d, st.ms_advect, st.ms_d2h);
	}
	//
	virtual void configure( configuration &config ) override {
		//
		shkz_b200_advect_default_params(&m_param);
		bool use_maccormack (m_param.maccormack != 0), weno_interpolation (m_param.weno != 0);
		config.get_bool("MacCormack",use_maccormack,"Whether to use MacCormack method");
		config.get_bool("WENO",weno_interpolation,"Whether to use WENO interpolation for advection");
		config.get_unsigned("TrimNarrowBand",m_param.trim_narrowband,"Narrow band count to turn to semi-Lagrangian advection");
		config.get_integer("GPU",m_device,"CUDA device index");
		m_param.maccormack = use_maccormack ? 1 : 0;
		m_param.weno = weno_interpolation ? 1 : 0;
	}
	//
	virtual void initialize( const shape3 &shape, double dx ) override {
		//
		release();
		m_shape = shape;
		m_dx = dx;
	}
	virtual void post_initialize() override {
		ensure_handle();
	}
	void release() {
		if( m_advect ) shkz_b200_advect_destroy(m_advect);
		m_advect = nullptr;
	}
	virtual ~b200advection3() {
		release();
		for( auto &s : m_slot ) { s.values.release(); s.active.release(); }
	}
	//
	shkz_b200_advect_params m_param;
	shkz_b200_advect *m_advect {nullptr};
	slot m_slot[5]; // velocity x, y, z; the advected scalar; the liquid level set
	shape3 m_shape;
	double m_dx {0.0};
	int m_device {0};
};
//
extern "C" module * create_instance() {
	return new b200advection3();
}
//
extern "C" const char *license() {
	return "MIT";
}
//
