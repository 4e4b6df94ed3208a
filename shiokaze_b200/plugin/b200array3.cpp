/*
 * b200array3 — a dense, page-locked array core for Shiokaze (`Array=b200array3`): SURVEY.md section 8(f) rank 1.
 *
 * array3 / macarray3 keep their cells in an `array_core3` module loaded by name (include/shiokaze/array/array3.h:91-102 ->
 * array_core3::quick_load_module -> module::alloc_module, whose name the command line overrides: src/core/module.cpp:103-111), so
 * choosing this core is a run-time flag, like choosing the projection module. The interface is include/shiokaze/array/array_core3.h:24-140;
 * the model for its semantics is the reference's own dense core, src/array/lineararray3.cpp.
 *
 * Why: the projection module's bridge between the host's grids and the C-ABI (gather into dense buffers, scatter back) goes through one
 * std::function call per cell on the stock cores and was 90 % of a project() through the module. This core stores
 *     values   nx*ny*nz elements, x fastest — the layout the C-ABI takes (include/shkz_b200.h) —, in PAGE-LOCKED memory (shkz_b200_host_alloc)
 *     active   one byte per cell — the activity-mask format the C-ABI takes
 *     filled   one byte per cell (flood fill), allocated on first use
 * and answers the message "b200:dense" (array3::send_message forwards to the core, array3.h:125-150) with a descriptor of those buffers:
 * b200pressure3 then hands the pointers straight to shkz_b200_project_host — the host-to-device copies read the grids where they live and the
 * device-to-host copies write the results in place. Without a CUDA device the core falls back to ordinary aligned memory (it holds data, it does
 * not compute): the stock CPU modules run on it unchanged, which is how tests/ pin its semantics against tiledarray3 here.
 */
#include <shiokaze/array/array_core3.h>
#include <shiokaze/math/shape.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stack>
#include <vector>
//
#include "../../include/shkz_b200.h"
#include "b200dense.h"
//
SHKZ_BEGIN_NAMESPACE
//
class b200array3 : public array_core3 {
public:
	b200array3 () = default;
protected:
	//
	LONG_NAME("B200 Dense Array 3D")
	ARGUMENT_NAME("B200Array")
	//
	virtual void initialize( unsigned nx, unsigned ny, unsigned nz, unsigned element_bytes ) override {
		dealloc();
		m_nx = nx; m_ny = ny; m_nz = nz;
		m_element_bytes = element_bytes;
		m_count = (size_t)nx*ny*nz;
		if( element_bytes ) m_buffer = static_cast<unsigned char *>(alloc(m_count*element_bytes,m_buffer_pinned));
		m_active = static_cast<uint8_t *>(alloc(m_count,m_active_pinned));
		std::memset(m_active,0,m_count);
	}
	virtual void get( unsigned &nx, unsigned &ny, unsigned &nz, unsigned &element_bytes ) const override {
		nx = m_nx; ny = m_ny; nz = m_nz; element_bytes = m_element_bytes;
	}
	virtual ~b200array3() { dealloc(); }
	//
	// the zero-copy hook: include/b200dense.h documents the descriptor
	virtual bool const_send_message( std::string message, void *ptr ) const override {
		if( message == B200_DENSE_MESSAGE && ptr ) {
			b200_dense_descriptor *d = static_cast<b200_dense_descriptor *>(ptr);
			d->nx = m_nx; d->ny = m_ny; d->nz = m_nz; d->element_bytes = m_element_bytes;
			d->values = m_buffer; d->active = m_active; d->filled = m_filled;
			d->pinned = (m_buffer_pinned && m_active_pinned) ? 1 : 0;
			return true;
		}
		return false;
	}
	virtual bool send_message( std::string message, void *ptr ) override { return const_send_message(message,ptr); }
	//
	virtual size_t count( const parallel_driver &parallel ) const override {
		std::vector<size_t> partial (parallel.get_thread_num(),0);
		parallel.for_each(m_nz,[&]( size_t k, int tn ) {
			const uint8_t *a = m_active+k*(size_t)m_nx*m_ny;
			size_t sum (0);
			for( size_t n=0; n<(size_t)m_nx*m_ny; ++n ) sum += a[n] ? 1 : 0;
			partial[tn] += sum;
		});
		size_t total (0);
		for( size_t v : partial ) total += v;
		return total;
	}
	//
	virtual void copy( const array_core3 &array, std::function<void(void *target, const void *src)> copy_func, const parallel_driver &parallel ) override {
		unsigned nx, ny, nz, element_bytes;
		array.get(nx,ny,nz,element_bytes);
		initialize(nx,ny,nz,element_bytes);
		auto mate = dynamic_cast<const b200array3 *>(&array);
		if( mate ) {
			std::memcpy(m_active,mate->m_active,m_count);
			if( mate->m_filled && m_element_bytes ) {
				ensure_filled();
				std::memcpy(m_filled,mate->m_filled,m_count);
			}
			if( m_buffer ) parallel.for_each(m_nz,[&]( size_t k ) {
				const size_t plane = (size_t)m_nx*m_ny;
				for( size_t n=k*plane; n<(k+1)*plane; ++n ) if( m_active[n] ) copy_func(m_buffer+n*m_element_bytes,mate->m_buffer+n*m_element_bytes);
			});
		} else {
			array.const_serial_actives([&](int i, int j, int k, const void *value_ptr, const bool &filled ) {
				const size_t n = encode(i,j,k);
				m_active[n] = 1;
				copy_func(m_buffer ? m_buffer+n*m_element_bytes : nullptr,value_ptr);
				return false;
			});
			if( m_element_bytes ) {
				array.const_serial_inside([&](int i, int j, int k, const void *value_ptr, const bool &active ) {
					if( ! active ) {
						ensure_filled();
						m_filled[encode(i,j,k)] = 1;
					}
					return false;
				});
			}
		}
	}
	//
	virtual void set( int i, int j, int k, std::function<void(void *value_ptr, bool &active)> func ) override {
		const size_t n = encode(i,j,k);
		bool active = m_active[n] != 0;
		func(m_buffer ? m_buffer+n*m_element_bytes : nullptr,active);
		m_active[n] = active ? 1 : 0;
	}
	virtual const void * operator()( int i, int j, int k, bool &filled ) const override {
		const size_t n = encode(i,j,k);
		filled = m_filled ? m_filled[n] != 0 : false;
		static char tmp_ptr;
		if( m_active[n] ) return m_buffer ? m_buffer+n*m_element_bytes : (void *)&tmp_ptr;
		return nullptr;
	}
	//
	// ---- loops: one z-plane per task (rows of a plane are contiguous; a byte per cell, so no two tasks share a mask word)
	template <class F> void loop_planes( const parallel_driver *parallel, F body ) const {
		if( parallel ) parallel->for_each(m_nz,[&]( size_t k, int tn ) { body((int)k,tn); });
		else for( unsigned k=0; k<m_nz; ++k ) if( body((int)k,0)) break;
	}
	virtual void parallel_actives ( std::function<void(int i, int j, int k, void *value_ptr, bool &active, const bool &filled, int thread_index )> func, const parallel_driver &parallel ) override {
		loop_planes(&parallel,[&]( int k, int tn ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_active[n] ) {
				bool active (true);
				func(i,j,k,ptr(n),active,is_filled(n),tn);
				if( ! active ) m_active[n] = 0;
			}
			return false;
		});
	}
	virtual void serial_actives ( std::function<bool(int i, int j, int k, void *value_ptr, bool &active, const bool &filled )> func ) override {
		loop_planes(nullptr,[&]( int k, int ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_active[n] ) {
				bool active (true);
				const bool stop = func(i,j,k,ptr(n),active,is_filled(n));
				if( stop ) return true;
				if( ! active ) m_active[n] = 0;
			}
			return false;
		});
	}
	virtual void const_parallel_actives ( std::function<void(int i, int j, int k, const void *value_ptr, const bool &filled, int thread_index )> func, const parallel_driver &parallel ) const override {
		loop_planes(&parallel,[&]( int k, int tn ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_active[n] ) func(i,j,k,ptr(n),is_filled(n),tn);
			return false;
		});
	}
	virtual void const_serial_actives ( std::function<bool(int i, int j, int k, const void *value_ptr, const bool &filled )> func ) const override {
		loop_planes(nullptr,[&]( int k, int ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_active[n] ) if( func(i,j,k,ptr(n),is_filled(n))) return true;
			return false;
		});
	}
	virtual void parallel_all ( std::function<void(int i, int j, int k, void *value_ptr, bool &active, const bool &filled, int thread_index )> func, const parallel_driver &parallel ) override {
		loop_planes(&parallel,[&]( int k, int tn ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) {
				bool active = m_active[n] != 0;
				func(i,j,k,ptr(n),active,is_filled(n),tn);
				m_active[n] = active ? 1 : 0;
			}
			return false;
		});
	}
	virtual void serial_all ( std::function<bool(int i, int j, int k, void *value_ptr, bool &active, const bool &filled )> func ) override {
		loop_planes(nullptr,[&]( int k, int ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) {
				bool active = m_active[n] != 0;
				const bool stop = func(i,j,k,ptr(n),active,is_filled(n));
				m_active[n] = active ? 1 : 0;
				if( stop ) return true;
			}
			return false;
		});
	}
	virtual void const_parallel_all ( std::function<void(int i, int j, int k, const void *value_ptr, const bool &active, const bool &filled, int thread_index )> func, const parallel_driver &parallel ) const override {
		loop_planes(&parallel,[&]( int k, int tn ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) func(i,j,k,ptr(n),m_active[n] != 0,is_filled(n),tn);
			return false;
		});
	}
	virtual void const_serial_all ( std::function<bool(int i, int j, int k, const void *value_ptr, const bool &active, const bool &filled )> func ) const override {
		loop_planes(nullptr,[&]( int k, int ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( func(i,j,k,ptr(n),m_active[n] != 0,is_filled(n))) return true;
			return false;
		});
	}
	virtual void const_parallel_inside ( std::function<void(int i, int j, int k, const void *value_ptr, const bool &active, int thread_index )> func, const parallel_driver &parallel ) const override {
		if( ! m_filled ) return;
		loop_planes(&parallel,[&]( int k, int tn ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_filled[n] ) func(i,j,k,ptr(n),m_active[n] != 0,tn);
			return false;
		});
	}
	virtual void const_serial_inside ( std::function<bool(int i, int j, int k, const void *value_ptr, const bool &active )> func ) const override {
		if( ! m_filled ) return;
		loop_planes(nullptr,[&]( int k, int ) {
			size_t n = (size_t)k*m_nx*m_ny;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( m_filled[n] ) if( func(i,j,k,ptr(n),m_active[n] != 0)) return true;
			return false;
		});
	}
	//
	// One layer of inactive cells around the active set is offered to `func` (on a zeroed scratch value, inactive), ascending in the linear index;
	// those it activates are set afterwards — the contract of src/array/dilate3.h, which every stock core uses.
	virtual void dilate( std::function<void(int i, int j, int k, void *value_ptr, bool &active, const bool &filled, int thread_index)> func, const parallel_driver &parallel ) override {
		const size_t plane = (size_t)m_nx*m_ny;
		std::vector<std::vector<size_t> > found (m_nz);
		parallel.for_each(m_nz,[&]( size_t k ) {
			// candidates of plane k: inactive cells with an active face neighbour
			size_t n = k*plane;
			for( int j=0; j<(int)m_ny; ++j ) for( int i=0; i<(int)m_nx; ++i, ++n ) if( ! m_active[n] ) {
				const bool near = (i > 0 && m_active[n-1]) || (i+1 < (int)m_nx && m_active[n+1]) || (j > 0 && m_active[n-m_nx]) || (j+1 < (int)m_ny && m_active[n+m_nx])
					|| (k > 0 && m_active[n-plane]) || (k+1 < m_nz && m_active[n+plane]);
				if( near ) found[k].push_back(n);
			}
		});
		std::vector<size_t> cells;
		for( const auto &f : found ) cells.insert(cells.end(),f.begin(),f.end());
		std::vector<unsigned char> scratch (cells.size()*(size_t)m_element_bytes,0);
		std::vector<uint8_t> accepted (cells.size(),0);
		parallel.for_each(cells.size(),[&]( size_t q, int tn ) {
			int i, j, k; decode(cells[q],i,j,k);
			bool active (false);
			func(i,j,k,m_element_bytes ? scratch.data()+q*m_element_bytes : nullptr,active,is_filled(cells[q]),tn);
			accepted[q] = active ? 1 : 0;
		});
		for( size_t q=0; q<cells.size(); ++q ) if( accepted[q] ) {
			m_active[cells[q]] = 1;
			if( m_element_bytes ) std::memcpy(m_buffer+cells[q]*m_element_bytes,scratch.data()+q*m_element_bytes,m_element_bytes);
		}
	}
	//
	// Flood fill from every active cell that `inside_func` accepts, through inactive cells and accepted active ones (src/array/lineararray3.cpp:206-257)
	virtual void flood_fill( std::function<bool(void *value_ptr)> inside_func, const parallel_driver &parallel ) override {
		if( ! m_element_bytes ) return;
		ensure_filled();
		std::memset(m_filled,0,m_count);
		const size_t plane = (size_t)m_nx*m_ny;
		auto markable = [&]( size_t n, bool default_result ) {
			if( m_filled[n] ) return false;
			return m_active[n] ? inside_func(m_buffer+n*m_element_bytes) : default_result;
		};
		std::stack<size_t> queue;
		for( size_t n0=0; n0<m_count; ++n0 ) if( m_active[n0] && markable(n0,false)) {
			queue.push(n0);
			while( ! queue.empty()) {
				const size_t n = queue.top();
				queue.pop();
				m_filled[n] = 1;
				int i, j, k; decode(n,i,j,k);
				if( i > 0 && markable(n-1,true)) queue.push(n-1);
				if( i+1 < (int)m_nx && markable(n+1,true)) queue.push(n+1);
				if( j > 0 && markable(n-m_nx,true)) queue.push(n-m_nx);
				if( j+1 < (int)m_ny && markable(n+m_nx,true)) queue.push(n+m_nx);
				if( k > 0 && markable(n-plane,true)) queue.push(n-plane);
				if( k+1 < (int)m_nz && markable(n+plane,true)) queue.push(n+plane);
			}
		}
	}
	//
private:
	//
	unsigned char *m_buffer {nullptr};
	uint8_t *m_active {nullptr}, *m_filled {nullptr};
	bool m_buffer_pinned {false}, m_active_pinned {false}, m_filled_pinned {false};
	unsigned m_nx {0}, m_ny {0}, m_nz {0}, m_element_bytes {0};
	size_t m_count {0};
	//
	size_t encode( int i, int j, int k ) const { return i + (size_t)m_nx*(j + (size_t)m_ny*k); }
	void decode( size_t n, int &i, int &j, int &k ) const {
		const size_t plane = (size_t)m_nx*m_ny;
		k = (int)(n/plane); j = (int)((n%plane)/m_nx); i = (int)((n%plane)%m_nx);
	}
	void * ptr( size_t n ) const { return m_buffer ? m_buffer+n*m_element_bytes : nullptr; }
	bool is_filled( size_t n ) const { return m_filled ? m_filled[n] != 0 : false; }
	void ensure_filled() {
		if( ! m_filled ) {
			m_filled = static_cast<uint8_t *>(alloc(m_count,m_filled_pinned));
			std::memset(m_filled,0,m_count);
		}
	}
	// page-locked when a CUDA device is there (the copies to and from the device then run at PCIe speed and need no staging), ordinary memory otherwise
	static void * alloc( size_t bytes, bool &pinned ) {
		void *p (nullptr);
		static const bool have_device = shkz_b200_device_count() > 0;
		if( have_device && shkz_b200_host_alloc(bytes ? bytes : 1,&p) == SHKZ_B200_OK ) { pinned = true; return p; }
		pinned = false;
		if( posix_memalign(&p,256,bytes ? bytes : 1)) { fprintf(stderr,"b200array3: out of memory (%zu bytes)\n",bytes); exit(-1); }
		return p;
	}
	static void release( void *p, bool pinned ) {
		if( ! p ) return;
		if( pinned ) shkz_b200_host_free(p); else free(p);
	}
	void dealloc() {
		release(m_buffer,m_buffer_pinned); release(m_active,m_active_pinned); release(m_filled,m_filled_pinned);
		m_buffer = nullptr; m_active = m_filled = nullptr;
		m_count = 0;
	}
};
//
extern "C" module * create_instance() {
	return new b200array3();
}
//
extern "C" const char *license() {
	return "MIT";
}
//
SHKZ_END_NAMESPACE
//
