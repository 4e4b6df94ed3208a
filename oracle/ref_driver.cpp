/*
 * TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * Headless host for ONE macproject3 module call. It links against the *unmodified* reference
 * libraries that oracle/Makefile builds from /root/reference into oracle/_ref/, fills the
 * reference's own sparse grids (array3 / macarray3) from a dense scene file, calls
 * macproject3_interface::project() on whichever module `Projection=<name>` selects
 * (default: the reference's macpressuresolver3) and dumps dense results.
 *
 * It replaces src/ui/ui.cpp (which needs boost posix_time + GL) with the same
 * load -> configure -> initialize sequence (ui.cpp:632-702) and mirrors what the simulators do
 * before their first project() call:
 *   - level sets: activate |phi| < band, set_as_levelset(band), flood_fill()
 *       (src/utility/macutility3.cpp:336-374)
 *   - smoke: fluid = constant -1 background, no actives (src/smoke/macsmoke3.cpp:106)
 *   - accuracy test: solid = cell-shaped constant +1 (src/examples/accuracytest3-example.cpp:90)
 *
 * File formats are documented in oracle/refio.py (the only reader/writer on the Python side).
 */
#include <shiokaze/core/cmdparser.h>
#include <shiokaze/core/console.h>
#include <shiokaze/array/array3.h>
#include <shiokaze/array/macarray3.h>
#include <shiokaze/array/shared_array_core3.h>
#include <shiokaze/projection/macproject3_interface.h>
#include <shiokaze/advection/macadvection3_interface.h>
#include <shiokaze/array/shared_array3.h>
#include <shiokaze/utility/macutility3_interface.h>
#include <shiokaze/utility/gridutility3_interface.h>
#include <shiokaze/utility/utility.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
//
SHKZ_USING_NAMESPACE
//
namespace {
//
struct scene_header {
	char magic[8];
	int32_t nx, ny, nz;
	int32_t solid_mode; // 0: nodal, nothing active   1: nodal narrow band from raw values   2: cell-shaped constant +1
	int32_t fluid_mode; // 0: constant -1 background (all fluid)   1: narrow band from raw values
	int32_t repeat;     // number of project() calls on the same inputs (timing); results are from the last
	int32_t pad0, pad1;
	double dx, dt, surface_tension, band, current_volume, target_volume;
};
//
struct host : public recursive_configurable {
	array3<Real> fluid{this};
	array3<Real> solid{this};
	macarray3<Real> velocity{this};
	macproject3_driver proj{this,"macpressuresolver3"};
	macutility3_driver util{this,"macutility3"};
	macadvection3_driver adv{this,"macadvection3"};
	gridutility3_driver grid{this,"gridutility3"};
	shape3 shape;
	double dx;
	host( const scene_header &h ) {
		shape = shape3(h.nx,h.ny,h.nz);
		dx = h.dx;
		set_environment("shape",&shape);
		set_environment("dx",&dx);
	}
};
//
template <class T> std::vector<T> read_block( FILE *fp, size_t count ) {
	std::vector<T> buf(count);
	if( count && std::fread(buf.data(),sizeof(T),count,fp) != count ) {
		std::fprintf(stderr,"ref_driver: short read\n");
		std::exit(2);
	}
	return buf;
}
template <class T> void write_block( FILE *fp, const std::vector<T> &buf ) {
	if( buf.size() && std::fwrite(buf.data(),sizeof(T),buf.size(),fp) != buf.size() ) {
		std::fprintf(stderr,"ref_driver: short write\n");
		std::exit(2);
	}
}
// Dense read exactly as a bridge would do it: array3::operator() = active value / fill value / background
// (include/shiokaze/array/array3.h:796-801).
void dump_dense( FILE *fp, const array3<Real> &a, bool with_active ) {
	const shape3 s = a.shape();
	std::vector<double> values(s.count());
	std::vector<uint8_t> active(with_active ? s.count() : 0);
	size_t n (0);
	for( int k=0; k<(int)s.d; ++k ) for( int j=0; j<(int)s.h; ++j ) for( int i=0; i<(int)s.w; ++i, ++n ) {
		values[n] = a(i,j,k);
		if( with_active ) active[n] = a.active(i,j,k) ? 1 : 0;
	}
	int32_t dims[4] = {(int32_t)s.w,(int32_t)s.h,(int32_t)s.d,with_active ? 1 : 0};
	std::fwrite(dims,sizeof(int32_t),4,fp);
	write_block(fp,values);
	write_block(fp,active);
}
//
} // namespace
//
int main( int argc, const char *argv[] ) {
	//
	std::string in_path, out_path;
	bool dump_fractions (false), skip_project (false);
	int extrapolate_constrain (-1);
	std::string advect_mode;
	int volume_repeat (0);
	for( int i=1; i<argc; ++i ) {
		if( ! std::strncmp(argv[i],"in=",3)) in_path = argv[i]+3;
		if( ! std::strncmp(argv[i],"out=",4)) out_path = argv[i]+4;
		if( ! std::strcmp(argv[i],"DumpFractions=1")) dump_fractions = true;
		// RecordDir=<existing directory>: console::write records (src/core/console.cpp:259-281) land in <dir>/record/<name>.out, as under ./run with Log=
		if( ! std::strncmp(argv[i],"RecordDir=",10)) console::set_root_path(argv[i]+10);
		// RefExtrapolate=<width>: after the projection, the step the simulators run next (macliquid3.cpp:309-319) through the REFERENCE's macutility3:
		// extrapolate_and_constrain_velocity(solid,velocity,width). RefSkipProject=1: that step alone, on the input velocity.
		if( ! std::strncmp(argv[i],"RefExtrapolate=",15)) extrapolate_constrain = std::atoi(argv[i]+15);
		if( ! std::strcmp(argv[i],"RefSkipProject=1")) skip_project = true;
		// RefAdvect=vector|density|levelset: INSTEAD of the projection, one call of the step the simulators run right before it, through whichever module
		// `Advection=<name>` selects (default: the reference's macadvection3), with the scene's dt:
		//   vector    advect_vector(velocity, copy of velocity, fluid, dt)                      src/liquid/macliquid3.cpp:345-346, src/smoke/macsmoke3.cpp:280-281
		//   density   advect_scalar(q, velocity, fluid, dt), q = a sparse cell grid (background 0) built from the scene: active where the y-face of the same
		//             index is, with that face's value                                          src/smoke/macsmoke3.cpp:274
		//   levelset  advect_scalar(fluid, velocity, copy of fluid, dt)                         src/surfacetracker/maclevelsetsurfacetracker3.cpp:49-51
		// The advected scalar is dumped in the result's pressure slot.
		if( ! std::strncmp(argv[i],"RefAdvect=",10)) advect_mode = argv[i]+10;
		// RefVolume=<n>: print gridutility3::get_volume(solid,fluid) (src/utility/gridutility3.cpp:318-346) n times — the value macliquid3 feeds
		// set_target_volume with every step (DESIGN.md section 9 on why it cannot carry a parity bar)
		if( ! std::strncmp(argv[i],"RefVolume=",10)) volume_repeat = std::atoi(argv[i]+10);
	}
	if( in_path.empty() || out_path.empty()) {
		std::fprintf(stderr,"usage: ref_driver in=<scene> out=<result> [DumpFractions=1] [RecordDir=<dir>] [Projection=<module>] [flag=value ...]\n");
		return 2;
	}
	FILE *fp = std::fopen(in_path.c_str(),"rb");
	if( ! fp ) { std::perror(in_path.c_str()); return 2; }
	scene_header h;
	if( std::fread(&h,sizeof(h),1,fp) != 1 || std::memcmp(h.magic,"SHKZIN02",8)) {
		std::fprintf(stderr,"ref_driver: bad scene header\n");
		return 2;
	}
	//
	cmdparser parser(argc,argv);
	configuration &config = configurable::set_global_configuration(parser);
	config.push_group("Root","Root");
	host H(h);
	H.setup_now(config);
	//
	const shape3 shape = H.shape;
	const int repeat = h.repeat > 0 ? h.repeat : 1;
	//
	std::vector<float> vel[DIM3];
	std::vector<uint8_t> vel_active[DIM3];
	for( int dim : DIMS3 ) vel[dim] = read_block<float>(fp,shape.face(dim).count());
	for( int dim : DIMS3 ) vel_active[dim] = read_block<uint8_t>(fp,shape.face(dim).count());
	//
	if( h.solid_mode == 2 ) {
		H.solid.initialize(shape,1.0);
	} else {
		H.solid.initialize(shape.nodal());
		if( h.solid_mode == 1 ) {
			std::vector<float> raw = read_block<float>(fp,shape.nodal().count());
			const shape3 ns = shape.nodal();
			H.solid.parallel_all([&]( int i, int j, int k, auto &it ) {
				double value = raw[ns.encode(i,j,k)];
				if( std::abs(value) < h.band ) it.set(value);
			});
		}
		H.solid.set_as_levelset(h.band);
		H.solid.flood_fill();
	}
	if( h.fluid_mode == 0 ) {
		H.fluid.initialize(shape,-1.0);
	} else {
		std::vector<float> raw = read_block<float>(fp,shape.count());
		H.fluid.initialize(shape);
		H.fluid.set_as_levelset(h.band);
		H.fluid.parallel_all([&]( int i, int j, int k, auto &it ) {
			double value = raw[shape.encode(i,j,k)];
			if( std::abs(value) < h.band ) it.set(value);
			else it.set_off();
		});
		H.fluid.flood_fill();
	}
	std::fclose(fp);
	//
	for( int n=0; n<volume_repeat; ++n ) std::printf("REFDRIVER volume=%.17g\n",H.grid->get_volume(H.solid,H.fluid));
	double ms_last (0.0), ms_sum (0.0);
	array3<Real> scalar; // RefAdvect=density
	for( int rep=0; rep<repeat; ++rep ) {
		H.velocity.initialize(shape);
		for( int dim : DIMS3 ) {
			const shape3 fs = shape.face(dim);
			const std::vector<float> &src = vel[dim];
			const std::vector<uint8_t> &act = vel_active[dim];
			H.velocity[dim].parallel_all([&]( int i, int j, int k, auto &it ) {
				size_t n = fs.encode(i,j,k);
				if( act[n] ) it.set(src[n]);
			});
		}
		if( h.target_volume ) H.proj->set_target_volume(h.current_volume,h.target_volume);
		double t0 = utility::get_milliseconds();
		if( ! advect_mode.empty()) {
			if( advect_mode == "vector" ) {
				shared_macarray3<Real> velocity_save(H.velocity);
				t0 = utility::get_milliseconds();
				H.adv->advect_vector(H.velocity,velocity_save(),H.fluid,h.dt,"velocity");
			} else if( advect_mode == "density" ) {
				scalar.initialize(shape);
				const shape3 fs = shape.face(1);
				scalar.parallel_all([&]( int i, int j, int k, auto &it ) {
					size_t n = fs.encode(i,j,k);
					if( vel_active[1][n] ) it.set(vel[1][n]);
				});
				t0 = utility::get_milliseconds();
				H.adv->advect_scalar(scalar,H.velocity,H.fluid,h.dt,"density");
			} else if( advect_mode == "levelset" ) {
				shared_array3<Real> fluid_save(H.fluid);
				t0 = utility::get_milliseconds();
				H.adv->advect_scalar(H.fluid,H.velocity,fluid_save(),h.dt,"levelset");
			} else {
				std::fprintf(stderr,"ref_driver: RefAdvect=%s ?\n",advect_mode.c_str());
				return 2;
			}
		} else if( ! skip_project ) H.proj->project(h.dt,H.velocity,H.solid,H.fluid,h.surface_tension);
		ms_last = utility::get_milliseconds()-t0;
		if( extrapolate_constrain >= 0 ) H.util->extrapolate_and_constrain_velocity(H.solid,H.velocity,extrapolate_constrain);
		ms_sum += ms_last;
	}
	std::printf("REFDRIVER project_ms_last=%.3f project_ms_mean=%.3f repeat=%d sizeof_Real=%d\n",ms_last,ms_sum/repeat,repeat,(int)sizeof(Real));
	//
	FILE *out = std::fopen(out_path.c_str(),"wb");
	if( ! out ) { std::perror(out_path.c_str()); return 2; }
	const char magic[8] = {'S','H','K','Z','O','U','T','2'};
	std::fwrite(magic,1,8,out);
	int32_t info[4] = {h.nx,h.ny,h.nz,(int32_t)sizeof(Real)};
	std::fwrite(info,sizeof(int32_t),4,out);
	double ms[2] = {ms_last,ms_sum/repeat};
	std::fwrite(ms,sizeof(double),2,out);
	for( int dim : DIMS3 ) dump_dense(out,H.velocity[dim],true);
	if( advect_mode == "density" ) dump_dense(out,scalar,true);
	else if( advect_mode == "levelset" ) dump_dense(out,H.fluid,true);
	else dump_dense(out,*H.proj->get_pressure(),true);
	dump_dense(out,H.fluid,true);
	dump_dense(out,H.solid,true);
	int32_t has_fractions = dump_fractions ? 1 : 0;
	std::fwrite(&has_fractions,sizeof(int32_t),1,out);
	if( dump_fractions ) {
		// The two macutility3 methods on the path, called directly (src/utility/macutility3.cpp:94-194).
		macarray3<Real> areas(shape), rhos(shape);
		H.util->compute_area_fraction(H.solid,areas);
		H.util->compute_fluid_fraction(H.fluid,rhos);
		for( int dim : DIMS3 ) dump_dense(out,areas[dim],false);
		for( int dim : DIMS3 ) dump_dense(out,rhos[dim],false);
	}
	std::fclose(out);
	shared_array_core3::clear();
	return 0;
}
