/*
 * host_capi.cpp -- include/pbr_host.h on top of the C++ host mirror.
 */
#include "../../include/pbr_host.h"

#include <fstream>
#include <stdexcept>
#include <stdio.h>
#include <string.h>
#include <functional>
#include <string>
#include <vector>

#include "Cfg.h"
#include "ModelLoader.h"
#include "PathTracer.h"
#include "accelstructures/BVH.h"
#include "qt/GLWidget.h"

using std::string;
using std::vector;

namespace {

thread_local string gLastError;

int failMsg( const string& m ) {
	gLastError = m;
	return 1;
}

#define PBRH_TRY try {
#define PBRH_CATCH } catch( const std::exception& e ) { return failMsg( e.what() ); } catch( ... ) { return failMsg( "unknown error" ); }

}

struct pbrh_scene {
	ModelLoader* ml;
};

struct pbrh_flat {
	vector<bvhNode_cl> nodes;
	vector<cl_uint4> facesV, facesN;
	int64_t info[6];
	double buildSeconds;
};

struct pbrh_renderer {
	GLWidget* widget;
};


extern "C" {

const char* pbrh_last_error( void ) { return gLastError.c_str(); }

/* ---- Cfg ------------------------------------------------------------------------------------- */

int pbrh_config_reset( void ) {
	PBRH_TRY
	Cfg::get().loadDefaults();
	return 0;
	PBRH_CATCH
}

int pbrh_config_load_file( const char* path ) {
	PBRH_TRY
	Cfg::get().loadConfigFile( path );
	return 0;
	PBRH_CATCH
}

int pbrh_config_load_string( const char* json ) {
	PBRH_TRY
	Cfg::get().loadConfigString( json );
	return 0;
	PBRH_CATCH
}

int pbrh_config_set( const char* key, const char* value ) {
	PBRH_TRY
	Cfg::get().value<string>( key, string( value ) );
	return 0;
	PBRH_CATCH
}

int pbrh_config_get( const char* key, char* out, size_t out_len ) {
	PBRH_TRY
	const string v = Cfg::get().value<string>( key );
	if( out && out_len ) {
		strncpy( out, v.c_str(), out_len - 1 );
		out[out_len - 1] = 0;
	}
	return 0;
	PBRH_CATCH
}

/* ---- scene ----------------------------------------------------------------------------------- */

int pbrh_scene_load( const char* filepath, const char* filename, pbrh_scene** out ) {
	PBRH_TRY
	ModelLoader* ml = new ModelLoader();
	ml->loadModel( filepath, filename );
	*out = new pbrh_scene();
	( *out )->ml = ml;
	return 0;
	PBRH_CATCH
}

int pbrh_scene_from_arrays(
	const float* vertices, int64_t numVertexFloats, const float* normals, int64_t numNormalFloats,
	const uint32_t* facesV, int64_t numFaceIdx, const uint32_t* facesVN, int64_t numFaceVNIdx,
	const int32_t* facesMtl, int64_t numFaces,
	const uint32_t* objFaceCounts, const uint32_t* objNormalFaceCounts, int32_t numObjects,
	const float* materials24, int32_t numMaterials, const char* materialNames,
	const float* lights10, int32_t numLights, pbrh_scene** out
) {
	PBRH_TRY
	vector<cl_float> v( vertices, vertices + numVertexFloats );
	vector<cl_float> n( normals, normals + numNormalFloats );
	vector<cl_uint> fv( facesV, facesV + numFaceIdx );
	vector<cl_uint> fvn( facesVN, facesVN + numFaceVNIdx );
	vector<cl_int> fm( facesMtl, facesMtl + numFaces );

	vector<object3D> objects( (size_t) numObjects );
	size_t offV = 0, offN = 0;
	for( int32_t i = 0; i < numObjects; i++ ) {
		char name[32];
		snprintf( name, 32, "object%d", i );
		objects[i].oName = name;
		const size_t cv = (size_t) objFaceCounts[i] * 3, cn = (size_t) objNormalFaceCounts[i] * 3;
		if( offV + cv > fv.size() || offN + cn > fvn.size() ) {
			return failMsg( "pbrh_scene_from_arrays: object face counts exceed the face lists" );
		}
		objects[i].facesV.assign( fv.begin() + (long) offV, fv.begin() + (long) ( offV + cv ) );
		objects[i].facesVN.assign( fvn.begin() + (long) offN, fvn.begin() + (long) ( offN + cn ) );
		offV += cv;
		offN += cn;
	}

	vector<string> names;
	{
		string all = materialNames ? materialNames : "";
		size_t pos = 0;
		while( pos <= all.size() && (int32_t) names.size() < numMaterials ) {
			size_t nl = all.find( '\n', pos );
			if( nl == string::npos ) { nl = all.size(); }
			names.push_back( all.substr( pos, nl - pos ) );
			pos = nl + 1;
		}
		while( (int32_t) names.size() < numMaterials ) { names.push_back( "" ); }
	}
	vector<material_t> materials( (size_t) numMaterials );
	for( int32_t i = 0; i < numMaterials; i++ ) {
		const float* m = materials24 + (size_t) i * 24;
		material_t& t = materials[i];
		t = MtlParser::getEmptyMaterial();
		t.mtlName = names[i];
		t.Ka.x = m[0]; t.Ka.y = m[1]; t.Ka.z = m[2]; t.Ka.w = m[3];
		t.Kd.x = m[4]; t.Kd.y = m[5]; t.Kd.z = m[6]; t.Kd.w = m[7];
		t.Ks.x = m[8]; t.Ks.y = m[9]; t.Ks.z = m[10]; t.Ks.w = m[11];
		t.d = m[12]; t.Ni = m[13]; t.Ns = m[14]; t.illum = (cl_char) m[15]; t.light = (cl_char) m[16];
		t.rough = m[17]; t.p = m[18]; t.nu = m[19]; t.nv = m[20]; t.Rs = m[21]; t.Rd = m[22];
	}
	vector<light_t> lights( (size_t) numLights );
	for( int32_t i = 0; i < numLights; i++ ) {
		const float* l = lights10 + (size_t) i * 10;
		light_t& t = lights[i];
		t = LightParser::getEmptyLight();
		char name[32];
		snprintf( name, 32, "light%d", i );
		t.lightName = name;
		t.type = (cl_uint) l[0];
		t.pos.x = l[1]; t.pos.y = l[2]; t.pos.z = l[3]; t.pos.w = l[4];
		t.rgb.x = l[5]; t.rgb.y = l[6]; t.rgb.z = l[7]; t.rgb.w = l[8];
		t.radius = l[9];
	}

	ModelLoader* ml = new ModelLoader();
	ml->getObjParser()->setScene( v, n, fv, fvn, fm, objects, materials, lights );
	*out = new pbrh_scene();
	( *out )->ml = ml;
	return 0;
	PBRH_CATCH
}

void pbrh_scene_free( pbrh_scene* s ) {
	if( s ) {
		delete s->ml;
		delete s;
	}
}

int64_t pbrh_scene_get( pbrh_scene* s, int32_t what, void* dst ) {
	ObjParser* p = s->ml->getObjParser();
#define COPY_VEC( v ) do { if( dst && !( v ).empty() ) memcpy( dst, ( v ).data(), ( v ).size() * sizeof( ( v )[0] ) ); return (int64_t) ( v ).size(); } while( 0 )
	switch( what ) {
		case 0: COPY_VEC( p->vertices() );
		case 1: COPY_VEC( p->normals() );
		case 2: COPY_VEC( p->facesV() );
		case 3: COPY_VEC( p->facesVN() );
		case 4: COPY_VEC( p->facesMtl() );
		case 12: { vector<cl_uint> t = p->getFacesVT(); COPY_VEC( t ); }
		case 13: { vector<cl_float> t = p->getTextureCoordinates(); COPY_VEC( t ); }
		case 5: case 8: {
			const vector<object3D>& o = p->objects();
			if( dst ) for( size_t i = 0; i < o.size(); i++ ) {
				( (uint32_t*) dst )[i] = (uint32_t) ( ( what == 5 ? o[i].facesV.size() : o[i].facesVN.size() ) / 3 );
			}
			return (int64_t) o.size();
		}
		case 6: case 7: {
			const vector<object3D>& o = p->objects();
			int64_t n = 0;
			for( size_t i = 0; i < o.size(); i++ ) {
				const vector<cl_uint>& v = ( what == 6 ) ? o[i].facesV : o[i].facesVN;
				if( dst && !v.empty() ) memcpy( (uint32_t*) dst + n, v.data(), v.size() * 4 );
				n += (int64_t) v.size();
			}
			return n;
		}
		case 9: {
			vector<material_t> mats = p->getMaterials();
			if( dst ) for( size_t i = 0; i < mats.size(); i++ ) {
				const material_t& m = mats[i];
				float* o = (float*) dst + i * 24;
				o[0] = m.Ka.x; o[1] = m.Ka.y; o[2] = m.Ka.z; o[3] = m.Ka.w;
				o[4] = m.Kd.x; o[5] = m.Kd.y; o[6] = m.Kd.z; o[7] = m.Kd.w;
				o[8] = m.Ks.x; o[9] = m.Ks.y; o[10] = m.Ks.z; o[11] = m.Ks.w;
				o[12] = m.d; o[13] = m.Ni; o[14] = m.Ns; o[15] = (float) m.illum; o[16] = (float) m.light;
				o[17] = m.rough; o[18] = m.p; o[19] = m.nu; o[20] = m.nv; o[21] = m.Rs; o[22] = m.Rd; o[23] = 0.0f;
			}
			return (int64_t) mats.size();
		}
		case 10: {
			vector<light_t> ls = p->getLights();
			if( dst ) for( size_t i = 0; i < ls.size(); i++ ) {
				const light_t& l = ls[i];
				float* o = (float*) dst + i * 10;
				o[0] = (float) l.type;
				o[1] = l.pos.x; o[2] = l.pos.y; o[3] = l.pos.z; o[4] = l.pos.w;
				o[5] = l.rgb.x; o[6] = l.rgb.y; o[7] = l.rgb.z; o[8] = l.rgb.w;
				o[9] = l.radius;
			}
			return (int64_t) ls.size();
		}
	}
#undef COPY_VEC
	return -1;
}

const char* pbrh_scene_name( pbrh_scene* s, int32_t kind, int32_t idx ) {
	static thread_local string name;
	name = "";
	ObjParser* p = s->ml->getObjParser();
	if( kind == 0 && idx >= 0 && (size_t) idx < p->objects().size() ) { name = p->objects()[idx].oName; }
	if( kind == 1 ) { vector<material_t> m = p->getMaterials(); if( idx >= 0 && (size_t) idx < m.size() ) { name = m[idx].mtlName; } }
	if( kind == 2 ) { vector<light_t> l = p->getLights(); if( idx >= 0 && (size_t) idx < l.size() ) { name = l[idx].lightName; } }
	return name.c_str();
}

/* ---- BVH + flatten --------------------------------------------------------------------------- */

int pbrh_flat_build( pbrh_scene* s, pbrh_flat** out ) {
	PBRH_TRY
	ObjParser* op = s->ml->getObjParser();
	BVH bvh( op->objects(), op->vertices(), op->normals() );
	pbrh_flat* f = new pbrh_flat();
	if( bvh.getRoot() == NULL ) {
		delete f;
		return failMsg( "BVH: no objects with faces" );
	}
	PathTracer::flattenBVH( &bvh, op, op->facesV(), &f->nodes, &f->facesV, &f->facesN );
	f->info[0] = (int64_t) bvh.nodes().size();
	f->info[1] = (int64_t) bvh.getLeafNodes().size();
	f->info[2] = (int64_t) bvh.getDepth();
	f->info[3] = (int64_t) bvh.getNumSkipped();
	f->info[4] = (int64_t) f->nodes.size();
	f->info[5] = (int64_t) f->facesV.size();
	f->buildSeconds = bvh.getBuildSeconds();
	*out = f;
	return 0;
	PBRH_CATCH
}

void pbrh_flat_info( pbrh_flat* f, int64_t info[6], double* build_seconds ) {
	for( int i = 0; i < 6; i++ ) { info[i] = f->info[i]; }
	if( build_seconds ) { *build_seconds = f->buildSeconds; }
}

void pbrh_flat_get( pbrh_flat* f, pbr_bvh_node* nodes, pbr_uint4* facesV, pbr_uint4* facesN ) {
	if( nodes ) memcpy( nodes, f->nodes.data(), f->nodes.size() * sizeof( pbr_bvh_node ) );
	if( facesV ) memcpy( facesV, f->facesV.data(), f->facesV.size() * sizeof( pbr_uint4 ) );
	if( facesN ) memcpy( facesN, f->facesN.data(), f->facesN.size() * sizeof( pbr_uint4 ) );
}

void pbrh_flat_free( pbrh_flat* f ) { delete f; }

/* ---- renderer -------------------------------------------------------------------------------- */

void pbrh_set_device( int device ) { CL::setDefaultDevice( device ); }

int pbrh_renderer_create( pbrh_renderer** out ) {
	PBRH_TRY
	pbrh_renderer* r = new pbrh_renderer();
	r->widget = new GLWidget();
	*out = r;
	return 0;
	PBRH_CATCH
}

void pbrh_renderer_destroy( pbrh_renderer* r ) {
	if( r ) {
		delete r->widget;
		delete r;
	}
}

int pbrh_renderer_load_scene( pbrh_renderer* r, pbrh_scene* s ) {
	PBRH_TRY
	ModelLoader* ml = s->ml;
	s->ml = NULL;
	delete s;
	r->widget->loadModel( ml );
	return 0;
	PBRH_CATCH
}

int pbrh_renderer_load_model( pbrh_renderer* r, const char* filepath, const char* filename ) {
	PBRH_TRY
	r->widget->loadModel( string( filepath ), string( filename ) );
	return 0;
	PBRH_CATCH
}

#define NEED_READY if( !r || !r->widget->isReady() ) { return failMsg( "renderer has no model loaded" ); }

int pbrh_renderer_set_deterministic( pbrh_renderer* r, int32_t enabled ) {
	r->widget->getPathTracer()->setDeterministicSeeds( enabled != 0 );
	return 0;
}

int pbrh_renderer_set_seed_schedule( pbrh_renderer* r, uint32_t stride, uint32_t offset ) {
	r->widget->getPathTracer()->setSeedSchedule( stride, offset );
	return 0;
}

int pbrh_renderer_set_tile_stripes( pbrh_renderer* r, int32_t stripe_rows, int32_t world, int32_t rank ) {
	NEED_READY
	r->widget->getPathTracer()->setTileStripes( stripe_rows, world, rank );
	return 0;
}

int pbrh_comm_unique_id( void* id128 ) {
	if( !id128 ) { return failMsg( "pbrh_comm_unique_id: null pointer" ); }
	if( !CL::commUniqueId( id128 ) ) { return failMsg( "pbrh_comm_unique_id: NCCL is not available (libnccl.so.2 could not be loaded)" ); }
	return 0;
}

int pbrh_renderer_set_ranks( pbrh_renderer* r, int32_t rank, int32_t world, const void* id128, int32_t sharding ) {
	NEED_READY
	if( world < 1 || rank < 0 || rank >= world || sharding < 0 || sharding > 2 || ( world > 1 && !id128 ) ) {
		return failMsg( "pbrh_renderer_set_ranks: bad rank / world / sharding" );
	}
	if( !r->widget->getPathTracer()->setRanks( rank, world, id128, sharding ) ) {
		return failMsg( "PathTracer::setRanks failed (see the log)" );
	}
	return 0;
}

int pbrh_renderer_set_sharding( pbrh_renderer* r, int32_t sharding ) {
	NEED_READY
	if( !r->widget->getPathTracer()->setSharding( sharding ) ) { return failMsg( "PathTracer::setSharding failed (see the log)" ); }
	return 0;
}

int pbrh_renderer_comm_fence( pbrh_renderer* r ) {
	NEED_READY
	r->widget->getPathTracer()->commFence();
	return 0;
}

int pbrh_renderer_set_traversal( pbrh_renderer* r, int32_t mode ) {
	NEED_READY
	r->widget->getPathTracer()->setTraversal( mode );
	return 0;
}

int pbrh_renderer_set_render_ahead( pbrh_renderer* r, int32_t depth ) {
	r->widget->getPathTracer()->setRenderAhead( depth );
	return 0;
}

int pbrh_renderer_set_frame_time_ms( pbrh_renderer* r, uint32_t ms ) {
	r->widget->getPathTracer()->setFrameTimeMs( ms );
	return 0;
}

int pbrh_renderer_set_tile( pbrh_renderer* r, int32_t y0, int32_t y1 ) {
	NEED_READY
	r->widget->getPathTracer()->setTileRows( y0, y1 );
	return 0;
}

int pbrh_renderer_generate_image( pbrh_renderer* r, float* out, float* debug ) {
	NEED_READY
	PBRH_TRY
	r->widget->getPathTracer()->generateImageInto( out, debug );
	return 0;
	PBRH_CATCH
}

int pbrh_renderer_render_frames( pbrh_renderer* r, int32_t n ) {
	NEED_READY
	PBRH_TRY
	r->widget->getPathTracer()->renderFrames( (cl_uint) n );
	return 0;
	PBRH_CATCH
}

int pbrh_renderer_read_image( pbrh_renderer* r, float* out, float* debug ) {
	NEED_READY
	r->widget->getPathTracer()->readImage( out, debug );
	return 0;
}

int pbrh_renderer_write_image( pbrh_renderer* r, const float* image, uint32_t sample_count ) {
	NEED_READY
	r->widget->getPathTracer()->writeImage( image, sample_count );
	return 0;
}

int pbrh_renderer_finish( pbrh_renderer* r ) {
	NEED_READY
	r->widget->getPathTracer()->getCL()->finish();
	return 0;
}

int pbrh_renderer_reset_sample_count( pbrh_renderer* r ) {
	r->widget->getPathTracer()->resetSampleCount();
	return 0;
}

int pbrh_renderer_set_focus( pbrh_renderer* r, int32_t x, int32_t y ) {
	r->widget->getPathTracer()->setFocus( x, y );
	return 0;
}

int pbrh_renderer_set_eye( pbrh_renderer* r, float x, float y, float z ) {
	r->widget->getCamera()->setEye( x, y, z );
	return 0;
}

int pbrh_renderer_rotate_camera( pbrh_renderer* r, int32_t move_x, int32_t move_y ) {
	r->widget->getCamera()->updateCameraRot( move_x, move_y );
	return 0;
}

int pbrh_renderer_move_camera( pbrh_renderer* r, int32_t direction ) {
	Camera* c = r->widget->getCamera();
	switch( direction ) {
		case 0: c->cameraMoveForward(); break;
		case 1: c->cameraMoveBackward(); break;
		case 2: c->cameraMoveLeft(); break;
		case 3: c->cameraMoveRight(); break;
		case 4: c->cameraMoveUp(); break;
		case 5: c->cameraMoveDown(); break;
		case 6: c->cameraReset(); break;
		default: return failMsg( "move_camera: direction must be 0..6" );
	}
	return 0;
}

int pbrh_renderer_info( pbrh_renderer* r, int64_t info[8], double* bvh_build_seconds, double* last_kernel_ms ) {
	NEED_READY
	PathTracer* pt = r->widget->getPathTracer();
	info[0] = pt->getWidth();
	info[1] = pt->getHeight();
	info[2] = pt->getSampleCount();
	info[3] = r->widget->getBvhNumNodes();
	info[4] = (int64_t) pt->getFlatNodes().size();
	info[5] = (int64_t) pt->getFlatFacesV().size();
	info[6] = pt->getNumLights();
	info[7] = r->widget->getBvhNumSkipped();
	if( bvh_build_seconds ) { *bvh_build_seconds = r->widget->getBvhBuildSeconds(); }
	if( last_kernel_ms ) { *last_kernel_ms = pt->getLastKernelMs(); }
	return 0;
}

int pbrh_renderer_stats( pbrh_renderer* r, uint64_t out[6], int32_t reset ) {
	NEED_READY
	r->widget->getPathTracer()->getCL()->getStats( out, reset != 0 );
	return 0;
}

int pbrh_renderer_flat_get( pbrh_renderer* r, pbr_bvh_node* nodes, pbr_uint4* facesV, pbr_uint4* facesN ) {
	NEED_READY
	PathTracer* pt = r->widget->getPathTracer();
	if( nodes ) memcpy( nodes, pt->getFlatNodes().data(), pt->getFlatNodes().size() * sizeof( pbr_bvh_node ) );
	if( facesV ) memcpy( facesV, pt->getFlatFacesV().data(), pt->getFlatFacesV().size() * sizeof( pbr_uint4 ) );
	if( facesN ) memcpy( facesN, pt->getFlatFacesN().data(), pt->getFlatFacesN().size() * sizeof( pbr_uint4 ) );
	return 0;
}

int pbrh_renderer_camera( pbrh_renderer* r, pbr_camera* cam, float* px_dim ) {
	NEED_READY
	PathTracer* pt = r->widget->getPathTracer();
	if( cam ) { *cam = pt->getCameraStruct(); }
	if( px_dim ) { *px_dim = pt->getPxDim(); }
	return 0;
}

int pbrh_renderer_trace( pbrh_renderer* r, const pbr_ray* rays, int64_t n, int32_t any_hit, pbr_hit* hits ) {
	NEED_READY
	PathTracer* pt = r->widget->getPathTracer();
	pbr_ctx* ctx = pt->getCL()->getContext();
	const int err = pbr_trace(
		ctx, pt->getBufBVH(), pt->getBufFacesV(), pt->getBufVertices(), pt->getBufLights(),
		(int32_t) pt->getNumLights(), rays, n, any_hit, hits
	);
	if( err != PBR_OK ) { return failMsg( pbr_last_error( ctx ) ); }
	return 0;
}

int pbrh_renderer_handles( pbrh_renderer* r, void** pbr_ctx_out, uint64_t handles[6] ) {
	NEED_READY
	PathTracer* pt = r->widget->getPathTracer();
	if( pbr_ctx_out ) { *pbr_ctx_out = pt->getCL()->getContext(); }
	handles[0] = pt->getBufBVH();
	handles[1] = pt->getBufFacesV();
	handles[2] = pt->getBufVertices();
	handles[3] = pt->getBufLights();
	handles[4] = pt->getImageHandle();
	handles[5] = 1;
	return 0;
}

/* ---- image files ----------------------------------------------------------------------------- */

/** Portable float map, RGB, little endian, bottom row first -- the kernel's row order. */
/* Files are written next to their final name and renamed into place once every byte has reached the file: a full disk
 * or an I/O error leaves the previous file (the checkpoint a long render resumes from) untouched and is reported. */
static int writeAtomically( const char* path, const std::function<bool( FILE* )>& body ) {
	const string tmp = string( path ) + ".tmp";
	FILE* f = fopen( tmp.c_str(), "wb" );
	if( !f ) { return failMsg( string( "cannot write " ) + tmp ); }
	bool ok = body( f );
	ok = ( fflush( f ) == 0 ) && ok;
	ok = ( fclose( f ) == 0 ) && ok;
	if( !ok ) {
		remove( tmp.c_str() );
		return failMsg( string( "write error on " ) + tmp + " (disk full?)" );
	}
	if( rename( tmp.c_str(), path ) != 0 ) {
		remove( tmp.c_str() );
		return failMsg( string( "cannot rename " ) + tmp + " to " + path );
	}
	return 0;
}

int pbrh_write_pfm( const char* path, const float* rgba, int32_t width, int32_t height ) {
	return writeAtomically( path, [&]( FILE* f ) {
		if( fprintf( f, "PF\n%d %d\n-1.0\n", width, height ) < 0 ) { return false; }
		vector<float> row( (size_t) width * 3 );
		for( int32_t y = 0; y < height; y++ ) {
			for( int32_t x = 0; x < width; x++ ) {
				const float* p = rgba + ( (size_t) y * width + x ) * 4;
				row[3 * x] = p[0]; row[3 * x + 1] = p[1]; row[3 * x + 2] = p[2];
			}
			if( fwrite( row.data(), sizeof( float ), row.size(), f ) != row.size() ) { return false; }
		}
		return true;
	} );
}

/** Accumulation checkpoint: "PBRACC1\n" width height sampleCount, then W*H*4 floats. */
int pbrh_write_checkpoint( const char* path, const float* rgba, int32_t width, int32_t height, uint32_t sample_count ) {
	return writeAtomically( path, [&]( FILE* f ) {
		const size_t n = (size_t) width * height * 4;
		return fprintf( f, "PBRACC1\n%d %d %u\n", width, height, sample_count ) >= 0 && fwrite( rgba, sizeof( float ), n, f ) == n;
	} );
}

int pbrh_read_checkpoint( const char* path, float* rgba, int32_t width, int32_t height, uint32_t* sample_count ) {
	FILE* f = fopen( path, "rb" );
	if( !f ) { return failMsg( string( "cannot read " ) + path ); }
	int w = 0, h = 0;
	unsigned sc = 0;
	char magic[16] = { 0 };
	if( fscanf( f, "%15s %d %d %u", magic, &w, &h, &sc ) != 4 || strcmp( magic, "PBRACC1" ) != 0 || w != width || h != height ) {
		fclose( f );
		return failMsg( "checkpoint header mismatch" );
	}
	fgetc( f );
	const size_t n = (size_t) width * height * 4;
	const size_t got = fread( rgba, sizeof( float ), n, f );
	fclose( f );
	if( got != n ) { return failMsg( "checkpoint truncated" ); }
	if( sample_count ) { *sample_count = sc; }
	return 0;
}

} /* extern "C" */
