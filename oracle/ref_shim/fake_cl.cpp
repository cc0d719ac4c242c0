/*
 * fake_cl.cpp -- the reference's `class CL` (declared in /root/reference/source/CL.h) implemented on the host,
 * so that the reference's own PathTracer.cpp can run unmodified in the tests.  TEST INFRASTRUCTURE.
 *
 * The reference's CL.cpp talks to an OpenCL platform; none exists here.  This file implements the same
 * member functions with buffers and images in host memory, kernel arguments recorded per slot, and
 * execute() forwarding to the reference kernel built for the host by oracle/build_ref.py: the program text
 * values CL::setValues would splice in are exposed through fakecl_program_values(), the test builds the
 * matching oracle/_ref/pt_ref_<key>.so and hands its path back with fakecl_set_kernel_library().
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "CL.h"

struct _cl_mem {
	std::vector<unsigned char> data;
	size_t width, height;
	bool image;
};

struct _cl_kernel {
	std::string name;
	std::vector<unsigned char> args[16];
};

namespace {

typedef int (*ref_path_tracing_fn)(
	float, float, float, const void*, const void*, long long, const void*, const void*, long long, const void*, long long,
	const void*, long long, const void*, long long, const void*, long long, const float*, float*, float*, int, int, int, int, int);

ref_path_tracing_fn gKernelEntry = NULL;
int gThreads = 4;
std::string gProgramValues;
_cl_kernel* gLastKernel = NULL;

cl_mem argMem(const _cl_kernel* k, int slot) {
	cl_mem m = NULL;
	if (k->args[slot].size() == sizeof(cl_mem)) memcpy(&m, &k->args[slot][0], sizeof(cl_mem));
	return m;
}

} /* namespace */

extern "C" {

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t size, void* host_ptr, cl_int* errcode_ret) {
	_cl_mem* m = new _cl_mem();
	m->width = m->height = 0;
	m->image = false;
	m->data.resize(size);
	if (host_ptr && size) memcpy(&m->data[0], host_ptr, size);
	if (errcode_ret) *errcode_ret = CL_SUCCESS;
	return m;
}

int fakecl_set_kernel_library(const char* path, int threads) {
	void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!h) { fprintf(stderr, "fakecl: %s\n", dlerror()); return 1; }
	gKernelEntry = (ref_path_tracing_fn) dlsym(h, "ref_path_tracing");
	gThreads = threads > 0 ? threads : 1;
	return gKernelEntry ? 0 : 2;
}

/* "NAME=text\n" per placeholder, as CL::setValues would substitute them (CL.cpp:626-705) */
const char* fakecl_program_values(void) { return gProgramValues.c_str(); }

/* the argument bytes of the last launch: slot 0..3 by value, 4..13 the buffer / image contents */
long long fakecl_kernel_arg(int slot, void* dst) {
	if (!gLastKernel || slot < 0 || slot > 13) return -1;
	const _cl_kernel* k = gLastKernel;
	if (slot < 4) {
		if (dst && !k->args[slot].empty()) memcpy(dst, &k->args[slot][0], k->args[slot].size());
		return (long long) k->args[slot].size();
	}
	cl_mem m = argMem(k, slot);
	if (!m) return -1;
	if (dst && !m->data.empty()) memcpy(dst, &m->data[0], m->data.size());
	return (long long) m->data.size();
}

} /* extern "C" */

CL::CL( const bool silent ) {
	mDoCheckErrors = true;
	mWorkWidth = Cfg::get().value<cl_uint>( Cfg::WINDOW_WIDTH );
	mWorkHeight = Cfg::get().value<cl_uint>( Cfg::WINDOW_HEIGHT );
	mKernel = NULL;
}

CL::~CL() {
	this->freeBuffers();
	for( size_t i = 0; i < mKernels.size(); i++ ) {
		if( gLastKernel == mKernels[i] ) { gLastKernel = NULL; }
		delete mKernels[i];
	}
}

bool CL::checkError( cl_int err, const char* functionName ) {
	if( err != CL_SUCCESS ) {
		fprintf( stderr, "fakecl: error %d in %s\n", err, functionName );
		return false;
	}
	return true;
}

cl_mem CL::createEmptyBuffer( size_t size, cl_mem_flags flags ) {
	cl_mem m = clCreateBuffer( NULL, flags, size, NULL, NULL );
	mMemObjects.push_back( m );
	return m;
}

cl_mem CL::createImage2DReadOnly( size_t width, size_t height, cl_float* data ) {
	cl_mem m = clCreateBuffer( NULL, CL_MEM_READ_ONLY, width * height * 16, data, NULL );
	m->image = true; m->width = width; m->height = height;
	mMemObjects.push_back( m );
	return m;
}

cl_mem CL::createImage2DWriteOnly( size_t width, size_t height ) {
	cl_mem m = clCreateBuffer( NULL, CL_MEM_WRITE_ONLY, width * height * 16, NULL, NULL );
	m->image = true; m->width = width; m->height = height;
	mMemObjects.push_back( m );
	return m;
}

cl_kernel CL::createKernel( const char* functionName ) {
	_cl_kernel* k = new _cl_kernel();
	k->name = functionName;
	mKernels.push_back( k );
	mKernelNames[k] = k->name;
	mKernel = k;
	gLastKernel = k;
	return k;
}

void CL::setKernelArg( cl_kernel kernel, cl_uint index, size_t size, void* data ) {
	if( index >= 16 ) { return; }
	kernel->args[index].assign( (unsigned char*) data, (unsigned char*) data + size );
}

void CL::setReplacement( string before, string after ) {
	mReplaceString[before] = after;
}

/* The values CL::setValues substitutes (CL.cpp:626-705): integers "%u", floats "%ff", then the strings. */
void CL::loadProgram( string filepath ) {
	char text[64];
	std::string out;
	const float phongAlpha = Cfg::get().value<cl_float>( Cfg::RENDER_PHONGTESS );
	const char* intNames[9] = { "ACCEL_STRUCT", "BRDF", "IMG_HEIGHT", "IMG_WIDTH", "SHADOW_RAYS", "MAX_DEPTH", "MAX_ADDED_DEPTH", "PHONGTESS", "SAMPLES" };
	const cl_uint intValues[9] = {
		Cfg::get().value<cl_uint>( Cfg::ACCEL_STRUCT ), Cfg::get().value<cl_uint>( Cfg::RENDER_BRDF ),
		Cfg::get().value<cl_uint>( Cfg::WINDOW_HEIGHT ), Cfg::get().value<cl_uint>( Cfg::WINDOW_WIDTH ),
		Cfg::get().value<cl_uint>( Cfg::RENDER_SHADOWRAYS ), Cfg::get().value<cl_uint>( Cfg::RENDER_MAXDEPTH ),
		Cfg::get().value<cl_uint>( Cfg::RENDER_MAXADDEDDEPTH ), (cl_uint) ( phongAlpha > 0.0f ? 1 : 0 ),
		Cfg::get().value<cl_uint>( Cfg::RENDER_SAMPLES )
	};
	for( int i = 0; i < 9; i++ ) {
		snprintf( text, 64, "%u", intValues[i] );
		out += std::string( intNames[i] ) + "=" + text + "\n";
	}
	snprintf( text, 64, "%ff", Cfg::get().value<cl_float>( Cfg::RENDER_ANTIALIAS ) );
	out += std::string( "ANTI_ALIASING=" ) + text + "\n";
	snprintf( text, 64, "%ff", phongAlpha );
	out += std::string( "PHONGTESS_ALPHA=" ) + text + "\n";
	for( map<string, string>::iterator it = mReplaceString.begin(); it != mReplaceString.end(); it++ ) {
		std::string name = it->first;                       /* "#NAME#" */
		if( name.size() >= 2 ) { name = name.substr( 1, name.size() - 2 ); }
		out += name + "=" + it->second + "\n";
	}
	gProgramValues = out;
}

void CL::execute( cl_kernel kernel ) {
	if( !gKernelEntry ) { fprintf( stderr, "fakecl: execute() before fakecl_set_kernel_library()\n" ); return; }
	const _cl_kernel* k = kernel;
	float seed = 0.0f, weight = 0.0f, pxDim = 0.0f;
	memcpy( &seed, &k->args[0][0], 4 );
	memcpy( &weight, &k->args[1][0], 4 );
	memcpy( &pxDim, &k->args[2][0], 4 );
	cl_mem bvh = argMem( k, 4 ), fv = argMem( k, 5 ), fn = argMem( k, 6 ), vtx = argMem( k, 7 ), nrm = argMem( k, 8 );
	cl_mem mat = argMem( k, 9 ), lights = argMem( k, 10 ), in = argMem( k, 11 ), out = argMem( k, 12 ), dbg = argMem( k, 13 );
	const bool schlick = Cfg::get().value<cl_uint>( Cfg::RENDER_BRDF ) == 0;
	gKernelEntry(
		seed, weight, pxDim, &k->args[3][0],
		&bvh->data[0], (long long) ( bvh->data.size() / 32 ), &fv->data[0], &fn->data[0], (long long) ( fv->data.size() / 16 ),
		&vtx->data[0], (long long) ( vtx->data.size() / 16 ), &nrm->data[0], (long long) ( nrm->data.size() / 16 ),
		&mat->data[0], (long long) ( mat->data.size() / ( schlick ? 48 : 64 ) ), &lights->data[0], (long long) ( lights->data.size() / 48 ),
		(const float*) &in->data[0], (float*) &out->data[0], (float*) &dbg->data[0],
		(int) mWorkWidth, (int) mWorkHeight, 0, (int) mWorkHeight, gThreads
	);
}

void CL::finish() {}

void CL::freeBuffers() {
	for( size_t i = 0; i < mMemObjects.size(); i++ ) { delete mMemObjects[i]; }
	mMemObjects.clear();
}

map<cl_kernel, string> CL::getKernelNames() { return mKernelNames; }
map<cl_kernel, double> CL::getKernelTimes() { return mKernelTime; }

void CL::readImageOutput( cl_mem image, size_t width, size_t height, cl_float* outputTarget ) {
	memcpy( outputTarget, &image->data[0], width * height * 16 );
}

cl_mem CL::updateBuffer( cl_mem buffer, size_t size, void* data ) {
	buffer->data.assign( (unsigned char*) data, (unsigned char*) data + size );
	return buffer;
}

cl_mem CL::updateImageReadOnly( cl_mem image, size_t width, size_t height, cl_float* data ) {
	image->data.assign( (unsigned char*) data, (unsigned char*) data + width * height * 16 );
	return image;
}
