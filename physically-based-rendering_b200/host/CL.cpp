#include "CL.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int CL::sDefaultDevice = -1;


/**
 * Constructor: picks the device and creates the context + queue
 * (reference: CL::CL, getDefaultPlatform, getDefaultDevice, initContext, initCommandQueue;
 * CL.cpp:10-24, 338-473, 513-547).
 */
CL::CL( const bool silent ) {
	mContext = NULL;
	mDoCheckErrors = Cfg::get().value<bool>( Cfg::OPENCL_CHECKERRORS );
	mWorkWidth = Cfg::get().value<cl_uint>( Cfg::WINDOW_WIDTH );
	mWorkHeight = Cfg::get().value<cl_uint>( Cfg::WINDOW_HEIGHT );

	const int err = pbr_create( sDefaultDevice, &mContext );
	if( err != PBR_OK || mContext == NULL ) {
		Logger::logError( "[OpenCL] No usable CUDA device found (pbr_create). There is no CPU fallback." );
		exit( EXIT_FAILURE );
	}

	if( !silent ) {
		char name[256];
		int sms = 0;
		size_t mem = 0;
		pbr_device_info( mContext, name, sizeof( name ), &sms, &mem );
		float memFloat;
		string unit;
		utils::formatBytes( mem, &memFloat, &unit );
		char msg[512];
		snprintf( msg, 512, "[OpenCL] Using device %s (%d SMs, %.2f %s global memory).", name, sms, memFloat, unit.c_str() );
		Logger::logInfo( msg );
	}
}


/** Destructor: releases buffers, images, kernels, stream (reference: CL.cpp:30-52). */
CL::~CL() {
	if( mContext ) {
		this->checkError( pbr_destroy( mContext ), "clReleaseContext" );
	}
}


/** Reference: CL.cpp:89-99. */
bool CL::checkError( int err, const char* functionName ) {
	if( mDoCheckErrors && err != PBR_OK ) {
		char msg[512];
		snprintf(
			msg, 512, "[OpenCL] Error in function %s: %s (code %d)",
			functionName, mContext ? pbr_last_error( mContext ) : "", err
		);
		Logger::logError( msg );
		return false;
	}
	return true;
}


cl_mem CL::createBufferFromPtr( const void* data, size_t objectSize ) {
	cl_mem buffer = 0;
	this->checkError( pbr_buffer_create( mContext, data, objectSize, &buffer ), "clCreateBuffer" );
	return buffer;
}


/** Reference: CL.cpp:136-144. */
cl_mem CL::createEmptyBuffer( size_t size, int flags ) {
	(void) flags;
	cl_mem buffer = 0;
	this->checkError( pbr_buffer_create_empty( mContext, size, &buffer ), "clCreateBuffer" );
	return buffer;
}


/** Reference: CL.cpp:153-178. */
cl_mem CL::createImage2DReadOnly( size_t width, size_t height, cl_float* data ) {
	cl_mem image = 0;
	this->checkError( pbr_image_create( mContext, width, height, data, &image ), "clCreateImage2D" );
	return image;
}


/** Reference: CL.cpp:186-197. */
cl_mem CL::createImage2DWriteOnly( size_t width, size_t height ) {
	cl_mem image = 0;
	this->checkError( pbr_image_create( mContext, width, height, NULL, &image ), "clCreateImage2D" );
	return image;
}


/** Reference: CL.cpp:205-217 -- a missing kernel is fatal. */
cl_kernel CL::createKernel( const char* functionName ) {
	cl_kernel kernel = 0;
	const int err = pbr_kernel_get( mContext, functionName, &kernel );
	if( !this->checkError( err, "clCreateKernel" ) || err != PBR_OK ) {
		exit( EXIT_FAILURE );
	}
	mKernels.push_back( kernel );
	mKernelNames[kernel] = string( functionName );
	return kernel;
}


/**
 * Execute a kernel over window.width x window.height (reference: CL.cpp:289-306).  The launch is
 * asynchronous; its device time is available through getKernelTimes() after finish().
 */
void CL::execute( cl_kernel kernel ) {
	this->checkError( pbr_kernel_launch( mContext, kernel ), "clEnqueueNDRangeKernel" );
	mKernelTime[kernel] = -1.0;
}


/**
 * Additive: `frames` consecutive frames of `kernel` in one call (pbr_kernel_launch_batch) -- what a loop of
 * setKernelArg( 0, seed ), setKernelArg( 1, weight ), execute(), "output becomes input" does, without the
 * device draining between frames.
 */
void CL::executeBatch( cl_kernel kernel, cl_uint frames, const cl_float* seeds, const cl_float* pixelWeights ) {
	this->checkError( pbr_kernel_launch_batch( mContext, kernel, (int32_t) frames, seeds, pixelWeights ), "clEnqueueNDRangeKernel" );
	mKernelTime[kernel] = -1.0;
}


/** Reference: CL.cpp:312-316. */
void CL::finish() {
	this->checkError( pbr_finish( mContext ), "clFinish" );
}


/** Reference: CL.cpp:322-331. */
void CL::freeBuffers() {
	this->checkError( pbr_free_buffers( mContext ), "clReleaseMemObject" );
}


map<cl_kernel, string> CL::getKernelNames() {
	return mKernelNames;
}


/** Reference: CL.cpp:480-506 (event END - START of the last launch, in ms). */
map<cl_kernel, double> CL::getKernelTimes() {
	for( map<cl_kernel, double>::iterator it = mKernelTime.begin(); it != mKernelTime.end(); it++ ) {
		double ms = 0.0;
		if( pbr_kernel_time_ms( mContext, it->first, &ms ) == PBR_OK ) {
			it->second = ms;
		}
	}
	return mKernelTime;
}


/** The values CL::setValues splices into pt_header.cl (reference: CL.cpp:626-705). */
pbr_defines CL::getValues() {
	pbr_defines d;
	memset( &d, 0, sizeof( d ) );
	const float phongTessAlpha = Cfg::get().value<cl_float>( Cfg::RENDER_PHONGTESS );
	d.accel_struct = (int32_t) Cfg::get().value<cl_uint>( Cfg::ACCEL_STRUCT );
	d.brdf = (int32_t) Cfg::get().value<cl_uint>( Cfg::RENDER_BRDF );
	d.img_height = (int32_t) mWorkHeight;
	d.img_width = (int32_t) mWorkWidth;
	d.shadow_rays = (int32_t) Cfg::get().value<cl_uint>( Cfg::RENDER_SHADOWRAYS );
	d.max_depth = (int32_t) Cfg::get().value<cl_uint>( Cfg::RENDER_MAXDEPTH );
	d.max_added_depth = (int32_t) Cfg::get().value<cl_uint>( Cfg::RENDER_MAXADDEDDEPTH );
	d.phongtess = phongTessAlpha > 0.0f ? 1 : 0;
	d.samples = (int32_t) Cfg::get().value<cl_uint>( Cfg::RENDER_SAMPLES );
	/* the reference prints floats with "%ff" into the source, i.e. six decimals */
	char text[32];
	snprintf( text, 32, "%f", Cfg::get().value<cl_float>( Cfg::RENDER_ANTIALIAS ) );
	d.anti_aliasing = (float) atof( text );
	snprintf( text, 32, "%f", phongTessAlpha );
	d.phongtess_alpha = (float) atof( text );
	d.sky_light.x = d.sky_light.y = d.sky_light.z = 1.0f;
	return d;
}


/**
 * "Load the program" (reference: CL.cpp:554-571).  The path argument names the reference's OpenCL
 * source; nothing is read from it -- the kernel is precompiled for sm_100a and the config values the
 * reference would substitute into the source are handed over as a struct.  Failure is fatal.
 */
void CL::loadProgram( string filepath ) {
	(void) filepath;
	pbr_defines d = this->getValues();
	const int err = pbr_program_load( mContext, &d );
	if( err != PBR_OK ) {
		this->checkError( err, "clBuildProgram" );
		Logger::logError( string( "[OpenCL] " ) + pbr_last_error( mContext ) );
		exit( EXIT_FAILURE );
	}
}


/** Reference: CL.cpp:581-594. */
void CL::readImageOutput( cl_mem image, size_t width, size_t height, cl_float* outputTarget ) {
	this->checkError( pbr_image_read( mContext, image, width, height, outputTarget ), "clEnqueueReadImage" );
}


/** Additive: readImageOutput in two halves; what is enqueued in between overlaps the copy. */
void CL::readImageOutputBegin( cl_mem image, size_t width, size_t height, cl_float* outputTarget ) {
	this->checkError( pbr_image_read_begin( mContext, image, width, height, outputTarget ), "clEnqueueReadImage" );
}


void CL::readImageOutputEnd() {
	this->checkError( pbr_image_read_end( mContext ), "clWaitForEvents" );
}


/** Reference: CL.cpp:604-607. */
void CL::setKernelArg( cl_kernel kernel, cl_uint index, size_t size, void* data ) {
	this->checkError( pbr_kernel_set_arg( mContext, kernel, index, size, data ), "clSetKernelArg" );
}


/** Reference: CL.cpp:615-617. */
void CL::setReplacement( string before, string after ) {
	this->checkError( pbr_set_define( mContext, before.c_str(), after.c_str() ), "setReplacement" );
}


/** Reference: CL.cpp:715-728. */
cl_mem CL::updateBuffer( cl_mem buffer, size_t size, void* data ) {
	this->checkError( pbr_buffer_update( mContext, buffer, size, data ), "clEnqueueWriteBuffer" );
	return buffer;
}


/** Reference: CL.cpp:738-753. */
cl_mem CL::updateImageReadOnly( cl_mem image, size_t width, size_t height, cl_float* data ) {
	this->checkError( pbr_image_write( mContext, image, width, height, data ), "clEnqueueWriteImage" );
	return image;
}


void CL::copyImage( cl_mem dst, cl_mem src ) {
	this->checkError( pbr_image_copy( mContext, dst, src ), "clEnqueueCopyImage" );
}


void CL::setTile( int y0, int y1 ) {
	this->checkError( pbr_set_tile( mContext, y0, y1 ), "pbr_set_tile" );
}


void CL::setTileStripes( int stripeRows, int world, int rank ) {
	this->checkError( pbr_set_tile_stripes( mContext, stripeRows, world, rank ), "pbr_set_tile_stripes" );
}


void CL::setDebugImage( bool enabled ) {
	this->checkError( pbr_set_debug_image( mContext, enabled ? 1 : 0 ), "pbr_set_debug_image" );
}


void CL::getStats( uint64_t out[6], bool reset ) {
	this->checkError( pbr_stats( mContext, out, reset ? 1 : 0 ), "pbr_stats" );
}


bool CL::commUniqueId( void* id128 ) {
	return pbr_comm_unique_id( id128 ) == 0;
}


bool CL::commInit( const void* id128, int rank, int world ) {
	return this->checkError( pbr_comm_init( mContext, id128, rank, world ), "pbr_comm_init (ncclCommInitRank)" );
}


void CL::frameCombine( cl_mem image, int mode, cl_mem out ) {
	this->checkError( pbr_frame_combine( mContext, image, mode, out ), "pbr_frame_combine" );
}


void CL::setBatchCombine( int mode, const cl_mem* outs, int numOuts, int first ) {
	this->checkError( pbr_set_batch_combine( mContext, mode, outs, numOuts, first ), "pbr_set_batch_combine" );
}


void CL::commFence() {
	this->checkError( pbr_comm_fence( mContext ), "pbr_comm_fence" );
}


void CL::setTraversal( int mode ) {
	this->checkError( pbr_set_traversal( mContext, mode ), "pbr_set_traversal" );
}


void* CL::allocHost( size_t bytes ) {
	void* p = NULL;
	this->checkError( pbr_host_alloc( mContext, bytes, &p ), "pbr_host_alloc" );
	return p;
}


void CL::freeHost( void* ptr ) {
	this->checkError( pbr_host_free( mContext, ptr ), "pbr_host_free" );
}
