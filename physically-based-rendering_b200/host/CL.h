/*
 * CL -- the reference's device-runtime class (source/CL.h:20-83) re-targeted from the OpenCL 1.1 C API
 * onto the C ABI of libpbr_b200.so (include/pbr_b200.h).  Method names, argument meaning and error
 * behaviour are the reference's, so PathTracer drives it with the same call sequence:
 *     createBuffer<T>, createImage2DReadOnly / WriteOnly, setReplacement, loadProgram, createKernel,
 *     setKernelArg, execute, finish, readImageOutput, updateImageReadOnly, getKernelNames / Times.
 * Error convention (CL.cpp:89-99): a failing call logs "[OpenCL] Error in function <name>: ..." when
 * opencl.check_errors is set and execution continues; failures to obtain a device, a program or a
 * kernel end the process with EXIT_FAILURE like the reference (CL.cpp:209-211, 347-350, 438-448,
 * 523-525, 540-542, 564-566).
 */
#ifndef CL_H
#define CL_H

#include <map>
#include <string>
#include <vector>

#include "cl_types.h"
#include "Cfg.h"
#include "Logger.h"
#include "utils.h"

using std::map;
using std::string;
using std::vector;


class CL {
	public:
		CL( const bool silent = false );
		~CL();
		/** The device on which the next CL() is created (multi-GPU: one process per GPU). */
		static void setDefaultDevice( int device ) { sDefaultDevice = device; }
		pbr_ctx* getContext() { return mContext; }

		/* --- buffers and images (reference: CL.h:26-38, CL.cpp:135-197, 714-753) */
		template<typename T> cl_mem createBuffer( vector<T> object, size_t objectSize ) {
			return this->createBufferFromPtr( object.empty() ? NULL : &object[0], objectSize );
		}
		cl_mem createBufferFromPtr( const void* data, size_t objectSize );      /* additive: no by-value vector */
		cl_mem createEmptyBuffer( size_t size, int flags = 0 );
		cl_mem updateBuffer( cl_mem buffer, size_t size, void* data );
		cl_mem createImage2DReadOnly( size_t width, size_t height, cl_float* data );
		cl_mem createImage2DWriteOnly( size_t width, size_t height );
		cl_mem updateImageReadOnly( cl_mem image, size_t width, size_t height, cl_float* data );
		void readImageOutput( cl_mem image, size_t width, size_t height, cl_float* outputTarget );
		void freeBuffers();

		/* --- program and kernel (reference: CL.cpp:205-217, 289-316, 554-571, 604-617) */
		void setReplacement( string before, string after );
		void loadProgram( string filepath );
		cl_kernel createKernel( const char* functionName );
		void setKernelArg( cl_kernel kernel, cl_uint index, size_t size, void* data );
		void execute( cl_kernel kernel );
		void finish();
		map<cl_kernel, string> getKernelNames();
		map<cl_kernel, double> getKernelTimes();

		/* --- additive (include/pbr_b200.h): batches, split read-back, device-side copy, row sharding, counters,
		 *     pinned host memory */
		void executeBatch( cl_kernel kernel, cl_uint frames, const cl_float* seeds, const cl_float* pixelWeights );
		void readImageOutputBegin( cl_mem image, size_t width, size_t height, cl_float* outputTarget );
		void readImageOutputEnd();
		void copyImage( cl_mem dst, cl_mem src );
		void setTile( int y0, int y1 );
		void setTileStripes( int stripeRows, int world, int rank );
		void setDebugImage( bool enabled );
		void getStats( uint64_t out[6], bool reset );
		/* multi-GPU (pbr_comm_*): one CL per process and GPU, one collective per frame */
		static bool commUniqueId( void* id128 );
		bool commInit( const void* id128, int rank, int world );
		void frameCombine( cl_mem image, int mode, cl_mem out );
		void setBatchCombine( int mode, const cl_mem* outs, int numOuts, int first );
		void commFence();
		/* which walk finds the hits (pbr_set_traversal): -1 automatic, 0 reference order, 1 ordered */
		void setTraversal( int mode );
		void* allocHost( size_t bytes );
		void freeHost( void* ptr );

	protected:
		bool checkError( int err, const char* functionName );
		pbr_defines getValues();

	private:
		pbr_ctx* mContext;
		cl_uint mWorkWidth, mWorkHeight;
		bool mDoCheckErrors;
		vector<cl_kernel> mKernels;
		map<cl_kernel, string> mKernelNames;
		map<cl_kernel, double> mKernelTime;
		static int sDefaultDevice;
};

#endif
