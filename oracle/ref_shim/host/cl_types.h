/*
 * cl_types.h -- force-included (-include) when the reference's HOST sources are compiled for the tests
 * (oracle/build_ref_host.py).  TEST INFRASTRUCTURE.
 *
 * The reference includes a vendored Khronos "cl.hpp" next to its sources; that header needs the OpenCL C API,
 * which this image does not have.  Its include guard is defined here so that it is skipped, and the handful
 * of OpenCL host types the reference's parsers and BVH builder use are declared instead, laid out like the
 * Khronos ones (cl_platform.h): a vector type is a union whose first member is the array `s`, so that
 * `cl_float4 v = { a, b, c, d }` means the same thing.
 */
#ifndef PBR_REF_CL_TYPES_H
#define PBR_REF_CL_TYPES_H

#define CL_HPP_           /* skip /root/reference/source/cl.hpp */
#define GLWIDGET_H        /* skip /root/reference/source/qt/GLWidget.h (Qt, GLEW, GLUT); stand-in below */

/* what the Khronos headers would have brought in (cl_platform.h, cl.hpp) */
#include <float.h>
#include <limits.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>    /* uint (BVH.h) */
#ifdef __cplusplus
/* <cmath>, not <math.h>: with the toolchain the reference was written for, an unqualified tan( float ) is the C
 * function (binary64); libstdc++'s newer <math.h> wrapper would pull the float overloads into the global
 * namespace and silently change PathTracer::initKernelArgs' pixel size (PathTracer.cpp:90). */
#include <cmath>
#include <algorithm>
#include <iostream>
#include <limits>
#include <string>
#include <utility>
#include <vector>
#endif

typedef int8_t cl_char;
typedef uint8_t cl_uchar;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef float cl_float;
typedef double cl_double;
typedef uint64_t cl_ulong;

typedef union {
	cl_float s[4] __attribute__((aligned(16)));
	struct { cl_float x, y, z, w; };
	struct { cl_float s0, s1, s2, s3; };
} cl_float4;

typedef union {
	cl_uint s[4] __attribute__((aligned(16)));
	struct { cl_uint x, y, z, w; };
	struct { cl_uint s0, s1, s2, s3; };
} cl_uint4;

typedef union {
	cl_int s[4] __attribute__((aligned(16)));
	struct { cl_int x, y, z, w; };
	struct { cl_int s0, s1, s2, s3; };
} cl_int4;

typedef union {
	cl_float s[2] __attribute__((aligned(8)));
	struct { cl_float x, y; };
	struct { cl_float s0, s1; };
} cl_float2;

typedef union {
	cl_int s[2] __attribute__((aligned(8)));
	struct { cl_int x, y; };
	struct { cl_int s0, s1; };
} cl_int2;

typedef cl_float4 cl_float3;      /* as in cl_platform.h */

typedef union {
	cl_float s[8] __attribute__((aligned(32)));
	struct { cl_float s0, s1, s2, s3, s4, s5, s6, s7; };
} cl_float8;


/* ---- the OpenCL objects CL.h declares members of: opaque here, implemented by ref_shim/fake_cl.cpp ---- */
typedef struct _cl_mem* cl_mem;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_event* cl_event;
typedef struct _cl_program* cl_program;
typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_platform_id* cl_platform_id;
typedef cl_ulong cl_mem_flags;
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_SUCCESS 0
#ifdef __cplusplus
extern "C"
#endif
cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void* host_ptr, cl_int* errcode_ret);

#ifdef __cplusplus
/* ---- qt/GLWidget.h stand-in: what PathTracer.cpp and Camera.cpp call on their parent widget ---- */
class CL;
class PathTracer;
class GLWidget {
	public:
		GLWidget() : mPathTracer(NULL) {}
		void cameraUpdate();                       /* GLWidget.cpp:78-82: resets the sample count */
		void createKernelWindow( CL* ) {}
		void resetRenderTime() {}
		PathTracer* mPathTracer;
};
#endif

#endif
