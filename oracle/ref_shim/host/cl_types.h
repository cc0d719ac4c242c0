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

/* what the Khronos headers would have brought in (cl_platform.h, cl.hpp) */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>    /* uint (BVH.h) */
#ifdef __cplusplus
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

typedef union {
	cl_float s[8] __attribute__((aligned(32)));
	struct { cl_float s0, s1, s2, s3, s4, s5, s6, s7; };
} cl_float8;

#endif
