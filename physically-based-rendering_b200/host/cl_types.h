/*
 * cl_types.h -- the handful of `cl_*` names the reference's host code uses, without OpenCL.
 *
 * The reference includes the vendored Khronos cl.hpp (source/cl.hpp) only for these typedefs
 * (cl_float, cl_uint, cl_float4, cl_uint4, cl_mem, cl_kernel ...).  Here the vector types are the
 * POD records of include/pbr_types.h and the two handle types are the 64-bit handles of the C ABI.
 */
#ifndef PBR_HOST_CL_TYPES_H
#define PBR_HOST_CL_TYPES_H

#include <stdint.h>

#include "../../include/pbr_b200.h"

typedef float cl_float;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int8_t cl_char;
typedef pbr_float4 cl_float4;
typedef pbr_float4 cl_float3;    /* OpenCL: float3 occupies 16 bytes */
typedef pbr_uint4 cl_uint4;
typedef pbr_int2 cl_int2;
typedef pbr_float2 cl_float2;
typedef pbr_mem cl_mem;
typedef pbr_kernel cl_kernel;

#endif
