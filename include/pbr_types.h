/*
 * pbr_types.h -- plain-old-data records that cross the host <-> device boundary.
 *
 * These mirror, field for field, the host structs the reference's PathTracer uploads
 * (reference: source/PathTracer.h:25-73) and the device structs its kernel reads
 * (reference: source/opencl/pt_header.cl:41-109).  They replace the cl_float4 /
 * cl_uint4 typedefs of the vendored cl.hpp; sizes and alignments are identical
 * (16-byte vectors, 16-byte aligned) so a byte buffer packed for the reference's
 * kernel is valid input here.
 *
 * C and C++ (host + CUDA) compatible.  No torch, no CUDA types.
 */
#ifndef PBR_TYPES_H
#define PBR_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__CUDACC__) || defined(__GNUC__)
#define PBR_ALIGN16 __attribute__((aligned(16)))
#else
#define PBR_ALIGN16
#endif

typedef struct PBR_ALIGN16 { float x, y, z, w; } pbr_float4;
typedef struct PBR_ALIGN16 { uint32_t x, y, z, w; } pbr_uint4;
typedef struct { int32_t x, y; } pbr_int2;
typedef struct { float x, y; } pbr_float2;

/* reference: PathTracer.h:25-32 / pt_header.cl:41-48.  cl_float3 is 16 bytes. 80 B. */
typedef struct PBR_ALIGN16 {
	pbr_float4 eye;        /* xyz used */
	pbr_float4 w;          /* view direction */
	pbr_float4 u;          /* right */
	pbr_float4 v;          /* up */
	pbr_int2 focusPoint;   /* (-1,-1) = depth of field off */
	pbr_float2 lense;      /* x: focal length, y: aperture */
} pbr_camera;

/* reference: PathTracer.h:39-43 / pt_header.cl:55-59. 48 B. */
typedef struct PBR_ALIGN16 {
	pbr_float4 pos;
	pbr_float4 rgb;
	pbr_float4 data;       /* x: type (1 point, 2 orb), y: radius */
} pbr_light;

/* reference: PathTracer.h:45-53 / pt_header.cl:86-94 (BRDF 0, Schlick). 48 B. */
typedef struct PBR_ALIGN16 {
	pbr_float4 data;       /* d, Ni, p, rough */
	pbr_float4 rgbDiff;
	pbr_float4 rgbSpec;
} pbr_material_schlick;

/* reference: PathTracer.h:55-65 / pt_header.cl:99-109 (BRDF 1, Shirley-Ashikhmin). 64 B. */
typedef struct PBR_ALIGN16 {
	float data[8];         /* d, Ni, nu, nv, Rs, Rd, pad, pad */
	pbr_float4 rgbDiff;
	pbr_float4 rgbSpec;
} pbr_material_sa;

/* reference: PathTracer.h:70-73 / pt_header.cl:65-68. 32 B.
 * bbMin.w: index of the first face of a leaf, or -1.0f for an inner node.
 * bbMax.w: leaf: index of the second face or -1.0f;
 *          inner: index of the node to continue with if the box is missed, -1.0f = stop.
 * Nodes are stored in pre-order, left child at i+1; node 0 is the never-visited root. */
typedef struct PBR_ALIGN16 {
	pbr_float4 bbMin;
	pbr_float4 bbMax;
} pbr_bvh_node;

/* A ray for the explicit-ray entry points (additive API).
 * origin.w: unused.  dir.w: initial t (INFINITY for a primary ray, distance to the
 * light for a shadow ray). */
typedef struct PBR_ALIGN16 {
	pbr_float4 origin;
	pbr_float4 dir;
} pbr_ray;

/* Result of one explicit ray. */
typedef struct PBR_ALIGN16 {
	float t;               /* INFINITY = nothing hit */
	int32_t hitFace;       /* index into the leaf-ordered facesV; <0: -(light+1); 0 with t=INF: miss */
	int32_t leaf;          /* flattened index of the leaf node that holds the hit; -1 = none */
	uint32_t visits;       /* bits 0..19: BVH nodes visited, bits 20..31: triangle tests (saturating) */
} pbr_hit;

/* The values the reference splices into pt_header.cl as text before compiling
 * (reference: source/CL.cpp:626-705, PathTracer.cpp:210,338,472-516). */
typedef struct {
	int32_t accel_struct;     /* ACCEL_STRUCT      (accel_struct)              */
	int32_t brdf;             /* BRDF              (render.brdf) 0 Schlick, 1 Shirley-Ashikhmin */
	int32_t img_width;        /* IMG_WIDTH         (window.width)              */
	int32_t img_height;       /* IMG_HEIGHT        (window.height)             */
	int32_t shadow_rays;      /* SHADOW_RAYS       (render.shadow_rays)        */
	int32_t max_depth;        /* MAX_DEPTH         (render.max_depth)          */
	int32_t max_added_depth;  /* MAX_ADDED_DEPTH   (render.max_added_depth)    */
	int32_t phongtess;        /* PHONGTESS         (render.phong_tessellation > 0) */
	int32_t samples;          /* SAMPLES           (render.samples)            */
	float anti_aliasing;      /* ANTI_ALIASING     (render.antialiasing)       */
	float phongtess_alpha;    /* PHONGTESS_ALPHA   (render.phong_tessellation) */
	int32_t bvh_num_nodes;    /* BVH_NUM_NODES     (#BVH_NUM_NODES#)           */
	int32_t num_lights;       /* NUM_LIGHTS        (#NUM_LIGHTS#)              */
	pbr_float4 sky_light;     /* SKY_LIGHT         (#SKY_LIGHT#)               */
} pbr_defines;

#ifdef __cplusplus
}
#endif

#endif /* PBR_TYPES_H */
