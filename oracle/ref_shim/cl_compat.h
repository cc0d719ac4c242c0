/*
 * cl_compat.h -- just enough OpenCL C on top of C++ to compile the REFERENCE's own kernel source
 * (/root/reference/source/opencl/pathtracing.cl and the pt_*.cl files it splices in) for the host CPU.
 *
 * TEST INFRASTRUCTURE.  oracle/build_ref.py assembles the program text the way the reference's CL class does
 * (CL::combineParts, CL::setValues: CL.cpp:107-127, 626-705), wraps it in `namespace clref { ... }` after
 * this header and compiles it into oracle/_ref/ (git-ignored; nothing of the reference is copied into the
 * repository).  The result is the reference kernel itself, executed one work-item after the other, with
 * every OpenCL built-in given the meaning include/pbr_pinned_math.h pins for it -- which makes it the
 * yardstick for oracle/pt_oracle.cpp (the restatement) and, through it, for the CUDA kernels.
 *
 * What is here, and only that: the vector types and swizzles the kernel uses (float2/3/4/8, int2/3, uint4;
 * .x .y .z .w .s0-.s7 .xyz .yzx), component-wise operators, the built-ins it calls, 2-D float images with a
 * nearest / clamp-to-edge sampler, get_global_id.
 */
#ifndef PBR_REF_CL_COMPAT_H
#define PBR_REF_CL_COMPAT_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>       /* for ref_driver.inc: every system header comes in before the qualifier macros below */
#include <vector>

#include "pbr_pinned_math.h"

namespace clref {

typedef unsigned int uint;

/* address spaces and access qualifiers mean nothing on the host */
#define global
#define constant static const
#define kernel
#define __kernel
#define read_only
#define write_only

struct float3;
struct float4;

/* .xyz of a float3 / float4 and .yzx of a float3: views of the first three floats */
struct xyz_view {
	float d[3];
	inline operator float3() const;
	inline xyz_view& operator=(const float3& v);
};
struct yzx_view {
	float d[3];
	inline operator float3() const;
};

struct alignas(8) float2 {
	float x, y;
	float2() {}
	explicit float2(float s) : x(s), y(s) {}
	float2(float a, float b) : x(a), y(b) {}
};

struct alignas(8) int2 {
	int x, y;
	int2() {}
	int2(int a, int b) : x(a), y(b) {}
};

struct alignas(16) int3 {
	int x, y, z, pad;
};

struct alignas(16) uint4 {
	uint x, y, z, w;
};

/* OpenCL: sizeof(float3) == sizeof(float4) == 16 */
struct alignas(16) float3 {
	union {
		struct { float x, y, z, pad; };
		struct { float s0, s1, s2, pad_s; };
		xyz_view xyz;
		yzx_view yzx;
	};
	float3() {}
	explicit float3(float s) : x(s), y(s), z(s), pad(0.0f) {}
	float3(float a, float b, float c) : x(a), y(b), z(c), pad(0.0f) {}
};

struct alignas(16) float4 {
	union {
		struct { float x, y, z, w; };
		struct { float s0, s1, s2, s3; };
		xyz_view xyz;
	};
	float4() {}
	explicit float4(float s) : x(s), y(s), z(s), w(s) {}
	float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

struct alignas(32) float8 {
	float s0, s1, s2, s3, s4, s5, s6, s7;
};

inline xyz_view::operator float3() const { return float3(d[0], d[1], d[2]); }
inline xyz_view& xyz_view::operator=(const float3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; return *this; }
inline yzx_view::operator float3() const { return float3(d[1], d[2], d[0]); }

/* ---- component-wise operators (IEEE, no contraction: the file is compiled with -ffp-contract=off) ---- */

#define CLREF_OPS3(OP) \
	inline float3 operator OP(const float3& a, const float3& b) { return float3(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
	inline float3 operator OP(const float3& a, float s) { return float3(a.x OP s, a.y OP s, a.z OP s); } \
	inline float3 operator OP(float s, const float3& a) { return float3(s OP a.x, s OP a.y, s OP a.z); } \
	inline float3& operator OP##=(float3& a, const float3& b) { a = a OP b; return a; } \
	inline float3& operator OP##=(float3& a, float s) { a = a OP s; return a; }
CLREF_OPS3(+) CLREF_OPS3(-) CLREF_OPS3(*) CLREF_OPS3(/)
inline float3 operator-(const float3& a) { return float3(-a.x, -a.y, -a.z); }

#define CLREF_OPS4(OP) \
	inline float4 operator OP(const float4& a, const float4& b) { return float4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
	inline float4 operator OP(const float4& a, float s) { return float4(a.x OP s, a.y OP s, a.z OP s, a.w OP s); } \
	inline float4 operator OP(float s, const float4& a) { return float4(s OP a.x, s OP a.y, s OP a.z, s OP a.w); } \
	inline float4& operator OP##=(float4& a, const float4& b) { a = a OP b; return a; } \
	inline float4& operator OP##=(float4& a, float s) { a = a OP s; return a; }
CLREF_OPS4(+) CLREF_OPS4(-) CLREF_OPS4(*) CLREF_OPS4(/)
inline float4 operator-(const float4& a) { return float4(-a.x, -a.y, -a.z, -a.w); }

/* the swizzle views take part in expressions as float3 */
#define CLREF_VIEW_OPS(OP) \
	inline float3 operator OP(const xyz_view& a, const float3& b) { return float3(a) OP b; } \
	inline float3 operator OP(const float3& a, const xyz_view& b) { return a OP float3(b); } \
	inline float3 operator OP(const xyz_view& a, const xyz_view& b) { return float3(a) OP float3(b); }
CLREF_VIEW_OPS(+) CLREF_VIEW_OPS(-) CLREF_VIEW_OPS(*) CLREF_VIEW_OPS(/)

/* vector comparison: -1 where equal, 0 where not (OpenCL C 6.3.d) */
inline int3 operator==(const float3& a, const float3& b) {
	int3 r; r.x = (a.x == b.x) ? -1 : 0; r.y = (a.y == b.y) ? -1 : 0; r.z = (a.z == b.z) ? -1 : 0; r.pad = 0; return r;
}
inline int3 operator+(const int3& a, const int3& b) {
	int3 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; r.pad = 0; return r;
}

/* ---- built-ins, each with the one meaning the pinned arithmetic contract gives it ---- */

inline pm::vec3 to_pm(const float3& a) { return pm::v3(a.x, a.y, a.z); }
inline float3 from_pm(const pm::vec3& a) { return float3(a.x, a.y, a.z); }

inline float dot(const float3& a, const float3& b) { return pm::dot(to_pm(a), to_pm(b)); }
inline float3 cross(const float3& a, const float3& b) { return from_pm(pm::cross(to_pm(a), to_pm(b))); }
inline float3 fast_normalize(const float3& a) { return from_pm(pm::normalize(to_pm(a))); }
inline float length(const float3& a) { return pm::length(to_pm(a)); }

inline float fmax(float a, float b) { return pm::max_(a, b); }
inline float fmin(float a, float b) { return pm::min_(a, b); }
inline float3 fmax(const float3& a, const float3& b) { return float3(pm::max_(a.x, b.x), pm::max_(a.y, b.y), pm::max_(a.z, b.z)); }
inline float3 fmin(const float3& a, const float3& b) { return float3(pm::min_(a.x, b.x), pm::min_(a.y, b.y), pm::min_(a.z, b.z)); }
inline float max(float a, float b) { return pm::max_(a, b); }
inline float min(float a, float b) { return pm::min_(a, b); }
inline int max(int a, int b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline float fabs(float a) { return ::fabsf(a); }
inline float3 fabs(const float3& a) { return float3(::fabsf(a.x), ::fabsf(a.y), ::fabsf(a.z)); }

inline float clamp(float x, float lo, float hi) { return pm::clamp_(x, lo, hi); }
inline float4 clamp(const float4& v, float lo, float hi) {
	return float4(pm::clamp_(v.x, lo, hi), pm::clamp_(v.y, lo, hi), pm::clamp_(v.z, lo, hi), pm::clamp_(v.w, lo, hi));
}
inline float mix(float x, float y, float a) { return pm::mix_(x, y, a); }
inline float4 mix(const float4& x, const float4& y, float a) {
	return float4(pm::mix_(x.x, y.x, a), pm::mix_(x.y, y.y, a), pm::mix_(x.z, y.z, a), pm::mix_(x.w, y.w, a));
}
inline float fract(float x, float* ip) { *ip = ::floorf(x); return pm::fract_(x); }

inline float native_sin(float x) { return pm::sin_(x); }
inline float native_cos(float x) { return pm::cos_(x); }
inline float native_tan(float x) { return pm::tan_(x); }
inline float native_sqrt(float x) { return pm::sqrt_(x); }
inline float native_recip(float x) { return pm::rcp(x); }
inline float3 native_recip(const float3& v) { return float3(pm::rcp(v.x), pm::rcp(v.y), pm::rcp(v.z)); }
inline float native_divide(float a, float b) { return pm::divide(a, b); }
inline float pow(float x, float y) { return pm::pow_(x, y); }
inline float acos(float x) { return pm::acos_(x); }
inline float atan(float x) { return pm::atan_(x); }
inline float cbrt(float x) { return pm::cbrt_(x); }

/* fma: the reference passes the scalar in either of the first two places */
inline float3 fma(float s, const float3& a, const float3& c) { return from_pm(pm::fma3(to_pm(a), s, to_pm(c))); }
inline float3 fma(const float3& a, float s, const float3& c) { return from_pm(pm::fma3(to_pm(a), s, to_pm(c))); }

/* ---- images, sampler, work-item id ---- */

typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_FALSE = 0, CLK_ADDRESS_CLAMP_TO_EDGE = 0, CLK_FILTER_NEAREST = 0 };

struct image2d_t {
	float4* data;
	int width, height;
};

inline float4 read_imagef(const image2d_t& img, sampler_t, const int2& pos) {
	const int x = pos.x < 0 ? 0 : (pos.x >= img.width ? img.width - 1 : pos.x);
	const int y = pos.y < 0 ? 0 : (pos.y >= img.height ? img.height - 1 : pos.y);
	return img.data[(size_t) y * img.width + x];
}
inline void write_imagef(const image2d_t& img, const int2& pos, const float4& c) {
	if (pos.x < 0 || pos.y < 0 || pos.x >= img.width || pos.y >= img.height) return;
	img.data[(size_t) pos.y * img.width + pos.x] = c;
}

extern thread_local int g_global_id[2];
inline int get_global_id(int dim) { return g_global_id[dim]; }

} /* namespace clref */

#endif /* PBR_REF_CL_COMPAT_H */
