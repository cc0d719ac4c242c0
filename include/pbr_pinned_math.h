/*
 * pbr_pinned_math.h -- the pinned arithmetic contract of the path tracer.
 *
 * The reference kernel (source/opencl/pt_*.cl) is written against OpenCL C built-ins whose
 * results are implementation-defined: native_recip / native_divide / native_sqrt /
 * native_sin / native_cos / native_tan, fast_normalize, and the <=4..16 ulp acos / atan /
 * pow / cbrt (e.g. pt_bvh.cl:83, pt_intersect.cl:57,103,122, pt_utils.cl:43, pt_brdf.cl:297-310).
 * Two devices running the reference therefore do not agree bit-for-bit.  To make "same
 * scene + same rays => same hit index / same radiance" a testable statement, this header
 * gives every one of those built-ins ONE meaning, expressed only in
 * IEEE-754 +,-,*,/ and sqrt in binary32/binary64, round-to-nearest-even, no fused
 * multiply-add except where the reference itself writes fma().  Compiled with
 *     nvcc  -fmad=false  (default -prec-div=true -prec-sqrt=true -ftz=false)
 *     g++   -ffp-contract=off  (x86-64 SSE2, no -ffast-math)
 * the functions below return identical bits on the B200 and on the host.  Both the CUDA
 * kernels (csrc/) and the CPU oracle (oracle/) include this file: it is the specification of
 * the arithmetic, not an implementation of the algorithm.
 *
 * Transcendentals are evaluated in binary64 with plain Taylor/Maclaurin polynomials after an
 * exact or two-constant range reduction and rounded once to binary32; they are accurate to
 * well below one binary32 ulp for the argument ranges the renderer produces (checked against
 * libm in tests/test_pinned_math.py), which is inside every OpenCL native_* allowance.
 */
#ifndef PBR_PINNED_MATH_H
#define PBR_PINNED_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif

namespace pm {

/* ---------------------------------------------------------------- bit casts */

PM_HD uint64_t d2bits(double d) {
#if defined(__CUDA_ARCH__)
	return (uint64_t) __double_as_longlong(d);
#else
	uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
PM_HD double bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long) u);
#else
	double d; memcpy(&d, &u, 8); return d;
#endif
}
PM_HD uint32_t f2bits(float f) {
#if defined(__CUDA_ARCH__)
	return (uint32_t) __float_as_int(f);
#else
	uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PM_HD float bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
	return __int_as_float((int) u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}

#define PM_INF_F (pm::bits2f(0x7f800000u))
#define PM_NAN_F (pm::bits2f(0x7fc00000u))

/* ------------------------------------------------- binary32 scalar built-ins */

/* native_recip, native_divide, native_sqrt: IEEE, correctly rounded. */
PM_HD float rcp(float x) { return 1.0f / x; }
PM_HD float divide(float a, float b) { return a / b; }
PM_HD float sqrt_(float x) { return sqrtf(x); }
/* OpenCL max/min/clamp (gentype): comparisons, not the NaN-dropping fmin/fmax. */
PM_HD float max_(float x, float y) { return (x < y) ? y : x; }
PM_HD float min_(float x, float y) { return (y < x) ? y : x; }
PM_HD float clamp_(float x, float lo, float hi) { return min_(max_(x, lo), hi); }
/* OpenCL fract(x, &i): fmin(x - floor(x), 0x1.fffffep-1f). */
PM_HD float fract_(float x) { return fminf(x - floorf(x), 0x1.fffffep-1f); }
/* OpenCL mix(x, y, a) = x + (y - x) * a. */
PM_HD float mix_(float x, float y, float a) { return x + (y - x) * a; }

/* ------------------------------------------- binary64 kernels (internal use) */

/* Reduce x to r in [-pi/4, pi/4], return quadrant (x = q*pi/2 + r). */
PM_HD int rem_pio2(double x, double* r) {
	const double fn = rint(x * 6.36619772367581382433e-01);
	*r = (x - fn * 1.57079632673412561417e+00) - fn * 6.07710050650619224932e-11;
	return (int) (((long long) fn) & 3);
}

PM_HD double ksin(double r) {
	const double z = r * r;
	double p = 2.81145725434552075980e-15;              /*  1/17! */
	p = -7.64716373181981647590e-13 + z * p;             /* -1/15! */
	p = 1.60590438368216145994e-10 + z * p;              /*  1/13! */
	p = -2.50521083854417187751e-08 + z * p;             /* -1/11! */
	p = 2.75573192239858906526e-06 + z * p;              /*  1/9!  */
	p = -1.98412698412698412698e-04 + z * p;             /* -1/7!  */
	p = 8.33333333333333333333e-03 + z * p;              /*  1/5!  */
	p = -1.66666666666666666667e-01 + z * p;             /* -1/3!  */
	return r + r * (z * p);
}

PM_HD double kcos(double r) {
	const double z = r * r;
	double p = 1.56192069685862264622e-16;               /*  1/18! */
	p = -4.77947733238738529744e-14 + z * p;             /* -1/16! */
	p = 1.14707455977297247139e-11 + z * p;              /*  1/14! */
	p = -2.08767569878680989792e-09 + z * p;             /* -1/12! */
	p = 2.75573192239858906526e-07 + z * p;              /*  1/10! */
	p = -2.48015873015873015873e-05 + z * p;             /* -1/8!  */
	p = 1.38888888888888888889e-03 + z * p;              /*  1/6!  */
	p = -4.16666666666666666667e-02 + z * p;             /* -1/4!  */
	p = 5.00000000000000000000e-01 + z * p;              /*  1/2!  (sign applied below) */
	return 1.0 - z * p;
}

PM_HD void sincos_d(double x, double* s, double* c) {
	if (!(fabs(x) < 1.0e15)) { *s = *c = (double) PM_NAN_F; return; }
	double r;
	const int q = rem_pio2(x, &r);
	const double sr = ksin(r), cr = kcos(r);
	switch (q) {
		case 0: *s = sr; *c = cr; break;
		case 1: *s = cr; *c = -sr; break;
		case 2: *s = -sr; *c = -cr; break;
		default: *s = -cr; *c = sr; break;
	}
}

/* atan on binary64: |x|>1 -> pi/2 - atan(1/x); |x|>tan(pi/8) -> pi/4 + atan((x-1)/(x+1));
 * then the Maclaurin series to x^27. */
PM_HD double atan_d(double x) {
	if (x != x) return x;
	const bool neg = x < 0.0;
	double a = neg ? -x : x;
	const bool inv = a > 1.0;
	if (inv) a = 1.0 / a;
	const bool shift = a > 4.14213562373095048802e-01;
	if (shift) a = (a - 1.0) / (a + 1.0);
	const double z = a * a;
	double p = -1.0 / 27.0;
	p = 1.0 / 25.0 + z * p;
	p = -1.0 / 23.0 + z * p;
	p = 1.0 / 21.0 + z * p;
	p = -1.0 / 19.0 + z * p;
	p = 1.0 / 17.0 + z * p;
	p = -1.0 / 15.0 + z * p;
	p = 1.0 / 13.0 + z * p;
	p = -1.0 / 11.0 + z * p;
	p = 1.0 / 9.0 + z * p;
	p = -1.0 / 7.0 + z * p;
	p = 1.0 / 5.0 + z * p;
	p = -1.0 / 3.0 + z * p;
	double r = a + a * (z * p);
	if (shift) r = 7.85398163397448309616e-01 + r;
	if (inv) r = 1.57079632679489661923e+00 - r;
	return neg ? -r : r;
}

/* log2 of a finite, positive, normal binary64. */
PM_HD double log2_d(double a) {
	uint64_t b = d2bits(a);
	int e = (int) ((b >> 52) & 0x7ff) - 1023;
	double m = bits2d((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);   /* [1,2) */
	if (m > 1.41421356237309504880e+00) { m = m * 0.5; e += 1; }
	const double s = (m - 1.0) / (m + 1.0);
	const double z = s * s;
	double p = 1.0 / 21.0;
	p = 1.0 / 19.0 + z * p;
	p = 1.0 / 17.0 + z * p;
	p = 1.0 / 15.0 + z * p;
	p = 1.0 / 13.0 + z * p;
	p = 1.0 / 11.0 + z * p;
	p = 1.0 / 9.0 + z * p;
	p = 1.0 / 7.0 + z * p;
	p = 1.0 / 5.0 + z * p;
	p = 1.0 / 3.0 + z * p;
	const double lnm = 2.0 * (s + s * (z * p));
	return (double) e + lnm * 1.44269504088896340736e+00;
}

/* 2^z on binary64; results outside the binary32 range collapse to inf / 0. */
PM_HD double exp2_d(double z) {
	if (z != z) return z;
	if (z > 1100.0) return (double) PM_INF_F;
	if (z < -1100.0) return 0.0;
	const double fn = rint(z);
	const double t = (z - fn) * 6.93147180559945309417e-01;
	double p = 1.0 / 6227020800.0;                       /* 1/13! */
	p = 1.0 / 479001600.0 + t * p;
	p = 1.0 / 39916800.0 + t * p;
	p = 1.0 / 3628800.0 + t * p;
	p = 1.0 / 362880.0 + t * p;
	p = 1.0 / 40320.0 + t * p;
	p = 1.0 / 5040.0 + t * p;
	p = 1.0 / 720.0 + t * p;
	p = 1.0 / 120.0 + t * p;
	p = 1.0 / 24.0 + t * p;
	p = 1.0 / 6.0 + t * p;
	p = 0.5 + t * p;
	p = 1.0 + t * p;
	p = 1.0 + t * p;
	const int n = (int) fn;
	if (n > 1023) return (double) PM_INF_F;
	if (n < -1022) return 0.0;
	return p * bits2d((uint64_t) (n + 1023) << 52);
}

/* ------------------------------------------- binary32 transcendental built-ins */

/* native_sin / native_cos / native_tan / sin / cos. */
PM_HD float sin_(float x) { double s, c; sincos_d((double) x, &s, &c); return (float) s; }
PM_HD float cos_(float x) { double s, c; sincos_d((double) x, &s, &c); return (float) c; }
PM_HD float tan_(float x) { double s, c; sincos_d((double) x, &s, &c); return (float) (s / c); }
/* sin / cos / tan of a binary64 argument (OpenCL promotes `M_PI_2 * a` etc. to double). */
PM_HD float sin_d2f(double x) { double s, c; sincos_d(x, &s, &c); return (float) s; }
PM_HD float cos_d2f(double x) { double s, c; sincos_d(x, &s, &c); return (float) c; }
PM_HD double tan_dd(double x) { double s, c; sincos_d(x, &s, &c); return s / c; }

PM_HD double atan_dd(double x) { return atan_d(x); }
PM_HD float atan_(float x) { return (float) atan_d((double) x); }

/* acos(x) = 2 atan( sqrt((1-x)/(1+x)) ); NaN outside [-1,1]. */
PM_HD double acos_dd(double x) { return 2.0 * atan_d(sqrt((1.0 - x) / (1.0 + x))); }
PM_HD float acos_(float x) { return (float) acos_dd((double) x); }

/* pow(x, y), IEEE-754 special cases for the combinations the renderer can produce. */
PM_HD double pow_dd(double x, double y) {
	if (y == 0.0 || x == 1.0) return 1.0;
	if (x != x || y != y) return (double) PM_NAN_F;
	double ax = fabs(x);
	bool negResult = false;
	if (x < 0.0) {
		const bool yIsInt = (fabs(y) >= 9007199254740992.0) || (floor(y) == y);
		if (!yIsInt && ax != (double) PM_INF_F && fabs(y) != (double) PM_INF_F) return (double) PM_NAN_F;
		if (yIsInt && fabs(y) < 9007199254740992.0) {
			const double h = y * 0.5;
			negResult = (floor(h) != h);
		}
	}
	double r;
	if (fabs(y) == (double) PM_INF_F) {
		if (ax == 1.0) return 1.0;
		r = ((ax < 1.0) == (y > 0.0)) ? 0.0 : (double) PM_INF_F;
		return r;
	}
	if (ax == 0.0) r = (y > 0.0) ? 0.0 : (double) PM_INF_F;
	else if (ax == (double) PM_INF_F) r = (y > 0.0) ? (double) PM_INF_F : 0.0;
	else {
		/* binary64 subnormals cannot come from a binary32 input; scale anyway to stay total. */
		double e0 = 0.0;
		if (ax < 2.2250738585072014e-308) { ax = ax * 18014398509481984.0; e0 = -54.0; }
		r = exp2_d(y * (log2_d(ax) + e0));
	}
	return negResult ? -r : r;
}
PM_HD float pow_(float x, float y) { return (float) pow_dd((double) x, (double) y); }

PM_HD float cbrt_(float x) {
	if (x == 0.0f || x != x) return x;
	const double ax = fabs((double) x);
	if (ax == (double) PM_INF_F) return x;
	const double r = exp2_d(log2_d(ax) / 3.0);
	return (float) ((x < 0.0f) ? -r : r);
}

/* ---------------------------------------------------------------- 3-vectors */

struct vec3 {
	float x, y, z;
};

PM_HD vec3 v3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
PM_HD vec3 v3s(float s) { return v3(s, s, s); }
PM_HD vec3 operator+(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PM_HD vec3 operator-(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PM_HD vec3 operator*(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
PM_HD vec3 operator*(vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
PM_HD vec3 operator*(float s, vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
PM_HD vec3 operator-(vec3 a) { return v3(-a.x, -a.y, -a.z); }
PM_HD vec3 yzx(vec3 a) { return v3(a.y, a.z, a.x); }

/* dot / cross with a fixed evaluation order and no contraction (Appendix D). */
PM_HD float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PM_HD vec3 cross(vec3 a, vec3 b) {
	return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* fast_normalize(v) = v * (1 / sqrt(dot(v,v))). */
PM_HD vec3 normalize(vec3 a) {
	const float s = 1.0f / sqrtf(dot(a, a));
	return v3(a.x * s, a.y * s, a.z * s);
}
PM_HD float length(vec3 a) { return sqrtf(dot(a, a)); }
/* fma(a, b, c) per component -- the only fused operations (reference: pt_intersect.cl:97,
 * pathtracing.cl:189, pt_brdf.cl:349, pt_utils.cl:370). */
PM_HD vec3 fma3(vec3 a, float s, vec3 c) { return v3(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z)); }

} /* namespace pm */

#endif /* PBR_PINNED_MATH_H */
