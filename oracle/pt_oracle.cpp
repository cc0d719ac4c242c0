/*
 * pt_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, one-pixel-at-a-time restatement of the reference's OpenCL megakernel
 * `pathTracing` and every helper it splices in:
 *     source/opencl/pathtracing.cl, pt_header.cl, pt_utils.cl, pt_rgb.cl, pt_brdf.cl,
 *     pt_intersect.cl, pt_bvh.cl, pt_phongtess.cl
 * Each function below cites the reference lines it follows.  The compile-time macros the
 * reference splices in as text (CL.cpp:626-705) are read at run time from `pbr_defines`.
 * Implementation-defined OpenCL built-ins take the single meaning fixed in
 * include/pbr_pinned_math.h; everything else keeps the reference's
 * expression order, with -ffp-contract=off so nothing is fused that the reference does not fuse.
 *
 * PARITY STATUS: PINNED against outputs of the reference itself.  The reference ships no tests,
 * golden vectors or known answers for this path and its host program cannot be built here (no
 * OpenCL ICD, Boost, GLM, Qt) -- but its KERNEL SOURCE can: oracle/build_ref.py assembles the
 * program text the way CL::combineParts / CL::setValues do and compiles it for the host behind
 * oracle/ref_shim/cl_compat.h (OpenCL built-ins = the same pinned arithmetic), into oracle/_ref/.
 * tests/test_oracle_vs_reference.py requires this restatement to equal that build bit for bit:
 * image and debug image for 12 configurations (both BRDFs, shadow rays, SAMPLES > 1, depth of
 * field, Phong tessellation, extended depth, three scenes) and t / hitFace / node visits / face
 * tests for explicit closest-hit and shadow rays.  Also pinned: the soft known-answer of
 * pathtracing.cl:75-76 (suzanne.obj: 1082 faces) and brute-force cross-checks in tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (csrc/, host/) never links, imports or calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

#include "../include/pbr_pinned_math.h"
#include "../include/pbr_types.h"

using pm::vec3;
using pm::v3;

namespace {

#define EPSILON5 0.00001f
#define NI_AIR 1.00028f
#define PI_X2 6.28318530718f
/* OpenCL C: M_PI, M_PI_2, M_1_PI are binary64 constants. */
#define CL_M_PI 3.14159265358979323846
#define CL_M_PI_2 1.57079632679489661923
#define CL_M_1_PI 0.31830988618379067154
#define INF_F PM_INF_F

struct vec4 {
	float x, y, z, w;
};
inline vec4 v4(float x, float y, float z, float w) { vec4 r = {x, y, z, w}; return r; }
inline vec4 v4s(float s) { return v4(s, s, s, s); }
inline vec4 operator+(vec4 a, vec4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return v4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(vec4 a, float s) { return v4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, vec4 a) { return v4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline vec4 operator+(vec4 a, float s) { return v4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline vec4 operator/(vec4 a, float s) { return v4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4 from4(const pbr_float4& f) { return v4(f.x, f.y, f.z, f.w); }
inline vec3 xyz(const pbr_float4& f) { return v3(f.x, f.y, f.z); }
inline vec4 clamp4(vec4 a, float lo, float hi) {
	return v4(pm::clamp_(a.x, lo, hi), pm::clamp_(a.y, lo, hi), pm::clamp_(a.z, lo, hi), pm::clamp_(a.w, lo, hi));
}

/* pt_header.cl:24-30.  hitLeaf is an addition for the explicit-ray API (C5). */
struct ray4 {
	vec3 origin;
	vec3 dir;
	vec3 normal;
	float t;
	int hitFace;
	int hitLeaf;
};

/* pt_header.cl:33-38 */
struct rayPlanes {
	vec3 n1, n2;
	float o1, o2;
};

/* pt_header.cl:86-109: both material layouts widened to one struct. */
struct material {
	float data[8];
	vec4 rgbDiff;
	vec4 rgbSpec;
};

struct Stats {
	uint64_t traverseCalls, shadowCalls, nodeVisits, triTests, shadedHits, shadowNodeVisits;
};

/* pt_header.cl:70-78 */
struct Scene {
	const pbr_defines* D;
	const pbr_bvh_node* bvh;
	const pbr_light* lights;
	const pbr_uint4* facesV;
	const pbr_uint4* facesN;
	const pbr_float4* vertices;
	const pbr_float4* normals;
	vec4 debugColor;
	Stats* stats;
};

/* ------------------------------------------------------------------ pt_utils.cl */

/* pt_utils.cl:39-44 */
inline float rand_(float* seed) {
	*seed += 1.0f;
	return pm::fract_(pm::sin_(*seed) * 43758.5453123f);
}

/* pt_utils.cl:53-56 */
inline float fresnel(const float u, const float c) {
	const float v = 1.0f - u;
	return c + (1.0f - c) * v * v * v * v * v;
}

/* pt_utils.cl:65-68 */
inline vec4 fresnel4(const float u, const vec4 c) {
	const float v = 1.0f - u;
	const vec4 omc = v4(1.0f - c.x, 1.0f - c.y, 1.0f - c.z, 1.0f - c.w);
	return c + omc * v * v * v * v * v;
}

/* pt_utils.cl:76-80 */
inline void swap_(float* a, float* b) {
	const float tmp = *a;
	*a = *b;
	*b = tmp;
}

/* pt_utils.cl:89-96 */
inline bool extendDepth(const pbr_defines* D, const material* mtl, float* seed) {
	if (D->brdf == 1) {
		return (fmaxf(mtl->data[2], mtl->data[3]) >= 50.0f);
	}
	return (mtl->data[3] < rand_(seed));
}

/* pt_utils.cl:108-199 */
char solveCubic(const float a0, const float a1, const float a2, const float a3, float x[3]) {
	const float THIRD = 0.3333333333f;
	const float THIRD_HALF = 0.1666666666f;
	float w, p, q, dis, phi;

	if (fabsf(a0) > 0.0f) {
		w = pm::divide(a1, a0) * THIRD;
		p = pm::divide(a2, a0) * THIRD - w * w;
		p = p * p * p;
		q = 0.5f * pm::divide(a2 * w - a3, a0) - w * w * w;
		dis = q * q + p;

		if (dis < 0.0f) {
			phi = pm::acos_(pm::clamp_(pm::divide(q, pm::sqrt_(-p)), -1.0f, 1.0f));
			p = 2.0f * pm::pow_(-p, THIRD_HALF);

			const float u[3] = {
				p * pm::cos_(phi * THIRD) - w,
				p * pm::cos_((float) (((double) phi + 2.0f * CL_M_PI) * (double) THIRD)) - w,
				p * pm::cos_((float) (((double) phi + 4.0f * CL_M_PI) * (double) THIRD)) - w
			};

			x[0] = fminf(u[0], fminf(u[1], u[2]));
			x[1] = fmaxf(fminf(u[0], u[1]), fmaxf(fminf(u[0], u[2]), fminf(u[1], u[2])));
			x[2] = fmaxf(u[0], fmaxf(u[1], u[2]));
			for (int i = 0; i < 3; i++) {
				x[i] -= pm::divide(
					a3 + x[i] * (a2 + x[i] * (a1 + x[i] * a0)),
					a2 + x[i] * (2.0f * a1 + x[i] * 3.0f * a0)
				);
			}
			return 3;
		}
		else {
			dis = pm::sqrt_(dis);
			x[0] = pm::cbrt_(q + dis) + pm::cbrt_(q - dis) - w;
			x[0] -= pm::divide(
				a3 + x[0] * (a2 + x[0] * (a1 + x[0] * a0)),
				a2 + x[0] * (2.0f * a1 + x[0] * 3.0f * a0)
			);
			return 1;
		}
	}
	else if (fabsf(a1) > 0.0f) {
		p = 0.5f * pm::divide(a2, a1);
		dis = p * p - pm::divide(a3, a1);

		if (dis >= 0.0f) {
			const float dis_sqrt = pm::sqrt_(dis);
			x[0] = -p - dis_sqrt;
			x[1] = -p + dis_sqrt;
			x[0] -= pm::divide(a3 + x[0] * (a2 + x[0] * a1), a2 + x[0] * 2.0f * a1);
			x[1] -= pm::divide(a3 + x[1] * (a2 + x[1] * a1), a2 + x[1] * 2.0f * a1);
			return 2;
		}
	}
	else if (fabsf(a2) > 0.0f) {
		x[0] = pm::divide(-a3, a2);
		return 1;
	}
	return 0;
}

/* pt_utils.cl:208-218 */
rayPlanes getPlanesFromRay(const ray4* ray) {
	rayPlanes rp;
	rp.n1 = pm::normalize(pm::cross(ray->origin, ray->dir));
	rp.n2 = pm::normalize(pm::cross(rp.n1, ray->dir));
	rp.o1 = pm::dot(rp.n1, ray->origin);
	rp.o2 = pm::dot(rp.n2, ray->origin);
	return rp;
}

/* pt_utils.cl:231 */
inline vec3 getTriangleNormal(vec3 an, vec3 bn, vec3 cn, float u, float v, float w) {
	return pm::normalize(an * u + bn * v + cn * w);
}

/* pt_utils.cl:246-254 */
inline vec3 getTriangleNormalS(
	const float u, const float v, const float w,
	const vec3 C12, const vec3 C23, const vec3 C31, const vec3 E23, const vec3 E31
) {
	const vec3 du = (w - u) * C31 + v * (C12 - C23) + E31;
	const vec3 dv = (w - v) * C23 + u * (C12 - C31) - E23;
	return pm::normalize(pm::cross(du, dv));
}

/* pt_utils.cl:263-265 */
inline vec3 getTriangleReflectionVec(const vec3 view, const vec3 np) {
	return view - 2.0f * np * pm::dot(view, np);
}

/* pt_utils.cl:282-294 */
inline vec3 getPhongTessNormal(
	const vec3 an, const vec3 bn, const vec3 cn, const vec3 rayDir,
	const float u, const float v, const float w,
	const vec3 C1, const vec3 C2, const vec3 C3, const vec3 E12, const vec3 E20
) {
	const vec3 ns = getTriangleNormalS(u, v, w, C1, C2, C3, E12, E20);
	const vec3 np = getTriangleNormal(an, bn, cn, u, v, w);
	const vec3 r = getTriangleReflectionVec(rayDir, np);
	return (pm::dot(ns, r) < 0.0f) ? ns : np;
}

/* pt_utils.cl:306-318 */
vec3 jitter(const vec3 nl, const float phi, const float sina, const float cosa) {
	const vec3 u = pm::normalize(pm::cross(pm::yzx(nl), nl));
	const vec3 v = pm::normalize(pm::cross(nl, u));
	return pm::normalize(
		pm::normalize(u * pm::cos_(phi) + v * pm::sin_(phi)) * sina + nl * cosa
	);
}

/* pt_utils.cl:327-337 */
void antiAliasing(const pbr_defines* D, ray4* ray, const float pxDim, float* seed) {
	const float rnd = rand_(seed);
	const float phi = PI_X2 * rand_(seed);
	const vec3 aaDir = jitter(ray->dir, phi, pm::sqrt_(rnd), pm::sqrt_(1.0f - rnd));
	ray->dir = pm::normalize(ray->dir + aaDir * pxDim * D->anti_aliasing);
}

/* pt_utils.cl:349-373 */
void depthOfField(ray4* ray, const pbr_camera* cam, float tObject, float tFocus, float* seed) {
	if (tObject == INF_F) {
		tObject = 1000.0f;
	}
	if (tFocus == INF_F) {
		tFocus = 1000.0f;
	}
	if (tObject > 0.0f) {
		const float aperture = cam->lense.x / cam->lense.y;
		const float radius = rand_(seed) * aperture * 0.5f;
		const float angle = PI_X2 * rand_(seed);
		const float x = radius * pm::cos_(angle);
		const float y = radius * pm::sin_(angle);

		ray->origin = ray->origin + x * xyz(cam->u) + y * xyz(cam->v);

		const vec3 hitFocalPlane = pm::fma3(ray->dir, tFocus, xyz(cam->eye));
		ray->dir = pm::normalize(hitFocalPlane - ray->origin);
	}
}

/* pt_utils.cl:385-387 */
inline bool russianRoulette(const int depth, const int depthAdded, const float maxValColor, float* seed) {
	return (depth > 2 + depthAdded && maxValColor < rand_(seed));
}

/* pt_utils.cl:397-399 */
inline vec3 projectOnPlane(const vec3 q, const vec3 p, const vec3 n) {
	return q - pm::dot(q - p, n) * n;
}

/* pt_utils.cl:408 */
inline float lambert(vec3 n, vec3 l) { return fmaxf(pm::dot(n, l), 0.0f); }

/* pt_utils.cl:426 */
inline vec3 reflect(vec3 dir, vec3 normal) { return dir - 2.0f * pm::dot(normal, dir) * normal; }

/* pt_utils.cl:436-465 */
vec3 refract(const ray4* ray, const material* mtl, float* seed) {
	const bool into = (pm::dot(ray->normal, -ray->dir) > 0.0f);
	const vec3 nl = into ? ray->normal : -ray->normal;

	const float m1 = into ? NI_AIR : mtl->data[1];
	const float m2 = into ? mtl->data[1] : NI_AIR;
	const float m = pm::divide(m1, m2);

	const float cosI = -pm::dot(nl, ray->dir);
	const float sinT2 = m * m * (1.0f - cosI * cosI);

	if (sinT2 >= 1.0f) {
		return reflect(ray->dir, nl);
	}

	const float sqrtCosT = pm::sqrt_(1.0f - sinT2);
	const float r0 = pm::divide(m1 - m2, m1 + m2);
	const float c = (m1 > m2) ? sqrtCosT : cosI;
	const float reflectance = fresnel(c, r0 * r0);

	const vec3 newDir = (reflectance < rand_(seed)) ?
		m * ray->dir + (m * cosI - sqrtCosT) * nl :
		reflect(ray->dir, nl);

	return newDir;
}

/* ------------------------------------------------------------------ pt_brdf.cl */

/* pt_brdf.cl:11-14 */
inline float Z(const float t, const float r) {
	const float x = 1.0f + r * t * t - t * t;
	return (x == 0.0f) ? 0.0f : pm::divide(r, x * x);
}

/* pt_brdf.cl:23-28 */
inline float A(const float w, const float p) {
	const float p2 = p * p;
	const float w2 = w * w;
	const float x = p2 - p2 * w2 + w2;
	return (x == 0.0f) ? 0.0f : pm::sqrt_(pm::divide(p, x));
}

/* pt_brdf.cl:37-40 */
inline float G(const float v, const float r) {
	const float x = r - r * v + v;
	return (x == 0.0f) ? 0.0f : pm::divide(v, x);
}

/* pt_brdf.cl:71-80 */
inline float B2(const float t, const float vOut, const float vIn, const float w, const float r, const float p) {
	const float gp = G(vOut, r) * G(vIn, r);
	const float obstructed = gp * Z(t, r) * A(w, p);
	const float reemission = 1.0f - gp;
	return obstructed + reemission;
}

/* pt_brdf.cl:93-112 */
inline float Dfac(const float t, const float vOut, const float vIn, const float w, const float r, const float p) {
	const float b = 4.0f * r * (1.0f - r);
	const float a = (r < 0.5f) ? 0.0f : 1.0f - b;
	const float c = (r < 0.5f) ? 1.0f - b : 0.0f;

	const float d = (float) (4.0f * CL_M_PI * (double) vOut * (double) vIn);

	const float lam = (float) ((double) a * CL_M_1_PI);
	const float ani = (b == 0.0f || d == 0.0f)
		? 0.0f
		: pm::divide(b, d) * B2(t, vOut, vIn, w, r, p);
	const float fres = (vIn == 0.0f)
		? 0.0f
		: pm::divide(c, vIn);

	return lam + ani + fres;
}

/* pt_brdf.cl:125-149 */
float brdfSchlick(
	const material* mtl, const ray4* rayLightOut, const ray4* rayLightIn,
	const vec3* normal, float* u, float* pdf
) {
	const vec3 V_IN = rayLightIn->dir;
	const vec3 V_OUT = -rayLightOut->dir;

	const vec3 un = pm::normalize(pm::cross(pm::yzx(*normal), *normal));

	const vec3 h = pm::normalize(V_OUT + V_IN);
	const float t = pm::dot(h, *normal);
	const float vIn = pm::dot(V_IN, *normal);
	const float vOut = pm::dot(V_OUT, *normal);
	const vec3 hp = pm::normalize(pm::cross(pm::cross(h, *normal), *normal));
	const float w = pm::dot(un, hp);

	*u = pm::dot(h, V_OUT);
	*pdf = pm::divide(t, (float) (4.0f * CL_M_PI * (double) pm::dot(V_OUT, h)));

	return Dfac(t, vOut, vIn, w, mtl->data[3], mtl->data[2]);
}

/* pt_brdf.cl:159-208 */
vec3 newRaySchlick(const ray4* ray, const material* mtl, float* seed) {
	vec3 newRay;

	if (mtl->data[3] == 0.0f) {
		return reflect(ray->dir, ray->normal);
	}

	float a = rand_(seed);
	float b = rand_(seed);
	float iso2 = mtl->data[2] * mtl->data[2];
	float alpha = pm::acos_(pm::sqrt_(pm::divide(a, mtl->data[3] - a * mtl->data[3] + a)));
	float phi;

	if (b < 0.25f) {
		b = 1.0f - 4.0f * (0.25f - b);
		const float b2 = b * b;
		phi = (float) (CL_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
	}
	else if (b < 0.5f) {
		b = 1.0f - 4.0f * (0.5f - b);
		const float b2 = b * b;
		phi = (float) (CL_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (CL_M_PI - (double) phi);
	}
	else if (b < 0.75f) {
		b = 1.0f - 4.0f * (0.75f - b);
		const float b2 = b * b;
		phi = (float) (CL_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (CL_M_PI + (double) phi);
	}
	else {
		b = 1.0f - 4.0f * (1.0f - b);
		const float b2 = b * b;
		phi = (float) (CL_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (2.0f * CL_M_PI - (double) phi);
	}

	if (mtl->data[2] < 1.0f) {
		phi = (float) ((double) phi + CL_M_PI_2);
	}

	vec3 H = jitter(ray->normal, phi, pm::sin_(alpha), pm::cos_(alpha));
	newRay = reflect(ray->dir, H);

	if (pm::dot(newRay, ray->normal) <= 0.0f) {
		const float phi2 = PI_X2 * rand_(seed);
		newRay = jitter(ray->normal, phi2, pm::sqrt_(a), pm::sqrt_(1.0f - a));
	}

	return newRay;
}

/* pt_brdf.cl:228-268 */
void brdfShirleyAshikhmin(
	const float nu, const float nv, const float Rs, const float Rd,
	const ray4* rayLightOut, const ray4* rayLightIn, const vec3* normal,
	float* brdfSpec, float* brdfDiff, float* dotHK1, float* pdf
) {
	(void) Rs;
	const vec3 un = pm::normalize(pm::cross(pm::yzx(*normal), *normal));
	const vec3 vn = pm::normalize(pm::cross(*normal, un));

	const vec3 k1 = rayLightIn->dir;
	const vec3 k2 = -rayLightOut->dir;
	const vec3 h = pm::normalize(k1 + k2);

	const float dotHU = pm::dot(h, un);
	const float dotHV = pm::dot(h, vn);
	const float dotHN = pm::dot(h, *normal);
	const float dotNK1 = pm::dot(*normal, k1);
	const float dotNK2 = pm::dot(*normal, k2);
	*dotHK1 = pm::dot(h, k1);

	float ps_e = nu * dotHU * dotHU + nv * dotHV * dotHV;
	ps_e = (dotHN == 1.0f) ? 0.0f : pm::divide(ps_e, 1.0f - dotHN * dotHN);
	const float ps0 = (float) ((double) (pm::sqrt_((nu + 1.0f) * (nv + 1.0f)) * 0.125f) * CL_M_1_PI);
	const float ps1_num = pm::pow_(dotHN, ps_e);
	const float ps1 = pm::divide(ps1_num, (*dotHK1) * fmaxf(dotNK1, dotNK2));

	float pd = Rd * 0.38750768752f;
	const float a = 1.0f - dotNK1 * 0.5f;
	const float b = 1.0f - dotNK2 * 0.5f;
	pd *= 1.0f - a * a * a * a * a;
	pd *= 1.0f - b * b * b * b * b;

	*brdfSpec = ps0 * ps1;
	*brdfDiff = pd;

	const float ph = ps0 * ps1_num;
	*pdf = pm::divide(ph, (*dotHK1));
}

/* pt_brdf.cl:278-330 */
vec3 newRayShirleyAshikhmin(const ray4* ray, const material* mtl, float* seed) {
	float a = rand_(seed);
	const float b = rand_(seed);
	float phi_flip = (float) CL_M_PI;
	float phi_flipf = 1.0f;
	float aMax = 1.0f;

	if (a < 0.25f) {
		aMax = 0.25f;
		phi_flip = 0.0f;
	}
	else if (a < 0.5f) {
		aMax = 0.5f;
		phi_flipf = -1.0f;
	}
	else if (a < 0.75f) {
		aMax = 0.75f;
	}
	else {
		phi_flip = (float) (2.0f * CL_M_PI);
		phi_flipf = -1.0f;
	}

	a = 1.0f - 4.0f * (aMax - a);

	const float phi = pm::atan_(
		pm::sqrt_(pm::divide(mtl->data[2] + 1.0f, mtl->data[3] + 1.0f)) *
		pm::tan_((float) (CL_M_PI_2 * (double) a))
	);
	const float phi_full = phi_flip + phi_flipf * phi;

	const float cosphi = pm::cos_(phi);
	const float sinphi = pm::sin_(phi);
	const float theta_e = pm::rcp(mtl->data[2] * cosphi * cosphi + mtl->data[3] * sinphi * sinphi + 1.0f);
	const float theta = pm::acos_(pm::pow_(1.0f - b, theta_e));

	const vec3 normal = (mtl->data[0] < 1.0f || pm::dot(ray->normal, -ray->dir) >= 0.0f) ? ray->normal : -ray->normal;

	const vec3 h = jitter(normal, phi_full, pm::sin_(theta), pm::cos_(theta));
	const vec3 spec = reflect(ray->dir, h);
	const float phi3 = PI_X2 * rand_(seed);
	const vec3 diff = jitter(normal, phi3, pm::sqrt_(b), pm::sqrt_(1.0f - b));

	const vec3 newRay = (pm::dot(spec, normal) <= 0.0f) ? diff : spec;
	return newRay;
}

/* pt_brdf.cl:344-378.  The reference leaves newRay.normal / newRay.hitFace uninitialised;
 * they are defined as 0 here. */
ray4 getNewRay(const pbr_defines* D, const ray4* ray, const material* mtl, float* seed, bool* addDepth) {
	ray4 newRay;
	newRay.t = INF_F;
	newRay.origin = pm::fma3(ray->dir, ray->t, ray->origin);
	newRay.normal = v3(0.0f, 0.0f, 0.0f);
	newRay.hitFace = 0;
	newRay.hitLeaf = -1;

	bool doTransRefr = (mtl->data[0] < 1.0f && mtl->data[0] <= rand_(seed));

	*addDepth = (*addDepth || doTransRefr);

	if (doTransRefr) {
		newRay.dir = refract(ray, mtl, seed);
	}
	else if (D->brdf == 0) {
		newRay.dir = newRaySchlick(ray, mtl, seed);
	}
	else {
		newRay.dir = newRayShirleyAshikhmin(ray, mtl, seed);
	}

	return newRay;
}

/* ------------------------------------------------------------- pt_phongtess.cl */

/* pt_phongtess.cl:14-27 */
vec3 phongTessellation(
	const float alpha,
	const vec3 P1, const vec3 P2, const vec3 P3,
	const vec3 N1, const vec3 N2, const vec3 N3,
	const float u, const float v, const float w
) {
	const vec3 pBary = P1 * u + P2 * v + P3 * w;
	const vec3 pTessellated =
		u * projectOnPlane(pBary, P1, N1) +
		v * projectOnPlane(pBary, P2, N2) +
		w * projectOnPlane(pBary, P3, N3);

	return (1.0f - alpha) * pBary + alpha * pTessellated;
}

/* pt_phongtess.cl:36-45 */
inline char getBestRayDomain(const vec3 rd) {
	const vec3 d = v3(fabsf(rd.x), fabsf(rd.y), fabsf(rd.z));
	char domain = (d.y > d.z) ? 1 : 2;
	if (d.x > d.y) {
		domain = (d.x > d.z) ? 0 : 2;
	}
	return domain;
}

/* pt_phongtess.cl:56-212 */
vec3 phongTessTriAndRayIntersect(
	const float ALPHA,
	const vec3 P1, const vec3 P2, const vec3 P3,
	const vec3 N1, const vec3 N2, const vec3 N3,
	const ray4* ray, float* t, const float tNear, const float tFar
) {
	vec3 normal = v3(0.0f, 0.0f, 0.0f);
	*t = INF_F;

	const vec3 E01 = P2 - P1;
	const vec3 E12 = P3 - P2;
	const vec3 E20 = P1 - P3;

	const vec3 C1 = ALPHA * (pm::dot(N2, E01) * N2 - pm::dot(N1, E01) * N1);
	const vec3 C2 = ALPHA * (pm::dot(N3, E12) * N3 - pm::dot(N2, E12) * N2);
	const vec3 C3 = ALPHA * (pm::dot(N1, E20) * N1 - pm::dot(N3, E20) * N3);

	float a, b, c, d, e, f, l, m, n, o, p, q;
	{
		const rayPlanes rp = getPlanesFromRay(ray);
		a = pm::dot(-rp.n1, C3);
		b = pm::dot(-rp.n1, C2);
		c = pm::dot(rp.n1, P3) - rp.o1;
		d = pm::dot(rp.n1, C1 - C2 - C3) * 0.5f;
		e = pm::dot(rp.n1, C3 + E20) * 0.5f;
		f = pm::dot(rp.n1, C2 - E12) * 0.5f;
		l = pm::dot(-rp.n2, C3);
		m = pm::dot(-rp.n2, C2);
		n = pm::dot(rp.n2, P3) - rp.o2;
		o = pm::dot(rp.n2, C1 - C2 - C3) * 0.5f;
		p = pm::dot(rp.n2, C3 + E20) * 0.5f;
		q = pm::dot(rp.n2, C2 - E12) * 0.5f;
	}

	float xs[3] = { -1.0f, -1.0f, -1.0f };
	char numCubicRoots = 0;
	{
		const float a3 = (l*m*n + 2.0f*o*p*q) - (l*q*q + m*p*p + n*o*o);
		const float a2 = (a*m*n + l*b*n + l*m*c + 2.0f*(d*p*q + o*e*q + o*p*f)) -
		                 (a*q*q + b*p*p + c*o*o + 2.0f*(l*f*q + m*e*p + n*d*o));
		const float a1 = (a*b*n + a*m*c + l*b*c + 2.0f*(o*e*f + d*e*q + d*p*f)) -
		                 (l*f*f + m*e*e + n*d*d + 2.0f*(a*f*q + b*e*p + c*d*o));
		const float a0 = (a*b*c + 2.0f*d*e*f) - (a*f*f + b*e*e + c*d*d);

		numCubicRoots = solveCubic(a0, a1, a2, a3, xs);
	}

	if (0 == numCubicRoots) {
		return normal;
	}

	float x = 0.0f;
	float determinant = INF_F;
	float mA, mB, mC, mD, mE, mF;

	for (char i = 0; i < numCubicRoots; i++) {
		mA = a * xs[(int) i] + l;
		mB = b * xs[(int) i] + m;
		mD = d * xs[(int) i] + o;
		const float tmp = mD * mD - mA * mB;

		x = (determinant > tmp) ? xs[(int) i] : x;
		determinant = fminf(determinant, tmp);
	}

	if (0.0f >= determinant) {
		return normal;
	}

	const char domain = getBestRayDomain(ray->dir);

	mA = a * x + l;
	mB = b * x + m;
	mC = c * x + n;
	mD = d * x + o;
	mE = e * x + p;
	mF = f * x + q;

	const bool AlessB = fabsf(mA) < fabsf(mB);

	const float mBorA = AlessB ? mB : mA;
	mA = pm::divide(mA, mBorA);
	mB = pm::divide(mB, mBorA);
	mC = pm::divide(mC, mBorA);
	mD = pm::divide(mD, mBorA);
	mE = pm::divide(mE, mBorA);
	mF = pm::divide(mF, mBorA);

	const float mAorB = AlessB ? mA : mB;
	const float mEorF = AlessB ? 2.0f * mE : 2.0f * mF;
	const float mForE = AlessB ? mF : mE;
	const float ab = AlessB ? a : b;
	const float ba = AlessB ? b : a;
	const float ef = AlessB ? e : f;
	const float fe = AlessB ? f : e;

	const float sqrtAorB = pm::sqrt_(mD * mD - mAorB);
	const float sqrtC = pm::sqrt_(mForE * mForE - mC);
	const float lab1 = mD + sqrtAorB;
	const float lab2 = mD - sqrtAorB;
	float lc1 = mForE + sqrtC;
	float lc2 = mForE - sqrtC;

	if (fabsf(mEorF - lab1 * lc1 - lab2 * lc2) < fabsf(mEorF - lab1 * lc2 - lab2 * lc1)) {
		swap_(&lc1, &lc2);
	}

	for (char loop = 0; loop < 2; loop++) {
		const float g = (0 == loop) ? -lab1 : -lab2;
		const float h = (0 == loop) ? -lc1 : -lc2;

		const float c0 = ab + g * (2.0f * d + ba * g);
		const float c1 = 2.0f * (h * (d + ba * g) + ef + fe * g);
		const float c2 = h * (ba * h + 2.0f * fe) + c;
		const char numResults = solveCubic(0.0f, c0, c1, c2, xs);

		for (char i = 0; i < numResults; i++) {
			float u = xs[(int) i];
			float v = g * u + h;
			const float w = 1.0f - u - v;

			if (u < 0.0f || v < 0.0f || w < 0.0f) {
				continue;
			}

			if (!AlessB) {
				swap_(&u, &v);
			}

			const vec3 pTessellated = phongTessellation(ALPHA, P1, P2, P3, N1, N2, N3, u, v, w) - ray->origin;
			const float pT[3] = { pTessellated.x, pTessellated.y, pTessellated.z };
			const float rD[3] = { ray->dir.x, ray->dir.y, ray->dir.z };
			const float tParam = pm::divide(pT[(int) domain], rD[(int) domain]);

			if (tParam >= fabsf(tNear) && tParam <= fminf(*t, fminf(ray->t, tFar))) {
				*t = tParam;
				normal = getPhongTessNormal(N1, N2, N3, ray->dir, u, v, w, C1, C2, C3, E12, E20);
			}
		}
	}

	return normal;
}

/* ------------------------------------------------------------- pt_intersect.cl */

/* pt_intersect.cl:11-25 */
inline bool intersectBox(
	const ray4* ray, const vec3* invDir,
	const pbr_float4 bbMin, const pbr_float4 bbMax,
	float* tNear, float* tFar
) {
	const vec3 t1 = (xyz(bbMin) - ray->origin) * (*invDir);
	vec3 tMax = (xyz(bbMax) - ray->origin) * (*invDir);
	const vec3 tMin = v3(fminf(t1.x, tMax.x), fminf(t1.y, tMax.y), fminf(t1.z, tMax.z));
	tMax = v3(fmaxf(t1.x, tMax.x), fmaxf(t1.y, tMax.y), fmaxf(t1.z, tMax.z));

	*tNear = fmaxf(fmaxf(tMin.x, tMin.y), tMin.z);
	*tFar = fminf(fminf(tMax.x, tMax.y), fminf(tMax.z, *tFar));

	return (*tNear <= *tFar);
}

/* pt_intersect.cl:37-77 */
inline bool intersectSphere(
	ray4* ray, const vec3 pos, const float r,
	float* tNear, float* tFar
) {
	float t0, t1;

	vec3 L = pos - ray->origin;
	float tca = pm::dot(L, ray->dir);

	if (tca < 0.0f) {
		return false;
	}

	float d2 = pm::dot(L, L) - tca * tca;

	if (d2 > r) {
		return false;
	}

	float thc = pm::sqrt_(r - d2);
	t0 = tca - thc;
	t1 = tca + thc;

	if (t0 > t1) {
		swap_(&t0, &t1);
	}

	if (t0 < 0.0f) {
		t0 = t1;
		if (t0 < 0.0f) {
			return false;
		}
	}

	*tNear = t0;
	*tFar = t1;

	return true;
}

/* pt_intersect.cl:92-129 */
vec3 flatTriAndRayIntersect(
	const vec3 a, const vec3 b, const vec3 c,
	const ray4* ray, float* t, const float tNear
) {
	const float f = fmaxf(0.0f, tNear - 0.001f);
	const vec3 closeOrigin = pm::fma3(ray->dir, f, ray->origin);
	const vec3 edge1 = b - a;
	const vec3 edge2 = c - a;
	const vec3 tVec = closeOrigin - a;
	const vec3 pVec = pm::cross(ray->dir, edge2);
	const vec3 qVec = pm::cross(tVec, edge1);
	const float invDet = pm::rcp(pm::dot(edge1, pVec));

	*t = pm::dot(edge2, qVec) * invDet;

	if (*t >= ray->t || *t < EPSILON5) {
		*t = INF_F;
		return v3(0.0f, 0.0f, 0.0f);
	}

	const float u = pm::dot(tVec, pVec) * invDet;
	const float v = pm::dot(ray->dir, qVec) * invDet;

	if (u + v > 1.0f || fminf(u, v) < 0.0f) {
		*t = INF_F;
		return v3(0.0f, 0.0f, 0.0f);
	}

	*t += f;

	return pm::normalize(pm::cross(edge1, edge2));
}

/* pt_intersect.cl:142-176 */
vec3 checkFaceIntersection(
	const Scene* scene, const ray4* ray, const int fIndex, float* t,
	const float tNear, const float tFar
) {
	const pbr_uint4 fv = scene->facesV[fIndex];
	const vec3 a = xyz(scene->vertices[fv.x]);
	const vec3 b = xyz(scene->vertices[fv.y]);
	const vec3 c = xyz(scene->vertices[fv.z]);

	if (scene->D->phongtess == 1) {
		const pbr_uint4 fn = scene->facesN[fIndex];
		const vec3 an = xyz(scene->normals[fn.x]);
		const vec3 bn = xyz(scene->normals[fn.y]);
		const vec3 cn = xyz(scene->normals[fn.z]);
		/* `( an == bn ) + ( bn == cn )` summed to -6: all three normals component-wise equal. */
		const bool allEqual =
			an.x == bn.x && an.y == bn.y && an.z == bn.z &&
			bn.x == cn.x && bn.y == cn.y && bn.z == cn.z;

		if (!allEqual) {
			return phongTessTriAndRayIntersect(scene->D->phongtess_alpha, a, b, c, an, bn, cn, ray, t, tNear, tFar);
		}
	}

	return flatTriAndRayIntersect(a, b, c, ray, t, tNear);
}

/* ------------------------------------------------------------------- pt_bvh.cl */

/* pt_bvh.cl:10-24 */
void intersectFace(
	Scene* scene, ray4* ray, const int faceIndex, float* t,
	const float tNear, float tFar, const int leafIndex
) {
	const vec3 normal = checkFaceIntersection(scene, ray, faceIndex, t, tNear, tFar);

	if (ray->t > *t) {
		ray->normal = normal;
		ray->hitFace = faceIndex;
		ray->hitLeaf = leafIndex;
		ray->t = *t;
	}

	scene->debugColor.x += 1.0f;
	scene->stats->triTests++;
}

/* pt_bvh.cl:35-46 */
void intersectFaces(Scene* scene, ray4* ray, const pbr_bvh_node* node, const float tNear, float tFar, const int leafIndex) {
	float t = INF_F;

	intersectFace(scene, ray, (int) node->bbMin.w, &t, tNear, tFar, leafIndex);

	if (node->bbMax.w == -1) {
		return;
	}

	intersectFace(scene, ray, (int) node->bbMax.w, &t, tNear, tFar, leafIndex);
}

/* pt_bvh.cl:54-74 */
void traverseLights(const Scene* scene, ray4* ray) {
	const int NUM_LIGHTS = scene->D->num_lights;
	if (NUM_LIGHTS > 0) {
		float tNear = 0.0f;
		float tFar = INF_F;

		for (int i = 0; i < NUM_LIGHTS; i++) {
			pbr_light light = scene->lights[i];

			if (light.data.x == 2) {
				if (
					intersectSphere(ray, xyz(light.pos), light.data.y, &tNear, &tFar) &&
					tNear < ray->t
				) {
					ray->t = INF_F;
					ray->hitFace = -(i + 1);
				}
			}
		}
	}
}

/* pt_bvh.cl:82-123 */
void traverse(Scene* scene, ray4* ray) {
	const int BVH_NUM_NODES = scene->D->bvh_num_nodes;
	const vec3 invDir = v3(pm::rcp(ray->dir.x), pm::rcp(ray->dir.y), pm::rcp(ray->dir.z));
	int index = 1;

	scene->stats->traverseCalls++;
	traverseLights(scene, ray);

	do {
		scene->debugColor.y += 1.0f;
		scene->stats->nodeVisits++;
		const pbr_bvh_node node = scene->bvh[index];
		int currentIndex = index;

		index = (node.bbMin.w <= -1.0f) ? (int) node.bbMax.w : currentIndex + 1;

		float tNear = 0.0f;
		float tFar = INF_F;

		bool isNodeHit = (
			intersectBox(ray, &invDir, node.bbMin, node.bbMax, &tNear, &tFar) &&
			tFar > EPSILON5 && ray->t > tNear
		);

		if (!isNodeHit) {
			continue;
		}

		index = currentIndex + 1;

		if (node.bbMin.w >= 0.0f) {
			intersectFaces(scene, ray, &node, tNear, tFar, currentIndex);
		}
	} while (index > 0 && index < BVH_NUM_NODES);
}

/* pt_bvh.cl:133-177 */
void traverseShadows(Scene* scene, ray4* ray) {
	const int BVH_NUM_NODES = scene->D->bvh_num_nodes;
	float tLight = ray->t;
	const vec3 invDir = v3(pm::rcp(ray->dir.x), pm::rcp(ray->dir.y), pm::rcp(ray->dir.z));
	int index = 1;

	scene->stats->shadowCalls++;
	traverseLights(scene, ray);

	do {
		scene->stats->shadowNodeVisits++;
		const pbr_bvh_node node = scene->bvh[index];
		int currentIndex = index;

		index = (node.bbMin.w <= -1.0f) ? (int) node.bbMax.w : currentIndex + 1;

		float tNear = 0.0f;
		float tFar = INF_F;

		bool isNodeHit = (
			intersectBox(ray, &invDir, node.bbMin, node.bbMax, &tNear, &tFar) &&
			tFar > EPSILON5
		);

		if (!isNodeHit) {
			continue;
		}

		index = currentIndex + 1;

		/* Dead in practice: the host never writes -2.0f. */
		if (node.bbMin.w == -2.0f) {
			index++;
		}

		if (node.bbMin.w >= 0.0f) {
			intersectFaces(scene, ray, &node, tNear, tFar, currentIndex);

			if (ray->t < tLight) {
				break;
			}
		}
	} while (index > 0 && index < BVH_NUM_NODES);
}

/* --------------------------------------------------------------- pathtracing.cl */

/* pathtracing.cl:25-48 */
ray4 initRay(
	const pbr_defines* D, const int px, const int py,
	const float pxDim, const pbr_camera* cam, float* seed, float tFocus, float tObject
) {
	const vec3 camU = xyz(cam->u), camV = xyz(cam->v), camW = xyz(cam->w);
	const float IMG_WIDTH = (float) D->img_width;
	const float IMG_HEIGHT = (float) D->img_height;

	const vec3 initialRay = camW + pxDim * 0.5f * (
		camU - IMG_WIDTH * camU + 2.0f * (float) px * camU +
		camV - IMG_HEIGHT * camV + 2.0f * (float) py * camV
	);

	ray4 ray;
	ray.t = INF_F;
	ray.origin = xyz(cam->eye);
	ray.dir = pm::normalize(initialRay);
	ray.normal = v3(0.0f, 0.0f, 0.0f);
	ray.hitFace = 0;
	ray.hitLeaf = -1;

	antiAliasing(D, &ray, pxDim, seed);

	if (tFocus >= 0.0f && tObject >= 0.0f) {
		depthOfField(&ray, cam, tObject, tFocus, seed);
	}

	return ray;
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* pathtracing.cl:58-65 (CLK_ADDRESS_CLAMP_TO_EDGE) */
void getPreviousFocus(const pbr_defines* D, const pbr_camera* cam, const float* imageIn, int px, int py, float* tObject, float* tFocus) {
	const int W = D->img_width, H = D->img_height;
	*tObject = imageIn[((size_t) py * W + px) * 4 + 3];
	const int fx = clampi(cam->focusPoint.x, 0, W - 1), fy = clampi(cam->focusPoint.y, 0, H - 1);
	*tFocus = imageIn[((size_t) fy * W + fx) * 4 + 3];
}

/* pathtracing.cl:92-178 */
void updateColor(
	const pbr_defines* D,
	const ray4* ray, const ray4* newRay, const material* mtl,
	const ray4* lightRay, const vec4 lightRaySource, uint32_t* secondaryPaths,
	vec4* color, vec4* finalColor
) {
	if (D->brdf == 0) {
		float brdf, pdf, u;

		if (D->shadow_rays == 1) {
			if (lightRaySource.x >= 0) {
				brdf = brdfSchlick(mtl, ray, lightRay, &(ray->normal), &u, &pdf);

				if (fabsf(pdf) > 0.00001f) {
					brdf *= lambert(ray->normal, lightRay->dir);
					brdf = pm::divide(brdf, pdf);

					*finalColor = *finalColor + *color * lightRaySource * mtl->rgbDiff *
						(fresnel4(u, mtl->rgbSpec) * brdf * mtl->data[0] + (1.0f - mtl->data[0]));

					*secondaryPaths += 1;
				}
			}
		}

		brdf = brdfSchlick(mtl, ray, newRay, &(ray->normal), &u, &pdf);
		brdf *= lambert(ray->normal, newRay->dir);
		brdf = pm::divide(brdf, pdf);

		*color = *color * (mtl->rgbDiff * (fresnel4(u, mtl->rgbSpec) * brdf * mtl->data[0] + (1.0f - mtl->data[0])));
	}
	else {
		float brdfDiff, brdfSpec, pdf;
		vec4 brdf_d, brdf_s;
		float dotHK1;

		if (D->shadow_rays == 1) {
			if (lightRaySource.x >= 0) {
				brdfShirleyAshikhmin(
					mtl->data[2], mtl->data[3], mtl->data[4], mtl->data[5],
					ray, lightRay, &(ray->normal), &brdfSpec, &brdfDiff, &dotHK1, &pdf
				);

				if (fabsf(pdf) > 0.00001f) {
					brdfSpec = pm::divide(brdfSpec, pdf);
					brdfDiff = pm::divide(brdfDiff, pdf);

					brdf_s = brdfSpec * mtl->rgbSpec * fresnel(dotHK1, mtl->data[4]);
					brdf_d = brdfDiff * mtl->rgbDiff * (1.0f - mtl->data[4]);

					vec4 brdfColor = (brdf_s + brdf_d) * mtl->data[0] + (1.0f - mtl->data[0]);
					float maxRGB = pm::max_(1.0f, pm::max_(brdfColor.x, pm::max_(brdfColor.y, brdfColor.z)));
					brdfColor = brdfColor / maxRGB;

					*finalColor = *finalColor + (clamp4(brdfColor, 0.0f, 1.0f) * lightRaySource * mtl->data[0] + (1.0f - mtl->data[0]));

					*secondaryPaths += 1;
				}
			}
		}

		brdfShirleyAshikhmin(
			mtl->data[2], mtl->data[3], mtl->data[4], mtl->data[5],
			ray, newRay, &(ray->normal), &brdfSpec, &brdfDiff, &dotHK1, &pdf
		);

		brdfSpec = pm::divide(brdfSpec, pdf);
		brdfDiff = pm::divide(brdfDiff, pdf);

		brdf_s = brdfSpec * mtl->rgbSpec * fresnel(dotHK1, mtl->data[4]);
		brdf_d = brdfDiff * mtl->rgbDiff * (1.0f - mtl->data[4]);

		vec4 brdfColor = (brdf_s + brdf_d) * mtl->data[0] + (1.0f - mtl->data[0]);
		float maxRGB = pm::max_(1.0f, pm::max_(brdfColor.x, pm::max_(brdfColor.y, brdfColor.z)));
		brdfColor = brdfColor / maxRGB;

		*color = *color * clamp4(brdfColor, 0.0f, 1.0f);
	}
}

/* pathtracing.cl:188-199 */
void shadowRayTest(Scene* scene, ray4* ray, ray4* lightRay, vec4* lightRaySource) {
	lightRay->origin = pm::fma3(ray->dir, ray->t, ray->origin);
	lightRay->dir = pm::normalize(xyz(scene->lights[0].pos) - lightRay->origin);
	float tLight = pm::length(xyz(scene->lights[0].pos) - lightRay->origin);
	lightRay->t = tLight;

	traverseShadows(scene, lightRay);

	if (lightRay->t >= tLight) {
		*lightRaySource = from4(scene->lights[0].rgb);
	}
}

/* Material fetch: `materials[scene.facesV[ray.hitFace].w]` (pathtracing.cl:268).  Faces parsed
 * before any matching `usemtl` carry index (uint)-1 (ObjParser.cpp:140,206), which reads out of
 * bounds in the reference; here any out-of-range index yields MtlParser::getEmptyMaterial()'s
 * defaults (MtlParser.cpp:11-36). */
material fetchMaterial(const pbr_defines* D, const void* materials, int numMaterials, uint32_t idx) {
	material m;
	memset(&m, 0, sizeof(m));
	if (idx >= (uint32_t) numMaterials) {
		m.data[0] = 1.0f;              /* d */
		m.data[1] = 1.0f;              /* Ni */
		if (D->brdf == 0) { m.data[2] = 1.0f; m.data[3] = 1.0f; }              /* p, rough */
		else { m.data[2] = 0.0f; m.data[3] = 0.0f; m.data[4] = 0.0f; m.data[5] = 1.0f; }   /* nu nv Rs Rd */
		m.rgbDiff = v4(1.0f, 1.0f, 1.0f, 0.0f);
		m.rgbSpec = v4(1.0f, 1.0f, 1.0f, 0.0f);
		return m;
	}
	if (D->brdf == 0) {
		const pbr_material_schlick* s = (const pbr_material_schlick*) materials + idx;
		m.data[0] = s->data.x; m.data[1] = s->data.y; m.data[2] = s->data.z; m.data[3] = s->data.w;
		m.rgbDiff = from4(s->rgbDiff);
		m.rgbSpec = from4(s->rgbSpec);
	}
	else {
		const pbr_material_sa* s = (const pbr_material_sa*) materials + idx;
		for (int i = 0; i < 8; i++) m.data[i] = s->data[i];
		m.rgbDiff = from4(s->rgbDiff);
		m.rgbSpec = from4(s->rgbSpec);
	}
	return m;
}

/* pathtracing.cl:207-334, one work-item. */
void pathTracingPixel(
	const pbr_defines* D, const int px, const int py,
	float seed, const float pixelWeight, const float pxDim, const pbr_camera cam,
	const pbr_bvh_node* bvh, const pbr_uint4* facesV, const pbr_uint4* facesN,
	const pbr_float4* vertices, const pbr_float4* normals,
	const void* materials, const int numMaterials, const pbr_light* lights,
	const float* imageIn, float* imageOut, float* imageDebug, Stats* stats
) {
	const int MAX_DEPTH = D->max_depth, MAX_ADDED_DEPTH = D->max_added_depth;
	const uint32_t SAMPLES = (uint32_t) D->samples;
	const vec4 SKY_LIGHT = from4(D->sky_light);

	vec4 finalColor = v4s(0.0f);

	Scene scene = { D, bvh, lights, facesV, facesN, vertices, normals, v4s(0.0f), stats };

	float focus = 0.0f;
	float prevFocusX = -1.0f, prevFocusY = -1.0f;   /* x: tObject, y: tFocus */

	if (cam.focusPoint.x >= 0 && cam.focusPoint.y >= 0) {
		getPreviousFocus(D, &cam, imageIn, px, py, &prevFocusX, &prevFocusY);
	}

	bool addDepth;
	uint32_t secondaryPaths = 1;

	for (uint32_t sample = 0; sample < SAMPLES; sample++) {
		vec4 color = v4s(1.0f);
		vec4 light = v4s(-1.0f);

		ray4 ray = initRay(D, px, py, pxDim, &cam, &seed, prevFocusY, prevFocusX);
		int depthAdded = 0;

		for (uint32_t depth = 0; depth < (uint32_t) (MAX_DEPTH + depthAdded); depth++) {
			traverse(&scene, &ray);

			focus = (sample + depth == 0) ? ray.t : focus;

			if (ray.t == INF_F) {
				light = (ray.hitFace < 0) ? from4(scene.lights[-(ray.hitFace + 1)].rgb) : SKY_LIGHT;
				break;
			}

			material mtl = fetchMaterial(D, materials, numMaterials, scene.facesV[ray.hitFace].w);
			stats->shadedHits++;

			addDepth = extendDepth(D, &mtl, &seed);

			if (mtl.data[0] == 1.0f && !addDepth && depth == (uint32_t) (MAX_DEPTH + depthAdded - 1)) {
				break;
			}

			seed += ray.t;

			vec4 lightRaySource = v4s(-1.0f);
			ray4 lightRay;
			memset(&lightRay, 0, sizeof(lightRay));
			lightRay.t = INF_F;

			if (D->shadow_rays == 1 && D->num_lights > 0) {
				if (mtl.data[0] > 0.0f) {
					shadowRayTest(&scene, &ray, &lightRay, &lightRaySource);
				}
			}

			ray4 newRay = getNewRay(D, &ray, &mtl, &seed, &addDepth);

			if (pm::dot(ray.normal, -ray.dir) <= 0.0f) {
				ray.normal = -ray.normal;
			}

			updateColor(D, &ray, &newRay, &mtl, &lightRay, lightRaySource, &secondaryPaths, &color, &finalColor);

			depthAdded += (addDepth && depthAdded < MAX_ADDED_DEPTH);

			float maxValColor = fmaxf(color.x, fmaxf(color.y, color.z));

			if (russianRoulette((int) depth, depthAdded, maxValColor, &seed)) {
				break;
			}

			ray = newRay;
		}

		if (light.x > -1.0f) {
			color = color * light;
			finalColor = finalColor + color;
		}
	}

	finalColor = finalColor / (float) secondaryPaths;

	if (SAMPLES > 1) {
		finalColor = finalColor / (float) SAMPLES;
	}

	/* setColors, pt_rgb.cl:9-21 */
	const size_t o = ((size_t) py * D->img_width + px) * 4;
	const pbr_float4 imagePixel = { imageIn[o], imageIn[o + 1], imageIn[o + 2], imageIn[o + 3] };
	imageOut[o + 0] = pm::mix_(finalColor.x, imagePixel.x, pixelWeight);
	imageOut[o + 1] = pm::mix_(finalColor.y, imagePixel.y, pixelWeight);
	imageOut[o + 2] = pm::mix_(finalColor.z, imagePixel.z, pixelWeight);
	imageOut[o + 3] = focus;

	/* writeDebugImage, pathtracing.cl:73-78 */
	if (imageDebug) {
		imageDebug[o + 0] = scene.debugColor.x / 1082.0f;
		imageDebug[o + 1] = scene.debugColor.y / 1265.0f;
		imageDebug[o + 2] = scene.debugColor.z;
		imageDebug[o + 3] = scene.debugColor.w;
	}
}

template <typename F>
void parallelFor(int n, int nthreads, F fn) {
	if (nthreads <= 1 || n <= 1) {
		fn(0, n, 0);
		return;
	}
	if (nthreads > n) nthreads = n;
	std::vector<std::thread> th;
	for (int k = 0; k < nthreads; k++) {
		th.emplace_back([=]() {
			/* interleaved blocks so threads get similar work */
			fn(k, n, nthreads);
		});
	}
	for (auto& t : th) t.join();
}

} /* namespace */

extern "C" {

/* One launch of the kernel `pathTracing` over rows [y0, y1) of the W x H frame.
 * stats (may be NULL) receives 6 counters: traverse calls, shadow-ray calls, BVH nodes visited
 * by traverse, triangle tests (both traversals), shaded hits, nodes visited by traverseShadows. */
void oracle_path_tracing(
	const pbr_defines* D, float seed, float pixelWeight, float pxDim, const pbr_camera* cam,
	const pbr_bvh_node* bvh, const pbr_uint4* facesV, const pbr_uint4* facesN,
	const pbr_float4* vertices, const pbr_float4* normals,
	const void* materials, int numMaterials, const pbr_light* lights,
	const float* imageIn, float* imageOut, float* imageDebug,
	int y0, int y1, int nthreads, uint64_t* stats
) {
	const int rows = y1 - y0;
	if (nthreads < 1) nthreads = 1;
	std::vector<Stats> perThread((size_t) nthreads);
	memset(perThread.data(), 0, sizeof(Stats) * perThread.size());

	parallelFor(rows, nthreads, [&](int k, int n, int stride) {
		Stats* st = &perThread[(size_t) k];
		const int step = stride == 0 ? 1 : stride;
		for (int r = (stride == 0 ? 0 : k); r < n; r += step) {
			const int py = y0 + r;
			for (int px = 0; px < D->img_width; px++) {
				pathTracingPixel(
					D, px, py, seed, pixelWeight, pxDim, *cam, bvh, facesV, facesN, vertices, normals,
					materials, numMaterials, lights, imageIn, imageOut, imageDebug, st
				);
			}
		}
	});

	if (stats) {
		for (int i = 0; i < 6; i++) stats[i] = 0;
		for (const Stats& s : perThread) {
			stats[0] += s.traverseCalls; stats[1] += s.shadowCalls; stats[2] += s.nodeVisits;
			stats[3] += s.triTests; stats[4] += s.shadedHits; stats[5] += s.shadowNodeVisits;
		}
	}
}

/* Explicit rays through traverse() (anyHit = 0) or traverseShadows() (anyHit = 1).
 * rays[i].dir.w is the initial ray.t. */
void oracle_trace(
	const pbr_defines* D,
	const pbr_bvh_node* bvh, const pbr_uint4* facesV, const pbr_uint4* facesN,
	const pbr_float4* vertices, const pbr_float4* normals, const pbr_light* lights,
	const pbr_ray* rays, int64_t n, int anyHit, pbr_hit* out, int nthreads, uint64_t* stats
) {
	if (nthreads < 1) nthreads = 1;
	std::vector<Stats> perThread((size_t) nthreads);
	memset(perThread.data(), 0, sizeof(Stats) * perThread.size());
	const int64_t chunk = 4096;
	const int nchunks = (int) ((n + chunk - 1) / chunk);

	parallelFor(nchunks, nthreads, [&](int k, int nc, int stride) {
		Stats* st = &perThread[(size_t) k];
		const int step = stride == 0 ? 1 : stride;
		for (int c = (stride == 0 ? 0 : k); c < nc; c += step) {
			const int64_t i1 = ((int64_t) c + 1) * chunk < n ? ((int64_t) c + 1) * chunk : n;
			for (int64_t i = (int64_t) c * chunk; i < i1; i++) {
				Scene scene = { D, bvh, lights, facesV, facesN, vertices, normals, v4s(0.0f), st };
				ray4 ray;
				ray.origin = xyz(rays[i].origin);
				ray.dir = xyz(rays[i].dir);
				ray.normal = v3(0.0f, 0.0f, 0.0f);
				ray.t = rays[i].dir.w;
				ray.hitFace = 0;
				ray.hitLeaf = -1;
				const uint64_t n0 = st->nodeVisits + st->shadowNodeVisits, t0 = st->triTests;
				if (anyHit) traverseShadows(&scene, &ray);
				else traverse(&scene, &ray);
				uint64_t nv = st->nodeVisits + st->shadowNodeVisits - n0, tt = st->triTests - t0;
				if (nv > 0xfffffu) nv = 0xfffffu;
				if (tt > 0xfffu) tt = 0xfffu;
				out[i].t = ray.t;
				out[i].hitFace = ray.hitFace;
				out[i].leaf = ray.hitLeaf;
				out[i].visits = (uint32_t) nv | ((uint32_t) tt << 20);
			}
		}
	});

	if (stats) {
		for (int i = 0; i < 6; i++) stats[i] = 0;
		for (const Stats& s : perThread) {
			stats[0] += s.traverseCalls; stats[1] += s.shadowCalls; stats[2] += s.nodeVisits;
			stats[3] += s.triTests; stats[4] += s.shadedHits; stats[5] += s.shadowNodeVisits;
		}
	}
}

/* Closest hit by testing every face with the reference's flat triangle test (tNear = 0), for
 * cross-checking the stackless traversal. */
void oracle_trace_bruteforce(
	const pbr_uint4* facesV, int64_t numFaces, const pbr_float4* vertices,
	const pbr_ray* rays, int64_t n, pbr_hit* out
) {
	for (int64_t i = 0; i < n; i++) {
		ray4 ray;
		ray.origin = xyz(rays[i].origin);
		ray.dir = xyz(rays[i].dir);
		ray.t = rays[i].dir.w;
		ray.hitFace = 0;
		ray.hitLeaf = -1;
		for (int64_t f = 0; f < numFaces; f++) {
			float t;
			const pbr_uint4 fv = facesV[f];
			flatTriAndRayIntersect(xyz(vertices[fv.x]), xyz(vertices[fv.y]), xyz(vertices[fv.z]), &ray, &t, 0.0f);
			if (ray.t > t) {
				ray.t = t;
				ray.hitFace = (int) f;
			}
		}
		out[i].t = ray.t;
		out[i].hitFace = ray.hitFace;
		out[i].leaf = -1;
		out[i].visits = 0;
	}
}

/* Analysis helper: how often each BVH node is visited by traverse() for a set of rays. */
void oracle_visit_histogram(
	const pbr_defines* D, const pbr_bvh_node* bvh, const pbr_ray* rays, int64_t n, uint32_t* counts
) {
	const int N = D->bvh_num_nodes;
	for (int64_t i = 0; i < n; i++) {
		const vec3 o = xyz(rays[i].origin), d = xyz(rays[i].dir);
		ray4 ray;
		ray.origin = o; ray.dir = d; ray.t = rays[i].dir.w;
		const vec3 invDir = v3(pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z));
		int index = 1;
		int prevLine = -1;
		do {
			counts[index]++;
			if ((index >> 2) != prevLine) { counts[0]++; prevLine = index >> 2; }   /* counts[0]: 128-byte line switches */
			const pbr_bvh_node node = bvh[index];
			const int cur = index;
			index = (node.bbMin.w <= -1.0f) ? (int) node.bbMax.w : cur + 1;
			float tNear = 0.0f, tFar = INF_F;
			if (!(intersectBox(&ray, &invDir, node.bbMin, node.bbMax, &tNear, &tFar) && tFar > EPSILON5)) continue;
			index = cur + 1;   /* no t pruning: upper bound of the visit pattern */
		} while (index > 0 && index < N);
	}
}

/* Analysis helper: the same histogram for the walk traverse() really does (t pruning and face tests included).
 * counts[0] receives the number of leaf visits whose box was hit. */
void oracle_visit_histogram_pruned(
	const pbr_defines* D, const pbr_bvh_node* bvh, const pbr_uint4* facesV, const pbr_uint4* facesN,
	const pbr_float4* vertices, const pbr_float4* normals, const pbr_ray* rays, int64_t n, uint32_t* counts
) {
	const int N = D->bvh_num_nodes;
	Stats st;
	memset(&st, 0, sizeof(st));
	for (int64_t i = 0; i < n; i++) {
		Scene scene = { D, bvh, nullptr, facesV, facesN, vertices, normals, v4s(0.0f), &st };
		ray4 ray;
		ray.origin = xyz(rays[i].origin);
		ray.dir = xyz(rays[i].dir);
		ray.normal = v3(0.0f, 0.0f, 0.0f);
		ray.t = rays[i].dir.w;
		ray.hitFace = 0;
		ray.hitLeaf = -1;
		const vec3 invDir = v3(pm::rcp(ray.dir.x), pm::rcp(ray.dir.y), pm::rcp(ray.dir.z));
		int index = 1;
		do {
			counts[index]++;
			const pbr_bvh_node node = bvh[index];
			const int cur = index;
			index = (node.bbMin.w <= -1.0f) ? (int) node.bbMax.w : cur + 1;
			float tNear = 0.0f, tFar = INF_F;
			if (!(intersectBox(&ray, &invDir, node.bbMin, node.bbMax, &tNear, &tFar) && tFar > EPSILON5 && ray.t > tNear)) continue;
			index = cur + 1;
			if (node.bbMin.w >= 0.0f) {
				counts[0]++;
				intersectFaces(&scene, &ray, &node, tNear, tFar, cur);
			}
		} while (index > 0 && index < N);
	}
}

/* Analysis helper: the node indices ray i visits, in order, into trace[i * cap ...] (at most cap; lengths[i] = how
 * many there were).  Same walk as oracle_visit_histogram_pruned; a face handed to the triangle test appears as
 * -(face + 1) behind its leaf. */
void oracle_visit_trace(
	const pbr_defines* D, const pbr_bvh_node* bvh, const pbr_uint4* facesV, const pbr_uint4* facesN,
	const pbr_float4* vertices, const pbr_float4* normals, const pbr_ray* rays, int64_t n, int32_t cap,
	int32_t* trace, int32_t* lengths
) {
	const int N = D->bvh_num_nodes;
	Stats st;
	memset(&st, 0, sizeof(st));
	for (int64_t i = 0; i < n; i++) {
		Scene scene = { D, bvh, nullptr, facesV, facesN, vertices, normals, v4s(0.0f), &st };
		ray4 ray;
		ray.origin = xyz(rays[i].origin);
		ray.dir = xyz(rays[i].dir);
		ray.normal = v3(0.0f, 0.0f, 0.0f);
		ray.t = rays[i].dir.w;
		ray.hitFace = 0;
		ray.hitLeaf = -1;
		const vec3 invDir = v3(pm::rcp(ray.dir.x), pm::rcp(ray.dir.y), pm::rcp(ray.dir.z));
		int index = 1;
		int32_t len = 0;
		do {
			if (len < cap) trace[i * cap + len] = index;
			len++;
			const pbr_bvh_node node = bvh[index];
			const int cur = index;
			index = (node.bbMin.w <= -1.0f) ? (int) node.bbMax.w : cur + 1;
			float tNear = 0.0f, tFar = INF_F;
			if (!(intersectBox(&ray, &invDir, node.bbMin, node.bbMax, &tNear, &tFar) && tFar > EPSILON5 && ray.t > tNear)) continue;
			index = cur + 1;
			if (node.bbMin.w >= 0.0f) {
				/* the faces this leaf hands to the triangle test, as -(face + 1) */
				if (len < cap) trace[i * cap + len] = -((int) node.bbMin.w + 1);
				len++;
				if (node.bbMax.w >= 0.0f) {
					if (len < cap) trace[i * cap + len] = -((int) node.bbMax.w + 1);
					len++;
				}
				intersectFaces(&scene, &ray, &node, tNear, tFar, cur);
			}
		} while (index > 0 && index < N);
		lengths[i] = len;
	}
}

/* Scalar entry points of the pinned math, for tests/test_pinned_math.py. */
float oracle_pm_sin(float x) { return pm::sin_(x); }
float oracle_pm_cos(float x) { return pm::cos_(x); }
float oracle_pm_tan(float x) { return pm::tan_(x); }
float oracle_pm_acos(float x) { return pm::acos_(x); }
float oracle_pm_atan(float x) { return pm::atan_(x); }
float oracle_pm_pow(float x, float y) { return pm::pow_(x, y); }
float oracle_pm_cbrt(float x) { return pm::cbrt_(x); }
float oracle_rand(float* seed) { return rand_(seed); }

} /* extern "C" */
