/*
 * pt_device.cuh -- device-side building blocks of the B200 path tracer.
 *
 * What the reference computes inside its one OpenCL megakernel (source/opencl/pt_*.cl) is cut here
 * into the pieces a wavefront pipeline needs:
 *     traverseClosest / traverseAny   <- traverse / traverseShadows   (pt_bvh.cl:82-177)
 *     beginSample                     <- initRay + antiAliasing + DoF  (pathtracing.cl:25-48, pt_utils.cl:327-373)
 *     bounce                          <- the body of the depth loop    (pathtracing.cl:258-317)
 *     endSample / finishPixel         <- pathtracing.cl:320-333, setColors (pt_rgb.cl:9-21)
 * Arithmetic follows include/pbr_pinned_math.h; compile with -fmad=false.
 *
 * Device data layout (built once per scene by the repack kernels in pbr_capi.cu from the
 * reference-layout buffers the host uploads; a lossless transform):
 *   nodes  : 32 B per BVH node, one 256-bit load.  xyz lanes as uploaded; the two w lanes hold
 *            INTEGER bit patterns instead of float-encoded indices:
 *              lo.w  >= 0: leaf, index of its first face;   < 0: inner node
 *              hi.w  leaf: index of its second face or -1;  inner: miss link or -1 (stop)
 *   tris   : 36 B per face in leaf order, in two arrays: one 256-bit load (a.xyz, edge1.xyz, edge2.xy) and one
 *            32-bit load (edge2.z), with edge1 = b - a, edge2 = c - a -- the reference fetches facesV[f] and
 *            then three dependent vertices (pt_intersect.cl:146-149).  The material index, which the walk never
 *            needs, is a third array read once per shaded hit.
 *   wide   : the 4-wide BVH of the ordered walk (wide_bvh.h, pt_wide.cuh): 128 B per node = 4 children x
 *            (box min, box max, ref, aux), four 256-bit loads per node visit.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbr_pinned_math.h"
#include "../../include/pbr_types.h"

namespace ptd {

using pm::vec3;
using pm::v3;

#define PT_EPSILON5 0.00001f
#define PT_NI_AIR 1.00028f
#define PT_PI_X2 6.28318530718f
#define PT_M_PI 3.14159265358979323846
#define PT_M_PI_2 1.57079632679489661923
#define PT_M_1_PI 0.31830988618379067154

struct SceneDev {
	const float4* nodes;          /* 2 x float4 per node */
	const float4* tris;           /* 2 x float4 (32 B) per face; PHONGTESS: 6 x float4 (a b c an bn cn) */
	const float* trisB;           /* + 4 B per face (edge2.z); unused with PHONGTESS */
	const uint32_t* triMat;       /* material index per face, read once per shaded hit; unused with PHONGTESS */
	const float4* wide;           /* ordered walk: 8 x float4 per wide node (NULL: not built) */
	int wideTop;                  /* wide nodes [0, wideTop) are staged in shared memory by the wide kernels */
	const int* faceLeaf;          /* ordered walk: face -> index of its leaf in the reference's node array */
	const pbr_light* lights;
	float phongAlpha;             /* PHONGTESS_ALPHA */
	int numNodes;
	int numLights;
	int nodePhaseMin;             /* traversal engine: leave the node phase below this many stepping lanes */
	int refillMin;                /* traversal engine: claim new rays once this many lanes are idle */
};

struct Material {                 /* both reference layouts, widened */
	float d, Ni, a2, a3, Rs, Rd;  /* a2/a3: p/rough (BRDF 0) or nu/nv (BRDF 1) */
	vec3 rgbDiff, rgbSpec;
};

#define PT_MAX_BATCH 1            /* frames per launch (the in-kernel batch of round 1 measured slower and is gone) */

struct FrameParams {
	SceneDev scene;
	const void* materials;
	int numMaterials;
	pbr_camera cam;
	float seed, pixelWeight, pxDim;
	int width, height;            /* IMG_WIDTH / IMG_HEIGHT */
	int y0, y1;                   /* rows rendered by this launch */
	/* interleaved row sharding: stripeRows > 0 means local row r (y0 = 0, y1 = HEIGHT / stripeWorld) is image row
	 * (r / stripeRows) * stripeRows * stripeWorld + stripeRank * stripeRows + r % stripeRows */
	int stripeRows, stripeWorld, stripeRank;
	int maxDepth, maxAddedDepth, samples;
	float antiAliasing;
	float4 skyLight;
	const float4* imageIn;
	float4* imageOut;
	float4* imageDebug;           /* may be NULL */
	float4* frameOut;             /* != NULL (frames in flight): a finished pixel leaves (its frame's radiance, focus) here
	                                 and mixFrameKernel folds it into the accumulation image later, in frame order */
	unsigned long long* stats;    /* 6 counters, see pbr_stats */
	/* A launch may cover several consecutive frames (pbr_kernel_launch_batch): frame f of the batch uses
	 * frameSeed[f] / frameWeight[f]; f = 0 reads imageIn, later frames read what the frame before wrote to
	 * imageOut.  A pixel moves on to its next frame as soon as it has finished one (the frames of one pixel
	 * depend on each other, different pixels do not), so the device never drains between frames.
	 * seed / pixelWeight above are frame 0's. */
	int frameCount;
	float frameSeed[PT_MAX_BATCH];
	float frameWeight[PT_MAX_BATCH];
};

/* ------------------------------------------------------------------ loads */

__device__ __forceinline__ void loadNode(const float4* nodes, int index, float4& lo, float4& hi) {
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
		: "l"(nodes + 2 * (size_t) index));
}

/*
 * The two index words of a node record (the reference's array order, pre-order): lo.w = first face (leaf), -1
 * (inner node) or -2 (inner node carrying traverseShadows' "skip the next left child" flag, pt_bvh.cl:157-159 --
 * the reference's host never sets it, a foreign buffer may); hi.w = second face or -1 (leaf) / miss link (inner
 * node).  The node after a box hit, and after a leaf, is cur + 1.
 */
struct NodeWords {
	bool leaf;
	int afterMiss;     /* next node when the box is missed -- and after a leaf, hit or not */
	int afterHit;      /* next node when the box of an inner node is hit */
	int face0, face1;  /* leaf: its faces (face1 = -1: only one) */
};

template <bool ANY_HIT>
__device__ __forceinline__ NodeWords decodeNode(const int cur, const int loW, const int hiW) {
	NodeWords w;
	w.leaf = loW >= 0;
	w.afterMiss = w.leaf ? cur + 1 : hiW;
	w.afterHit = cur + 1 + ((ANY_HIT && loW == -2) ? 1 : 0);
	w.face0 = loW;
	w.face1 = hiW;
	return w;
}

#define PT_TRI_STRIDE 2           /* float4s per triangle record in `tris` */

/* 36 bytes per face in two arrays, so that the randomly accessed part of the scene stays as small as it can
 * (the walk is served from L2, and its hit rate falls off beyond ~60 MB, scripts/micro/chase.cu):
 * (a.xyz, edge1.xyz, edge2.xy) with one 256-bit load, edge2.z with one 32-bit load.  The material index is not
 * needed by the walk and lives in SceneDev::triMat. */
__device__ __forceinline__ void loadTri(const SceneDev& S, int face, float4& A, float4& E1, float4& E2) {
	const float4* p = S.tris + PT_TRI_STRIDE * (size_t) face;
	float e1y, e1z, e2x, e2y;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(A.x), "=f"(A.y), "=f"(A.z), "=f"(E1.x), "=f"(e1y), "=f"(e1z), "=f"(e2x), "=f"(e2y)
		: "l"(p));
	const float e2z = __ldg(S.trisB + face);
	A.w = 0.0f;
	E1.y = e1y; E1.z = e1z; E1.w = 0.0f;
	E2 = make_float4(e2x, e2y, e2z, 0.0f);
}

__device__ __forceinline__ vec3 f4xyz(const float4& f) { return v3(f.x, f.y, f.z); }
/* OpenCL `vector + scalar` */
__device__ __forceinline__ vec3 adds(const vec3 a, const float s) { return v3(a.x + s, a.y + s, a.z + s); }
__device__ __forceinline__ vec3 p4xyz(const pbr_float4& f) { return v3(f.x, f.y, f.z); }

/* ------------------------------------------------------------ intersection */

/* flatTriAndRayIntersect (pt_intersect.cl:92-129) + intersectFace (pt_bvh.cl:10-24) on a
 * pre-gathered triangle record.  Updates (rt, hitFace, hitLeaf) when the face is closer. */
__device__ __forceinline__ void intersectFaceLoaded(
	const float4 A, const float4 E1, const float4 E2, const int face, const int leaf,
	const vec3 o, const vec3 d, const float tNear,
	float& rt, int& hitFace, int& hitLeaf
) {
	const float f = fmaxf(0.0f, tNear - 0.001f);
	const vec3 closeOrigin = pm::fma3(d, f, o);
	const vec3 edge1 = f4xyz(E1);
	const vec3 edge2 = f4xyz(E2);
	const vec3 tVec = closeOrigin - f4xyz(A);
	const vec3 pVec = pm::cross(d, edge2);
	const vec3 qVec = pm::cross(tVec, edge1);
	const float invDet = pm::rcp(pm::dot(edge1, pVec));

	float t = pm::dot(edge2, qVec) * invDet;

	if (t >= rt || t < PT_EPSILON5) {
		return;
	}

	const float u = pm::dot(tVec, pVec) * invDet;
	const float v = pm::dot(d, qVec) * invDet;

	if (u + v > 1.0f || fminf(u, v) < 0.0f) {
		return;
	}

	t += f;

	if (rt > t) {
		rt = t;
		hitFace = face;
		hitLeaf = leaf;
	}
}

/* The same test as a function of the face alone (the ordered walk, pt_wide.cuh): the t flatTriAndRayIntersect
 * leaves behind for a ray whose ray.t is still INFINITY, or INFINITY when it rejects the face.  With a finite
 * ray.t the reference additionally rejects t' >= ray.t before the barycentric tests (pt_intersect.cl:107); such a
 * face has t = t' + f >= ray.t and loses the `ray->t > t` comparison of intersectFace anyway (pt_bvh.cl:17), so
 * "accepted" is exactly "returned t < ray.t". */
__device__ __forceinline__ float faceTLoaded(
	const float4 A, const float4 E1, const float4 E2, const vec3 o, const vec3 d, const float tNear
) {
	const float f = fmaxf(0.0f, tNear - 0.001f);
	const vec3 closeOrigin = pm::fma3(d, f, o);
	const vec3 edge1 = f4xyz(E1);
	const vec3 edge2 = f4xyz(E2);
	const vec3 tVec = closeOrigin - f4xyz(A);
	const vec3 pVec = pm::cross(d, edge2);
	const vec3 qVec = pm::cross(tVec, edge1);
	const float invDet = pm::rcp(pm::dot(edge1, pVec));

	float t = pm::dot(edge2, qVec) * invDet;
	if (t >= PM_INF_F || t < PT_EPSILON5) return PM_INF_F;

	const float u = pm::dot(tVec, pVec) * invDet;
	const float v = pm::dot(d, qVec) * invDet;
	if (u + v > 1.0f || fminf(u, v) < 0.0f) return PM_INF_F;

	t += f;
	return (t < PM_INF_F) ? t : PM_INF_F;          /* NaN can never be accepted (`ray->t > NaN` is false) */
}

__device__ __forceinline__ void intersectFace(
	const SceneDev& S, const int face, const int leaf,
	const vec3 o, const vec3 d, const float tNear,
	float& rt, int& hitFace, int& hitLeaf
) {
	float4 A, E1, E2;
	loadTri(S, face, A, E1, E2);
	intersectFaceLoaded(A, E1, E2, face, leaf, o, d, tNear, rt, hitFace, hitLeaf);
}

/* ------------------------------------------------------------ Phong tessellation (pt_phongtess.cl) */

__device__ __forceinline__ void swapf(float& a, float& b) { const float t = a; a = b; b = t; }

/* solveCubic (pt_utils.cl:108-199): a0 x^3 + a1 x^2 + a2 x + a3 = 0 */
__device__ __noinline__ int solveCubic(const float a0, const float a1, const float a2, const float a3, float x[3]) {
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

			const float u0 = p * pm::cos_(phi * THIRD) - w;
			const float u1 = p * pm::cos_((float) (((double) phi + 2.0f * PT_M_PI) * (double) THIRD)) - w;
			const float u2 = p * pm::cos_((float) (((double) phi + 4.0f * PT_M_PI) * (double) THIRD)) - w;

			x[0] = fminf(u0, fminf(u1, u2));
			x[1] = fmaxf(fminf(u0, u1), fmaxf(fminf(u0, u2), fminf(u1, u2)));
			x[2] = fmaxf(u0, fmaxf(u1, u2));
			#pragma unroll
			for (int i = 0; i < 3; i++) {
				x[i] -= pm::divide(
					a3 + x[i] * (a2 + x[i] * (a1 + x[i] * a0)),
					a2 + x[i] * (2.0f * a1 + x[i] * 3.0f * a0));
			}
			return 3;
		}
		dis = pm::sqrt_(dis);
		x[0] = pm::cbrt_(q + dis) + pm::cbrt_(q - dis) - w;
		x[0] -= pm::divide(
			a3 + x[0] * (a2 + x[0] * (a1 + x[0] * a0)),
			a2 + x[0] * (2.0f * a1 + x[0] * 3.0f * a0));
		return 1;
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

/* projectOnPlane (pt_utils.cl:397-399) */
__device__ __forceinline__ vec3 projectOnPlane(const vec3 q, const vec3 p, const vec3 n) {
	return q - pm::dot(q - p, n) * n;
}

/* phongTessellation (pt_phongtess.cl:14-27) */
__device__ __forceinline__ vec3 phongTessellation(
	const float alpha, const vec3 P1, const vec3 P2, const vec3 P3, const vec3 N1, const vec3 N2, const vec3 N3,
	const float u, const float v, const float w
) {
	const vec3 pBary = P1 * u + P2 * v + P3 * w;
	const vec3 pTessellated =
		u * projectOnPlane(pBary, P1, N1) +
		v * projectOnPlane(pBary, P2, N2) +
		w * projectOnPlane(pBary, P3, N3);
	return (1.0f - alpha) * pBary + alpha * pTessellated;
}

/* getPhongTessNormal with getTriangleNormalS / getTriangleNormal / getTriangleReflectionVec
 * (pt_utils.cl:231-294) */
__device__ __forceinline__ vec3 getPhongTessNormal(
	const vec3 an, const vec3 bn, const vec3 cn, const vec3 rayDir,
	const float u, const float v, const float w,
	const vec3 C12, const vec3 C23, const vec3 C31, const vec3 E23, const vec3 E31
) {
	const vec3 du = (w - u) * C31 + v * (C12 - C23) + E31;
	const vec3 dv = (w - v) * C23 + u * (C12 - C31) - E23;
	const vec3 ns = pm::normalize(pm::cross(du, dv));
	const vec3 np = pm::normalize(an * u + bn * v + cn * w);
	const vec3 r = rayDir - 2.0f * np * pm::dot(rayDir, np);
	return (pm::dot(ns, r) < 0.0f) ? ns : np;
}

/* phongTessTriAndRayIntersect (pt_phongtess.cl:56-212): direct ray tracing of Phong tessellation
 * after Ogaki & Tokuyoshi.  rayT is the current ray.t. */
__device__ __noinline__ vec3 phongTessTriAndRayIntersect(
	const float ALPHA,
	const vec3 P1, const vec3 P2, const vec3 P3, const vec3 N1, const vec3 N2, const vec3 N3,
	const vec3 rayO, const vec3 rayD, const float rayT, float* t, const float tNear, const float tFar
) {
	vec3 normal = v3(0.0f, 0.0f, 0.0f);
	*t = PM_INF_F;

	const vec3 E01 = P2 - P1;
	const vec3 E12 = P3 - P2;
	const vec3 E20 = P1 - P3;

	const vec3 C1 = ALPHA * (pm::dot(N2, E01) * N2 - pm::dot(N1, E01) * N1);
	const vec3 C2 = ALPHA * (pm::dot(N3, E12) * N3 - pm::dot(N2, E12) * N2);
	const vec3 C3 = ALPHA * (pm::dot(N1, E20) * N1 - pm::dot(N3, E20) * N3);

	float a, b, c, d, e, f, l, m, n, o, p, q;
	{
		/* getPlanesFromRay (pt_utils.cl:208-218) */
		const vec3 n1 = pm::normalize(pm::cross(rayO, rayD));
		const vec3 n2 = pm::normalize(pm::cross(n1, rayD));
		const float o1 = pm::dot(n1, rayO);
		const float o2 = pm::dot(n2, rayO);

		a = pm::dot(-n1, C3);
		b = pm::dot(-n1, C2);
		c = pm::dot(n1, P3) - o1;
		d = pm::dot(n1, C1 - C2 - C3) * 0.5f;
		e = pm::dot(n1, C3 + E20) * 0.5f;
		f = pm::dot(n1, C2 - E12) * 0.5f;
		l = pm::dot(-n2, C3);
		m = pm::dot(-n2, C2);
		n = pm::dot(n2, P3) - o2;
		o = pm::dot(n2, C1 - C2 - C3) * 0.5f;
		p = pm::dot(n2, C3 + E20) * 0.5f;
		q = pm::dot(n2, C2 - E12) * 0.5f;
	}

	float xs[3] = { -1.0f, -1.0f, -1.0f };
	int numCubicRoots = 0;
	{
		const float a3 = (l*m*n + 2.0f*o*p*q) - (l*q*q + m*p*p + n*o*o);
		const float a2 = (a*m*n + l*b*n + l*m*c + 2.0f*(d*p*q + o*e*q + o*p*f)) -
		                 (a*q*q + b*p*p + c*o*o + 2.0f*(l*f*q + m*e*p + n*d*o));
		const float a1 = (a*b*n + a*m*c + l*b*c + 2.0f*(o*e*f + d*e*q + d*p*f)) -
		                 (l*f*f + m*e*e + n*d*d + 2.0f*(a*f*q + b*e*p + c*d*o));
		const float a0 = (a*b*c + 2.0f*d*e*f) - (a*f*f + b*e*e + c*d*d);
		numCubicRoots = solveCubic(a0, a1, a2, a3, xs);
	}

	if (0 == numCubicRoots) return normal;

	float x = 0.0f;
	float determinant = PM_INF_F;
	float mA, mB, mC, mD, mE, mF;

	for (int i = 0; i < numCubicRoots; i++) {
		mA = a * xs[i] + l;
		mB = b * xs[i] + m;
		mD = d * xs[i] + o;
		const float tmp = mD * mD - mA * mB;
		x = (determinant > tmp) ? xs[i] : x;
		determinant = fminf(determinant, tmp);
	}

	if (0.0f >= determinant) return normal;

	/* getBestRayDomain (pt_phongtess.cl:36-45) */
	const float dx = fabsf(rayD.x), dy = fabsf(rayD.y), dz = fabsf(rayD.z);
	int domain = (dy > dz) ? 1 : 2;
	if (dx > dy) domain = (dx > dz) ? 0 : 2;

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

	if (fabsf(mEorF - lab1 * lc1 - lab2 * lc2) < fabsf(mEorF - lab1 * lc2 - lab2 * lc1)) swapf(lc1, lc2);

	for (int loop = 0; loop < 2; loop++) {
		const float g = (0 == loop) ? -lab1 : -lab2;
		const float h = (0 == loop) ? -lc1 : -lc2;

		const float c0 = ab + g * (2.0f * d + ba * g);
		const float c1 = 2.0f * (h * (d + ba * g) + ef + fe * g);
		const float c2 = h * (ba * h + 2.0f * fe) + c;
		const int numResults = solveCubic(0.0f, c0, c1, c2, xs);

		for (int i = 0; i < numResults; i++) {
			float u = xs[i];
			float v = g * u + h;
			const float w = 1.0f - u - v;

			if (u < 0.0f || v < 0.0f || w < 0.0f) continue;
			if (!AlessB) swapf(u, v);

			const vec3 pT = phongTessellation(ALPHA, P1, P2, P3, N1, N2, N3, u, v, w) - rayO;
			const float num = (domain == 0) ? pT.x : ((domain == 1) ? pT.y : pT.z);
			const float den = (domain == 0) ? rayD.x : ((domain == 1) ? rayD.y : rayD.z);
			const float tParam = pm::divide(num, den);

			if (tParam >= fabsf(tNear) && tParam <= fminf(*t, fminf(rayT, tFar))) {
				*t = tParam;
				normal = getPhongTessNormal(N1, N2, N3, rayD, u, v, w, C1, C2, C3, E12, E20);
			}
		}
	}
	return normal;
}

/* checkFaceIntersection + intersectFace with PHONGTESS == 1 (pt_intersect.cl:142-176, pt_bvh.cl:10-24).
 * In this mode a face record is 6 x float4: (a, material), (b, allNormalsEqual), (c, 0), an, bn, cn.
 * Faces whose three vertex normals are component-wise equal take the flat test, the others the Phong
 * test; the hit normal is part of the result. */
#define PT_TRI_STRIDE_PHONG 6

__device__ __forceinline__ void intersectFacePhong(
	const SceneDev& S, const int face, const int leaf,
	const vec3 o, const vec3 d, const float tNear, const float tFar,
	float& rt, int& hitFace, int& hitLeaf, vec3& hitNormal
) {
	const float4* p = S.tris + PT_TRI_STRIDE_PHONG * (size_t) face;
	const float4 A = __ldg(p), B = __ldg(p + 1), C = __ldg(p + 2);
	const vec3 a = f4xyz(A), b = f4xyz(B), c = f4xyz(C);
	float t;
	vec3 normal;
	if (B.w != 0.0f) {
		/* flatTriAndRayIntersect (pt_intersect.cl:92-129) */
		const float f = fmaxf(0.0f, tNear - 0.001f);
		const vec3 closeOrigin = pm::fma3(d, f, o);
		const vec3 edge1 = b - a, edge2 = c - a;
		const vec3 tVec = closeOrigin - a;
		const vec3 pVec = pm::cross(d, edge2);
		const vec3 qVec = pm::cross(tVec, edge1);
		const float invDet = pm::rcp(pm::dot(edge1, pVec));
		t = pm::dot(edge2, qVec) * invDet;
		if (t >= rt || t < PT_EPSILON5) return;
		const float u = pm::dot(tVec, pVec) * invDet;
		const float v = pm::dot(d, qVec) * invDet;
		if (u + v > 1.0f || fminf(u, v) < 0.0f) return;
		t += f;
		normal = pm::normalize(pm::cross(edge1, edge2));
	}
	else {
		const float4 NA = __ldg(p + 3), NB = __ldg(p + 4), NC = __ldg(p + 5);
		normal = phongTessTriAndRayIntersect(S.phongAlpha, a, b, c, f4xyz(NA), f4xyz(NB), f4xyz(NC), o, d, rt, &t, tNear, tFar);
	}
	if (rt > t) {
		rt = t;
		hitFace = face;
		hitLeaf = leaf;
		hitNormal = normal;
	}
}

/* intersectSphere (pt_intersect.cl:37-77), radius deliberately NOT squared (reference quirk). */
__device__ __forceinline__ bool intersectSphere(const vec3 o, const vec3 d, const vec3 pos, const float r, float* tNear) {
	const vec3 L = pos - o;
	const float tca = pm::dot(L, d);
	if (tca < 0.0f) return false;
	const float d2 = pm::dot(L, L) - tca * tca;
	if (d2 > r) return false;
	const float thc = pm::sqrt_(r - d2);
	float t0 = tca - thc;
	float t1 = tca + thc;
	if (t0 > t1) { const float tmp = t0; t0 = t1; t1 = tmp; }
	if (t0 < 0.0f) {
		t0 = t1;
		if (t0 < 0.0f) return false;
	}
	*tNear = t0;
	return true;
}

/* traverseLights (pt_bvh.cl:54-74): an orb that is hit closer than ray.t resets t to INFINITY and
 * tags the ray with -(light+1); any triangle hit afterwards overrides it. */
__device__ __forceinline__ void traverseLights(const SceneDev& S, const vec3 o, const vec3 d, float& rt, int& hitFace) {
	float tNear = 0.0f;
	for (int i = 0; i < S.numLights; i++) {
		const pbr_light light = S.lights[i];
		if (light.data.x == 2) {
			if (intersectSphere(o, d, p4xyz(light.pos), light.data.y, &tNear) && tNear < rt) {
				rt = PM_INF_F;
				hitFace = -(i + 1);
			}
		}
	}
}

/* intersectBox (pt_intersect.cl:11-25) with tFar entering as INFINITY. */
__device__ __forceinline__ bool intersectBox(
	const vec3 o, const vec3 invDir, const float4 lo, const float4 hi, float& tNear, float& tFar
) {
	const float t1x = (lo.x - o.x) * invDir.x, t1y = (lo.y - o.y) * invDir.y, t1z = (lo.z - o.z) * invDir.z;
	const float t2x = (hi.x - o.x) * invDir.x, t2y = (hi.y - o.y) * invDir.y, t2z = (hi.z - o.z) * invDir.z;
	const float tMinX = fminf(t1x, t2x), tMinY = fminf(t1y, t2y), tMinZ = fminf(t1z, t2z);
	const float tMaxX = fmaxf(t1x, t2x), tMaxY = fmaxf(t1y, t2y), tMaxZ = fmaxf(t1z, t2z);
	tNear = fmaxf(fmaxf(tMinX, tMinY), tMinZ);
	tFar = fminf(fminf(tMaxX, tMaxY), fminf(tMaxZ, PM_INF_F));
	return (tNear <= tFar);
}

/* traverse (pt_bvh.cl:82-123): stackless closest hit, one thread per ray, reference visiting order. */
template <bool PHONG>
__device__ __forceinline__ void traverseClosest(
	const SceneDev& S, const vec3 o, const vec3 d,
	float& rt, int& hitFace, int& hitLeaf, uint32_t& nNodes, uint32_t& nTris, vec3& hitNormal
) {
	const vec3 invDir = v3(pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z));
	int index = 1;

	if (S.numLights > 0) traverseLights(S, o, d, rt, hitFace);

	do {
		nNodes++;
		float4 lo, hi;
		loadNode(S.nodes, index, lo, hi);
		const int cur = index;
		const NodeWords w = decodeNode<false>(cur, __float_as_int(lo.w), __float_as_int(hi.w));

		index = w.afterMiss;

		float tNear, tFar;
		const bool isNodeHit = intersectBox(o, invDir, lo, hi, tNear, tFar) && tFar > PT_EPSILON5 && rt > tNear;

		if (!isNodeHit) continue;

		index = w.afterHit;

		if (w.leaf) {
			if (PHONG) intersectFacePhong(S, w.face0, cur, o, d, tNear, tFar, rt, hitFace, hitLeaf, hitNormal);
			else intersectFace(S, w.face0, cur, o, d, tNear, rt, hitFace, hitLeaf);
			nTris++;
			if (w.face1 != -1) {
				if (PHONG) intersectFacePhong(S, w.face1, cur, o, d, tNear, tFar, rt, hitFace, hitLeaf, hitNormal);
				else intersectFace(S, w.face1, cur, o, d, tNear, rt, hitFace, hitLeaf);
				nTris++;
			}
		}
	} while (index > 0 && index < S.numNodes);
}

/* traverseShadows (pt_bvh.cl:133-177): any hit closer than the light ends the walk; the box test
 * has no `ray.t > tNear` prune (reference behaviour). */
template <bool PHONG>
__device__ __forceinline__ void traverseAny(
	const SceneDev& S, const vec3 o, const vec3 d,
	float& rt, int& hitFace, int& hitLeaf, uint32_t& nNodes, uint32_t& nTris
) {
	vec3 hitNormal = v3(0.0f, 0.0f, 0.0f);
	const float tLight = rt;
	const vec3 invDir = v3(pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z));
	int index = 1;

	if (S.numLights > 0) traverseLights(S, o, d, rt, hitFace);

	do {
		nNodes++;
		float4 lo, hi;
		loadNode(S.nodes, index, lo, hi);
		const int cur = index;
		const NodeWords w = decodeNode<true>(cur, __float_as_int(lo.w), __float_as_int(hi.w));

		index = w.afterMiss;

		float tNear, tFar;
		const bool isNodeHit = intersectBox(o, invDir, lo, hi, tNear, tFar) && tFar > PT_EPSILON5;

		if (!isNodeHit) continue;

		index = w.afterHit;

		if (w.leaf) {
			if (PHONG) intersectFacePhong(S, w.face0, cur, o, d, tNear, tFar, rt, hitFace, hitLeaf, hitNormal);
			else intersectFace(S, w.face0, cur, o, d, tNear, rt, hitFace, hitLeaf);
			nTris++;
			if (w.face1 != -1) {
				if (PHONG) intersectFacePhong(S, w.face1, cur, o, d, tNear, tFar, rt, hitFace, hitLeaf, hitNormal);
				else intersectFace(S, w.face1, cur, o, d, tNear, rt, hitFace, hitLeaf);
				nTris++;
			}
			if (rt < tLight) break;
		}
	} while (index > 0 && index < S.numNodes);
}

/* ------------------------------------------------------------------ pt_utils.cl */

/* rand (pt_utils.cl:39-44) */
__device__ __forceinline__ float rnd(float& seed) {
	seed += 1.0f;
	return pm::fract_(pm::sin_(seed) * 43758.5453123f);
}

/* fresnel (pt_utils.cl:53-56) */
__device__ __forceinline__ float fresnel(const float u, const float c) {
	const float v = 1.0f - u;
	return c + (1.0f - c) * v * v * v * v * v;
}

/* reflect (pt_utils.cl:426) */
__device__ __forceinline__ vec3 reflect(const vec3 dir, const vec3 normal) {
	return dir - 2.0f * pm::dot(normal, dir) * normal;
}

/* jitter (pt_utils.cl:306-318) */
__device__ __noinline__ vec3 jitter(const vec3 nl, const float phi, const float sina, const float cosa) {
	const vec3 u = pm::normalize(pm::cross(pm::yzx(nl), nl));
	const vec3 v = pm::normalize(pm::cross(nl, u));
	double s, c;
	pm::sincos_d((double) phi, &s, &c);
	return pm::normalize(pm::normalize(u * (float) c + v * (float) s) * sina + nl * cosa);
}

/* refract (pt_utils.cl:436-465) */
__device__ __forceinline__ vec3 refractDir(const vec3 dir, const vec3 normal, const Material& mtl, float& seed) {
	const bool into = (pm::dot(normal, -dir) > 0.0f);
	const vec3 nl = into ? normal : -normal;

	const float m1 = into ? PT_NI_AIR : mtl.Ni;
	const float m2 = into ? mtl.Ni : PT_NI_AIR;
	const float m = pm::divide(m1, m2);

	const float cosI = -pm::dot(nl, dir);
	const float sinT2 = m * m * (1.0f - cosI * cosI);

	if (sinT2 >= 1.0f) {
		return reflect(dir, nl);
	}

	const float sqrtCosT = pm::sqrt_(1.0f - sinT2);
	const float r0 = pm::divide(m1 - m2, m1 + m2);
	const float c = (m1 > m2) ? sqrtCosT : cosI;
	const float reflectance = fresnel(c, r0 * r0);

	return (reflectance < rnd(seed)) ? m * dir + (m * cosI - sqrtCosT) * nl : reflect(dir, nl);
}

/* ------------------------------------------------------------------ pt_brdf.cl, BRDF 0 (Schlick) */

__device__ __forceinline__ float schlickZ(const float t, const float r) {
	const float x = 1.0f + r * t * t - t * t;
	return (x == 0.0f) ? 0.0f : pm::divide(r, x * x);
}
__device__ __forceinline__ float schlickA(const float w, const float p) {
	const float p2 = p * p;
	const float w2 = w * w;
	const float x = p2 - p2 * w2 + w2;
	return (x == 0.0f) ? 0.0f : pm::sqrt_(pm::divide(p, x));
}
__device__ __forceinline__ float schlickG(const float v, const float r) {
	const float x = r - r * v + v;
	return (x == 0.0f) ? 0.0f : pm::divide(v, x);
}

/* brdfSchlick (pt_brdf.cl:125-149) with D / B2 (pt_brdf.cl:71-112) folded in. */
__device__ __forceinline__ float brdfSchlick(
	const Material& mtl, const vec3 dirOut, const vec3 dirIn, const vec3 normal, float* u, float* pdf
) {
	const vec3 V_IN = dirIn;
	const vec3 V_OUT = -dirOut;

	const vec3 un = pm::normalize(pm::cross(pm::yzx(normal), normal));
	const vec3 h = pm::normalize(V_OUT + V_IN);
	const float t = pm::dot(h, normal);
	const float vIn = pm::dot(V_IN, normal);
	const float vOut = pm::dot(V_OUT, normal);
	const vec3 hp = pm::normalize(pm::cross(pm::cross(h, normal), normal));
	const float w = pm::dot(un, hp);

	*u = pm::dot(h, V_OUT);
	*pdf = pm::divide(t, (float) (4.0f * PT_M_PI * (double) pm::dot(V_OUT, h)));

	const float r = mtl.a3, p = mtl.a2;
	const float b = 4.0f * r * (1.0f - r);
	const float a = (r < 0.5f) ? 0.0f : 1.0f - b;
	const float c = (r < 0.5f) ? 1.0f - b : 0.0f;
	const float dd = (float) (4.0f * PT_M_PI * (double) vOut * (double) vIn);
	const float lam = (float) ((double) a * PT_M_1_PI);
	float ani = 0.0f;
	if (!(b == 0.0f || dd == 0.0f)) {
		const float gp = schlickG(vOut, r) * schlickG(vIn, r);
		const float obstructed = gp * schlickZ(t, r) * schlickA(w, p);
		const float reemission = 1.0f - gp;
		ani = pm::divide(b, dd) * (obstructed + reemission);
	}
	const float fres = (vIn == 0.0f) ? 0.0f : pm::divide(c, vIn);
	return lam + ani + fres;
}

/* newRaySchlick (pt_brdf.cl:159-208) */
__device__ __forceinline__ vec3 newRaySchlick(const vec3 dir, const vec3 normal, const Material& mtl, float& seed) {
	if (mtl.a3 == 0.0f) {
		return reflect(dir, normal);
	}

	float a = rnd(seed);
	float b = rnd(seed);
	const float iso2 = mtl.a2 * mtl.a2;
	const float alpha = pm::acos_(pm::sqrt_(pm::divide(a, mtl.a3 - a * mtl.a3 + a)));
	float phi;

	if (b < 0.25f) {
		b = 1.0f - 4.0f * (0.25f - b);
		const float b2 = b * b;
		phi = (float) (PT_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
	}
	else if (b < 0.5f) {
		b = 1.0f - 4.0f * (0.5f - b);
		const float b2 = b * b;
		phi = (float) (PT_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (PT_M_PI - (double) phi);
	}
	else if (b < 0.75f) {
		b = 1.0f - 4.0f * (0.75f - b);
		const float b2 = b * b;
		phi = (float) (PT_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (PT_M_PI + (double) phi);
	}
	else {
		b = 1.0f - 4.0f * (1.0f - b);
		const float b2 = b * b;
		phi = (float) (PT_M_PI_2 * (double) pm::sqrt_(pm::divide(iso2 * b2, 1.0f - b2 + b2 * iso2)));
		phi = (float) (2.0f * PT_M_PI - (double) phi);
	}

	if (mtl.a2 < 1.0f) {
		phi = (float) ((double) phi + PT_M_PI_2);
	}

	const vec3 H = jitter(normal, phi, pm::sin_(alpha), pm::cos_(alpha));
	vec3 newRay = reflect(dir, H);

	if (pm::dot(newRay, normal) <= 0.0f) {
		const float phi2 = PT_PI_X2 * rnd(seed);
		newRay = jitter(normal, phi2, pm::sqrt_(a), pm::sqrt_(1.0f - a));
	}
	return newRay;
}

/* ------------------------------------------------------------------ pt_brdf.cl, BRDF 1 (Shirley-Ashikhmin) */

/* brdfShirleyAshikhmin (pt_brdf.cl:228-268) */
__device__ __forceinline__ void brdfShirleyAshikhmin(
	const float nu, const float nv, const float Rd,
	const vec3 dirOut, const vec3 dirIn, const vec3 normal,
	float* brdfSpec, float* brdfDiff, float* dotHK1, float* pdf
) {
	const vec3 un = pm::normalize(pm::cross(pm::yzx(normal), normal));
	const vec3 vn = pm::normalize(pm::cross(normal, un));

	const vec3 k1 = dirIn;
	const vec3 k2 = -dirOut;
	const vec3 h = pm::normalize(k1 + k2);

	const float dotHU = pm::dot(h, un);
	const float dotHV = pm::dot(h, vn);
	const float dotHN = pm::dot(h, normal);
	const float dotNK1 = pm::dot(normal, k1);
	const float dotNK2 = pm::dot(normal, k2);
	*dotHK1 = pm::dot(h, k1);

	float ps_e = nu * dotHU * dotHU + nv * dotHV * dotHV;
	ps_e = (dotHN == 1.0f) ? 0.0f : pm::divide(ps_e, 1.0f - dotHN * dotHN);
	const float ps0 = (float) ((double) (pm::sqrt_((nu + 1.0f) * (nv + 1.0f)) * 0.125f) * PT_M_1_PI);
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

/* newRayShirleyAshikhmin (pt_brdf.cl:278-330) */
__device__ __forceinline__ vec3 newRayShirleyAshikhmin(const vec3 dir, const vec3 rayNormal, const Material& mtl, float& seed) {
	float a = rnd(seed);
	const float b = rnd(seed);
	float phi_flip = (float) PT_M_PI;
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
		phi_flip = (float) (2.0f * PT_M_PI);
		phi_flipf = -1.0f;
	}

	a = 1.0f - 4.0f * (aMax - a);

	const float phi = pm::atan_(
		pm::sqrt_(pm::divide(mtl.a2 + 1.0f, mtl.a3 + 1.0f)) * pm::tan_((float) (PT_M_PI_2 * (double) a))
	);
	const float phi_full = phi_flip + phi_flipf * phi;

	double sphi, cphi;
	pm::sincos_d((double) phi, &sphi, &cphi);
	const float cosphi = (float) cphi;
	const float sinphi = (float) sphi;
	const float theta_e = pm::rcp(mtl.a2 * cosphi * cosphi + mtl.a3 * sinphi * sinphi + 1.0f);
	const float theta = pm::acos_(pm::pow_(1.0f - b, theta_e));

	const vec3 normal = (mtl.d < 1.0f || pm::dot(rayNormal, -dir) >= 0.0f) ? rayNormal : -rayNormal;

	double sth, cth;
	pm::sincos_d((double) theta, &sth, &cth);
	const vec3 h = jitter(normal, phi_full, (float) sth, (float) cth);
	const vec3 spec = reflect(dir, h);
	const float phi3 = PT_PI_X2 * rnd(seed);
	const vec3 diff = jitter(normal, phi3, pm::sqrt_(b), pm::sqrt_(1.0f - b));

	return (pm::dot(spec, normal) <= 0.0f) ? diff : spec;
}

/* ------------------------------------------------------------------ material fetch */

/* materials[facesV[hitFace].w] (pathtracing.cl:268); out-of-range index -> MtlParser defaults
 * (see oracle/pt_oracle.cpp fetchMaterial for the rationale). */
template <int BRDF>
__device__ __forceinline__ Material fetchMaterial(const void* materials, const int numMaterials, const uint32_t idx) {
	Material m;
	if (idx >= (uint32_t) numMaterials) {
		m.d = 1.0f; m.Ni = 1.0f;
		if (BRDF == 0) { m.a2 = 1.0f; m.a3 = 1.0f; m.Rs = 0.0f; m.Rd = 0.0f; }
		else { m.a2 = 0.0f; m.a3 = 0.0f; m.Rs = 0.0f; m.Rd = 1.0f; }
		m.rgbDiff = v3(1.0f, 1.0f, 1.0f);
		m.rgbSpec = v3(1.0f, 1.0f, 1.0f);
		return m;
	}
	if (BRDF == 0) {
		const float4* p = (const float4*) materials + 3 * (size_t) idx;
		const float4 d0 = __ldg(p), d1 = __ldg(p + 1), d2 = __ldg(p + 2);
		m.d = d0.x; m.Ni = d0.y; m.a2 = d0.z; m.a3 = d0.w; m.Rs = 0.0f; m.Rd = 0.0f;
		m.rgbDiff = f4xyz(d1);
		m.rgbSpec = f4xyz(d2);
	}
	else {
		const float4* p = (const float4*) materials + 4 * (size_t) idx;
		const float4 d0 = __ldg(p), d1 = __ldg(p + 1), d2 = __ldg(p + 2), d3 = __ldg(p + 3);
		m.d = d0.x; m.Ni = d0.y; m.a2 = d0.z; m.a3 = d0.w; m.Rs = d1.x; m.Rd = d1.y;
		m.rgbDiff = f4xyz(d2);
		m.rgbSpec = f4xyz(d3);
	}
	return m;
}

/* ------------------------------------------------------------------ per-pixel path state */

struct PathState {
	vec3 o, d;                /* current ray */
	float t;                  /* ray.t after traversal */
	int hitFace;
	vec3 color;               /* throughput of the current sample */
	vec3 finalColor;          /* sum over samples (and shadow-ray terms) */
	float seed;
	float focus;
	uint32_t depth;
	int depthAdded;
	uint32_t sample;
	uint32_t secondaryPaths;
	uint32_t nNodes, nTris;   /* debugColor.y / debugColor.x */
	uint32_t frame;           /* frame of the batch this pixel is working on */
	vec3 hitNormal;           /* PHONGTESS only: ray.normal of the hit (flat faces: recomputed from the edges) */
};

enum BounceResult { PATH_CONTINUE = 0, PATH_SAMPLE_DONE = 1 };

/* Pixel of a path index.  Rows [y0,y1) are walked in 8x4 pixel blocks so that a warp covers a
 * compact screen patch (the reference launches 8x8 work-groups, CL.cpp:293-297); falls back to
 * row-major when the tile is not a multiple of the block. */
__device__ __forceinline__ void pathToPixel(const int p, const int width, const int y0, const int rows, int& px, int& py) {
	if ((width & 7) == 0 && (rows & 3) == 0) {
		const int blocksPerRow = width >> 3;
		const int b = p >> 5, l = p & 31;
		px = ((b % blocksPerRow) << 3) + (l & 7);
		py = y0 + ((b / blocksPerRow) << 2) + (l >> 3);
	}
	else {
		px = p % width;
		py = y0 + p / width;
	}
}

/* initRay + antiAliasing + depthOfField (pathtracing.cl:25-48, pt_utils.cl:327-373) and the reset
 * at the top of the sample loop (pathtracing.cl:251-256). */
__device__ __forceinline__ void beginSample(const FrameParams& P, PathState& s, const int px, const int py) {
	const vec3 camU = p4xyz(P.cam.u), camV = p4xyz(P.cam.v), camW = p4xyz(P.cam.w);
	const float W = (float) P.width, H = (float) P.height;

	const vec3 initialRay = camW + P.pxDim * 0.5f * (
		camU - W * camU + 2.0f * (float) px * camU +
		camV - H * camV + 2.0f * (float) py * camV
	);

	s.color = v3(1.0f, 1.0f, 1.0f);
	s.t = PM_INF_F;
	s.o = p4xyz(P.cam.eye);
	s.d = pm::normalize(initialRay);
	s.hitFace = 0;
	s.depth = 0;
	s.depthAdded = 0;

	/* antiAliasing */
	const float r = rnd(s.seed);
	const float phi = PT_PI_X2 * rnd(s.seed);
	const vec3 aaDir = jitter(s.d, phi, pm::sqrt_(r), pm::sqrt_(1.0f - r));
	s.d = pm::normalize(s.d + aaDir * P.pxDim * P.antiAliasing);

	/* depthOfField, only with a focus point set (PathTracer::setFocus) */
	if (P.cam.focusPoint.x >= 0 && P.cam.focusPoint.y >= 0) {
		float tObject = P.imageIn[(size_t) py * P.width + px].w;
		const int fx = min(max(P.cam.focusPoint.x, 0), P.width - 1), fy = min(max(P.cam.focusPoint.y, 0), P.height - 1);
		float tFocus = P.imageIn[(size_t) fy * P.width + fx].w;
		if (tFocus >= 0.0f && tObject >= 0.0f) {
			if (tObject == PM_INF_F) tObject = 1000.0f;
			if (tFocus == PM_INF_F) tFocus = 1000.0f;
			if (tObject > 0.0f) {
				const float aperture = P.cam.lense.x / P.cam.lense.y;
				const float radius = rnd(s.seed) * aperture * 0.5f;
				const float angle = PT_PI_X2 * rnd(s.seed);
				const float x = radius * pm::cos_(angle);
				const float y = radius * pm::sin_(angle);
				s.o = s.o + x * camU + y * camV;
				const vec3 hitFocalPlane = pm::fma3(s.d, tFocus, p4xyz(P.cam.eye));
				s.d = pm::normalize(hitFocalPlane - s.o);
			}
		}
	}
}

/* Pixel of a path of this launch: pathToPixel over the launch's rows, then the stripe interleave if one is set. */
__device__ __forceinline__ void pixelOf(const FrameParams& P, const int p, int& px, int& py) {
	pathToPixel(p, P.width, P.y0, P.y1 - P.y0, px, py);
	if (P.stripeRows > 0) {
		const int r = py - P.y0;
		py = (r / P.stripeRows) * (P.stripeRows * P.stripeWorld) + P.stripeRank * P.stripeRows + (r % P.stripeRows);
	}
}

/* Everything the kernel does before the sample loop (pathtracing.cl:235-249). */
__device__ __forceinline__ void initPath(const FrameParams& P, PathState& s) {
	s.finalColor = v3(0.0f, 0.0f, 0.0f);
	s.seed = P.frameSeed[s.frame];
	s.focus = 0.0f;
	s.sample = 0;
	s.secondaryPaths = 1;
	s.nNodes = 0;
	s.nTris = 0;
}

/* `if( light.x > -1.0f ) { color *= light; finalColor += color; }` (pathtracing.cl:320-323) */
__device__ __forceinline__ void endSampleWithLight(PathState& s, const vec3 light) {
	if (light.x > -1.0f) {
		s.color = s.color * light;
		s.finalColor = s.finalColor + s.color;
	}
}

/* The body of the depth loop after traverse() (pathtracing.cl:261-317).  On PATH_CONTINUE the
 * state holds the next ray (t = INFINITY, hitFace = 0) and depth has been advanced and checked
 * against MAX_DEPTH + depthAdded.  On PATH_SAMPLE_DONE the sample's light has been applied. */
/* The part of the depth loop's body between traverse() and the shadow ray (pathtracing.cl:261-278): miss -> the
 * sample ends with the sky or an orb light; otherwise material, extendDepth, the "last round" exit, seed += t.
 * Shared by bounce() and by the kernel that generates the shadow rays of a wavefront (which runs it on a copy of
 * the path state), so that both take exactly the same decisions. */
enum PrefixResult { PREFIX_SAMPLE_DONE = 0, PREFIX_HIT = 1 };

template <int BRDF, bool PHONG>
__device__ __forceinline__ PrefixResult bouncePrefix(
	const FrameParams& P, PathState& s, Material& mtl, vec3& normal, bool& addDepth, vec3& hitPoint
) {
	const SceneDev& S = P.scene;

	s.focus = (s.sample + s.depth == 0) ? s.t : s.focus;

	if (s.t == PM_INF_F) {
		const vec3 light = (s.hitFace < 0) ? p4xyz(S.lights[-(s.hitFace + 1)].rgb) : v3(P.skyLight.x, P.skyLight.y, P.skyLight.z);
		endSampleWithLight(s, light);
		return PREFIX_SAMPLE_DONE;
	}

	uint32_t mtlIndex;
	if (PHONG) {
		mtlIndex = (uint32_t) __float_as_int(__ldg(S.tris + PT_TRI_STRIDE_PHONG * (size_t) s.hitFace).w);
		normal = s.hitNormal;
	}
	else {
		float4 A, E1, E2;
		loadTri(S, s.hitFace, A, E1, E2);
		mtlIndex = __ldg(S.triMat + s.hitFace);
		normal = pm::normalize(pm::cross(f4xyz(E1), f4xyz(E2)));
	}
	mtl = fetchMaterial<BRDF>(P.materials, P.numMaterials, mtlIndex);

	/* extendDepth (pt_utils.cl:89-96) */
	if (BRDF == 1) addDepth = (fmaxf(mtl.a2, mtl.a3) >= 50.0f);
	else addDepth = (mtl.a3 < rnd(s.seed));

	if (mtl.d == 1.0f && !addDepth && s.depth == (uint32_t) (P.maxDepth + s.depthAdded - 1)) {
		return PREFIX_SAMPLE_DONE;
	}

	s.seed += s.t;

	hitPoint = pm::fma3(s.d, s.t, s.o);
	return PREFIX_HIT;
}

/* Does this hit send a shadow ray (pathtracing.cl:282-288), and which one (shadowRayTest, :188-192)? */
__device__ __forceinline__ bool shadowRayOf(const SceneDev& S, const Material& mtl, const vec3 hitPoint, vec3& lightDir, float& tLight) {
	if (!(S.numLights > 0 && mtl.d > 0.0f)) return false;
	const vec3 toLight = p4xyz(S.lights[0].pos) - hitPoint;
	lightDir = pm::normalize(toLight);
	tLight = pm::length(toLight);
	return true;
}

/* SHADOW_PRE: the shadow ray of this hit was walked by a separate launch; `shadowT` is its ray.t afterwards. */
template <int BRDF, bool SHADOW, bool PHONG, bool SHADOW_PRE = false>
__device__ __forceinline__ BounceResult bounce(
	const FrameParams& P, PathState& s, uint32_t& shadowNodes, uint32_t& shadowRays, const float shadowT = 0.0f
) {
	const SceneDev& S = P.scene;
	Material mtl;
	vec3 normal, hitPoint;
	bool addDepth;
	if (bouncePrefix<BRDF, PHONG>(P, s, mtl, normal, addDepth, hitPoint) == PREFIX_SAMPLE_DONE) {
		return PATH_SAMPLE_DONE;
	}

	/* shadowRayTest (pathtracing.cl:188-199) */
	bool lit = false;
	vec3 lightDir = v3(0.0f, 0.0f, 0.0f);
	vec3 lightRgb = v3(-1.0f, -1.0f, -1.0f);
	if (SHADOW) {
		float tLight;
		if (shadowRayOf(S, mtl, hitPoint, lightDir, tLight)) {
			float lt = tLight;
			if (SHADOW_PRE) {
				lt = shadowT;
			}
			else {
				int lf = 0, ll = -1;
				uint32_t nn = 0;
				traverseAny<PHONG>(S, hitPoint, lightDir, lt, lf, ll, nn, s.nTris);
				shadowNodes += nn;
				shadowRays++;
			}
			if (lt >= tLight) {
				lightRgb = p4xyz(S.lights[0].rgb);
				lit = true;
			}
		}
	}

	/* getNewRay (pt_brdf.cl:344-378) */
	const bool doTransRefr = (mtl.d < 1.0f && mtl.d <= rnd(s.seed));
	addDepth = (addDepth || doTransRefr);
	vec3 newDir;
	if (doTransRefr) newDir = refractDir(s.d, normal, mtl, s.seed);
	else if (BRDF == 0) newDir = newRaySchlick(s.d, normal, mtl, s.seed);
	else newDir = newRayShirleyAshikhmin(s.d, normal, mtl, s.seed);

	if (pm::dot(normal, -s.d) <= 0.0f) {
		normal = -normal;
	}

	/* updateColor (pathtracing.cl:92-178) */
	if (BRDF == 0) {
		float brdf, pdf, u;
		if (SHADOW && lit && lightRgb.x >= 0) {
			brdf = brdfSchlick(mtl, s.d, lightDir, normal, &u, &pdf);
			if (fabsf(pdf) > 0.00001f) {
				brdf *= fmaxf(pm::dot(normal, lightDir), 0.0f);
				brdf = pm::divide(brdf, pdf);
				const float v = 1.0f - u;
				const vec3 fr = v3(
					mtl.rgbSpec.x + (1.0f - mtl.rgbSpec.x) * v * v * v * v * v,
					mtl.rgbSpec.y + (1.0f - mtl.rgbSpec.y) * v * v * v * v * v,
					mtl.rgbSpec.z + (1.0f - mtl.rgbSpec.z) * v * v * v * v * v);
				const vec3 term = adds(fr * brdf * mtl.d, 1.0f - mtl.d);
				s.finalColor = s.finalColor + s.color * lightRgb * mtl.rgbDiff * term;
				s.secondaryPaths += 1;
			}
		}
		brdf = brdfSchlick(mtl, s.d, newDir, normal, &u, &pdf);
		brdf *= fmaxf(pm::dot(normal, newDir), 0.0f);
		brdf = pm::divide(brdf, pdf);
		const float v = 1.0f - u;
		const vec3 fr = v3(
			mtl.rgbSpec.x + (1.0f - mtl.rgbSpec.x) * v * v * v * v * v,
			mtl.rgbSpec.y + (1.0f - mtl.rgbSpec.y) * v * v * v * v * v,
			mtl.rgbSpec.z + (1.0f - mtl.rgbSpec.z) * v * v * v * v * v);
		const vec3 term = adds(fr * brdf * mtl.d, 1.0f - mtl.d);
		s.color = s.color * (mtl.rgbDiff * term);
	}
	else {
		float brdfDiff, brdfSpec, pdf, dotHK1;
		if (SHADOW && lit && lightRgb.x >= 0) {
			brdfShirleyAshikhmin(mtl.a2, mtl.a3, mtl.Rd, s.d, lightDir, normal, &brdfSpec, &brdfDiff, &dotHK1, &pdf);
			if (fabsf(pdf) > 0.00001f) {
				brdfSpec = pm::divide(brdfSpec, pdf);
				brdfDiff = pm::divide(brdfDiff, pdf);
				const vec3 brdf_s = brdfSpec * mtl.rgbSpec * fresnel(dotHK1, mtl.Rs);
				const vec3 brdf_d = brdfDiff * mtl.rgbDiff * (1.0f - mtl.Rs);
				vec3 brdfColor = adds((brdf_s + brdf_d) * mtl.d, 1.0f - mtl.d);
				const float maxRGB = pm::max_(1.0f, pm::max_(brdfColor.x, pm::max_(brdfColor.y, brdfColor.z)));
				brdfColor = v3(brdfColor.x / maxRGB, brdfColor.y / maxRGB, brdfColor.z / maxRGB);
				const vec3 cl = v3(pm::clamp_(brdfColor.x, 0.0f, 1.0f), pm::clamp_(brdfColor.y, 0.0f, 1.0f), pm::clamp_(brdfColor.z, 0.0f, 1.0f));
				s.finalColor = s.finalColor + adds(cl * lightRgb * mtl.d, 1.0f - mtl.d);
				s.secondaryPaths += 1;
			}
		}
		brdfShirleyAshikhmin(mtl.a2, mtl.a3, mtl.Rd, s.d, newDir, normal, &brdfSpec, &brdfDiff, &dotHK1, &pdf);
		brdfSpec = pm::divide(brdfSpec, pdf);
		brdfDiff = pm::divide(brdfDiff, pdf);
		const vec3 brdf_s = brdfSpec * mtl.rgbSpec * fresnel(dotHK1, mtl.Rs);
		const vec3 brdf_d = brdfDiff * mtl.rgbDiff * (1.0f - mtl.Rs);
		vec3 brdfColor = adds((brdf_s + brdf_d) * mtl.d, 1.0f - mtl.d);
		const float maxRGB = pm::max_(1.0f, pm::max_(brdfColor.x, pm::max_(brdfColor.y, brdfColor.z)));
		brdfColor = v3(brdfColor.x / maxRGB, brdfColor.y / maxRGB, brdfColor.z / maxRGB);
		s.color = s.color * v3(pm::clamp_(brdfColor.x, 0.0f, 1.0f), pm::clamp_(brdfColor.y, 0.0f, 1.0f), pm::clamp_(brdfColor.z, 0.0f, 1.0f));
	}

	s.depthAdded += (addDepth && s.depthAdded < P.maxAddedDepth);

	/* russianRoulette (pt_utils.cl:385-387) */
	const float maxValColor = fmaxf(s.color.x, fmaxf(s.color.y, s.color.z));
	if ((int) s.depth > 2 + s.depthAdded && maxValColor < rnd(s.seed)) {
		return PATH_SAMPLE_DONE;
	}

	/* ray = newRay; depth++ */
	s.o = hitPoint;
	s.d = newDir;
	s.t = PM_INF_F;
	s.hitFace = 0;
	s.depth++;

	return (s.depth < (uint32_t) (P.maxDepth + s.depthAdded)) ? PATH_CONTINUE : PATH_SAMPLE_DONE;
}

/* finalColor /= secondaryPaths [/= SAMPLES]; setColors (pt_rgb.cl:9-21); writeDebugImage
 * (pathtracing.cl:73-78). */
__device__ __forceinline__ void finishPixel(const FrameParams& P, PathState& s, const int px, const int py) {
	const float sp = (float) s.secondaryPaths;
	vec3 fc = v3(s.finalColor.x / sp, s.finalColor.y / sp, s.finalColor.z / sp);
	if (P.samples > 1) {
		const float n = (float) P.samples;
		fc = v3(fc.x / n, fc.y / n, fc.z / n);
	}
	const size_t o = (size_t) py * P.width + px;
	if (P.frameOut) {
		P.frameOut[o] = make_float4(fc.x, fc.y, fc.z, s.focus);
		return;
	}
	const float4 in = (s.frame == 0u) ? P.imageIn[o] : P.imageOut[o];
	const float pixelWeight = P.frameWeight[s.frame];
	float4 out;
	out.x = pm::mix_(fc.x, in.x, pixelWeight);
	out.y = pm::mix_(fc.y, in.y, pixelWeight);
	out.z = pm::mix_(fc.z, in.z, pixelWeight);
	out.w = s.focus;
	P.imageOut[o] = out;
	if (P.imageDebug) {
		P.imageDebug[o] = make_float4((float) s.nTris / 1082.0f, (float) s.nNodes / 1265.0f, 0.0f, 0.0f);
	}
}

/* What follows a bounce: next bounce, next sample (pathtracing.cl:251 loop), or -- pixel written -- the next
 * frame of the batch.  Returns false when the pixel has nothing left to do in this launch. */
__device__ __forceinline__ bool advancePath(const FrameParams& P, PathState& s, const BounceResult r, const int px, const int py) {
	if (r == PATH_CONTINUE) return true;
	s.sample++;
	if (s.sample < (uint32_t) P.samples) {
		beginSample(P, s, px, py);
		return true;
	}
	finishPixel(P, s, px, py);
	s.frame++;
	if (s.frame < (uint32_t) P.frameCount) {
		initPath(P, s);
		beginSample(P, s, px, py);
		return true;
	}
	return false;
}

} /* namespace ptd */
