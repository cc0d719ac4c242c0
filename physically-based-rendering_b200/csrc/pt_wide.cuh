/*
 * pt_wide.cuh -- the ordered walk: closest-hit / any-hit traversal of the 4-wide BVH (wide_bvh.h).
 *
 * Why.  The reference's walk (pt_bvh.cl:82-123) is stackless and visits the tree in a fixed pre-order: every node
 * whose box the ray hits is visited, near or far, one dependent 32-byte fetch each -- 111 (primary) to 260
 * (incoherent) fetches per ray on the 1 M-triangle soup.  The hit it finds does not depend on that order (see
 * "Exactness"); only its visit counters do.  So when nobody looks at the counters (no debug image), the same hit is
 * found by an ordered walk with a stack:
 *
 *   - a work item is an inner node of the 4-wide tree or a leaf, with the tNear of its box;
 *   - an inner node is one 128-byte line: four 256-bit loads bring its four children (box, reference), four box tests
 *     -- the reference's intersectBox on the reference's bits -- decide which are still in reach; the nearest becomes
 *     the next item, the others go on the stack with their tNear (leaves too: a leaf's tNear is an input of the
 *     triangle test, pt_intersect.cl:96);
 *   - a leaf item tests its one or two triangles; a popped item whose tNear has fallen out of reach is dropped without
 *     a fetch;
 *   - the engine is the persistent-warp one of pt_kernels.cuh: lanes step through inner nodes until fewer than
 *     nodePhaseMin of them have one, then all lanes with a leaf test it together; finished lanes claim new rays;
 *   - the stack lives in shared memory ([entry][thread], conflict-free), entries beyond PT_WIDE_STACK_SM in local memory;
 *   - the top of the tree (SceneDev::wideTop nodes, breadth-first) is staged in shared memory once per block: the
 *     first steps of every ray cost a shared-memory read instead of a trip to L2.
 *   14 (primary) / 32 (incoherent) node visits and 6 / 13 triangle tests per ray instead of 111 / 260 and 15 / 34 on
 *   the soup (tests/test_wide_walk.py).
 *   (A first version gave each ray four lanes, one per child: one L1 wavefront per node, but 40 warp instructions per
 *   ray-visit against 3 for the reference-order engine -- it was slower than the walk it replaced; profiles/r02_*.)
 *
 * Exactness (DESIGN.md section 4 gives the full argument).  Let a ray's candidate set be every face of a leaf whose
 * box the ray hits (intersectBox && tFar > 1e-5: the very test, on the very bits, the reference applies) that
 * flatTriAndRayIntersect accepts for ray.t = INFINITY, with t_F its returned t -- a function of (ray, face, leaf box)
 * only.  The reference accepts a face iff t_F < ray.t at the time it tests it, tests faces in index order, and skips
 * a leaf iff ray.t <= tNear(leaf) when it gets there.  Because boxes nest exactly (checked by the builder) and ray.t
 * only shrinks, it follows that the reference returns (m, F0) = the smallest t_F and, among equals, the lowest face
 * index -- provided t_F0 >= tNear of F0's own leaf.  The shifted-origin arithmetic (pt_intersect.cl:96-120) can put
 * t_F below its leaf's tNear, but by no more than 0.001 (t' >= 1e-5, f = tNear - 0.001).  So the ordered walk
 *   (1) prunes a box only when tNear > ray.t + 0.0022 + 2e-6 ray.t (twice that bound, with room for rounding): it sees
 *       every face that could be F0 and every face within F0's leaf-box gap;
 *   (2) keeps the best hit with the reference's tie rule and the second-best t from a different leaf;
 *   (3) if the winner lies in front of its own leaf box AND another leaf has a hit no farther than that box's tNear,
 *       the reference's answer depends on its visiting order: the ray is walked again in the reference's order
 *       (traverseClosest).  0 rays on the soup, 1.7 % on a scene built of coplanar, overlapping axis-aligned quads
 *       (tests/test_wide_walk.py).
 * Any-hit (shadow rays inside a frame): the reference's traverseShadows reports "occluded" iff some candidate face has
 * t_F < tLight (it has no ray.t prune, pt_bvh.cl:150-153), which does not depend on the order at all.  Only that
 * boolean is consumed (pathtracing.cl:196-198); WHICH face ends the reference's walk does depend on the order, so
 * explicit any-hit rays (pbr_trace) always take the reference-order engine.
 */
#pragma once

#include "pt_kernels.cuh"

namespace ptk {

#define PT_WIDE_BLOCK 128                  /* threads per block */
#define PT_WIDE_STACK_SM 16                /* stack entries per ray in shared memory ... */
#define PT_WIDE_STACK_LOCAL 48             /* ... and behind them in local memory (the soup needs 20, tests/test_wide_walk.py) */
#ifndef PT_WIDE_MIN_BLOCKS
#define PT_WIDE_MIN_BLOCKS 8
#endif
#define PT_WIDE_EMPTY 0x7fffffff

__host__ __device__ inline size_t wideSharedBytes(const int topNodes) {
	return (size_t) topNodes * 128 + (size_t) PT_WIDE_BLOCK * PT_WIDE_STACK_SM * sizeof(uint2);
}

/* A box farther than this cannot hold the winner or a face that makes the winner ambiguous (header, (1)). */
__device__ __forceinline__ float widePruneLimit(const float rt) { return rt + (0.0022f + 2e-6f * rt); }

/* The reference-order walk for the rare ambiguous ray.
 * (Inlined: as a __noinline__ call inside the engine's loop it crashes ptxas 12.9, whatever the signature.) */
struct WideRewalk { float rt; int hitFace, hitLeaf; uint32_t nn, nt; };
__device__ __forceinline__ WideRewalk wideStrictRewalk(const SceneDev& S, const vec3 o, const vec3 d, const float rt0, const int hitFace0) {
	WideRewalk r = {rt0, hitFace0, -1, 0u, 0u};
	vec3 n = v3(0.0f, 0.0f, 0.0f);
	if (S.numNodes >= 2) traverseClosest<false>(S, o, d, r.rt, r.hitFace, r.hitLeaf, r.nn, r.nt, n);
	return r;
}

struct WideLane {
	vec3 o, d, inv;
	float rt;                 /* best t so far (ray.t) */
	float lim;                /* widePruneLimit(rt); any-hit: of the light's distance */
	float bestTn;             /* tNear of the leaf the best hit is in (-INF: no face accepted yet) */
	float t2;                 /* smallest t seen that is not the best hit's (other leaves, the ray's initial t) */
	float tLight, rt0;
	int hitFace, hitFace0;
	int item;                 /* >= 0: inner node;  < 0: leaf reference */
	float itemTn;
	int sp;
	uint32_t nn, nt;
	bool overflow;
};

enum { WIDE_IDLE = 0, WIDE_POP = 1, WIDE_NODE = 2, WIDE_LEAF = 3, WIDE_FINISHED = 4 };

/* intersectBox (pt_intersect.cl:11-25) on a child record (min.x, max.x, min.y, max.y, min.z, max.z): the same
 * operations on the same operands in the same order -- (bound - o) * invDir, fmin / fmax per axis, max of the mins,
 * min of the maxes -- with the two planes of an axis in one packed FADD2 / FMUL2 (IEEE round-to-nearest per half,
 * so the bits are those of the scalar instructions). */
__device__ __forceinline__ bool wideBoxTest(const vec3 o, const vec3 inv, const float4 a, const float4 b, float& tNear, float& tFar) {
	float t1x, t2x, t1y, t2y, t1z, t2z;
	asm("{\n\t.reg .b64 p, q, r;\n\t"
	    "mov.b64 p, {%2, %3};\n\tmov.b64 q, {%4, %4};\n\tsub.rn.f32x2 r, p, q;\n\tmov.b64 q, {%5, %5};\n\tmul.rn.f32x2 r, r, q;\n\t"
	    "mov.b64 {%0, %1}, r;\n\t}" : "=f"(t1x), "=f"(t2x) : "f"(a.x), "f"(a.y), "f"(o.x), "f"(inv.x));
	asm("{\n\t.reg .b64 p, q, r;\n\t"
	    "mov.b64 p, {%2, %3};\n\tmov.b64 q, {%4, %4};\n\tsub.rn.f32x2 r, p, q;\n\tmov.b64 q, {%5, %5};\n\tmul.rn.f32x2 r, r, q;\n\t"
	    "mov.b64 {%0, %1}, r;\n\t}" : "=f"(t1y), "=f"(t2y) : "f"(a.z), "f"(a.w), "f"(o.y), "f"(inv.y));
	asm("{\n\t.reg .b64 p, q, r;\n\t"
	    "mov.b64 p, {%2, %3};\n\tmov.b64 q, {%4, %4};\n\tsub.rn.f32x2 r, p, q;\n\tmov.b64 q, {%5, %5};\n\tmul.rn.f32x2 r, r, q;\n\t"
	    "mov.b64 {%0, %1}, r;\n\t}" : "=f"(t1z), "=f"(t2z) : "f"(b.x), "f"(b.y), "f"(o.z), "f"(inv.z));
	const float tMinX = fminf(t1x, t2x), tMinY = fminf(t1y, t2y), tMinZ = fminf(t1z, t2z);
	const float tMaxX = fmaxf(t1x, t2x), tMaxY = fmaxf(t1y, t2y), tMaxZ = fmaxf(t1z, t2z);
	tNear = fmaxf(fmaxf(tMinX, tMinY), tMinZ);
	tFar = fminf(fminf(tMaxX, tMaxY), fminf(tMaxZ, PM_INF_F));
	return (tNear <= tFar);
}

/* compare-exchange of two (tNear, reference) pairs: the nearer one first, ties keep their order */
__device__ __forceinline__ void wideOrder(float& ka, int& ra, float& kb, int& rb) {
	const bool swap = kb < ka;
	const float k0 = swap ? kb : ka, k1 = swap ? ka : kb;
	const int r0 = swap ? rb : ra, r1 = swap ? ra : rb;
	ka = k0; kb = k1; ra = r0; rb = r1;
}

template <bool ANY_HIT, typename RaySource, typename Counter>
__device__ __forceinline__ void wideEngine(
	const SceneDev& S, const float4* smTop, uint2* smStack, RaySource& src, const Counter count, Counter* cursor,
	uint32_t& totalNodes, uint32_t& totalTris, uint32_t& totalRays, uint32_t& totalRewalks
) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	uint2* const stackSm = smStack + threadIdx.x;                 /* entry e at stackSm[e * PT_WIDE_BLOCK] */
	uint2 stackLocal[PT_WIDE_STACK_LOCAL];

	WideLane L;
	L.sp = 0;
	int state = WIDE_IDLE;
	bool exhausted = false;
	Counter slot = 0;

	while (true) {
		/* ---- retire */
		if (state == WIDE_FINISHED) {
			int hitLeaf = -1;
			if (!ANY_HIT) {
				if (L.overflow || (L.rt < L.bestTn && L.t2 <= L.bestTn)) {
					/* (the lights were consulted when the ray started: rt0 / hitFace0 are what traverseLights left) */
					const WideRewalk r = wideStrictRewalk(S, L.o, L.d, L.rt0, L.hitFace0);
					L.rt = r.rt; L.hitFace = r.hitFace; hitLeaf = r.hitLeaf; L.nn += r.nn; L.nt += r.nt;
					totalRewalks++;
				}
				else if (L.bestTn > -PM_INF_F && src.wantsLeaf()) hitLeaf = __ldg(S.faceLeaf + L.hitFace);
			}
			LaneRay R;
			R.rt = L.rt; R.hitFace = L.hitFace; R.hitLeaf = hitLeaf; R.nn = L.nn; R.nt = L.nt;
			R.normal = v3(0.0f, 0.0f, 0.0f);
			src.store(slot, R);
			totalNodes += L.nn; totalTris += L.nt; totalRays++;
			state = WIDE_IDLE;
		}
		/* ---- refill */
		const unsigned need = __ballot_sync(FULL, state == WIDE_IDLE);
		if (!exhausted && (__popc(need) >= S.refillMin || need == FULL)) {
			const int leader = __ffs(need) - 1;
			const int n = __popc(need);
			Counter base = 0;
			if (lane == leader) base = atomicAdd(cursor, (Counter) n);
			base = __shfl_sync(FULL, base, leader);
			if (state == WIDE_IDLE) {
				const Counter i = base + (Counter) __popc(need & ltMask);
				if (i < count) {
					float r0;
					int hf;
					src.fetch(i, L.o, L.d, r0, hf);
					L.inv = v3(pm::rcp(L.d.x), pm::rcp(L.d.y), pm::rcp(L.d.z));
					L.rt = r0; L.tLight = r0; L.hitFace = hf;
					if (S.numLights > 0) traverseLights(S, L.o, L.d, L.rt, L.hitFace);
					L.rt0 = L.rt; L.hitFace0 = L.hitFace;
					L.lim = widePruneLimit(ANY_HIT ? L.tLight : L.rt);
					L.bestTn = -PM_INF_F;
					L.t2 = L.rt;
					L.nn = 0; L.nt = 0;
					L.item = 0; L.itemTn = 0.0f; L.sp = 0;
					L.overflow = false;
					slot = i;
					state = WIDE_NODE;
				}
			}
			if (base + (Counter) n >= count) exhausted = true;
		}
		if (__ballot_sync(FULL, state != WIDE_IDLE) == 0u) break;

		/* ---- node phase: lanes without an item take the nearest stack entry still in reach (one attempt per round),
		 * lanes whose item is an inner node visit it */
		while (true) {
			if (state == WIDE_POP) {
				if (L.sp == 0) state = WIDE_FINISHED;
				else {
					L.sp--;
					const uint2 e = (L.sp < PT_WIDE_STACK_SM) ? stackSm[L.sp * PT_WIDE_BLOCK] : stackLocal[L.sp - PT_WIDE_STACK_SM];
					if (__uint_as_float(e.y) <= L.lim) {
						L.item = (int) e.x; L.itemTn = __uint_as_float(e.y);
						state = (L.item >= 0) ? WIDE_NODE : WIDE_LEAF;
					}
				}
			}
			if (state == WIDE_NODE) {
				L.nn++;
				float4 a[4], b[4];
				if (L.item < S.wideTop) {
					const float4* p = smTop + L.item * 8;
					#pragma unroll
					for (int k = 0; k < 4; k++) { a[k] = p[2 * k]; b[k] = p[2 * k + 1]; }
				}
				else {
					#pragma unroll
					for (int k = 0; k < 4; k++) loadNode(S.wide, L.item * 4 + k, a[k], b[k]);
				}
				/* the four children: entry distance if still in reach, else INFINITY (reference -> PT_WIDE_EMPTY) */
				float key[4];
				int ref[4];
				#pragma unroll
				for (int k = 0; k < 4; k++) {
					float tNear, tFar;
					const int r = __float_as_int(b[k].z);
					const bool in = wideBoxTest(L.o, L.inv, a[k], b[k], tNear, tFar) && tFar > PT_EPSILON5 && r != PT_WIDE_EMPTY &&
						tNear <= L.lim && tNear < PM_INF_F;
					key[k] = in ? tNear : PM_INF_F;
					ref[k] = in ? r : PT_WIDE_EMPTY;
				}
				/* nearest first (five compare-exchanges); the others go on the stack, farthest first */
				wideOrder(key[0], ref[0], key[1], ref[1]);
				wideOrder(key[2], ref[2], key[3], ref[3]);
				wideOrder(key[0], ref[0], key[2], ref[2]);
				wideOrder(key[1], ref[1], key[3], ref[3]);
				wideOrder(key[1], ref[1], key[2], ref[2]);
				const int n = (ref[0] != PT_WIDE_EMPTY) + (ref[1] != PT_WIDE_EMPTY) + (ref[2] != PT_WIDE_EMPTY) + (ref[3] != PT_WIDE_EMPTY);
				if (L.sp + 3 <= PT_WIDE_STACK_SM) {
					int at = L.sp;
					if (n > 3) { stackSm[at * PT_WIDE_BLOCK] = make_uint2((uint32_t) ref[3], __float_as_uint(key[3])); at++; }
					if (n > 2) { stackSm[at * PT_WIDE_BLOCK] = make_uint2((uint32_t) ref[2], __float_as_uint(key[2])); at++; }
					if (n > 1) { stackSm[at * PT_WIDE_BLOCK] = make_uint2((uint32_t) ref[1], __float_as_uint(key[1])); at++; }
					L.sp = at;
				}
				else {
					#pragma unroll
					for (int k = 3; k >= 1; k--) {
						if (n > k) {
							const uint2 e = make_uint2((uint32_t) ref[k], __float_as_uint(key[k]));
							if (L.sp < PT_WIDE_STACK_SM) stackSm[L.sp * PT_WIDE_BLOCK] = e;
							else if (L.sp < PT_WIDE_STACK_SM + PT_WIDE_STACK_LOCAL) stackLocal[L.sp - PT_WIDE_STACK_SM] = e;
							else L.overflow = true;
							if (!L.overflow) L.sp++;
						}
					}
				}
				L.item = ref[0]; L.itemTn = key[0];
				state = L.overflow ? WIDE_FINISHED : (n == 0 ? WIDE_POP : (ref[0] >= 0 ? WIDE_NODE : WIDE_LEAF));
			}
			if (__popc(__ballot_sync(FULL, state == WIDE_NODE || state == WIDE_POP)) < S.nodePhaseMin) break;
		}

		/* ---- triangle phase: lanes whose item is a leaf test its faces (reference order inside the leaf) */
		if (state == WIDE_LEAF) {
			const int f0 = L.item & 0x3fffffff;
			const bool two = (L.item & 0x40000000) != 0;
			float4 A0, E10, E20;
			float4 A1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), E11 = A1, E21 = A1;
			loadTri(S, f0, A0, E10, E20);
			if (two) loadTri(S, f0 + 1, A1, E11, E21);
			float tl = faceTLoaded(A0, E10, E20, L.o, L.d, L.itemTn);
			int fl = f0;
			L.nt++;
			if (two) {
				const float t1 = faceTLoaded(A1, E11, E21, L.o, L.d, L.itemTn);
				L.nt++;
				if (t1 < tl) { tl = t1; fl = f0 + 1; }
			}
			state = WIDE_POP;
			if (ANY_HIT) {
				/* occluded iff some candidate face is closer than the light (pt_bvh.cl:170-172) */
				if (tl < L.tLight) { L.rt = tl; state = WIDE_FINISHED; }
			}
			else if (tl < PM_INF_F) {
				if (tl < L.rt || (tl == L.rt && L.bestTn > -PM_INF_F && fl < L.hitFace)) {
					L.t2 = fminf(L.t2, L.rt);
					L.rt = tl; L.hitFace = fl; L.bestTn = L.itemTn;
					L.lim = widePruneLimit(tl);
				}
				else L.t2 = fminf(L.t2, tl);
			}
		}
	}
}

/* Stage the top of the tree; every thread of the block must call. */
__device__ __forceinline__ void wideStageTop(const SceneDev& S, float4* smTop) {
	const int n = S.wideTop * 8;
	for (int i = threadIdx.x; i < n; i += blockDim.x) smTop[i] = __ldg(S.wide + i);
	__syncthreads();
}

/* Closest hit of every live path (the stage traverseKernel is in the reference-order pipeline). */
__global__ void __launch_bounds__(PT_WIDE_BLOCK, PT_WIDE_MIN_BLOCKS) traverseWideKernel(
	const SceneDev S, const WaveState W, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ countPtr,
	uint32_t* cursor, uint32_t* countToReset, unsigned long long* stats
) {
	extern __shared__ float4 smWide[];
	float4* smTop = smWide;
	uint2* smStack = (uint2*) (smWide + (size_t) S.wideTop * 8);
	const uint32_t count = *countPtr;
	if (blockIdx.x == 0 && threadIdx.x == 0) *countToReset = 0u;
	wideStageTop(S, smTop);
	uint32_t nodes = 0, tris = 0, rays = 0, rewalks = 0;
	WaveRaySource src = {W, queue, 0u};
	wideEngine<false>(S, smTop, smStack, src, count, cursor, nodes, tris, rays, rewalks);
	warpAddStat(stats + 0, rays);
	warpAddStat(stats + 2, nodes);
	warpAddStat(stats + 3, tris);
	warpAddStat(stats + 6, rewalks);
	warpAddStat(stats + 7, rays);
}

/* Shadow rays of a wavefront (the stage traverseShadowKernel is in the reference-order pipeline). */
__global__ void __launch_bounds__(PT_WIDE_BLOCK, PT_WIDE_MIN_BLOCKS) traverseWideShadowKernel(
	const SceneDev S, const WaveState W, float4* shadowO, const float4* __restrict__ shadowD,
	const uint32_t* __restrict__ shadowQ, const uint32_t* __restrict__ countPtr, uint32_t* cursor, unsigned long long* stats
) {
	extern __shared__ float4 smWide[];
	float4* smTop = smWide;
	uint2* smStack = (uint2*) (smWide + (size_t) S.wideTop * 8);
	const uint32_t count = *countPtr;
	wideStageTop(S, smTop);
	uint32_t nodes = 0, tris = 0, rays = 0, rewalks = 0;
	ShadowRaySource src = {W, shadowO, shadowD, shadowQ, 0u};
	wideEngine<true>(S, smTop, smStack, src, count, cursor, nodes, tris, rays, rewalks);
	warpAddStat(stats + 1, rays);
	warpAddStat(stats + 5, nodes);
	warpAddStat(stats + 3, tris);
	warpAddStat(stats + 7, rays);
}

/* Explicit closest-hit rays (BASELINE config 5). */
__global__ void __launch_bounds__(PT_WIDE_BLOCK, PT_WIDE_MIN_BLOCKS) traceRaysWideKernel(
	const SceneDev S, const pbr_ray* __restrict__ rays, const long long n, pbr_hit* __restrict__ hits,
	unsigned long long* cursor, unsigned long long* stats
) {
	extern __shared__ float4 smWide[];
	float4* smTop = smWide;
	uint2* smStack = (uint2*) (smWide + (size_t) S.wideTop * 8);
	wideStageTop(S, smTop);
	uint32_t nodes = 0, tris = 0, cnt = 0, rewalks = 0;
	ExplicitRaySource src = {rays, hits};
	wideEngine<false>(S, smTop, smStack, src, (unsigned long long) n, cursor, nodes, tris, cnt, rewalks);
	warpAddStat(stats + 0, cnt);
	warpAddStat(stats + 2, nodes);
	warpAddStat(stats + 3, tris);
	warpAddStat(stats + 6, rewalks);
	warpAddStat(stats + 7, cnt);
}

} /* namespace ptk */
