/*
 * pt_kernels.cuh -- the sm_100a kernels behind libpbr_b200.so.
 *
 * The reference launches ONE kernel per frame, one work-item per pixel, that loops over samples
 * and bounces (source/opencl/pathtracing.cl:207-334).  Here a frame is a wavefront:
 *
 *     raygen            every pixel of the tile: initPath + beginSample          -> ray + path state
 *     repeat (at most SAMPLES * (MAX_DEPTH + MAX_ADDED_DEPTH) times)
 *         traverse      persistent warps pull 32 live paths at a time from the queue and walk the
 *                       BVH (small register footprint -> many warps per SM to hide L2 latency)
 *         [shadowGen, traverseShadow   with render.shadow_rays: the shadow ray of every hit as a stage of its own]
 *         shade         per live path: bounce(); paths that finish a sample start the next one
 *                       (seed carried over) or write their pixel (setColors) and, in a batch, go on to
 *                       their next frame, or leave; survivors are appended to the next queue with one
 *                       warp-aggregated atomic per warp
 *
 * Per-pixel arithmetic and its order are those of the reference kernel, so every pixel gets the
 * same bits whatever the scheduling.  `megaKernel` keeps the reference's launch structure (one
 * thread per pixel, everything inline): the on-device cross-check of the wavefront and the faster of
 * the two on small scenes (pbr_capi.cu chooses by measurement).  Also here: the explicit-ray kernels and the
 * scene repack.  The traverse stage exists twice: here in the reference's visiting order (one lane per ray, the
 * stackless pre-order walk: visit counters and debug image bit-exact), and in pt_wide.cuh as the ordered walk
 * over a 4-wide BVH (four lanes per ray: same hits, same image, a fraction of the node fetches).
 *
 * Path state lives in HBM as 16-byte SoA records (coalesced 128-bit accesses):
 *     rayO (o.xyz, t)   rayD (d.xyz, hitFace)   colS (color.xyz, seed)   finF (finalColor.xyz, focus)
 *     misc (depth | depthAdded << 16, sample, secondaryPaths, frame of the batch)
 *     dbg (nodes visited, triangle tests)
 */
#pragma once

#include "pt_device.cuh"

namespace ptk {

using namespace ptd;

struct WaveState {
	float4* rayO;
	float4* rayD;
	float4* colS;
	float4* finF;
	uint4* misc;
	uint2* dbg;
	float4* hitN;             /* PHONGTESS only: normal of the hit found by traverse */
};

/* ctrl[0], ctrl[1]: element counts of queue 0 / 1;  ctrl[2]: work cursor of the traverse kernel */
struct QueueCtl {
	uint32_t* ctrl;
	uint32_t* queue[2];
};

__device__ __forceinline__ void storePath(const WaveState& W, const uint32_t p, const PathState& s) {
	W.rayO[p] = make_float4(s.o.x, s.o.y, s.o.z, s.t);
	W.rayD[p] = make_float4(s.d.x, s.d.y, s.d.z, __int_as_float(s.hitFace));
	W.colS[p] = make_float4(s.color.x, s.color.y, s.color.z, s.seed);
	W.finF[p] = make_float4(s.finalColor.x, s.finalColor.y, s.finalColor.z, s.focus);
	W.misc[p] = make_uint4(s.depth | ((uint32_t) s.depthAdded << 16), s.sample, s.secondaryPaths, s.frame);
}

__device__ __forceinline__ void loadPath(const WaveState& W, const uint32_t p, PathState& s) {
	const float4 o = W.rayO[p], d = W.rayD[p], c = W.colS[p], f = W.finF[p];
	const uint4 m = W.misc[p];
	const uint2 g = W.dbg[p];
	s.o = v3(o.x, o.y, o.z); s.t = o.w;
	s.d = v3(d.x, d.y, d.z); s.hitFace = __float_as_int(d.w);
	s.color = v3(c.x, c.y, c.z); s.seed = c.w;
	s.finalColor = v3(f.x, f.y, f.z); s.focus = f.w;
	s.depth = m.x & 0xffffu; s.depthAdded = (int) (m.x >> 16);
	s.sample = m.y; s.secondaryPaths = m.z; s.frame = m.w;
	s.nNodes = g.x; s.nTris = g.y;
	if (W.hitN) { const float4 n = W.hitN[p]; s.hitNormal = v3(n.x, n.y, n.z); }
	else s.hitNormal = v3(0.0f, 0.0f, 0.0f);
}

/* Sum a per-thread counter over the warp and add it to a global 64-bit counter once. */
__device__ __forceinline__ void warpAddStat(unsigned long long* dst, uint32_t v) {
	unsigned long long x = v;
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
	if ((threadIdx.x & 31) == 0 && x) atomicAdd(dst, x);
}

/* ------------------------------------------------------------------ raygen */

__global__ void __launch_bounds__(256) raygenKernel(const FrameParams P, const WaveState W, const QueueCtl Q, const int nPaths) {
	const int stride = gridDim.x * blockDim.x;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nPaths; p += stride) {
		int px, py;
		pixelOf(P, p, px, py);
		PathState s;
		s.frame = 0u;
		initPath(P, s);
		beginSample(P, s, px, py);
		storePath(W, (uint32_t) p, s);
		W.dbg[p] = make_uint2(0u, 0u);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		Q.ctrl[0] = (uint32_t) nPaths;   /* queue 0 = identity */
		Q.ctrl[1] = 0u;
		Q.ctrl[2] = 0u;
	}
}

/* setColors (pt_rgb.cl:9-21) for a frame whose pixels were finished into FrameParams::frameOut: the same mix on the
 * same operands, only later -- once the frame before it has been folded in. */
__global__ void __launch_bounds__(256) mixFrameKernel(const FrameParams P, const float4* __restrict__ frameOut, const int nPaths) {
	const int stride = gridDim.x * blockDim.x;
	const float pixelWeight = P.frameWeight[0];
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nPaths; p += stride) {
		int px, py;
		pixelOf(P, p, px, py);
		const size_t o = (size_t) py * P.width + px;
		const float4 in = P.imageIn[o], fr = frameOut[o];
		float4 out;
		out.x = pm::mix_(fr.x, in.x, pixelWeight);
		out.y = pm::mix_(fr.y, in.y, pixelWeight);
		out.z = pm::mix_(fr.z, in.z, pixelWeight);
		out.w = fr.w;
		P.imageOut[o] = out;
	}
}

/* ------------------------------------------------------------------ traverse */

/*
 * Persistent-warp traversal engine shared by the wavefront kernel and the explicit-ray kernel.
 *
 * The reference walks the BVH with one work-item per ray and tests a leaf's triangles the moment
 * the leaf is reached (pt_bvh.cl:82-123).  Run like that on a 32-wide warp, almost every loop trip
 * has SOME lane in a leaf, so the whole warp pays for the (long) triangle code while only one or two
 * lanes use it, and a warp lives as long as its slowest ray: ncu showed 5-9 active threads per warp.
 * Each ray still visits its nodes and triangles in exactly the reference order here -- only the
 * interleaving BETWEEN rays changes:
 *   node phase     lanes step through nodes until they reach a leaf whose box is hit ("pending");
 *                  the phase ends as soon as fewer than SceneDev::nodePhaseMin lanes are still stepping
 *   triangle phase all pending lanes test their one or two faces together
 *   retire/refill  lanes whose ray is finished store the result; once SceneDev::refillMin lanes are
 *                  idle they claim the next rays from the queue together (one warp-aggregated
 *                  atomic), so lanes do not idle while the slowest ray of a batch finishes
 */

struct LaneRay {
	vec3 o, d, invDir;
	float rt;
	float tLight;             /* any-hit only: the initial ray.t */
	int hitFace, hitLeaf;
	int index;                /* next node to visit; out of (0, numNodes) = finished */
	uint32_t nn, nt;
	/* pending leaf */
	int leafCur, leafF0, leafF1;
	float leafTNear, leafTFar;
	vec3 normal;              /* PHONGTESS only */
};

template <bool ANY_HIT>
__device__ __forceinline__ void startRay(const SceneDev& S, LaneRay& L, const vec3 o, const vec3 d, const float rt, const int hitFace) {
	L.o = o;
	L.d = d;
	L.invDir = v3(pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z));
	L.rt = rt;
	L.tLight = rt;
	L.hitFace = hitFace;
	L.hitLeaf = -1;
	L.index = 1;
	L.nn = 0;
	L.nt = 0;
	L.normal = v3(0.0f, 0.0f, 0.0f);
	if (S.numLights > 0) traverseLights(S, o, d, L.rt, L.hitFace);
}

/* One node of the stackless walk (pt_bvh.cl:88-121 without the face tests), node already loaded.
 * Returns true when the node is a leaf whose box was hit: the faces are tested in the triangle phase. */
template <bool ANY_HIT>
__device__ __forceinline__ bool nodeVisit(LaneRay& L, const float4 lo, const float4 hi) {
	L.nn++;
	const int cur = L.index;
	const NodeWords w = decodeNode<ANY_HIT>(cur, __float_as_int(lo.w), __float_as_int(hi.w));

	L.index = w.afterMiss;

	float tNear, tFar;
	bool isNodeHit = intersectBox(L.o, L.invDir, lo, hi, tNear, tFar) && tFar > PT_EPSILON5;
	if (!ANY_HIT) isNodeHit = isNodeHit && (L.rt > tNear);
	if (!isNodeHit) return false;

	L.index = w.afterHit;
	if (!w.leaf) return false;
	L.leafCur = cur;
	L.leafF0 = w.face0;
	L.leafF1 = w.face1;
	L.leafTNear = tNear;
	L.leafTFar = tFar;
	return true;
}

template <bool ANY_HIT>
__device__ __forceinline__ bool nodeStep(const SceneDev& S, LaneRay& L) {
	float4 lo, hi;
	loadNode(S.nodes, L.index, lo, hi);
	return nodeVisit<ANY_HIT>(L, lo, hi);
}

/* intersectFaces (pt_bvh.cl:35-46) for the pending leaf. */
#ifndef PT_TRAVERSE_MIN_BLOCKS
#define PT_TRAVERSE_MIN_BLOCKS 9           /* resident blocks of 128 threads per SM the traversal kernels are compiled for */
#endif
template <bool ANY_HIT, bool PHONG>
__device__ __forceinline__ void leafStep(const SceneDev& S, LaneRay& L) {
	if (!PHONG) {
		/* both records are requested before the first test, so a two-face leaf costs one trip to L2 instead of
		 * two; the tests themselves run in the reference's order */
		const bool two = L.leafF1 != -1;
		float4 A0, E10, E20;
		float4 A1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), E11 = A1, E21 = A1;
		loadTri(S, L.leafF0, A0, E10, E20);
		if (two) loadTri(S, L.leafF1, A1, E11, E21);
		intersectFaceLoaded(A0, E10, E20, L.leafF0, L.leafCur, L.o, L.d, L.leafTNear, L.rt, L.hitFace, L.hitLeaf);
		L.nt++;
		if (two) {
			intersectFaceLoaded(A1, E11, E21, L.leafF1, L.leafCur, L.o, L.d, L.leafTNear, L.rt, L.hitFace, L.hitLeaf);
			L.nt++;
		}
	}
	else {
		intersectFacePhong(S, L.leafF0, L.leafCur, L.o, L.d, L.leafTNear, L.leafTFar, L.rt, L.hitFace, L.hitLeaf, L.normal);
		L.nt++;
		if (L.leafF1 != -1) {
			intersectFacePhong(S, L.leafF1, L.leafCur, L.o, L.d, L.leafTNear, L.leafTFar, L.rt, L.hitFace, L.hitLeaf, L.normal);
			L.nt++;
		}
	}
	if (ANY_HIT && L.rt < L.tLight) L.index = -1;     /* `break` of traverseShadows (pt_bvh.cl:170-172) */
}

/* RaySource: bool fetch(i, o, d, rt, hitFace) / void store(i, L).  count = number of rays. */
/* Lane states of the traversal engine. */
enum { LANE_IDLE = 0, LANE_STEPPING = 1, LANE_PENDING = 2, LANE_FINISHED = 3 };

template <bool ANY_HIT, bool PHONG, typename RaySource, typename Counter>
__device__ __forceinline__ void traverseEngine(
	const SceneDev& S, RaySource& src, const Counter count, Counter* cursor,
	uint32_t& totalNodes, uint32_t& totalTris, uint32_t& totalRays
) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	const unsigned lastNode = (unsigned) (S.numNodes - 1);      /* valid indices: 1 .. numNodes-1 */

	LaneRay L;
	L.index = 0;
	int state = LANE_IDLE;
	bool exhausted = false;
	Counter slot = 0;

	while (true) {
		/* retire finished rays, refill idle lanes */
		if (state == LANE_FINISHED) {
			src.store(slot, L);
			totalNodes += L.nn; totalTris += L.nt; totalRays++;
			state = LANE_IDLE;
		}
		const unsigned need = __ballot_sync(FULL, state == LANE_IDLE);
		if (!exhausted && (__popc(need) >= S.refillMin || need == FULL)) {
			const int leader = __ffs(need) - 1;
			const int n = __popc(need);
			Counter base = 0;
			if (lane == leader) base = atomicAdd(cursor, (Counter) n);
			base = __shfl_sync(FULL, base, leader);
			if (state == LANE_IDLE) {
				const Counter i = base + (Counter) __popc(need & ltMask);
				if (i < count) {
					vec3 o, d;
					float rt;
					int hf;
					src.fetch(i, o, d, rt, hf);
					startRay<ANY_HIT>(S, L, o, d, rt, hf);
					slot = i;
					state = (lastNode >= 1u) ? LANE_STEPPING : LANE_FINISHED;
				}
			}
			if (base + (Counter) n >= count) exhausted = true;
		}
		if (__ballot_sync(FULL, state != LANE_IDLE) == 0u) break;

		/* node phase: index stays inside [1, numNodes) while a lane is stepping */
		while (true) {
			if (state == LANE_STEPPING) {
				const bool leaf = nodeStep<ANY_HIT>(S, L);
				const bool inside = (unsigned) (L.index - 1) < lastNode;
				state = leaf ? LANE_PENDING : (inside ? LANE_STEPPING : LANE_FINISHED);
			}
			if (__popc(__ballot_sync(FULL, state == LANE_STEPPING)) < S.nodePhaseMin) break;
		}

		/* triangle phase */
		if (state == LANE_PENDING) {
			leafStep<ANY_HIT, PHONG>(S, L);
			state = ((unsigned) (L.index - 1) < lastNode) ? LANE_STEPPING : LANE_FINISHED;
		}
	}
}

struct WaveRaySource {
	const WaveState& W;
	const uint32_t* queue;
	uint32_t p;
	__device__ __forceinline__ void fetch(uint32_t i, vec3& o, vec3& d, float& rt, int& hf) {
		p = queue ? queue[i] : i;
		const float4 a = W.rayO[p], b = W.rayD[p];
		o = v3(a.x, a.y, a.z); rt = a.w;
		d = v3(b.x, b.y, b.z); hf = __float_as_int(b.w);
	}
	__device__ __forceinline__ bool wantsLeaf() const { return false; }
	__device__ __forceinline__ void store(uint32_t, const LaneRay& L) {
		W.rayO[p].w = L.rt;
		W.rayD[p].w = __int_as_float(L.hitFace);
		uint2 g = W.dbg[p];
		g.x += L.nn; g.y += L.nt;
		W.dbg[p] = g;
		if (W.hitN) W.hitN[p] = make_float4(L.normal.x, L.normal.y, L.normal.z, 0.0f);
	}
};

/* Closest-hit traversal of every live path.  queue pointer NULL = identity (first bounce). */
template <bool PHONG>
__global__ void __launch_bounds__(128, PHONG ? 1 : PT_TRAVERSE_MIN_BLOCKS) traverseKernel(
	const SceneDev S, const WaveState W, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ countPtr,
	uint32_t* cursor, uint32_t* countToReset, unsigned long long* stats
) {
	const uint32_t count = *countPtr;
	uint32_t nodes = 0, tris = 0, rays = 0;
	/* the other queue was consumed by the previous shade launch: empty it for the next one */
	if (blockIdx.x == 0 && threadIdx.x == 0) *countToReset = 0u;

	WaveRaySource src = {W, queue, 0u};
	traverseEngine<false, PHONG>(S, src, count, cursor, nodes, tris, rays);

	warpAddStat(stats + 0, rays);
	warpAddStat(stats + 2, nodes);
	warpAddStat(stats + 3, tris);
}


/* Append path p to a queue for every lane with `push` set; all 32 lanes must call. */
__device__ __forceinline__ void queueAppend(uint32_t* queue, uint32_t* count, const bool push, const uint32_t p) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned m = __ballot_sync(FULL, push);
	if (m == 0u) return;
	const int leader = __ffs(m) - 1;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(count, (uint32_t) __popc(m));
	base = __shfl_sync(FULL, base, leader);
	if (push) queue[base + (uint32_t) __popc(m & ((1u << lane) - 1u))] = p;
}

/* ------------------------------------------------------------------ shade */

/* SHADOW_PRE: the shadow rays of this wavefront were walked by traverseShadowKernel; shadowO[p].w is the result. */
template <int BRDF, bool SHADOW, bool PHONG, bool SHADOW_PRE = false>
__global__ void __launch_bounds__(128) shadeKernel(
	const FrameParams P, const WaveState W, const uint32_t* __restrict__ queueIn, const uint32_t* __restrict__ countInPtr,
	uint32_t* __restrict__ queueOut, uint32_t* countOutPtr, uint32_t* cursorToReset,
	uint32_t* reset1 = nullptr, uint32_t* reset2 = nullptr, const float4* __restrict__ shadowO = nullptr
) {
	const uint32_t count = *countInPtr;
	const uint32_t stride = gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	uint32_t shaded = 0, shadowNodes = 0, shadowRays = 0, trisAfter = 0;

	/* round the loop bound up to a warp multiple so that the ballot below is convergent */
	const uint32_t countUp = (count + 31u) & ~31u;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < countUp; i += stride) {
		bool alive = false;
		uint32_t p = 0;
		if (i < count) {
			p = queueIn ? queueIn[i] : i;
			PathState s;
			loadPath(W, p, s);
			int px, py;
			pixelOf(P, (int) p, px, py);

			if (s.t != PM_INF_F) shaded++;
			const uint32_t nTrisIn = s.nTris;
			const BounceResult r = SHADOW_PRE
				? bounce<BRDF, SHADOW, PHONG, true>(P, s, shadowNodes, shadowRays, shadowO[p].w)
				: bounce<BRDF, SHADOW, PHONG, false>(P, s, shadowNodes, shadowRays);
			trisAfter += s.nTris - nTrisIn;                /* shadow-ray triangle tests (advancePath may reset nTris) */
			alive = advancePath(P, s, r, px, py);
			if (alive) {
				storePath(W, p, s);
				W.dbg[p] = make_uint2(s.nNodes, s.nTris);
			}
		}
		/* stream compaction: one atomic per warp */
		const uint32_t m = __ballot_sync(0xffffffffu, alive);
		if (m) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(countOutPtr, (uint32_t) __popc(m));
			base = __shfl_sync(0xffffffffu, base, 0);
			if (alive) queueOut[base + __popc(m & ((1u << lane) - 1u))] = p;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		*cursorToReset = 0u;
		if (reset1) *reset1 = 0u;
		if (reset2) *reset2 = 0u;
	}
	warpAddStat(P.stats + 4, shaded);
	if (SHADOW) {
		warpAddStat(P.stats + 1, shadowRays);
		warpAddStat(P.stats + 5, shadowNodes);
		warpAddStat(P.stats + 3, trisAfter);
	}
}

/* ------------------------------------------------------------------ shadow rays as their own wavefront stage */

/*
 * With render.shadow_rays the reference shoots one shadow ray per hit from inside the bounce
 * (pathtracing.cl:282-288).  Walked inside the shade kernel, one thread per path, those rays run at the lane
 * utilisation of a plain per-thread loop (and double the frame time of the interior scene).  Instead:
 *     shadowGenKernel       per hit path: the bounce's prefix on a COPY of the path state decides whether a shadow
 *                           ray leaves this hit and which one (bouncePrefix + shadowRayOf: the same code bounce()
 *                           runs); ray -> shadowO / shadowD, path -> shadow queue
 *     traverseShadowKernel  the traversal engine in any-hit mode over that queue (traverseShadows, pt_bvh.cl:133-177);
 *                           result ray.t -> shadowO[p].w, its triangle tests -> the path's debug counter
 *     shadeKernel<.., SHADOW_PRE>  the bounce with the walked result plugged in
 */
template <int BRDF, bool PHONG>
__global__ void __launch_bounds__(128) shadowGenKernel(
	const FrameParams P, const WaveState W, const uint32_t* __restrict__ queueIn, const uint32_t* __restrict__ countInPtr,
	float4* __restrict__ shadowO, float4* __restrict__ shadowD, uint32_t* __restrict__ shadowQ, uint32_t* shadowCount
) {
	const uint32_t count = *countInPtr;
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t countUp = (count + 31u) & ~31u;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < countUp; i += stride) {
		bool shoot = false;
		uint32_t p = 0;
		if (i < count) {
			p = queueIn ? queueIn[i] : i;
			PathState s;
			loadPath(W, p, s);
			Material mtl;
			vec3 normal, hitPoint, lightDir;
			bool addDepth;
			float tLight;
			if (bouncePrefix<BRDF, PHONG>(P, s, mtl, normal, addDepth, hitPoint) == PREFIX_HIT &&
			    shadowRayOf(P.scene, mtl, hitPoint, lightDir, tLight)) {
				shadowO[p] = make_float4(hitPoint.x, hitPoint.y, hitPoint.z, tLight);
				shadowD[p] = make_float4(lightDir.x, lightDir.y, lightDir.z, 0.0f);
				shoot = true;
			}
		}
		queueAppend(shadowQ, shadowCount, shoot, p);
	}
}

struct ShadowRaySource {
	const WaveState& W;
	float4* shadowO;
	const float4* shadowD;
	const uint32_t* queue;
	uint32_t p;
	__device__ __forceinline__ void fetch(uint32_t i, vec3& o, vec3& d, float& rt, int& hf) {
		p = queue[i];
		const float4 a = shadowO[p], b = shadowD[p];
		o = v3(a.x, a.y, a.z); rt = a.w;
		d = v3(b.x, b.y, b.z); hf = 0;
	}
	__device__ __forceinline__ bool wantsLeaf() const { return false; }
	__device__ __forceinline__ void store(uint32_t, const LaneRay& L) {
		shadowO[p].w = L.rt;
		W.dbg[p].y += L.nt;                 /* shadow-ray face tests count into debugColor.x; node visits do not */
	}
};

template <bool PHONG>
__global__ void __launch_bounds__(128, PHONG ? 1 : PT_TRAVERSE_MIN_BLOCKS) traverseShadowKernel(
	const SceneDev S, const WaveState W, float4* shadowO, const float4* __restrict__ shadowD,
	const uint32_t* __restrict__ shadowQ, const uint32_t* __restrict__ countPtr, uint32_t* cursor, unsigned long long* stats
) {
	const uint32_t count = *countPtr;
	uint32_t nodes = 0, tris = 0, rays = 0;
	ShadowRaySource src = {W, shadowO, shadowD, shadowQ, 0u};
	traverseEngine<true, PHONG>(S, src, count, cursor, nodes, tris, rays);
	warpAddStat(stats + 1, rays);
	warpAddStat(stats + 5, nodes);
	warpAddStat(stats + 3, tris);
}

/* ------------------------------------------------------------------ megakernel (cross-check) */

template <int BRDF, bool SHADOW, bool PHONG>
__global__ void __launch_bounds__(128) megaKernel(const FrameParams P, const int nPaths) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t rays = 0, shaded = 0, shadowNodes = 0, shadowRays = 0, nodes = 0, tris = 0;
	PathState s;
	s.nNodes = 0; s.nTris = 0;
	s.hitNormal = v3(0.0f, 0.0f, 0.0f);
	if (p < nPaths) {
		int px, py;
		pixelOf(P, p, px, py);
		for (s.frame = 0u; s.frame < (uint32_t) P.frameCount; s.frame++) {
			initPath(P, s);
			for (; s.sample < (uint32_t) P.samples; s.sample++) {
				beginSample(P, s, px, py);
				while (true) {
					int hitLeaf = -1;
					traverseClosest<PHONG>(P.scene, s.o, s.d, s.t, s.hitFace, hitLeaf, s.nNodes, s.nTris, s.hitNormal);
					rays++;
					if (s.t != PM_INF_F) shaded++;
					if (bounce<BRDF, SHADOW, PHONG>(P, s, shadowNodes, shadowRays) != PATH_CONTINUE) break;
				}
			}
			finishPixel(P, s, px, py);
			nodes += s.nNodes; tris += s.nTris;
		}
	}
	warpAddStat(P.stats + 0, rays);
	warpAddStat(P.stats + 1, shadowRays);
	warpAddStat(P.stats + 2, nodes);
	warpAddStat(P.stats + 3, tris);
	warpAddStat(P.stats + 4, shaded);
	warpAddStat(P.stats + 5, shadowNodes);
}

/* ------------------------------------------------------------------ explicit rays (C5) */

struct ExplicitRaySource {
	const pbr_ray* __restrict__ rays;
	pbr_hit* __restrict__ hits;
	__device__ __forceinline__ void fetch(unsigned long long i, vec3& o, vec3& d, float& rt, int& hf) {
		const float4 a = __ldg((const float4*) &rays[i].origin);
		const float4 b = __ldg((const float4*) &rays[i].dir);
		o = v3(a.x, a.y, a.z);
		d = v3(b.x, b.y, b.z);
		rt = b.w;
		hf = 0;
	}
	__device__ __forceinline__ bool wantsLeaf() const { return true; }
	__device__ __forceinline__ void store(unsigned long long i, const LaneRay& L) {
		int4 out;
		out.x = __float_as_int(L.rt);
		out.y = L.hitFace;
		out.z = L.hitLeaf;
		out.w = (int) (min(L.nn, 0xfffffu) | (min(L.nt, 0xfffu) << 20));
		*((int4*) &hits[i]) = out;
	}
};

template <bool ANY_HIT>
__global__ void __launch_bounds__(128, PT_TRAVERSE_MIN_BLOCKS) traceRaysKernel(
	const SceneDev S, const pbr_ray* __restrict__ rays, const long long n, pbr_hit* __restrict__ hits,
	unsigned long long* cursor, unsigned long long* stats
) {
	uint32_t nodes = 0, tris = 0, cnt = 0;
	ExplicitRaySource src = {rays, hits};
	traverseEngine<ANY_HIT, false>(S, src, (unsigned long long) n, cursor, nodes, tris, cnt);
	warpAddStat(stats + (ANY_HIT ? 1 : 0), cnt);
	warpAddStat(stats + (ANY_HIT ? 5 : 2), nodes);
	warpAddStat(stats + 3, tris);
}

/* ------------------------------------------------------------------ scene repack */

/* bvhNode_cl[] (float-encoded indices, PathTracer.cpp:267-268,299,305) -> integer w lanes.
 * Entries past numNodes (padding so that index 1 always exists) become never-hit inner nodes. */
__global__ void repackNodesKernel(const float4* __restrict__ src, const int numSrc, float4* __restrict__ dst, const int numDst) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numDst) return;
	float4 lo, hi;
	if (i < numSrc) {
		lo = src[2 * (size_t) i];
		hi = src[2 * (size_t) i + 1];
		/* pt_bvh.cl:100,117: `bbMin.w <= -1` takes the miss link, `bbMin.w >= 0` tests faces; anything in between
		 * (or NaN) does neither -- the walk steps over such a node to cur + 1 whether its box is hit or not */
		const bool inner = (lo.w <= -1.0f), leaf = (lo.w >= 0.0f);
		const int loW = inner ? ((lo.w == -2.0f) ? -2 : -1) : (leaf ? (int) lo.w : -1);
		const int hiW = (inner || leaf) ? (int) hi.w : i + 1;
		lo.w = __int_as_float(loW);
		hi.w = __int_as_float(hiW);
	}
	else {
		lo = make_float4(PM_INF_F, PM_INF_F, PM_INF_F, __int_as_float(-1));
		hi = make_float4(-PM_INF_F, -PM_INF_F, -PM_INF_F, __int_as_float(-1));
	}
	dst[2 * (size_t) i] = lo;
	dst[2 * (size_t) i + 1] = hi;
}

/* facesV[f] + vertices[] -> a, b - a, c - a; material index   (pt_intersect.cl:146-149, :98-99) */
__global__ void repackTrisKernel(
	const uint4* __restrict__ facesV, const int numFaces, const float4* __restrict__ vertices, const int numVertices,
	float4* __restrict__ tris, float* __restrict__ trisB, uint32_t* __restrict__ triMat
) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= numFaces) return;
	const uint4 fv = facesV[f];
	const uint32_t last = (uint32_t) (numVertices - 1);
	const float4 a = vertices[min(fv.x, last)], b = vertices[min(fv.y, last)], c = vertices[min(fv.z, last)];
	tris[PT_TRI_STRIDE * (size_t) f] = make_float4(a.x, a.y, a.z, b.x - a.x);
	tris[PT_TRI_STRIDE * (size_t) f + 1] = make_float4(b.y - a.y, b.z - a.z, c.x - a.x, c.y - a.y);
	trisB[f] = c.z - a.z;
	triMat[f] = fv.w;
}

/* PHONGTESS: facesV[f], facesN[f] + vertices[], normals[] -> (a, material), (b, allNormalsEqual), (c, 0),
 * an, bn, cn  (pt_intersect.cl:146-160).  `an == bn` etc. are OpenCL component-wise float compares. */
__global__ void repackTrisPhongKernel(
	const uint4* __restrict__ facesV, const uint4* __restrict__ facesN, const int numFaces,
	const float4* __restrict__ vertices, const int numVertices, const float4* __restrict__ normals, const int numNormals,
	float4* __restrict__ tris
) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= numFaces) return;
	const uint4 fv = facesV[f], fn = facesN[f];
	const uint32_t lastV = (uint32_t) (numVertices - 1), lastN = (uint32_t) (numNormals - 1);
	const float4 a = vertices[min(fv.x, lastV)], b = vertices[min(fv.y, lastV)], c = vertices[min(fv.z, lastV)];
	const float4 an = normals[min(fn.x, lastN)], bn = normals[min(fn.y, lastN)], cn = normals[min(fn.z, lastN)];
	const bool allEqual = an.x == bn.x && an.y == bn.y && an.z == bn.z && bn.x == cn.x && bn.y == cn.y && bn.z == cn.z;
	float4* o = tris + PT_TRI_STRIDE_PHONG * (size_t) f;
	o[0] = make_float4(a.x, a.y, a.z, __int_as_float((int) fv.w));
	o[1] = make_float4(b.x, b.y, b.z, allEqual ? 1.0f : 0.0f);
	o[2] = make_float4(c.x, c.y, c.z, 0.0f);
	o[3] = make_float4(an.x, an.y, an.z, 0.0f);
	o[4] = make_float4(bn.x, bn.y, bn.z, 0.0f);
	o[5] = make_float4(cn.x, cn.y, cn.z, 0.0f);
}

/* ------------------------------------------------------------------ pinned-math probe */

__global__ void pinnedMathKernel(const int op, const float* __restrict__ x, const float* __restrict__ y, const long long n, float* __restrict__ out) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float a = x[i];
	float r = 0.0f;
	switch (op) {
		case 0: r = pm::sin_(a); break;
		case 1: r = pm::cos_(a); break;
		case 2: r = pm::tan_(a); break;
		case 3: r = pm::acos_(a); break;
		case 4: r = pm::atan_(a); break;
		case 5: r = pm::pow_(a, y[i]); break;
		case 6: r = pm::cbrt_(a); break;
		case 7: { float s = a; r = rnd(s); break; }
	}
	out[i] = r;
}

} /* namespace ptk */
