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
 *         shade         per live path: bounce(); paths that finish a sample start the next one
 *                       (seed carried over) or write their pixel (setColors) and leave; survivors
 *                       are appended to the next queue with one warp-aggregated atomic per warp
 *
 * Per-pixel arithmetic and its order are those of the reference kernel, so every pixel gets the
 * same bits whatever the scheduling.  `megakernel` keeps the reference's launch structure (one
 * thread per pixel, everything inline) as an on-device cross-check of the wavefront.
 *
 * Path state lives in HBM as 16-byte SoA records (coalesced 128-bit accesses):
 *     rayO (o.xyz, t)   rayD (d.xyz, hitFace)   colS (color.xyz, seed)   finF (finalColor.xyz, focus)
 *     misc (depth | depthAdded << 16, sample, secondaryPaths, -)   dbg (nodes visited, triangle tests)
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
	W.misc[p] = make_uint4(s.depth | ((uint32_t) s.depthAdded << 16), s.sample, s.secondaryPaths, 0u);
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
	s.sample = m.y; s.secondaryPaths = m.z;
	s.nNodes = g.x; s.nTris = g.y;
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
		pathToPixel(p, P.width, P.y0, P.y1 - P.y0, px, py);
		PathState s;
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

/* ------------------------------------------------------------------ traverse */

/* Persistent warps: each warp claims 32 queue slots with one atomic, every lane walks one ray.
 * qsel: which queue holds the live paths; queue pointer NULL = identity (first bounce). */
__global__ void __launch_bounds__(128) traverseKernel(
	const SceneDev S, const WaveState W, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ countPtr,
	uint32_t* cursor, uint32_t* countToReset, unsigned long long* stats
) {
	const uint32_t count = *countPtr;
	const int lane = threadIdx.x & 31;
	uint32_t nodes = 0, tris = 0, rays = 0;
	/* the other queue was consumed by the previous shade launch: empty it for the next one */
	if (blockIdx.x == 0 && threadIdx.x == 0) *countToReset = 0u;

	while (true) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= count) break;
		const uint32_t i = base + lane;
		if (i < count) {
			const uint32_t p = queue ? queue[i] : i;
			const float4 o = W.rayO[p], d = W.rayD[p];
			float rt = o.w;
			int hitFace = __float_as_int(d.w), hitLeaf = -1;
			uint32_t nn = 0, nt = 0;
			traverseClosest(S, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), rt, hitFace, hitLeaf, nn, nt);
			W.rayO[p].w = rt;
			W.rayD[p].w = __int_as_float(hitFace);
			uint2 g = W.dbg[p];
			g.x += nn; g.y += nt;
			W.dbg[p] = g;
			nodes += nn; tris += nt; rays++;
		}
	}
	warpAddStat(stats + 0, rays);
	warpAddStat(stats + 2, nodes);
	warpAddStat(stats + 3, tris);
}

/* ------------------------------------------------------------------ shade */

template <int BRDF, bool SHADOW>
__global__ void __launch_bounds__(128) shadeKernel(
	const FrameParams P, const WaveState W, const uint32_t* __restrict__ queueIn, const uint32_t* __restrict__ countInPtr,
	uint32_t* __restrict__ queueOut, uint32_t* countOutPtr, uint32_t* cursorToReset
) {
	const uint32_t count = *countInPtr;
	const uint32_t stride = gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	uint32_t shaded = 0, shadowNodes = 0, shadowRays = 0, trisBefore = 0, trisAfter = 0;

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
			pathToPixel((int) p, P.width, P.y0, P.y1 - P.y0, px, py);
			trisBefore += s.nTris;

			if (s.t != PM_INF_F) shaded++;
			const BounceResult r = bounce<BRDF, SHADOW>(P, s, shadowNodes, shadowRays);
			if (r == PATH_CONTINUE) {
				alive = true;
			}
			else {
				s.sample++;
				if (s.sample < (uint32_t) P.samples) {
					beginSample(P, s, px, py);
					alive = true;
				}
				else {
					finishPixel(P, s, px, py);
				}
			}
			trisAfter += s.nTris;
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
	if (blockIdx.x == 0 && threadIdx.x == 0) *cursorToReset = 0u;
	warpAddStat(P.stats + 4, shaded);
	if (SHADOW) {
		warpAddStat(P.stats + 1, shadowRays);
		warpAddStat(P.stats + 5, shadowNodes);
		warpAddStat(P.stats + 3, trisAfter - trisBefore);
	}
}

/* ------------------------------------------------------------------ megakernel (cross-check) */

template <int BRDF, bool SHADOW>
__global__ void __launch_bounds__(128) megaKernel(const FrameParams P, const int nPaths) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t rays = 0, shaded = 0, shadowNodes = 0, shadowRays = 0;
	PathState s;
	s.nNodes = 0; s.nTris = 0;
	if (p < nPaths) {
		int px, py;
		pathToPixel(p, P.width, P.y0, P.y1 - P.y0, px, py);
		initPath(P, s);
		for (; s.sample < (uint32_t) P.samples; s.sample++) {
			beginSample(P, s, px, py);
			while (true) {
				int hitLeaf = -1;
				traverseClosest(P.scene, s.o, s.d, s.t, s.hitFace, hitLeaf, s.nNodes, s.nTris);
				rays++;
				if (s.t != PM_INF_F) shaded++;
				if (bounce<BRDF, SHADOW>(P, s, shadowNodes, shadowRays) != PATH_CONTINUE) break;
			}
		}
		finishPixel(P, s, px, py);
	}
	else {
		s.nNodes = 0; s.nTris = 0;
	}
	warpAddStat(P.stats + 0, rays);
	warpAddStat(P.stats + 1, shadowRays);
	warpAddStat(P.stats + 2, s.nNodes);
	warpAddStat(P.stats + 3, s.nTris);
	warpAddStat(P.stats + 4, shaded);
	warpAddStat(P.stats + 5, shadowNodes);
}

/* ------------------------------------------------------------------ explicit rays (C5) */

template <bool ANY_HIT>
__global__ void __launch_bounds__(128) traceRaysKernel(
	const SceneDev S, const pbr_ray* __restrict__ rays, const long long n, pbr_hit* __restrict__ hits,
	unsigned long long* cursor, unsigned long long* stats
) {
	const int lane = threadIdx.x & 31;
	uint32_t nodes = 0, tris = 0, cnt = 0;
	while (true) {
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(cursor, 32ull);
		base = __shfl_sync(0xffffffffu, base, 0);
		if ((long long) base >= n) break;
		const long long i = (long long) base + lane;
		if (i < n) {
			const float4 o = __ldg((const float4*) &rays[i].origin);
			const float4 d = __ldg((const float4*) &rays[i].dir);
			float rt = d.w;
			int hitFace = 0, hitLeaf = -1;
			uint32_t nn = 0, nt = 0;
			if (ANY_HIT) traverseAny(S, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), rt, hitFace, hitLeaf, nn, nt);
			else traverseClosest(S, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), rt, hitFace, hitLeaf, nn, nt);
			nodes += nn; tris += nt; cnt++;
			int4 out;
			out.x = __float_as_int(rt);
			out.y = hitFace;
			out.z = hitLeaf;
			out.w = (int) (min(nn, 0xfffffu) | (min(nt, 0xfffu) << 20));
			*((int4*) &hits[i]) = out;
		}
	}
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
		const bool inner = (lo.w <= -1.0f);
		const int loW = inner ? -1 : (int) lo.w;
		const int hiW = (int) hi.w;
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

/* facesV[f] + vertices[] -> (a, material), b - a, c - a   (pt_intersect.cl:146-149, :98-99) */
__global__ void repackTrisKernel(
	const uint4* __restrict__ facesV, const int numFaces, const float4* __restrict__ vertices, const int numVertices,
	float4* __restrict__ tris
) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= numFaces) return;
	const uint4 fv = facesV[f];
	const uint32_t last = (uint32_t) (numVertices - 1);
	const float4 a = vertices[min(fv.x, last)], b = vertices[min(fv.y, last)], c = vertices[min(fv.z, last)];
	tris[3 * (size_t) f] = make_float4(a.x, a.y, a.z, __int_as_float((int) fv.w));
	tris[3 * (size_t) f + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, 0.0f);
	tris[3 * (size_t) f + 2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, 0.0f);
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
