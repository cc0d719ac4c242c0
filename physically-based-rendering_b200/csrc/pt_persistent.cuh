/*
 * pt_persistent.cuh -- the persistent pipeline: one traversal kernel and one shading kernel that are
 * resident TOGETHER for a whole frame and hand paths to each other through two ring buffers in HBM.
 *
 * Why.  In the bounce-by-bounce wavefront (pt_kernels.cuh) every traverse launch ends with a drain:
 * the queue is empty, warps retire one by one, and the launch lasts until its slowest ray is done.
 * A 1080p frame is only ~12 rays per resident lane per bounce, and the number of nodes a ray visits has
 * a long tail (mean 110-200, max > 1 400 on the 1 M-triangle soup), so that drain costs about as much
 * as the useful part: traversal time fits  0.68 ms + n / 700 Mrays/s  for incoherent rays.  Here there
 * are no per-bounce barriers: a path whose ray is finished goes straight to a shading warp, the bounced
 * ray goes straight back to a traversal warp, and the only drain is the one at the end of the frame.
 *
 *      raygenKernel ──► rayRing ──► persistTraverseKernel ──► hitRing ──► persistShadeKernel ──┐
 *                          ▲                                                                    │
 *                          └───────────── next bounce / next sample ◄───────────────────────────┘
 *                                                                  finished pixels: setColors, counter
 *
 * Rings.  Entry = path index + 1, 0 = empty; capacity a power of two >= the number of paths (a path is
 * in at most one ring).  Producers reserve slots with one warp-aggregated atomicAdd on the tail, wait
 * for the slot to be empty, write the path's data, __threadfence(), then publish the entry.  Consumers
 * take tickets with one warp-aggregated atomicAdd on the head and poll their entry (non-blocking in the
 * traversal engine: lanes that already have a ray keep walking); entries arrive in ticket order.
 * Consumers read path data with ld.global.cg (L2): L1 is not coherent between SMs.  The BVH is read-only
 * and stays on the L1 path.  Both kernels leave when `finished == total`; a spin watchdog raises `abort`
 * instead of hanging if the other kernel never becomes resident.
 *
 * Per-pixel arithmetic is untouched -- the same beginSample / bounce / finishPixel and the same node and
 * triangle steps -- so frames are bit-identical to the other pipelines and to the oracle.
 */
#pragma once

#include "pt_kernels.cuh"

namespace ptk {

/* Every counter in a 128-byte line of its own: they are hammered by atomics from every warp, and lines
 * map to different L2 slices. */
struct PersistCtl {
	uint32_t rayHead; uint32_t pad0[31];     /* tickets taken in rayRing */
	uint32_t rayTail; uint32_t pad1[31];     /* slots reserved in rayRing */
	uint32_t hitHead; uint32_t pad2[31];
	uint32_t hitTail; uint32_t pad3[31];
	uint32_t finished; uint32_t pad4[31];    /* pixels written this frame */
	uint32_t total;                          /* paths of this frame */
	uint32_t abort;                          /* watchdog tripped */
	uint32_t pad5[30];
	unsigned long long idleT, idleS;         /* diagnostics: idle polls of the two kernels, summed over warps */
	uint32_t pad6[28];
};

#define PERSIST_SPIN_LIMIT (1u << 22)     /* idle polls (with nanosleep) before a warp gives up */

__device__ __forceinline__ uint32_t ldVolatile(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void stVolatile(uint32_t* p, uint32_t v) {
	asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v));
}

/* Publish path p in `ring` for every lane with `push` set.  Must be reached by all 32 lanes. */
__device__ __forceinline__ void ringPush(uint32_t* ring, uint32_t* tail, const uint32_t mask, const bool push, const uint32_t p) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned m = __ballot_sync(FULL, push);
	if (m == 0u) return;
	__threadfence();                                   /* path data before the entry */
	uint32_t base = 0;
	const int leader = __ffs(m) - 1;
	if (lane == leader) base = atomicAdd(tail, (uint32_t) __popc(m));
	base = __shfl_sync(FULL, base, leader);
	if (push) {
		uint32_t* e = ring + ((base + (uint32_t) __popc(m & ((1u << lane) - 1u))) & mask);
		while (ldVolatile(e) != 0u) { }                /* slot of the previous lap not consumed yet */
		stVolatile(e, p + 1u);
	}
}

__device__ __forceinline__ void loadPathCG(const WaveState& W, const uint32_t p, PathState& s) {
	const float4 o = __ldcg(W.rayO + p), d = __ldcg(W.rayD + p), c = __ldcg(W.colS + p), f = __ldcg(W.finF + p);
	const uint4 m = __ldcg(W.misc + p);
	const uint2 g = __ldcg(W.dbg + p);
	s.o = v3(o.x, o.y, o.z); s.t = o.w;
	s.d = v3(d.x, d.y, d.z); s.hitFace = __float_as_int(d.w);
	s.color = v3(c.x, c.y, c.z); s.seed = c.w;
	s.finalColor = v3(f.x, f.y, f.z); s.focus = f.w;
	s.depth = m.x & 0xffffu; s.depthAdded = (int) (m.x >> 16);
	s.sample = m.y; s.secondaryPaths = m.z; s.frame = m.w;
	s.nNodes = g.x; s.nTris = g.y;
	if (W.hitN) { const float4 n = __ldcg(W.hitN + p); s.hitNormal = v3(n.x, n.y, n.z); }
	else s.hitNormal = v3(0.0f, 0.0f, 0.0f);
}

/* ------------------------------------------------------------------ raygen into the ring */

__global__ void __launch_bounds__(256) persistRaygenKernel(
	const FrameParams P, const WaveState W, PersistCtl* ctl, uint32_t* __restrict__ rayRing, const int nPaths
) {
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
		rayRing[p] = (uint32_t) p + 1u;                /* ring capacity >= nPaths: slot p, lap 0 */
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		ctl->rayHead = 0u; ctl->rayTail = (uint32_t) nPaths;
		ctl->hitHead = 0u; ctl->hitTail = 0u;
		ctl->finished = 0u; ctl->total = (uint32_t) nPaths;
		ctl->abort = 0u;
	}
}

/* ------------------------------------------------------------------ traversal, persistent */

enum { PLANE_IDLE = 0, PLANE_STEPPING = 1, PLANE_PENDING = 2, PLANE_FINISHED = 3, PLANE_WAITING = 4 };

template <bool PHONG>
__global__ void __launch_bounds__(128) persistTraverseKernel(
	const SceneDev S, const WaveState W, PersistCtl* ctl, uint32_t* rayRing, uint32_t* hitRing, const uint32_t mask,
	unsigned long long* stats
) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	const unsigned lastNode = (unsigned) (S.numNodes - 1);
	const uint32_t total = ctl->total;

	LaneRay L;
	L.index = 0;
	int state = PLANE_IDLE;
	uint32_t p = 0, ticket = 0;
	uint32_t nodes = 0, tris = 0, rays = 0, idlePolls = 0, idleTotal = 0;

	while (true) {
		/* retire (a few lanes at a time: one fence and one atomic per batch): results to the path state,
		 * path to the shading ring */
		const unsigned fin = __ballot_sync(FULL, state == PLANE_FINISHED);
		if (fin != 0u && (__popc(fin) >= S.refillMin || __ballot_sync(FULL, state == PLANE_STEPPING || state == PLANE_PENDING) == 0u)) {
			if (state == PLANE_FINISHED) {
				W.rayO[p].w = L.rt;
				W.rayD[p].w = __int_as_float(L.hitFace);
				uint2 g = __ldcg(W.dbg + p);
				g.x += L.nn; g.y += L.nt;
				W.dbg[p] = g;
				if (PHONG) W.hitN[p] = make_float4(L.normal.x, L.normal.y, L.normal.z, 0.0f);
				nodes += L.nn; tris += L.nt; rays++;
			}
			ringPush(hitRing, &ctl->hitTail, mask, state == PLANE_FINISHED, p);
			if (state == PLANE_FINISHED) state = PLANE_IDLE;
		}

		/* idle lanes take tickets for the next rays (not knowing yet whether they exist) */
		const unsigned need = __ballot_sync(FULL, state == PLANE_IDLE);
		if (__popc(need) >= S.refillMin || need == FULL) {
			const int leader = __ffs(need) - 1;
			uint32_t base = 0;
			if (lane == leader) base = atomicAdd(&ctl->rayHead, (uint32_t) __popc(need));
			base = __shfl_sync(FULL, base, leader);
			if (state == PLANE_IDLE) {
				ticket = base + (uint32_t) __popc(need & ltMask);
				state = PLANE_WAITING;
			}
		}

		/* waiting lanes look once whether their ray has arrived */
		if (state == PLANE_WAITING) {
			uint32_t* e = rayRing + (ticket & mask);
			/* two lanes (tickets one lap apart) may watch the same slot: the exchange decides who gets it */
			const uint32_t v = (ldVolatile(e) != 0u) ? atomicExch(e, 0u) : 0u;
			if (v != 0u) {
				p = v - 1u;                                /* the loads below depend on p and go to L2 */
				const float4 a = __ldcg(W.rayO + p), b = __ldcg(W.rayD + p);
				startRay<false>(S, L, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w, __float_as_int(b.w));
				state = (lastNode >= 1u) ? PLANE_STEPPING : PLANE_FINISHED;
			}
		}

		/* nobody in this warp has a ray: done, or wait for the shading kernel to send more */
		if (__ballot_sync(FULL, state == PLANE_STEPPING || state == PLANE_PENDING || state == PLANE_FINISHED) == 0u) {
			idleTotal++;
			if ((++idlePolls & 7u) == 0u) {
				if (ldVolatile(&ctl->finished) >= total || ldVolatile(&ctl->abort) != 0u) break;
				if (idlePolls > PERSIST_SPIN_LIMIT) {
					if (lane == 0) atomicExch(&ctl->abort, 1u);
					break;
				}
			}
			__nanosleep(idlePolls < 8u ? 100u : 400u);
			continue;
		}
		idlePolls = 0;

		/* node phase */
		while (true) {
			if (state == PLANE_STEPPING) {
				const bool leaf = nodeStep<false>(S, L);
				const bool inside = (unsigned) (L.index - 1) < lastNode;
				state = leaf ? PLANE_PENDING : (inside ? PLANE_STEPPING : PLANE_FINISHED);
			}
			if (__popc(__ballot_sync(FULL, state == PLANE_STEPPING)) < S.nodePhaseMin) break;
		}

		/* triangle phase */
		if (state == PLANE_PENDING) {
			leafStep<false, PHONG>(S, L);
			state = ((unsigned) (L.index - 1) < lastNode) ? PLANE_STEPPING : PLANE_FINISHED;
		}
	}
	warpAddStat(stats + 0, rays);
	warpAddStat(stats + 2, nodes);
	warpAddStat(stats + 3, tris);
	if (lane == 0) atomicAdd(&ctl->idleT, (unsigned long long) idleTotal);
}

/* ------------------------------------------------------------------ shading, persistent */

/* fillPolls: once SOME lane of a warp has a hit to shade, how many more polls the warp spends waiting for
 * the entries of its other lanes before it shades with the lanes it has (a warp must never wait for ALL of
 * its lanes: near the end of a frame the only live paths may be the ones it is holding). */
template <int BRDF, bool SHADOW, bool PHONG>
__global__ void __launch_bounds__(128) persistShadeKernel(
	const FrameParams P, const WaveState W, PersistCtl* ctl, uint32_t* rayRing, uint32_t* hitRing, const uint32_t mask,
	const int fillPolls
) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	const uint32_t total = ctl->total;
	uint32_t shaded = 0, shadowNodes = 0, shadowRays = 0, shadowTris = 0, idlePolls = 0, idleTotal = 0;
	bool waiting = false;
	uint32_t ticket = 0;

	while (true) {
		/* lanes without a ticket take the next ones */
		const unsigned need = __ballot_sync(FULL, !waiting);
		if (need != 0u) {
			const int leader = __ffs(need) - 1;
			uint32_t base = 0;
			if (lane == leader) base = atomicAdd(&ctl->hitHead, (uint32_t) __popc(need));
			base = __shfl_sync(FULL, base, leader);
			if (!waiting) {
				ticket = base + (uint32_t) __popc(need & ltMask);
				waiting = true;
			}
		}
		uint32_t* e = hitRing + (ticket & mask);

		/* poll together until every lane has a hit, or some have and the others are slow to arrive */
		uint32_t v = 0;
		int tries = 0;
		bool over = false;
		while (true) {
			if (v == 0u && ldVolatile(e) != 0u) v = atomicExch(e, 0u);
			const unsigned got = __ballot_sync(FULL, v != 0u);
			if (got == FULL) break;
			if (got != 0u) {
				if (++tries > fillPolls) break;
				continue;
			}
			idleTotal++;
			if ((++idlePolls & 7u) == 0u) {
				bool end = false;
				if (lane == 0) end = ldVolatile(&ctl->finished) >= total || ldVolatile(&ctl->abort) != 0u;
				if (__any_sync(FULL, end)) { over = true; break; }
				if (idlePolls > PERSIST_SPIN_LIMIT) {
					if (lane == 0) atomicExch(&ctl->abort, 1u);
					over = true;
					break;
				}
			}
			__nanosleep(idlePolls < 8u ? 100u : 400u);
		}
		if (over) break;
		idlePolls = 0;

		bool alive = false, done = false;
		uint32_t p = 0;
		if (v != 0u) {
			waiting = false;
			p = v - 1u;
			PathState s;
			loadPathCG(W, p, s);
			int px, py;
			pixelOf(P, (int) p, px, py);
			const uint32_t trisBefore = s.nTris;
			if (s.t != PM_INF_F) shaded++;
			const BounceResult r = bounce<BRDF, SHADOW, PHONG>(P, s, shadowNodes, shadowRays);
			shadowTris += s.nTris - trisBefore;
			alive = advancePath(P, s, r, px, py);
			done = !alive;
			if (alive) {
				storePath(W, p, s);
				W.dbg[p] = make_uint2(s.nNodes, s.nTris);
			}
		}
		__syncwarp(FULL);
		ringPush(rayRing, &ctl->rayTail, mask, alive, p);
		const unsigned dm = __ballot_sync(FULL, done);
		if (dm != 0u) {
			__threadfence();                               /* the pixels before the count */
			if (lane == (__ffs(dm) - 1)) atomicAdd(&ctl->finished, (uint32_t) __popc(dm));
		}
	}
	warpAddStat(P.stats + 4, shaded);
	if (SHADOW) {
		warpAddStat(P.stats + 1, shadowRays);
		warpAddStat(P.stats + 5, shadowNodes);
		warpAddStat(P.stats + 3, shadowTris);
	}
	if (lane == 0) atomicAdd(&ctl->idleS, (unsigned long long) idleTotal);
}

} /* namespace ptk */
