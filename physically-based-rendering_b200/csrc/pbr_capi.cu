/*
 * pbr_capi.cu -- implementation of the C ABI declared in include/pbr_b200.h.
 *
 * Stands where the reference's `CL` class (source/CL.cpp) stands: owns the device, the buffers and
 * images, the "program" (a set of precompiled sm_100a kernel specialisations selected by the values
 * the reference would splice into pt_header.cl) and the one kernel `pathTracing` with its 14
 * argument slots (PathTracer.cpp:88-125).  No CPU fallback: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>

#include "../../include/pbr_b200.h"
#include "pt_kernels.cuh"
#include "pt_persistent.cuh"

using namespace ptk;

namespace {

struct Mem {
	void* dptr = nullptr;
	size_t bytes = 0;
	bool image = false;
	size_t width = 0, height = 0;
	bool alive = false;
};

struct KernelArgs {
	float seed = 0.0f, pixelWeight = 0.0f, pxDim = 0.0f;
	pbr_camera cam;
	pbr_mem mem[14] = {0};
	uint32_t setMask = 0;
};

} /* namespace */

struct pbr_ctx {
	int device = 0;
	int smCount = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t ownStream = nullptr;
	cudaStream_t copyStream = nullptr;         /* pbr_image_read_begin / _end */
	cudaEvent_t evCopy = nullptr;
	bool copyInFlight = false;
	cudaEvent_t evStart = nullptr, evStop = nullptr;

	/* per-kernel profiling */
	bool profiling = false;
	struct Timed { cudaEvent_t a, b; int kind; };
	std::vector<Timed> timedInFlight;
	std::vector<cudaEvent_t> eventPool;
	pbr_profile prof = {};
	bool timed = false;
	std::string lastError;
	std::vector<Mem> mems;
	std::vector<void*> pinned;

	/* program */
	bool programLoaded = false;
	pbr_defines defines;
	bool haveNumNodes = false, haveNumLights = false, haveSky = false;
	int defNumNodes = 0, defNumLights = 0;
	pbr_float4 defSky;

	KernelArgs args;
	int tileY0 = -1, tileY1 = -1;
	int stripeRows = 0, stripeWorld = 1, stripeRank = 0;
	pbr_mem scratchImage = 0;                  /* pbr_kernel_launch_batch with depth of field */
	int pipeline = 0;
	/* pipeline selection by measurement (pbr_set_pipeline(-1), the default): the first frame of a configuration
	 * runs as a wavefront, the second as the megakernel, both timed with events; whichever was faster renders
	 * the rest.  All pipelines write the same bits, so the switch is invisible in the image. */
	bool pipelineAuto = true;
	int autoState = 0;                          /* 0 warm-up frame (wavefront, untimed: allocations happen here); 1, 3 time
	                                               the wavefront; 2, 4 time the megakernel; 5 decide; 6 decided */
	int autoChoice = 0;
	unsigned long long autoKey = 0;
	cudaEvent_t evAuto[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	bool debugImage = true;

	/* repacked scene cache */
	float4* nodes = nullptr;
	float4* tris = nullptr;
	const float* trisB = nullptr;
	const uint32_t* triMat = nullptr;
#if PT_NODE_ORDER
	int* nodeOrig = nullptr;                   /* permuted position -> index in the reference's array */
	size_t nodeOrigCap = 0;
#endif
	size_t nodesCap = 0, trisCap = 0;
	pbr_mem cacheBvh = 0, cacheFacesV = 0, cacheVertices = 0, cacheFacesN = 0, cacheNormals = 0;
	bool cachePhong = false;
	int cacheNumNodes = -1;
	uint64_t sceneEpoch = 0, cacheEpoch = ~0ull;
	int numNodesDev = 0;

	/* wavefront state */
	WaveState wave = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	float4* hitN = nullptr;                    /* allocated with the wave state, used when PHONGTESS */
	QueueCtl qctl = {nullptr, {nullptr, nullptr}};
	size_t waveCap = 0;
	/* shadow rays as a wavefront stage (render.shadow_rays): ray per path, queue; qctl.ctrl[3] count, [4] cursor */
	float4* shadowO = nullptr;
	float4* shadowD = nullptr;
	uint32_t* shadowQ = nullptr;
	size_t shadowCap = 0;
	int shadowStage = 1;                       /* tuning "shadow_stage": 0 = walk shadow rays inside the shade kernel */

	/* carry-over wavefront (pipeline 3): resume nodes, hit queue, two carry queues, counters, host mailbox */
	int* waveNode = nullptr;
	uint32_t* hitQ = nullptr;
	uint32_t* carryQ[2] = {nullptr, nullptr};
	uint32_t* cctl = nullptr;                  /* nNew[2], nCarry[2], nHit[2], cursor, - */
	uint32_t* mailbox = nullptr;               /* pinned, MAILBOX_SLOTS words */
	cudaEvent_t evGroup[2] = {nullptr, nullptr};
	int tailStepsBulk = 64, tailStepsFlush = 256, flushGroup = 4;
	int prevGroupBegin = 0;
	int traverseBlocks = 0;                    /* tuning: cap on resident traverse blocks per SM (0 = all that fit) */
	int batchInterleave = 0;                   /* pbr_kernel_launch_batch: let pixels run ahead into later frames */

	/* can any material extend a path beyond MAX_DEPTH? (decides how many wavefront iterations to launch) */
	pbr_mem extendCacheMem = 0;
	uint64_t extendCacheEpoch = ~0ull;
	int extendCacheBrdf = -1;
	bool canExtendDepth = true;
	int nodePhaseMin = 16;                     /* PBR_NODE_PHASE_MIN overrides (tuning) */
	int refillMin = 4;                         /* PBR_REFILL_MIN overrides (tuning) */
	/* persistent pipeline (pipeline 2): rings between the two resident kernels, second stream */
	PersistCtl* pctl = nullptr;
	uint32_t* ring[2] = {nullptr, nullptr};    /* rayRing, hitRing */
	size_t ringCap = 0;                        /* entries, power of two */
	cudaStream_t shadeStream = nullptr;
	cudaEvent_t evFork = nullptr, evJoin = nullptr;
	bool persistUsed = false;                  /* a pipeline-2 frame ran since the last abort check */
	int persistTBlocks = 0, persistSBlocks = 2, persistFill = 4;   /* T: 0 = what fits;  PBR_PERSIST_T / _S / _FILL override */
	unsigned long long* stats = nullptr;       /* 6 counters */
	unsigned long long* cursor64 = nullptr;    /* work cursor of traceRaysKernel */
};

namespace {

int fail(pbr_ctx* ctx, int code, const std::string& msg) {
	if (ctx) ctx->lastError = msg;
	return code;
}

int cudaFail(pbr_ctx* ctx, cudaError_t e, const char* what) {
	if (e == cudaSuccess) return PBR_OK;
	std::string m = std::string(what) + ": " + cudaGetErrorString(e);
	if (ctx) ctx->lastError = m;
	return (int) e;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cudaFail(ctx, e_, #call); } while (0)

Mem* getMem(pbr_ctx* ctx, pbr_mem h) {
	if (h == 0 || h > ctx->mems.size()) return nullptr;
	Mem* m = &ctx->mems[h - 1];
	return m->alive ? m : nullptr;
}

int newMem(pbr_ctx* ctx, size_t bytes, pbr_mem* out, Mem** mp) {
	Mem m;
	m.bytes = bytes;
	m.alive = true;
	cudaError_t e = cudaMalloc(&m.dptr, bytes > 0 ? bytes : 16);
	if (e != cudaSuccess) return cudaFail(ctx, e, "cudaMalloc");
	ctx->mems.push_back(m);
	*out = (pbr_mem) ctx->mems.size();
	if (mp) *mp = &ctx->mems.back();
	ctx->sceneEpoch++;
	return PBR_OK;
}

enum { MAILBOX_SLOTS = 256 };

int gridFor(long long n, int block) { return (int) ((n + block - 1) / block); }

enum KernelKind { K_RAYGEN = 0, K_TRAVERSE = 1, K_SHADE = 2, K_OTHER = 3 };

cudaEvent_t takeEvent(pbr_ctx* ctx) {
	if (!ctx->eventPool.empty()) {
		cudaEvent_t e = ctx->eventPool.back();
		ctx->eventPool.pop_back();
		return e;
	}
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	return e;
}

/* Brackets one kernel launch: counts it and, when profiling is on, times it with two events. */
struct LaunchScope {
	pbr_ctx* ctx;
	cudaEvent_t a = nullptr, b = nullptr;
	int kind;
	cudaStream_t stream;
	LaunchScope(pbr_ctx* c, int k, cudaStream_t st = nullptr) : ctx(c), kind(k), stream(st ? st : c->stream) {
		ctx->prof.launches++;
		switch (kind) {
			case K_RAYGEN: ctx->prof.raygen_launches++; break;
			case K_TRAVERSE: ctx->prof.traverse_launches++; break;
			case K_SHADE: ctx->prof.shade_launches++; break;
			default: ctx->prof.other_launches++; break;
		}
		if (ctx->profiling) {
			a = takeEvent(ctx);
			b = takeEvent(ctx);
			cudaEventRecord(a, stream);
		}
	}
	~LaunchScope() {
		if (a) {
			cudaEventRecord(b, stream);
			pbr_ctx::Timed t = {a, b, kind};
			ctx->timedInFlight.push_back(t);
		}
	}
};

/* Fold finished event pairs into the profile (the stream must be idle). */
void drainTimed(pbr_ctx* ctx) {
	for (const pbr_ctx::Timed& t : ctx->timedInFlight) {
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
			switch (t.kind) {
				case K_RAYGEN: ctx->prof.raygen_ms += ms; break;
				case K_TRAVERSE: ctx->prof.traverse_ms += ms; break;
				case K_SHADE: ctx->prof.shade_ms += ms; break;
				default: ctx->prof.other_ms += ms; break;
			}
		}
		ctx->eventPool.push_back(t.a);
		ctx->eventPool.push_back(t.b);
	}
	ctx->timedInFlight.clear();
}

/* Rebuild the repacked node / triangle arrays when the bound buffers or BVH_NUM_NODES changed. */
int ensureScene(pbr_ctx* ctx, pbr_mem hBvh, pbr_mem hFacesV, pbr_mem hVertices, int numNodes,
                bool phong = false, pbr_mem hFacesN = 0, pbr_mem hNormals = 0) {
	if (ctx->cacheBvh == hBvh && ctx->cacheFacesV == hFacesV && ctx->cacheVertices == hVertices &&
	    ctx->cacheNumNodes == numNodes && ctx->cacheEpoch == ctx->sceneEpoch && ctx->cachePhong == phong &&
	    (!phong || (ctx->cacheFacesN == hFacesN && ctx->cacheNormals == hNormals))) {
		return PBR_OK;
	}
	Mem* bvh = getMem(ctx, hBvh);
	Mem* facesV = getMem(ctx, hFacesV);
	Mem* vertices = getMem(ctx, hVertices);
	if (!bvh || !facesV || !vertices) return fail(ctx, PBR_ERR_INVALID, "pathTracing: bvh / facesV / vertices argument is not a live buffer");

	const int numSrcNodes = (int) (bvh->bytes / sizeof(pbr_bvh_node));
	if (numNodes > numSrcNodes) return fail(ctx, PBR_ERR_INVALID, "BVH_NUM_NODES exceeds the size of the bvh buffer");
	if (numNodes > (1 << 24)) return fail(ctx, PBR_ERR_INVALID, "BVH_NUM_NODES exceeds 2^24: float-encoded indices are no longer exact");
	const int numFaces = (int) (facesV->bytes / sizeof(pbr_uint4));
	const int numVertices = (int) (vertices->bytes / sizeof(pbr_float4));
	const int numDst = numNodes < 2 ? 2 : numNodes;

	if ((size_t) numDst > ctx->nodesCap) {
		if (ctx->nodes) cudaFree(ctx->nodes);
		ctx->nodes = nullptr;
		CK(cudaMalloc(&ctx->nodes, (size_t) numDst * 32));
		ctx->nodesCap = (size_t) numDst;
	}
	Mem* facesN = phong ? getMem(ctx, hFacesN) : nullptr;
	Mem* normals = phong ? getMem(ctx, hNormals) : nullptr;
	if (phong && (!facesN || !normals || facesN->bytes < facesV->bytes || normals->bytes < sizeof(pbr_float4)))
		return fail(ctx, PBR_ERR_INVALID, "pathTracing: PHONGTESS needs facesN (one entry per face) and normals");
	/* float4s: PHONGTESS 6 per face; otherwise 2 per face, then a quarter float4 (edge2.z) per face behind them,
	 * then a quarter float4 (material index) per face behind those */
	const size_t facesAlloc = (size_t) (numFaces > 0 ? numFaces : 1);
	const size_t quarter = (facesAlloc + 3) / 4;
	const size_t wantTris = phong ? facesAlloc * PT_TRI_STRIDE_PHONG : facesAlloc * PT_TRI_STRIDE + 2 * quarter;
	if (wantTris > ctx->trisCap || !ctx->tris) {
		if (ctx->tris) cudaFree(ctx->tris);
		ctx->tris = nullptr;
		CK(cudaMalloc(&ctx->tris, wantTris * 16));
		ctx->trisCap = wantTris;
	}
	{
		LaunchScope ls(ctx, K_OTHER);
		repackNodesKernel<<<gridFor(numDst, 256), 256, 0, ctx->stream>>>((const float4*) bvh->dptr, numNodes, ctx->nodes, numDst);
	}
#if PT_NODE_ORDER
	{
		/* Experiment (scripts/node_permutation_proto.py): hot nodes dense, explicit links.  Done on the host, once per
		 * scene: positions 0 and 1 stay, the rest is ordered by surface area, ties in the reference's order. */
		const int N = numDst;
		std::vector<float> h((size_t) N * 8), out((size_t) N * 8);
		CK(cudaStreamSynchronize(ctx->stream));
		CK(cudaMemcpy(h.data(), ctx->nodes, (size_t) N * 32, cudaMemcpyDeviceToHost));
		std::vector<double> key((size_t) N);
		for (int i = 0; i < N; i++) {
			const float* r = &h[(size_t) i * 8];
			const float ex = fmaxf(r[4] - r[0], 0.0f), ey = fmaxf(r[5] - r[1], 0.0f), ez = fmaxf(r[6] - r[2], 0.0f);
			key[(size_t) i] = (double) ex * ey + (double) ey * ez + (double) ex * ez;
		}
		std::vector<int> orig((size_t) N), pos((size_t) N);
		std::iota(orig.begin(), orig.end(), 0);
		std::stable_sort(orig.begin() + (N > 2 ? 2 : N), orig.end(), [&](int a, int b) { return key[(size_t) a] > key[(size_t) b]; });
		for (int j = 0; j < N; j++) pos[(size_t) orig[(size_t) j]] = j;
		const auto link = [&](long long t) -> int { return (t > 0 && t < (long long) numNodes) ? pos[(size_t) t] : 0; };
		for (int j = 0; j < N; j++) {
			const int i = orig[(size_t) j];
			const float* r = &h[(size_t) i * 8];
			float* o = &out[(size_t) j * 8];
			memcpy(o, r, 32);
			int loW, hiW;
			memcpy(&loW, r + 3, 4);
			memcpy(&hiW, r + 7, 4);
			int nlo, nhi;
			if (loW < 0) { nlo = link((long long) i + 1); nhi = link(hiW); }
			else { nlo = loW; nhi = (int) ((unsigned) link((long long) i + 1) | 0x80000000u | (hiW != -1 ? 0x40000000u : 0u)); }
			memcpy(o + 3, &nlo, 4);
			memcpy(o + 7, &nhi, 4);
		}
		CK(cudaMemcpy(ctx->nodes, out.data(), (size_t) N * 32, cudaMemcpyHostToDevice));
		if ((size_t) N > ctx->nodeOrigCap) {
			if (ctx->nodeOrig) cudaFree(ctx->nodeOrig);
			ctx->nodeOrig = nullptr;
			CK(cudaMalloc(&ctx->nodeOrig, (size_t) N * sizeof(int)));
			ctx->nodeOrigCap = (size_t) N;
		}
		CK(cudaMemcpy(ctx->nodeOrig, orig.data(), (size_t) N * sizeof(int), cudaMemcpyHostToDevice));
	}
#endif
	if (numFaces > 0 && numVertices > 0) {
		LaunchScope ls(ctx, K_OTHER);
		if (phong) {
			repackTrisPhongKernel<<<gridFor(numFaces, 256), 256, 0, ctx->stream>>>(
				(const uint4*) facesV->dptr, (const uint4*) facesN->dptr, numFaces, (const float4*) vertices->dptr, numVertices,
				(const float4*) normals->dptr, (int) (normals->bytes / sizeof(pbr_float4)), ctx->tris);
		}
		else {
			repackTrisKernel<<<gridFor(numFaces, 256), 256, 0, ctx->stream>>>(
				(const uint4*) facesV->dptr, numFaces, (const float4*) vertices->dptr, numVertices, ctx->tris,
				(float*) (ctx->tris + PT_TRI_STRIDE * facesAlloc), (uint32_t*) (ctx->tris + PT_TRI_STRIDE * facesAlloc + quarter));
		}
	}
	CK(cudaGetLastError());
	ctx->cacheBvh = hBvh; ctx->cacheFacesV = hFacesV; ctx->cacheVertices = hVertices;
	ctx->cacheFacesN = hFacesN; ctx->cacheNormals = hNormals; ctx->cachePhong = phong;
	ctx->cacheNumNodes = numNodes;
	ctx->cacheEpoch = ctx->sceneEpoch;
	ctx->numNodesDev = numNodes;
	ctx->trisB = phong ? nullptr : (const float*) (ctx->tris + PT_TRI_STRIDE * facesAlloc);
	ctx->triMat = phong ? nullptr : (const uint32_t*) (ctx->tris + PT_TRI_STRIDE * facesAlloc + quarter);
	return PBR_OK;
}

int ensureWave(pbr_ctx* ctx, size_t nPaths) {
	if (nPaths <= ctx->waveCap) return PBR_OK;
	WaveState& W = ctx->wave;
	cudaFree(W.rayO); cudaFree(W.rayD); cudaFree(W.colS); cudaFree(W.finF); cudaFree(W.misc); cudaFree(W.dbg); cudaFree(ctx->hitN);
	cudaFree(ctx->qctl.queue[0]); cudaFree(ctx->qctl.queue[1]);
	cudaFree(ctx->waveNode); cudaFree(ctx->hitQ); cudaFree(ctx->carryQ[0]); cudaFree(ctx->carryQ[1]);
	ctx->waveNode = nullptr; ctx->hitQ = nullptr; ctx->carryQ[0] = ctx->carryQ[1] = nullptr;
	W = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	ctx->hitN = nullptr;
	ctx->qctl.queue[0] = ctx->qctl.queue[1] = nullptr;
	ctx->waveCap = 0;
	CK(cudaMalloc(&W.rayO, nPaths * 16));
	CK(cudaMalloc(&W.rayD, nPaths * 16));
	CK(cudaMalloc(&W.colS, nPaths * 16));
	CK(cudaMalloc(&W.finF, nPaths * 16));
	CK(cudaMalloc(&W.misc, nPaths * 16));
	CK(cudaMalloc(&W.dbg, nPaths * 8));
	CK(cudaMalloc(&ctx->hitN, nPaths * 16));
	CK(cudaMalloc(&ctx->qctl.queue[0], nPaths * 4));
	CK(cudaMalloc(&ctx->qctl.queue[1], nPaths * 4));
	CK(cudaMalloc(&ctx->waveNode, nPaths * 4));
	CK(cudaMalloc(&ctx->hitQ, nPaths * 4));
	CK(cudaMalloc(&ctx->carryQ[0], nPaths * 4));
	CK(cudaMalloc(&ctx->carryQ[1], nPaths * 4));
	if (!ctx->cctl) {
		CK(cudaMalloc(&ctx->cctl, 8 * sizeof(uint32_t)));
		CK(cudaMemset(ctx->cctl, 0, 8 * sizeof(uint32_t)));
		CK(cudaHostAlloc(&ctx->mailbox, MAILBOX_SLOTS * sizeof(uint32_t), cudaHostAllocMapped));
		CK(cudaEventCreateWithFlags(&ctx->evGroup[0], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&ctx->evGroup[1], cudaEventDisableTiming));
	}
	ctx->waveCap = nPaths;
	return PBR_OK;
}

int ensureShadow(pbr_ctx* ctx, size_t nPaths) {
	if (nPaths <= ctx->shadowCap) return PBR_OK;
	cudaFree(ctx->shadowO); cudaFree(ctx->shadowD); cudaFree(ctx->shadowQ);
	ctx->shadowO = ctx->shadowD = nullptr;
	ctx->shadowQ = nullptr;
	ctx->shadowCap = 0;
	CK(cudaMalloc(&ctx->shadowO, nPaths * 16));
	CK(cudaMalloc(&ctx->shadowD, nPaths * 16));
	CK(cudaMalloc(&ctx->shadowQ, nPaths * 4));
	ctx->shadowCap = nPaths;
	return PBR_OK;
}

int ensureRings(pbr_ctx* ctx, size_t nPaths) {
	if (nPaths <= ctx->ringCap / 2 && ctx->pctl) return PBR_OK;
	size_t cap = 1024;
	while (cap < 2 * nPaths) cap <<= 1;
	cudaFree(ctx->ring[0]); cudaFree(ctx->ring[1]);
	ctx->ring[0] = ctx->ring[1] = nullptr;
	ctx->ringCap = 0;
	if (!ctx->pctl) CK(cudaMalloc(&ctx->pctl, sizeof(PersistCtl)));
	if (!ctx->shadeStream) CK(cudaStreamCreateWithFlags(&ctx->shadeStream, cudaStreamNonBlocking));
	if (!ctx->evFork) CK(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
	if (!ctx->evJoin) CK(cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
	CK(cudaMalloc(&ctx->ring[0], cap * 4));
	CK(cudaMalloc(&ctx->ring[1], cap * 4));
	CK(cudaMemsetAsync(ctx->ring[0], 0, cap * 4, ctx->stream));
	CK(cudaMemsetAsync(ctx->ring[1], 0, cap * 4, ctx->stream));
	CK(cudaMemsetAsync(ctx->pctl, 0, sizeof(PersistCtl), ctx->stream));
	ctx->ringCap = cap;
	return PBR_OK;
}

/* After a pipeline-2 frame: did a watchdog fire?  (The stream must be idle.) */
int checkPersistAbort(pbr_ctx* ctx) {
	if (!ctx->persistUsed || !ctx->pctl) return PBR_OK;
	ctx->persistUsed = false;
	PersistCtl h;
	CK(cudaMemcpyAsync(&h, ctx->pctl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (getenv("PBR_PERSIST_DEBUG")) {
		fprintf(stderr, "[persist] idle polls: traverse %llu shade %llu\n", h.idleT, h.idleS);
		cudaMemsetAsync(&ctx->pctl->idleT, 0, 16, ctx->stream);
	}
	if (h.abort != 0u || h.finished != h.total) {
		cudaMemsetAsync(ctx->ring[0], 0, ctx->ringCap * 4, ctx->stream);
		cudaMemsetAsync(ctx->ring[1], 0, ctx->ringCap * 4, ctx->stream);
		cudaMemsetAsync(ctx->pctl, 0, sizeof(PersistCtl), ctx->stream);
		cudaStreamSynchronize(ctx->stream);
		char msg[160];
		snprintf(msg, sizeof(msg), "pathTracing: persistent pipeline gave up (abort=%u, %u of %u pixels written)", h.abort, h.finished, h.total);
		return fail(ctx, PBR_ERR_INVALID, msg);
	}
	return PBR_OK;
}

/* extendDepth (pt_utils.cl:89-96) and the transparency branch of getNewRay (pt_brdf.cl:352-354) are the
 * only places that set addDepth.  If no material can trigger either, depthAdded stays 0 and a path ends
 * after MAX_DEPTH bounces: the MAX_ADDED_DEPTH extra wavefront iterations would all be empty. */
int updateCanExtendDepth(pbr_ctx* ctx, pbr_mem hMaterials, int brdf) {
	if (ctx->extendCacheMem == hMaterials && ctx->extendCacheEpoch == ctx->sceneEpoch && ctx->extendCacheBrdf == brdf) return PBR_OK;
	Mem* m = getMem(ctx, hMaterials);
	if (!m) return fail(ctx, PBR_ERR_INVALID, "pathTracing: materials argument is not a live buffer");
	const size_t stride = brdf == 0 ? sizeof(pbr_material_schlick) : sizeof(pbr_material_sa);
	const size_t n = m->bytes / stride;
	std::vector<float> host(m->bytes / 4 + 1);
	if (n > 0) {
		CK(cudaMemcpyAsync(host.data(), m->dptr, n * stride, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	bool can = false;
	for (size_t i = 0; i < n; i++) {
		const float* d = host.data() + i * stride / 4;     /* d, Ni, (p|nu), (rough|nv) */
		if (!(d[0] >= 1.0f)) can = true;                    /* d < 1 (or NaN): transparency coin */
		if (brdf == 1 ? !(fmaxf(d[2], d[3]) < 50.0f) : !(d[3] >= 1.0f)) can = true;
	}
	ctx->canExtendDepth = can;
	ctx->extendCacheMem = hMaterials;
	ctx->extendCacheEpoch = ctx->sceneEpoch;
	ctx->extendCacheBrdf = brdf;
	return PBR_OK;
}

/* pipeline 2: two kernels resident together for the whole frame (pt_persistent.cuh) */
template <int BRDF, bool SHADOW, bool PHONG>
int runPersistent(pbr_ctx* ctx, const FrameParams& P, WaveState W, int nPaths) {
	int rc = ensureRings(ctx, (size_t) nPaths);
	if (rc) return rc;
	const uint32_t mask = (uint32_t) ctx->ringCap - 1u;
	/* Both kernels must be resident at once: give the shading kernel its blocks per SM and the
	 * traversal kernel what is left of the register file and the thread slots. */
	static int regsT = 0, regsS = 0;
	if (regsT == 0) {
		cudaFuncAttributes aT, aS;
		CK(cudaFuncGetAttributes(&aT, persistTraverseKernel<PHONG>));
		CK(cudaFuncGetAttributes(&aS, persistShadeKernel<BRDF, SHADOW, PHONG>));
		regsT = (aT.numRegs + 7) / 8 * 8 * 128;
		regsS = (aS.numRegs + 7) / 8 * 8 * 128;
	}
	int sB = ctx->persistSBlocks, tB = ctx->persistTBlocks;
	while (sB > 1 && sB * regsS + regsT > 65536) sB--;
	const int room = (65536 - sB * regsS) / regsT;
	if (tB <= 0 || tB > room) tB = room;
	if (tB + sB > 16) tB = 16 - sB;
	if (tB < 1) return fail(ctx, PBR_ERR_INVALID, "pathTracing: persistent pipeline does not fit on an SM");
	{
		LaunchScope ls(ctx, K_RAYGEN);
		persistRaygenKernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(P, W, ctx->pctl, ctx->ring[0], nPaths);
	}
	CK(cudaEventRecord(ctx->evFork, ctx->stream));
	CK(cudaStreamWaitEvent(ctx->shadeStream, ctx->evFork, 0));
	{
		LaunchScope ls(ctx, K_SHADE, ctx->shadeStream);
		persistShadeKernel<BRDF, SHADOW, PHONG><<<ctx->smCount * sB, 128, 0, ctx->shadeStream>>>(
			P, W, ctx->pctl, ctx->ring[0], ctx->ring[1], mask, ctx->persistFill);
	}
	CK(cudaEventRecord(ctx->evJoin, ctx->shadeStream));
	{
		LaunchScope ls(ctx, K_TRAVERSE);
		persistTraverseKernel<PHONG><<<ctx->smCount * tB, 128, 0, ctx->stream>>>(
			P.scene, W, ctx->pctl, ctx->ring[0], ctx->ring[1], mask, ctx->stats);
	}
	CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
	CK(cudaGetLastError());
	ctx->persistUsed = true;
	return PBR_OK;
}

/* pipeline 3: wavefront with carry-over (traverseCarryKernel) */
template <int BRDF, bool SHADOW, bool PHONG>
int runCarryOver(pbr_ctx* ctx, const FrameParams& P, WaveState W, int nPaths) {
	const QueueCtl& Q = ctx->qctl;
	/* wavefront with carry-over (traverseCarryKernel): the number of iterations depends on the rays,
	 * so launches go out in groups and the host looks at the mailbox of the group before the one it
	 * has just enqueued -- the device never waits for the host */
	W.node = ctx->waveNode;
	int occT = 0, occS = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occT, traverseCarryKernel<PHONG>, 128, 0));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, shadeKernel<BRDF, SHADOW, PHONG>, 128, 0));
	const int gridT = ctx->smCount * (occT > 0 ? occT : 1);
	const int gridS = ctx->smCount * (occS > 0 ? occS : 1);
	uint32_t* c = ctx->cctl;
	uint32_t* mailboxDev = nullptr;
	CK(cudaHostGetDevicePointer((void**) &mailboxDev, ctx->mailbox, 0));
	static const bool dump = getenv("PBR_PROFILE_DUMP") != nullptr;
	{
		LaunchScope ls(ctx, K_RAYGEN);
		raygenCarryKernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(P, W, c, nPaths);
	}
	const int firstGroup = P.frameCount * P.samples > 1 ? P.frameCount * P.samples : P.maxDepth;
	int it = 0;
	for (int group = 0; ; group++) {
		const int n = group == 0 ? firstGroup : ctx->flushGroup;
		const int itBegin = it;
		for (int j = 0; j < n; j++, it++) {
			const int a = it & 1, b = a ^ 1;
			ctx->mailbox[it % MAILBOX_SLOTS] = 0xffffffffu;
			CarryQueues Qc;
			Qc.qCarryIn = ctx->carryQ[a]; Qc.nCarryIn = c + 2 + a;
			Qc.qNew = (it == 0) ? nullptr : Q.queue[a]; Qc.nNew = c + 0 + a;
			Qc.qHit = ctx->hitQ; Qc.nHit = c + 4 + a;
			Qc.qCarryOut = ctx->carryQ[b]; Qc.nCarryOut = c + 2 + b;
			Qc.cursor = c + 6;
			Qc.zeroAtStart = c + 0 + b;
			Qc.mailbox = mailboxDev + (it % MAILBOX_SLOTS);
			cudaEvent_t d0 = nullptr, d1 = nullptr, d2 = nullptr;
			if (dump) { d0 = takeEvent(ctx); d1 = takeEvent(ctx); d2 = takeEvent(ctx); cudaEventRecord(d0, ctx->stream); }
			{
				LaunchScope ls(ctx, K_TRAVERSE);
				traverseCarryKernel<PHONG><<<gridT, 128, 0, ctx->stream>>>(P.scene, W, Qc, ctx->tailStepsBulk, ctx->tailStepsFlush, ctx->stats);
			}
			if (dump) cudaEventRecord(d1, ctx->stream);
			{
				LaunchScope ls(ctx, K_SHADE);
				shadeKernel<BRDF, SHADOW, PHONG><<<gridS, 128, 0, ctx->stream>>>(
					P, W, ctx->hitQ, c + 4 + a, Q.queue[b], c + 0 + b, c + 6, c + 2 + a, c + 4 + b);
			}
			if (dump) {
				cudaEventRecord(d2, ctx->stream);
				uint32_t h[8];
				cudaMemcpyAsync(h, c, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
				cudaStreamSynchronize(ctx->stream);
				float tMs = 0.0f, sMs = 0.0f;
				cudaEventElapsedTime(&tMs, d0, d1);
				cudaEventElapsedTime(&sMs, d1, d2);
				fprintf(stderr, "[carry] it %3d  start %8u  traverse %7.3f ms -> finished %8u parked %8u   shade %6.3f ms -> new %8u\n",
					it, ((volatile uint32_t*) ctx->mailbox)[it % MAILBOX_SLOTS], tMs, h[4 + a], h[2 + b], sMs, h[0 + b]);
				ctx->eventPool.push_back(d0); ctx->eventPool.push_back(d1); ctx->eventPool.push_back(d2);
			}
		}
		CK(cudaEventRecord(ctx->evGroup[group & 1], ctx->stream));
		CK(cudaGetLastError());
		if (group >= 1 || dump) {
			/* the group before this one: did one of its launches start with nothing to do? */
			const int gPrev = dump ? group : group - 1;
			CK(cudaEventSynchronize(ctx->evGroup[gPrev & 1]));
			bool done = false;
			const int pb = dump ? itBegin : ctx->prevGroupBegin, pe = dump ? it : itBegin;
			for (int k = pb; k < pe; k++) {
				const uint32_t live = ((volatile uint32_t*) ctx->mailbox)[k % MAILBOX_SLOTS];
				if (live == 0u) done = true;
			}
			if (done) break;
		}
		ctx->prevGroupBegin = itBegin;
		if (it > 1000000) return fail(ctx, PBR_ERR_INVALID, "pathTracing: carry-over wavefront does not terminate");
	}
	return PBR_OK;
}

/* pipeline 0: one traverse + one shade launch per bounce (what the measured choice picks on all but small scenes) */
template <int BRDF, bool SHADOW, bool PHONG>
int runWavefront(pbr_ctx* ctx, const FrameParams& P, WaveState W, int nPaths) {
	const QueueCtl& Q = ctx->qctl;
	int occT = 0, occS = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occT, traverseKernel<PHONG>, 128, 0));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, shadeKernel<BRDF, SHADOW, PHONG>, 128, 0));
	if (ctx->traverseBlocks > 0 && ctx->traverseBlocks < occT) occT = ctx->traverseBlocks;
	const int gridT = ctx->smCount * (occT > 0 ? occT : 1);
	const int gridS = ctx->smCount * (occS > 0 ? occS : 1);

	/* shadow rays through the traversal engine instead of one thread per path inside the shade kernel */
	const bool shadowStage = SHADOW && P.scene.numLights > 0 && ctx->shadowStage != 0;
	int gridG = 0;
	if (shadowStage) {
		int rc = ensureShadow(ctx, (size_t) nPaths);
		if (rc) return rc;
		int occG = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occG, shadowGenKernel<BRDF, PHONG>, 128, 0));
		gridG = ctx->smCount * (occG > 0 ? occG : 1);
	}
	{
		LaunchScope ls(ctx, K_RAYGEN);
		raygenKernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(P, W, Q, nPaths);
	}
	const int iterations = P.frameCount * P.samples * (P.maxDepth + (ctx->canExtendDepth ? P.maxAddedDepth : 0));
	static const bool dump = getenv("PBR_PROFILE_DUMP") != nullptr;     /* diagnostics: one line per iteration */
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
	if (dump) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
	for (int it = 0; it < iterations; it++) {
		const int in = it & 1, out = in ^ 1;
		const uint32_t* qIn = (it == 0) ? nullptr : Q.queue[in];
		if (dump) cudaEventRecord(e0, ctx->stream);
		{
			LaunchScope ls(ctx, K_TRAVERSE);
			traverseKernel<PHONG><<<gridT, 128, 0, ctx->stream>>>(P.scene, W, qIn, Q.ctrl + in, Q.ctrl + 2, Q.ctrl + out, ctx->stats);
		}
		if (dump) cudaEventRecord(e1, ctx->stream);
		if (shadowStage) {
			{
				LaunchScope ls(ctx, K_SHADE);
				shadowGenKernel<BRDF, PHONG><<<gridG, 128, 0, ctx->stream>>>(
					P, W, qIn, Q.ctrl + in, ctx->shadowO, ctx->shadowD, ctx->shadowQ, Q.ctrl + 3);
			}
			{
				LaunchScope ls(ctx, K_TRAVERSE);
				traverseShadowKernel<PHONG><<<gridT, 128, 0, ctx->stream>>>(
					P.scene, W, ctx->shadowO, ctx->shadowD, ctx->shadowQ, Q.ctrl + 3, Q.ctrl + 4, ctx->stats);
			}
			{
				LaunchScope ls(ctx, K_SHADE);
				shadeKernel<BRDF, SHADOW, PHONG, true><<<gridS, 128, 0, ctx->stream>>>(
					P, W, qIn, Q.ctrl + in, Q.queue[out], Q.ctrl + out, Q.ctrl + 2, Q.ctrl + 3, Q.ctrl + 4, ctx->shadowO);
			}
		}
		else {
			LaunchScope ls(ctx, K_SHADE);
			shadeKernel<BRDF, SHADOW, PHONG><<<gridS, 128, 0, ctx->stream>>>(P, W, qIn, Q.ctrl + in, Q.queue[out], Q.ctrl + out, Q.ctrl + 2);
		}
		if (dump) {
			cudaEventRecord(e2, ctx->stream);
			uint32_t c[4] = {0, 0, 0, 0};
			cudaMemcpyAsync(c, Q.ctrl, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream);
			cudaStreamSynchronize(ctx->stream);
			float tMs = 0.0f, sMs = 0.0f;
			cudaEventElapsedTime(&tMs, e0, e1);
			cudaEventElapsedTime(&sMs, e1, e2);
			fprintf(stderr, "[wavefront] it %3d  traverse %7.3f ms  shade%s %7.3f ms  alive after %u\n", it, tMs, shadowStage ? " + shadow stage" : "", sMs, c[out]);
		}
	}
	if (dump) { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); }
	CK(cudaGetLastError());
	return PBR_OK;
}

template <int BRDF, bool SHADOW, bool PHONG>
int runFrame(pbr_ctx* ctx, const FrameParams& P, int nPaths) {
	if (ctx->pipeline == 1) {
		{
			LaunchScope ls(ctx, K_OTHER);
			megaKernel<BRDF, SHADOW, PHONG><<<gridFor(nPaths, 128), 128, 0, ctx->stream>>>(P, nPaths);
		}
		CK(cudaGetLastError());
		return PBR_OK;
	}
	int rc = ensureWave(ctx, (size_t) nPaths);
	if (rc) return rc;
	WaveState W = ctx->wave;
	W.hitN = PHONG ? ctx->hitN : nullptr;
	switch (ctx->pipeline) {
		case 2: return runPersistent<BRDF, SHADOW, PHONG>(ctx, P, W, nPaths);
		case 3: return runCarryOver<BRDF, SHADOW, PHONG>(ctx, P, W, nPaths);
		default: return runWavefront<BRDF, SHADOW, PHONG>(ctx, P, W, nPaths);
	}
}

bool parseSky(const char* v, pbr_float4* out) {
	/* "(float4)( %f, %f, %f, 0.0f )" (PathTracer.cpp:470-472, 515) */
	const char* p = strstr(v, ")(");
	p = p ? p + 2 : v;
	float c[3];
	for (int i = 0; i < 3; i++) {
		char* end = nullptr;
		c[i] = strtof(p, &end);
		if (end == p) return false;
		p = end;
		while (*p == 'f' || *p == 'F' || *p == ' ' || *p == '\t') p++;   /* float-literal suffix */
		if (i < 2) {
			if (*p != ',') return false;
			p++;
		}
	}
	out->x = c[0]; out->y = c[1]; out->z = c[2]; out->w = 0.0f;
	return true;
}

} /* namespace */

extern "C" {

int pbr_create(int device, pbr_ctx** out) {
	if (!out) return PBR_ERR_INVALID;
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) return PBR_ERR_NO_DEVICE;
	if (device < 0) {
		if (cudaGetDevice(&device) != cudaSuccess) device = 0;
	}
	if (device >= n) return PBR_ERR_NO_DEVICE;
	pbr_ctx* ctx = new pbr_ctx();
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PBR_ERR_NO_DEVICE; }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return PBR_ERR_NO_DEVICE; }
	ctx->smCount = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PBR_ERR_NO_DEVICE; }
	ctx->stream = ctx->ownStream;
	cudaEventCreate(&ctx->evStart);
	cudaEventCreate(&ctx->evStop);
	if (cudaMalloc(&ctx->stats, 6 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMalloc(&ctx->cursor64, sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMalloc(&ctx->qctl.ctrl, 8 * sizeof(uint32_t)) != cudaSuccess) {
		delete ctx;
		return PBR_ERR_NO_DEVICE;
	}
	cudaMemset(ctx->stats, 0, 6 * sizeof(unsigned long long));
	cudaMemset(ctx->qctl.ctrl, 0, 8 * sizeof(uint32_t));
	if (const char* e = getenv("PBR_NODE_PHASE_MIN")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->nodePhaseMin = v;
	}
	if (const char* e = getenv("PBR_REFILL_MIN")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->refillMin = v;
	}
	if (const char* e = getenv("PBR_PERSIST_T")) { const int v = atoi(e); if (v >= 0 && v <= 16) ctx->persistTBlocks = v; }
	if (const char* e = getenv("PBR_PERSIST_S")) { const int v = atoi(e); if (v >= 1 && v <= 16) ctx->persistSBlocks = v; }
	if (const char* e = getenv("PBR_PERSIST_FILL")) { const int v = atoi(e); if (v >= 0 && v <= 100000) ctx->persistFill = v; }
	if (const char* e = getenv("PBR_PIPELINE")) { const int v = atoi(e); if (v >= 0 && v <= 3) { ctx->pipeline = v; ctx->pipelineAuto = false; } }
	memset(&ctx->defines, 0, sizeof(ctx->defines));
	memset(&ctx->args.cam, 0, sizeof(ctx->args.cam));
	*out = ctx;
	return PBR_OK;
}

int pbr_destroy(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (Mem& m : ctx->mems) if (m.alive && m.dptr) cudaFree(m.dptr);
	for (void* p : ctx->pinned) cudaFreeHost(p);
	cudaFree(ctx->nodes); cudaFree(ctx->tris);
#if PT_NODE_ORDER
	cudaFree(ctx->nodeOrig);
#endif
	WaveState& W = ctx->wave;
	cudaFree(W.rayO); cudaFree(W.rayD); cudaFree(W.colS); cudaFree(W.finF); cudaFree(W.misc); cudaFree(W.dbg); cudaFree(ctx->hitN);
	cudaFree(ctx->qctl.queue[0]); cudaFree(ctx->qctl.queue[1]); cudaFree(ctx->qctl.ctrl);
	cudaFree(ctx->stats); cudaFree(ctx->cursor64);
	cudaFree(ctx->shadowO); cudaFree(ctx->shadowD); cudaFree(ctx->shadowQ);
	cudaFree(ctx->pctl); cudaFree(ctx->ring[0]); cudaFree(ctx->ring[1]);
	cudaFree(ctx->waveNode); cudaFree(ctx->hitQ); cudaFree(ctx->carryQ[0]); cudaFree(ctx->carryQ[1]); cudaFree(ctx->cctl);
	if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
	if (ctx->evGroup[0]) cudaEventDestroy(ctx->evGroup[0]);
	if (ctx->evGroup[1]) cudaEventDestroy(ctx->evGroup[1]);
	if (ctx->shadeStream) { cudaStreamSynchronize(ctx->shadeStream); cudaStreamDestroy(ctx->shadeStream); }
	if (ctx->evFork) cudaEventDestroy(ctx->evFork);
	if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
	cudaEventDestroy(ctx->evStart); cudaEventDestroy(ctx->evStop);
	for (const pbr_ctx::Timed& t : ctx->timedInFlight) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
	for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
	if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
	if (ctx->evCopy) cudaEventDestroy(ctx->evCopy);
	for (int i = 0; i < 8; i++) if (ctx->evAuto[i]) cudaEventDestroy(ctx->evAuto[i]);
	cudaStreamDestroy(ctx->ownStream);
	delete ctx;
	return PBR_OK;
}

const char* pbr_last_error(pbr_ctx* ctx) { return ctx ? ctx->lastError.c_str() : "no context"; }

int pbr_device_info(pbr_ctx* ctx, char* name, size_t name_len, int* sm_count, size_t* total_mem) {
	if (!ctx) return PBR_ERR_INVALID;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, ctx->device));
	if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
	if (sm_count) *sm_count = prop.multiProcessorCount;
	if (total_mem) *total_mem = prop.totalGlobalMem;
	return PBR_OK;
}

int pbr_buffer_create(pbr_ctx* ctx, const void* host, size_t bytes, pbr_mem* out) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	Mem* m;
	int rc = newMem(ctx, bytes, out, &m);
	if (rc) return rc;
	if (host && bytes) {
		CK(cudaMemcpyAsync(m->dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return PBR_OK;
}

int pbr_buffer_create_empty(pbr_ctx* ctx, size_t bytes, pbr_mem* out) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	Mem* m;
	int rc = newMem(ctx, bytes, out, &m);
	if (rc) return rc;
	CK(cudaMemsetAsync(m->dptr, 0, bytes > 0 ? bytes : 16, ctx->stream));
	return PBR_OK;
}

int pbr_buffer_update(pbr_ctx* ctx, pbr_mem buf, size_t bytes, const void* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, buf);
	if (!m || bytes > m->bytes || !host) return fail(ctx, PBR_ERR_INVALID, "pbr_buffer_update: bad buffer or size");
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(m->dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->sceneEpoch++;
	return PBR_OK;
}

int pbr_buffer_read(pbr_ctx* ctx, pbr_mem buf, size_t bytes, void* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, buf);
	if (!m || bytes > m->bytes || !host) return fail(ctx, PBR_ERR_INVALID, "pbr_buffer_read: bad buffer or size");
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(host, m->dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return PBR_OK;
}

int pbr_image_create(pbr_ctx* ctx, size_t width, size_t height, const float* host, pbr_mem* out) {
	if (!ctx || !out || width == 0 || height == 0) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	const size_t bytes = width * height * 16;
	Mem* m;
	int rc = newMem(ctx, bytes, out, &m);
	if (rc) return rc;
	m->image = true; m->width = width; m->height = height;
	if (host) {
		CK(cudaMemcpyAsync(m->dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	else {
		CK(cudaMemsetAsync(m->dptr, 0, bytes, ctx->stream));
	}
	return PBR_OK;
}

int pbr_image_write(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, const float* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, image);
	if (!m || !m->image || width != m->width || height != m->height || !host)
		return fail(ctx, PBR_ERR_INVALID, "pbr_image_write: bad image or size");
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(m->dptr, host, m->bytes, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return PBR_OK;
}

int pbr_image_read(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, float* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, image);
	if (!m || !m->image || width != m->width || height != m->height || !host)
		return fail(ctx, PBR_ERR_INVALID, "pbr_image_read: bad image or size");
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(host, m->dptr, m->bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return checkPersistAbort(ctx);
}

int pbr_image_read_begin(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, float* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, image);
	if (!m || !m->image || width != m->width || height != m->height || !host)
		return fail(ctx, PBR_ERR_INVALID, "pbr_image_read_begin: bad image or size");
	if (ctx->copyInFlight) return fail(ctx, PBR_ERR_NOT_READY, "pbr_image_read_begin: a read is already in flight");
	CK(cudaSetDevice(ctx->device));
	if (!ctx->copyStream) {
		CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&ctx->evCopy, cudaEventDisableTiming));
	}
	CK(cudaEventRecord(ctx->evCopy, ctx->stream));
	CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evCopy, 0));
	CK(cudaMemcpyAsync(host, m->dptr, m->bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
	ctx->copyInFlight = true;
	return PBR_OK;
}

int pbr_image_read_end(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	if (!ctx->copyInFlight) return PBR_OK;
	CK(cudaSetDevice(ctx->device));
	ctx->copyInFlight = false;
	CK(cudaStreamSynchronize(ctx->copyStream));
	return PBR_OK;
}

int pbr_image_copy(pbr_ctx* ctx, pbr_mem dst, pbr_mem src) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* d = getMem(ctx, dst);
	Mem* s = getMem(ctx, src);
	if (!d || !s || d->bytes != s->bytes) return fail(ctx, PBR_ERR_INVALID, "pbr_image_copy: bad images");
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(d->dptr, s->dptr, s->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return PBR_OK;
}

int pbr_mem_device_ptr(pbr_ctx* ctx, pbr_mem mem, void** dev_ptr, size_t* bytes) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, mem);
	if (!m) return fail(ctx, PBR_ERR_INVALID, "pbr_mem_device_ptr: bad handle");
	if (dev_ptr) *dev_ptr = m->dptr;
	if (bytes) *bytes = m->bytes;
	return PBR_OK;
}

int pbr_free_buffers(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (Mem& m : ctx->mems) {
		if (m.alive && m.dptr) cudaFree(m.dptr);
		m.alive = false;
		m.dptr = nullptr;
	}
	ctx->sceneEpoch++;
	return PBR_OK;
}

int pbr_host_alloc(pbr_ctx* ctx, size_t bytes, void** out) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMallocHost(out, bytes));
	ctx->pinned.push_back(*out);
	return PBR_OK;
}

int pbr_host_free(pbr_ctx* ctx, void* ptr) {
	if (!ctx) return PBR_ERR_INVALID;
	for (size_t i = 0; i < ctx->pinned.size(); i++) {
		if (ctx->pinned[i] == ptr) {
			cudaFreeHost(ptr);
			ctx->pinned.erase(ctx->pinned.begin() + (long) i);
			return PBR_OK;
		}
	}
	return fail(ctx, PBR_ERR_INVALID, "pbr_host_free: unknown pointer");
}

int pbr_set_define(pbr_ctx* ctx, const char* name, const char* value) {
	if (!ctx || !name || !value) return PBR_ERR_INVALID;
	std::string n(name);
	if (n.size() >= 2 && n.front() == '#' && n.back() == '#') n = n.substr(1, n.size() - 2);
	if (n == "BVH_NUM_NODES") { ctx->defNumNodes = atoi(value); ctx->haveNumNodes = true; }
	else if (n == "NUM_LIGHTS") { ctx->defNumLights = atoi(value); ctx->haveNumLights = true; }
	else if (n == "SKY_LIGHT") {
		if (!parseSky(value, &ctx->defSky)) return fail(ctx, PBR_ERR_INVALID, "SKY_LIGHT: cannot parse \"" + std::string(value) + "\"");
		ctx->haveSky = true;
	}
	else return fail(ctx, PBR_ERR_UNSUPPORTED, "unknown define " + n);
	if (ctx->programLoaded) {
		if (ctx->haveNumNodes) ctx->defines.bvh_num_nodes = ctx->defNumNodes;
		if (ctx->haveNumLights) ctx->defines.num_lights = ctx->defNumLights;
		if (ctx->haveSky) ctx->defines.sky_light = ctx->defSky;
	}
	return PBR_OK;
}

int pbr_program_load(pbr_ctx* ctx, const pbr_defines* defines) {
	if (!ctx || !defines) return PBR_ERR_INVALID;
	pbr_defines d = *defines;
	if (ctx->haveNumNodes) d.bvh_num_nodes = ctx->defNumNodes;
	if (ctx->haveNumLights) d.num_lights = ctx->defNumLights;
	if (ctx->haveSky) d.sky_light = ctx->defSky;
	if (d.accel_struct != 0) return fail(ctx, PBR_ERR_UNSUPPORTED, "accel_struct: only 0 (BVH) exists");
	if (d.brdf != 0 && d.brdf != 1) return fail(ctx, PBR_ERR_UNSUPPORTED, "render.brdf must be 0 (Schlick) or 1 (Shirley-Ashikhmin)");
	if (d.img_width <= 0 || d.img_height <= 0 || d.samples <= 0 || d.max_depth < 0 || d.max_added_depth < 0)
		return fail(ctx, PBR_ERR_INVALID, "invalid image size / samples / depth");
	if (d.max_depth + d.max_added_depth > 0xffff) return fail(ctx, PBR_ERR_INVALID, "max_depth + max_added_depth too large");
	ctx->defines = d;
	ctx->programLoaded = true;
	return PBR_OK;
}

int pbr_kernel_get(pbr_ctx* ctx, const char* name, pbr_kernel* out) {
	if (!ctx || !name || !out) return PBR_ERR_INVALID;
	if (strcmp(name, "pathTracing") != 0) return fail(ctx, PBR_ERR_UNSUPPORTED, std::string("no kernel named ") + name);
	if (!ctx->programLoaded) return fail(ctx, PBR_ERR_NOT_READY, "createKernel before loadProgram");
	*out = 1;
	return PBR_OK;
}

int pbr_kernel_set_arg(pbr_ctx* ctx, pbr_kernel k, uint32_t index, size_t size, const void* data) {
	if (!ctx || k != 1 || !data) return PBR_ERR_INVALID;
	KernelArgs& a = ctx->args;
	switch (index) {
		case 0: if (size != 4) goto badsize; memcpy(&a.seed, data, 4); break;
		case 1: if (size != 4) goto badsize; memcpy(&a.pixelWeight, data, 4); break;
		case 2: if (size != 4) goto badsize; memcpy(&a.pxDim, data, 4); break;
		case 3: if (size != sizeof(pbr_camera)) goto badsize; memcpy(&a.cam, data, sizeof(pbr_camera)); break;
		default:
			if (index > 13) return fail(ctx, PBR_ERR_INVALID, "pathTracing has 14 arguments");
			if (size != sizeof(pbr_mem)) goto badsize;
			memcpy(&a.mem[index], data, sizeof(pbr_mem));
			break;
	}
	a.setMask |= (1u << index);
	return PBR_OK;
badsize:
	return fail(ctx, PBR_ERR_INVALID, "clSetKernelArg: wrong argument size for slot " + std::to_string(index));
}

/* n consecutive frames (n <= PT_MAX_BATCH) in one pass over the device: frame f uses seeds[f] / weights[f];
 * frame 0 reads hIn, every later frame reads what the one before wrote to hOut. */
static int launchFrames(pbr_ctx* ctx, int n, const float* seeds, const float* weights, pbr_mem hIn, pbr_mem hOut) {
	KernelArgs& a = ctx->args;
	const pbr_defines& D = ctx->defines;

	const bool phong = (D.phongtess == 1);
	int rc = ensureScene(ctx, a.mem[4], a.mem[5], a.mem[7], D.bvh_num_nodes, phong, a.mem[6], a.mem[8]);
	if (rc) return rc;

	rc = updateCanExtendDepth(ctx, a.mem[9], D.brdf);
	if (rc) return rc;
	Mem* materials = getMem(ctx, a.mem[9]);
	Mem* lights = getMem(ctx, a.mem[10]);
	Mem* imageIn = getMem(ctx, hIn);
	Mem* imageOut = getMem(ctx, hOut);
	Mem* imageDebug = getMem(ctx, a.mem[13]);
	if (!materials || !lights || !imageIn || !imageOut || !imageDebug)
		return fail(ctx, PBR_ERR_INVALID, "pathTracing: a buffer / image argument is not live");
	const size_t imgBytes = (size_t) D.img_width * D.img_height * 16;
	if (imageIn->bytes < imgBytes || imageOut->bytes < imgBytes || imageDebug->bytes < imgBytes)
		return fail(ctx, PBR_ERR_INVALID, "pathTracing: image smaller than IMG_WIDTH x IMG_HEIGHT");
	if ((size_t) D.num_lights * sizeof(pbr_light) > lights->bytes)
		return fail(ctx, PBR_ERR_INVALID, "NUM_LIGHTS exceeds the lights buffer");

	FrameParams P;
	P.scene.nodes = ctx->nodes;
	P.scene.tris = ctx->tris;
	P.scene.trisB = ctx->trisB;
	P.scene.triMat = ctx->triMat;
#if PT_NODE_ORDER
	P.scene.nodeOrig = ctx->nodeOrig;
#endif
	P.scene.lights = (const pbr_light*) lights->dptr;
	P.scene.numNodes = ctx->numNodesDev;
	P.scene.numLights = D.num_lights;
	P.scene.phongAlpha = D.phongtess_alpha;
	P.scene.nodePhaseMin = ctx->nodePhaseMin;
	P.scene.refillMin = ctx->refillMin;
	P.materials = materials->dptr;
	P.numMaterials = (int) (materials->bytes / (D.brdf == 0 ? sizeof(pbr_material_schlick) : sizeof(pbr_material_sa)));
	P.cam = a.cam;
	P.seed = seeds[0]; P.pixelWeight = weights[0]; P.pxDim = a.pxDim;
	P.frameCount = n;
	for (int i = 0; i < PT_MAX_BATCH; i++) {
		P.frameSeed[i] = i < n ? seeds[i] : 0.0f;
		P.frameWeight[i] = i < n ? weights[i] : 0.0f;
	}
	P.width = D.img_width; P.height = D.img_height;
	P.y0 = ctx->tileY0 < 0 ? 0 : ctx->tileY0;
	P.y1 = ctx->tileY1 < 0 ? D.img_height : ctx->tileY1;
	P.stripeRows = 0; P.stripeWorld = 1; P.stripeRank = 0;
	if (ctx->stripeRows > 0) {
		if (D.img_height % (ctx->stripeRows * ctx->stripeWorld) != 0)
			return fail(ctx, PBR_ERR_INVALID, "pbr_set_tile_stripes: IMG_HEIGHT is not a multiple of stripe_rows * world");
		P.stripeRows = ctx->stripeRows; P.stripeWorld = ctx->stripeWorld; P.stripeRank = ctx->stripeRank;
		P.y0 = 0;
		P.y1 = D.img_height / ctx->stripeWorld;
	}
	if (P.y0 < 0 || P.y1 > D.img_height || P.y0 >= P.y1) return fail(ctx, PBR_ERR_INVALID, "tile rows outside the image");
	P.maxDepth = D.max_depth; P.maxAddedDepth = D.max_added_depth; P.samples = D.samples;
	P.antiAliasing = D.anti_aliasing;
	P.skyLight = make_float4(D.sky_light.x, D.sky_light.y, D.sky_light.z, D.sky_light.w);
	P.imageIn = (const float4*) imageIn->dptr;
	P.imageOut = (float4*) imageOut->dptr;
	P.imageDebug = ctx->debugImage ? (float4*) imageDebug->dptr : nullptr;
	P.stats = ctx->stats;

	const int nPaths = D.img_width * (P.y1 - P.y0);
	const bool shadow = (D.shadow_rays == 1);

	const int variant = (D.brdf == 1 ? 4 : 0) | (shadow ? 2 : 0) | (phong ? 1 : 0);

	/* which pipeline: the caller's, or the one that measured faster for this configuration */
	const int callerPipeline = ctx->pipeline;
	int timing = -1;
	if (ctx->pipelineAuto && n == 1) {
		const unsigned long long key = ctx->sceneEpoch * 0x9e3779b97f4a7c15ull ^ ((unsigned long long) nPaths << 20) ^
			((unsigned long long) variant << 8) ^ ((unsigned long long) D.max_depth << 12) ^ (unsigned long long) D.samples;
		if (key != ctx->autoKey) { ctx->autoKey = key; ctx->autoState = 0; }
		if (!ctx->evAuto[0]) for (int i = 0; i < 8; i++) CK(cudaEventCreate(&ctx->evAuto[i]));
		if (ctx->autoState == 5) {
			/* two rounds, the better time of each pipeline counts (clocks may still be ramping up in the first) */
			float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};
			CK(cudaEventSynchronize(ctx->evAuto[7]));
			for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&t[i], ctx->evAuto[2 * i], ctx->evAuto[2 * i + 1]));
			const float tWave = fminf(t[0], t[2]), tMega = fminf(t[1], t[3]);
			ctx->autoChoice = (tMega < 0.9f * tWave) ? 1 : 0;
			ctx->autoState = 6;
		}
		if (ctx->autoState == 0) { ctx->pipeline = 0; }
		else if (ctx->autoState >= 1 && ctx->autoState <= 4) {
			ctx->pipeline = (ctx->autoState & 1) ? 0 : 1;
			timing = 2 * (ctx->autoState - 1);
		}
		else ctx->pipeline = ctx->autoChoice;
		if (timing >= 0) CK(cudaEventRecord(ctx->evAuto[timing], ctx->stream));
	}
	else if (ctx->pipelineAuto) {
		ctx->pipeline = 0;                      /* interleaved batches: the wavefront */
	}
	switch (variant) {
		case 0: rc = runFrame<0, false, false>(ctx, P, nPaths); break;
		case 1: rc = runFrame<0, false, true>(ctx, P, nPaths); break;
		case 2: rc = runFrame<0, true, false>(ctx, P, nPaths); break;
		case 3: rc = runFrame<0, true, true>(ctx, P, nPaths); break;
		case 4: rc = runFrame<1, false, false>(ctx, P, nPaths); break;
		case 5: rc = runFrame<1, false, true>(ctx, P, nPaths); break;
		case 6: rc = runFrame<1, true, false>(ctx, P, nPaths); break;
		default: rc = runFrame<1, true, true>(ctx, P, nPaths); break;
	}
	if (timing >= 0 && rc == PBR_OK) CK(cudaEventRecord(ctx->evAuto[timing + 1], ctx->stream));
	if (ctx->pipelineAuto && n == 1 && ctx->autoState < 5 && rc == PBR_OK) ctx->autoState++;
	ctx->pipeline = callerPipeline;
	return rc;
}

static int checkLaunchable(pbr_ctx* ctx, pbr_kernel k, uint32_t needed) {
	if (!ctx || k != 1) return PBR_ERR_INVALID;
	if (!ctx->programLoaded) return fail(ctx, PBR_ERR_NOT_READY, "execute before loadProgram");
	if ((ctx->args.setMask & needed) != needed) return fail(ctx, PBR_ERR_NOT_READY, "pathTracing: not all 14 arguments are set");
	return PBR_OK;
}

int pbr_kernel_launch(pbr_ctx* ctx, pbr_kernel k) {
	int rc = checkLaunchable(ctx, k, 0x3fffu);
	if (rc) return rc;
	CK(cudaSetDevice(ctx->device));
	KernelArgs& a = ctx->args;
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	rc = launchFrames(ctx, 1, &a.seed, &a.pixelWeight, a.mem[11], a.mem[12]);
	if (rc) return rc;
	CK(cudaEventRecord(ctx->evStop, ctx->stream));
	ctx->timed = true;
	return PBR_OK;
}

int pbr_kernel_launch_batch(pbr_ctx* ctx, pbr_kernel k, int32_t n_frames, const float* seeds, const float* pixel_weights) {
	int rc = checkLaunchable(ctx, k, 0x3ffcu);         /* slots 0 and 1 come with the call */
	if (rc) return rc;
	if (n_frames < 1 || !seeds || !pixel_weights) return fail(ctx, PBR_ERR_INVALID, "pbr_kernel_launch_batch: bad arguments");
	CK(cudaSetDevice(ctx->device));
	KernelArgs& a = ctx->args;
	const pbr_mem hIn = a.mem[11], hOut = a.mem[12];
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	const bool depthOfField = a.cam.focusPoint.x >= 0 && a.cam.focusPoint.y >= 0;
	if (!depthOfField) {
		/* pixels are independent of each other, so everything after frame 0 can run in place in imageOut:
		 * frame after frame (default: the rays of one bounce of one frame stay together, which the caches
		 * like), or -- "batch_interleave" -- PT_MAX_BATCH frames per pass, pixels running ahead */
		const int chunk = ctx->batchInterleave ? PT_MAX_BATCH : 1;
		for (int f = 0; f < n_frames; f += chunk) {
			const int n = n_frames - f < chunk ? n_frames - f : chunk;
			rc = launchFrames(ctx, n, seeds + f, pixel_weights + f, f == 0 ? hIn : hOut, hOut);
			if (rc) return rc;
		}
	}
	else {
		/* depthOfField reads ANOTHER pixel of the previous frame (pathtracing.cl:41-43): frame by frame,
		 * ping-pong between imageOut and a scratch image so that the last frame lands in imageOut */
		Mem* out = getMem(ctx, hOut);
		if (!out) return fail(ctx, PBR_ERR_INVALID, "pathTracing: imageOut is not live");
		if (ctx->scratchImage == 0 || !getMem(ctx, ctx->scratchImage) || getMem(ctx, ctx->scratchImage)->bytes < out->bytes) {
			const uint64_t epoch = ctx->sceneEpoch;
			Mem* m = nullptr;
			rc = newMem(ctx, out->bytes, &ctx->scratchImage, &m);
			if (rc) return rc;
			m->image = true; m->width = out->width; m->height = out->height;
			ctx->sceneEpoch = epoch;                   /* not a scene buffer: keep the repacked scene */
		}
		pbr_mem prev = hIn;
		for (int f = 0; f < n_frames; f++) {
			const pbr_mem dst = ((n_frames - 1 - f) % 2 == 0) ? hOut : ctx->scratchImage;
			rc = launchFrames(ctx, 1, seeds + f, pixel_weights + f, prev, dst);
			if (rc) return rc;
			prev = dst;
		}
	}
	CK(cudaEventRecord(ctx->evStop, ctx->stream));
	ctx->timed = true;
	return PBR_OK;
}

int pbr_finish(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	return checkPersistAbort(ctx);
}

int pbr_kernel_time_ms(pbr_ctx* ctx, pbr_kernel k, double* ms) {
	if (!ctx || k != 1 || !ms) return PBR_ERR_INVALID;
	*ms = 0.0;
	if (!ctx->timed) return PBR_OK;
	CK(cudaEventSynchronize(ctx->evStop));
	float f = 0.0f;
	CK(cudaEventElapsedTime(&f, ctx->evStart, ctx->evStop));
	*ms = (double) f;
	return PBR_OK;
}

int pbr_set_tile(pbr_ctx* ctx, int32_t y0, int32_t y1) {
	if (!ctx) return PBR_ERR_INVALID;
	ctx->tileY0 = y0;
	ctx->tileY1 = y1;
	return PBR_OK;
}

int pbr_set_tile_stripes(pbr_ctx* ctx, int32_t stripe_rows, int32_t world, int32_t rank) {
	if (!ctx) return PBR_ERR_INVALID;
	if (stripe_rows <= 0) { ctx->stripeRows = 0; ctx->stripeWorld = 1; ctx->stripeRank = 0; return PBR_OK; }
	if (world < 1 || rank < 0 || rank >= world) return fail(ctx, PBR_ERR_INVALID, "pbr_set_tile_stripes: bad world / rank");
	ctx->stripeRows = stripe_rows; ctx->stripeWorld = world; ctx->stripeRank = rank;
	return PBR_OK;
}

int pbr_set_pipeline(pbr_ctx* ctx, int32_t mode) {
	if (!ctx || mode < -1 || mode > 3) return PBR_ERR_INVALID;
	ctx->pipelineAuto = (mode == -1);
	ctx->pipeline = mode < 0 ? 0 : mode;
	ctx->autoState = 0;
	ctx->autoKey = 0;
	return PBR_OK;
}

int pbr_pipeline_in_use(pbr_ctx* ctx, int32_t* mode) {
	if (!ctx || !mode) return PBR_ERR_INVALID;
	*mode = !ctx->pipelineAuto ? ctx->pipeline : (ctx->autoState >= 6 ? ctx->autoChoice : -1);
	return PBR_OK;
}

int pbr_set_tuning(pbr_ctx* ctx, const char* key, int32_t value) {
	if (!ctx || !key) return PBR_ERR_INVALID;
	const std::string k(key);
	if (k == "node_phase_min" && value >= 1 && value <= 32) ctx->nodePhaseMin = value;
	else if (k == "refill_min" && value >= 1 && value <= 32) ctx->refillMin = value;
	else if (k == "persist_t" && value >= 0 && value <= 16) ctx->persistTBlocks = value;
	else if (k == "persist_s" && value >= 1 && value <= 16) ctx->persistSBlocks = value;
	else if (k == "persist_fill" && value >= 0 && value <= 100000) ctx->persistFill = value;
	else if (k == "tail_steps_bulk" && value >= 1 && value <= 1000000) ctx->tailStepsBulk = value;
	else if (k == "tail_steps_flush" && value >= 1 && value <= 1000000) ctx->tailStepsFlush = value;
	else if (k == "flush_group" && value >= 1 && value <= 64) ctx->flushGroup = value;
	else if (k == "batch_interleave" && (value == 0 || value == 1)) ctx->batchInterleave = value;
	else if (k == "traverse_blocks" && value >= 0 && value <= 32) ctx->traverseBlocks = value;
	else if (k == "shadow_stage" && (value == 0 || value == 1)) ctx->shadowStage = value;
	else return fail(ctx, PBR_ERR_INVALID, "pbr_set_tuning: unknown key or value out of range: " + k);
	return PBR_OK;
}

int pbr_set_debug_image(pbr_ctx* ctx, int32_t enabled) {
	if (!ctx) return PBR_ERR_INVALID;
	ctx->debugImage = enabled != 0;
	return PBR_OK;
}

int pbr_stats(pbr_ctx* ctx, uint64_t out[6], int32_t reset) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMemcpyAsync(out, ctx->stats, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
	if (reset) CK(cudaMemsetAsync(ctx->stats, 0, 6 * sizeof(uint64_t), ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return PBR_OK;
}

int pbr_set_stream(pbr_ctx* ctx, void* cuda_stream, int32_t own_stream) {
	if (!ctx) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->stream = own_stream ? ctx->ownStream : (cudaStream_t) cuda_stream;
	return PBR_OK;
}

int pbr_profile_enable(pbr_ctx* ctx, int32_t enabled) {
	if (!ctx) return PBR_ERR_INVALID;
	ctx->profiling = enabled != 0;
	return PBR_OK;
}

int pbr_profile_read(pbr_ctx* ctx, pbr_profile* out, int32_t reset) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	drainTimed(ctx);
	*out = ctx->prof;
	if (reset) memset(&ctx->prof, 0, sizeof(ctx->prof));
	return PBR_OK;
}

static int traceImpl(pbr_ctx* ctx, pbr_mem bvh, pbr_mem facesV, pbr_mem vertices, pbr_mem lights, int32_t num_lights,
                     const pbr_ray* dRays, int64_t n, int32_t any_hit, pbr_hit* dHits) {
	Mem* b = getMem(ctx, bvh);
	if (!b) return fail(ctx, PBR_ERR_INVALID, "pbr_trace: bad bvh buffer");
	const int numNodes = ctx->haveNumNodes ? ctx->defNumNodes : (int) (b->bytes / sizeof(pbr_bvh_node));
	int rc = ensureScene(ctx, bvh, facesV, vertices, numNodes);
	if (rc) return rc;
	SceneDev S;
	S.nodes = ctx->nodes;
	S.tris = ctx->tris;
	S.trisB = ctx->trisB;
	S.triMat = ctx->triMat;
#if PT_NODE_ORDER
	S.nodeOrig = ctx->nodeOrig;
#endif
	S.numNodes = ctx->numNodesDev;
	S.numLights = 0;
	S.lights = nullptr;
	S.phongAlpha = 0.0f;
	S.nodePhaseMin = ctx->nodePhaseMin;
	S.refillMin = ctx->refillMin;
	if (num_lights > 0) {
		Mem* l = getMem(ctx, lights);
		if (!l || (size_t) num_lights * sizeof(pbr_light) > l->bytes) return fail(ctx, PBR_ERR_INVALID, "pbr_trace: bad lights buffer");
		S.lights = (const pbr_light*) l->dptr;
		S.numLights = num_lights;
	}
	if (n <= 0) return PBR_OK;
	CK(cudaMemsetAsync(ctx->cursor64, 0, sizeof(unsigned long long), ctx->stream));
	int occ = 0;
	if (any_hit) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, traceRaysKernel<true>, 128, 0));
	else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, traceRaysKernel<false>, 128, 0));
	const int grid = ctx->smCount * (occ > 0 ? occ : 1);
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	{
		LaunchScope ls(ctx, K_TRAVERSE);
		if (any_hit) traceRaysKernel<true><<<grid, 128, 0, ctx->stream>>>(S, dRays, (long long) n, dHits, ctx->cursor64, ctx->stats);
		else traceRaysKernel<false><<<grid, 128, 0, ctx->stream>>>(S, dRays, (long long) n, dHits, ctx->cursor64, ctx->stats);
	}
	CK(cudaGetLastError());
	CK(cudaEventRecord(ctx->evStop, ctx->stream));
	ctx->timed = true;
	return PBR_OK;
}

int pbr_trace_device(pbr_ctx* ctx, pbr_mem bvh, pbr_mem facesV, pbr_mem vertices, pbr_mem lights, int32_t num_lights,
                     pbr_mem rays, int64_t n, int32_t any_hit, pbr_mem hits) {
	if (!ctx) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	Mem* r = getMem(ctx, rays);
	Mem* h = getMem(ctx, hits);
	if (!r || !h || n < 0 || (size_t) n * sizeof(pbr_ray) > r->bytes || (size_t) n * sizeof(pbr_hit) > h->bytes)
		return fail(ctx, PBR_ERR_INVALID, "pbr_trace_device: bad rays / hits buffer");
	return traceImpl(ctx, bvh, facesV, vertices, lights, num_lights, (const pbr_ray*) r->dptr, n, any_hit, (pbr_hit*) h->dptr);
}

int pbr_trace(pbr_ctx* ctx, pbr_mem bvh, pbr_mem facesV, pbr_mem vertices, pbr_mem lights, int32_t num_lights,
              const pbr_ray* rays, int64_t n, int32_t any_hit, pbr_hit* hits) {
	if (!ctx || n < 0 || (n > 0 && (!rays || !hits))) return PBR_ERR_INVALID;
	if (n == 0) return PBR_OK;
	CK(cudaSetDevice(ctx->device));
	pbr_ray* dRays = nullptr;
	pbr_hit* dHits = nullptr;
	CK(cudaMalloc(&dRays, (size_t) n * sizeof(pbr_ray)));
	cudaError_t e = cudaMalloc(&dHits, (size_t) n * sizeof(pbr_hit));
	if (e != cudaSuccess) { cudaFree(dRays); return cudaFail(ctx, e, "cudaMalloc"); }
	int rc = PBR_OK;
	e = cudaMemcpyAsync(dRays, rays, (size_t) n * sizeof(pbr_ray), cudaMemcpyHostToDevice, ctx->stream);
	if (e != cudaSuccess) rc = cudaFail(ctx, e, "cudaMemcpyAsync");
	if (!rc) rc = traceImpl(ctx, bvh, facesV, vertices, lights, num_lights, dRays, n, any_hit, dHits);
	if (!rc) {
		e = cudaMemcpyAsync(hits, dHits, (size_t) n * sizeof(pbr_hit), cudaMemcpyDeviceToHost, ctx->stream);
		if (e != cudaSuccess) rc = cudaFail(ctx, e, "cudaMemcpyAsync");
	}
	e = cudaStreamSynchronize(ctx->stream);
	if (!rc && e != cudaSuccess) rc = cudaFail(ctx, e, "cudaStreamSynchronize");
	cudaFree(dRays);
	cudaFree(dHits);
	return rc;
}

int pbr_pinned_math_eval(pbr_ctx* ctx, int32_t op, const float* x, const float* y, int64_t n, float* out) {
	if (!ctx || !x || !out || n < 0 || op < 0 || op > 7) return PBR_ERR_INVALID;
	if (n == 0) return PBR_OK;
	CK(cudaSetDevice(ctx->device));
	float *dx = nullptr, *dy = nullptr, *dout = nullptr;
	CK(cudaMalloc(&dx, (size_t) n * 4));
	CK(cudaMalloc(&dy, (size_t) n * 4));
	CK(cudaMalloc(&dout, (size_t) n * 4));
	CK(cudaMemcpyAsync(dx, x, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (y) CK(cudaMemcpyAsync(dy, y, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
	else CK(cudaMemsetAsync(dy, 0, (size_t) n * 4, ctx->stream));
	{
		LaunchScope ls(ctx, K_OTHER);
		pinnedMathKernel<<<gridFor(n, 256), 256, 0, ctx->stream>>>(op, dx, dy, (long long) n, dout);
	}
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, dout, (size_t) n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	cudaFree(dx); cudaFree(dy); cudaFree(dout);
	return PBR_OK;
}

} /* extern "C" */
