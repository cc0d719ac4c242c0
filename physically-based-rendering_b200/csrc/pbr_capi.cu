/*
 * pbr_capi.cu -- implementation of the C ABI declared in include/pbr_b200.h.
 *
 * Stands where the reference's `CL` class (source/CL.cpp) stands: owns the device, the buffers and
 * images, the "program" (a set of precompiled sm_100a kernel specialisations selected by the values
 * the reference would splice into pt_header.cl) and the one kernel `pathTracing` with its 14
 * argument slots (PathTracer.cpp:88-125).  No CPU fallback: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>                 /* types only: the library is dlopen'ed by pbr_comm_init, so one GPU needs no NCCL */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <deque>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>

#include "../../include/pbr_b200.h"
#include "pt_kernels.cuh"
#include "pt_wide.cuh"
#include "wide_bvh.h"

using namespace ptk;

namespace {

struct Mem {
	void* dptr = nullptr;
	size_t bytes = 0;
	bool image = false;
	size_t width = 0, height = 0;
	bool alive = false;
	uint64_t epoch = 0;               /* bumped whenever the contents may have changed (create, update, free) */
	cudaEvent_t evCombined = nullptr; /* pbr_frame_combine still reads (or all-gathers into) this image until then */
	bool combinePending = false;
	cudaEvent_t evOwnRowsFree = nullptr; /* ROWS combine with stripes: this rank's rows have been packed -- a frame that
	                                        touches only its own rows may overwrite them while the gather is still running */
	bool ownRowsPending = false;
	cudaEvent_t evWritten = nullptr;  /* recorded on the context's stream behind the last launch / copy that wrote this image:
	                                     an asynchronous read-back waits for THIS, not for whatever has been queued since */
	bool writtenValid = false;
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
	std::deque<Mem> mems;             /* a deque: Mem* stay valid when a handle is added */
	uint64_t epochCounter = 0;
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
	size_t nodesCap = 0, trisCap = 0;
	pbr_mem cacheBvh = 0, cacheFacesV = 0, cacheVertices = 0, cacheFacesN = 0, cacheNormals = 0;
	uint64_t cacheEpochBvh = ~0ull, cacheEpochFacesV = ~0ull, cacheEpochVertices = ~0ull, cacheEpochFacesN = ~0ull, cacheEpochNormals = ~0ull;
	bool cachePhong = false;
	int cacheNumNodes = -1;
	uint64_t geometryVersion = 0;              /* bumped by every repack: keys the pipeline choice */
	int numNodesDev = 0;

	/* the 4-wide BVH of the ordered walk (wide_bvh.h, pt_wide.cuh), built on demand from the uploaded node array */
	int traversal = -1;                        /* pbr_set_traversal: -1 automatic, 0 reference order, 1 ordered */
	int lastTraversal = 0;                     /* what the last launch used */
	float4* wide = nullptr;
	int* faceLeaf = nullptr;
	size_t wideCap = 0, faceLeafCap = 0;       /* wide nodes / faces allocated */
	int wideCount = 0, wideTop = 0, wideDepth = 0;
	int wideTopBudget = 21;                    /* tuning "wide_top": nodes staged in shared memory (1 + 4 + 16: measured +0.8 %
	                                              over none on C2; 85 costs occupancy and loses 0.6 %) */
	int wideNodePhaseMin = 20, wideRefillMin = 8;   /* the engine's two thresholds as measured best for the ordered walk */
	bool wideBuilt = false, wideOk = false;
	std::string wideWhy;
	uint64_t wideVersion = ~0ull;              /* geometryVersion the wide tree was built for */
	int wideBudgetBuilt = -1;
	double wideBuildMs = 0.0;

	/* wavefront state: one set per frame in flight (pbr_kernel_launch_batch overlaps the tracing of consecutive
	 * frames on streams of their own; a single launch uses set 0 on the context's stream) */
	struct WaveSet {
		WaveState wave = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
		float4* hitN = nullptr;                /* allocated with the wave state, used when PHONGTESS */
		QueueCtl qctl = {nullptr, {nullptr, nullptr}};
		size_t waveCap = 0;
		/* shadow rays as a wavefront stage (render.shadow_rays): ray per path, queue; qctl.ctrl[3] count, [4] cursor */
		float4* shadowO = nullptr;
		float4* shadowD = nullptr;
		uint32_t* shadowQ = nullptr;
		size_t shadowCap = 0;
		/* frames in flight: the frame's own radiance (rgb, focus) before it is mixed into the accumulation image */
		float4* frameOut = nullptr;
		size_t frameOutCap = 0;
		cudaStream_t stream = nullptr;
		cudaEvent_t evTraced = nullptr, evConsumed = nullptr;
		bool consumedPending = false;
	};
	enum { MAX_IN_FLIGHT = 8 };
	WaveSet sets[MAX_IN_FLIGHT];
	WaveSet serialSet;                         /* frames launched on the context's stream (debug image, depth of field, megakernel
	                                              timing rounds): a state of their own, so that they never share one with a frame
	                                              that is still being traced on a stream of its own */
	int framesInFlight = 4;                    /* tuning "frames_in_flight": 1 -> 2 -> 3 -> 4 = 1356 -> 1548 -> 1599 -> 1619 Mrays/s on C2 */
	cudaEvent_t evPrepared = nullptr;
	uint64_t preparedVersion = ~0ull, preparedWide = ~0ull;
	uint64_t overlapCounter = 0;               /* frames launched through launchOverlapped: picks the wave set */
	int batchCombineMode = -1;                 /* pbr_set_batch_combine */
	pbr_mem batchCombineOut[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	int batchCombineOuts = 0, batchCombineFirst = 0;
	int shadowStage = 1;                       /* tuning "shadow_stage": 0 = walk shadow rays inside the shade kernel */

	FrameParams lastFrameParams;               /* of the frame launched last (the deferred mix needs them) */
	int lastNumPaths = 0;
	bool launchUseWide = false;
	int launchSet = -1;                        /* which wave set / stream the frame being launched uses (-1: serialSet) */
	cudaStream_t launchStream = nullptr;
	float4* launchFrameOut = nullptr;          /* != NULL: the frame's radiance goes here, mixing is deferred */
	int traverseBlocks = 0;                    /* tuning: cap on resident traverse blocks per SM (0 = all that fit) */
	int wideBlocks = 0;                        /* tuning "wide_blocks": the same for the ordered walk's kernels */

	/* can any material extend a path beyond MAX_DEPTH? (decides how many wavefront iterations to launch) */
	pbr_mem extendCacheMem = 0;
	uint64_t extendCacheEpoch = ~0ull;         /* epoch of the materials buffer the answer was computed from */
	int extendCacheBrdf = -1;
	bool canExtendDepth = true;
	int nodePhaseMin = 16;                     /* PBR_NODE_PHASE_MIN overrides (tuning) */
	int refillMin = 4;                         /* PBR_REFILL_MIN overrides (tuning) */
	unsigned long long* stats = nullptr;       /* 8 counters: the six of pbr_stats, re-walked rays, rays of the ordered walk */
	unsigned long long* cursor64 = nullptr;    /* work cursor of traceRaysKernel */

	/* multi-GPU (pbr_comm_init): one process per GPU, one NCCL collective per frame on a stream of its own */
	ncclComm_t comm = nullptr;
	ncclComm_t commReduce = nullptr;           /* the all-reduce's communicator: split off `comm` with a cap on its blocks */
	int commRank = 0, commWorld = 1;
	cudaStream_t commStream = nullptr;
	cudaEvent_t evRendered = nullptr;
	float4* commSend = nullptr;
	float4* commRecv = nullptr;
	size_t commSendCap = 0, commRecvCap = 0;   /* float4s */
	uint64_t combines = 0;
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
	m.epoch = ++ctx->epochCounter;
	ctx->mems.push_back(m);
	*out = (pbr_mem) ctx->mems.size();
	if (mp) *mp = &ctx->mems.back();
	return PBR_OK;
}

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

/* Rebuild the repacked node / triangle arrays when the bound buffers, their contents or BVH_NUM_NODES changed. */
int ensureScene(pbr_ctx* ctx, pbr_mem hBvh, pbr_mem hFacesV, pbr_mem hVertices, int numNodes,
                bool phong = false, pbr_mem hFacesN = 0, pbr_mem hNormals = 0) {
	Mem* bvh = getMem(ctx, hBvh);
	Mem* facesV = getMem(ctx, hFacesV);
	Mem* vertices = getMem(ctx, hVertices);
	if (!bvh || !facesV || !vertices) return fail(ctx, PBR_ERR_INVALID, "pathTracing: bvh / facesV / vertices argument is not a live buffer");
	Mem* facesN = phong ? getMem(ctx, hFacesN) : nullptr;
	Mem* normals = phong ? getMem(ctx, hNormals) : nullptr;
	if (phong && (!facesN || !normals || facesN->bytes < facesV->bytes || normals->bytes < sizeof(pbr_float4)))
		return fail(ctx, PBR_ERR_INVALID, "pathTracing: PHONGTESS needs facesN (one entry per face) and normals");
	if (ctx->cacheBvh == hBvh && ctx->cacheFacesV == hFacesV && ctx->cacheVertices == hVertices &&
	    ctx->cacheEpochBvh == bvh->epoch && ctx->cacheEpochFacesV == facesV->epoch && ctx->cacheEpochVertices == vertices->epoch &&
	    ctx->cacheNumNodes == numNodes && ctx->cachePhong == phong &&
	    (!phong || (ctx->cacheFacesN == hFacesN && ctx->cacheNormals == hNormals &&
	                ctx->cacheEpochFacesN == facesN->epoch && ctx->cacheEpochNormals == normals->epoch))) {
		return PBR_OK;
	}

	/* frames still being traced on streams of their own read the arrays that are about to be rebuilt */
	for (pbr_ctx::WaveSet& T : ctx->sets) if (T.stream) CK(cudaStreamSynchronize(T.stream));

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
	ctx->cacheEpochBvh = bvh->epoch; ctx->cacheEpochFacesV = facesV->epoch; ctx->cacheEpochVertices = vertices->epoch;
	ctx->cacheEpochFacesN = facesN ? facesN->epoch : 0; ctx->cacheEpochNormals = normals ? normals->epoch : 0;
	ctx->cacheNumNodes = numNodes;
	ctx->geometryVersion++;
	ctx->numNodesDev = numNodes;
	ctx->trisB = phong ? nullptr : (const float*) (ctx->tris + PT_TRI_STRIDE * facesAlloc);
	ctx->triMat = phong ? nullptr : (const uint32_t*) (ctx->tris + PT_TRI_STRIDE * facesAlloc + quarter);
	return PBR_OK;
}

/* The 4-wide BVH for the scene ensureScene has just bound (wide_bvh.h): the node array is read back once, the tree is
 * recovered, checked and collapsed on the host, and the wide nodes go up.  When the builder refuses the array, wideOk
 * stays false and every launch keeps the reference-order walk; pbr_traversal_info tells why. */
int ensureWide(pbr_ctx* ctx) {
	if (ctx->wideBuilt && ctx->wideVersion == ctx->geometryVersion && ctx->wideBudgetBuilt == ctx->wideTopBudget) return PBR_OK;
	for (pbr_ctx::WaveSet& T : ctx->sets) if (T.stream) CK(cudaStreamSynchronize(T.stream));
	ctx->wideBuilt = true;
	ctx->wideVersion = ctx->geometryVersion;
	ctx->wideBudgetBuilt = ctx->wideTopBudget;
	ctx->wideOk = false;
	ctx->wideWhy.clear();
	ctx->wideCount = ctx->wideTop = ctx->wideDepth = 0;
	if (ctx->cachePhong) { ctx->wideWhy = "PHONGTESS: the ordered walk handles flat triangles only"; return PBR_OK; }
	Mem* bvh = getMem(ctx, ctx->cacheBvh);
	Mem* facesV = getMem(ctx, ctx->cacheFacesV);
	if (!bvh || !facesV) { ctx->wideWhy = "no scene bound"; return PBR_OK; }
	const int numNodes = ctx->cacheNumNodes;
	if (numNodes < 2) { ctx->wideWhy = "fewer than two nodes"; return PBR_OK; }
	const int numFaces = (int) (facesV->bytes / sizeof(pbr_uint4));
	std::vector<float> host((size_t) numNodes * 8);
	cudaEvent_t e0 = takeEvent(ctx), e1 = takeEvent(ctx);
	cudaEventRecord(e0, ctx->stream);
	CK(cudaMemcpyAsync(host.data(), bvh->dptr, (size_t) numNodes * 32, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	wbvh::Result R = wbvh::build(host.data(), numNodes, numFaces, ctx->wideTopBudget);
	if (!R.ok) { ctx->wideWhy = R.why; ctx->eventPool.push_back(e0); ctx->eventPool.push_back(e1); return PBR_OK; }
	if (R.nodes.size() > ctx->wideCap) {
		if (ctx->wide) cudaFree(ctx->wide);
		ctx->wide = nullptr;
		ctx->wideCap = 0;
		CK(cudaMalloc(&ctx->wide, R.nodes.size() * sizeof(wbvh::Node)));
		ctx->wideCap = R.nodes.size();
	}
	if (R.faceLeaf.size() > ctx->faceLeafCap) {
		if (ctx->faceLeaf) cudaFree(ctx->faceLeaf);
		ctx->faceLeaf = nullptr;
		ctx->faceLeafCap = 0;
		CK(cudaMalloc(&ctx->faceLeaf, R.faceLeaf.size() * sizeof(int32_t)));
		ctx->faceLeafCap = R.faceLeaf.size();
	}
	CK(cudaMemcpyAsync(ctx->wide, R.nodes.data(), R.nodes.size() * sizeof(wbvh::Node), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->faceLeaf, R.faceLeaf.data(), R.faceLeaf.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
	cudaEventRecord(e1, ctx->stream);
	CK(cudaStreamSynchronize(ctx->stream));
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, e0, e1);
	ctx->wideBuildMs = (double) ms;
	ctx->eventPool.push_back(e0);
	ctx->eventPool.push_back(e1);
	ctx->wideCount = (int) R.nodes.size();
	ctx->wideTop = R.topCount;
	ctx->wideDepth = R.depth;
	ctx->wideOk = true;
	return PBR_OK;
}

/* Which walk does this launch take?  Automatic: the ordered walk whenever the reference's visit counters cannot be
 * observed (no debug image) and the wide tree exists; the caller can force either. */
int chooseTraversal(pbr_ctx* ctx, bool countersObservable, bool* useWide) {
	*useWide = false;
	if (ctx->traversal == 0 || (ctx->traversal < 0 && countersObservable)) return PBR_OK;
	int rc = ensureWide(ctx);
	if (rc) return rc;
	if (!ctx->wideOk) {
		if (ctx->traversal == 1) return fail(ctx, PBR_ERR_UNSUPPORTED, "pbr_set_traversal(1): the ordered walk is not available for this scene: " + ctx->wideWhy);
		return PBR_OK;
	}
	*useWide = true;
	return PBR_OK;
}

void fillScene(pbr_ctx* ctx, SceneDev& S, bool useWide) {
	S.nodes = ctx->nodes;
	S.tris = ctx->tris;
	S.trisB = ctx->trisB;
	S.triMat = ctx->triMat;
	S.wide = useWide ? ctx->wide : nullptr;
	S.wideTop = useWide ? ctx->wideTop : 0;
	S.faceLeaf = useWide ? ctx->faceLeaf : nullptr;
	S.numNodes = ctx->numNodesDev;
	S.nodePhaseMin = useWide ? ctx->wideNodePhaseMin : ctx->nodePhaseMin;
	S.refillMin = useWide ? ctx->wideRefillMin : ctx->refillMin;
}

template <typename K>
int wideLaunchShape(pbr_ctx* ctx, K kernel, int* grid, size_t* shared) {
	const size_t bytes = wideSharedBytes(ctx->wideTop);
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
	int occ = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, PT_WIDE_BLOCK, bytes));
	if (occ < 1) return fail(ctx, PBR_ERR_INVALID, "the ordered walk does not fit on an SM (wide_top too large?)");
	*grid = ctx->smCount * occ;
	*shared = bytes;
	return PBR_OK;
}

int ensureWave(pbr_ctx* ctx, pbr_ctx::WaveSet& T, size_t nPaths) {
	if (!T.qctl.ctrl) {
		CK(cudaMalloc(&T.qctl.ctrl, 8 * sizeof(uint32_t)));
		CK(cudaMemsetAsync(T.qctl.ctrl, 0, 8 * sizeof(uint32_t), ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	if (nPaths <= T.waveCap) return PBR_OK;
	WaveState& W = T.wave;
	cudaFree(W.rayO); cudaFree(W.rayD); cudaFree(W.colS); cudaFree(W.finF); cudaFree(W.misc); cudaFree(W.dbg); cudaFree(T.hitN);
	cudaFree(T.qctl.queue[0]); cudaFree(T.qctl.queue[1]);
	W = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	T.hitN = nullptr;
	T.qctl.queue[0] = T.qctl.queue[1] = nullptr;
	T.waveCap = 0;
	CK(cudaMalloc(&W.rayO, nPaths * 16));
	CK(cudaMalloc(&W.rayD, nPaths * 16));
	CK(cudaMalloc(&W.colS, nPaths * 16));
	CK(cudaMalloc(&W.finF, nPaths * 16));
	CK(cudaMalloc(&W.misc, nPaths * 16));
	CK(cudaMalloc(&W.dbg, nPaths * 8));
	CK(cudaMalloc(&T.hitN, nPaths * 16));
	CK(cudaMalloc(&T.qctl.queue[0], nPaths * 4));
	CK(cudaMalloc(&T.qctl.queue[1], nPaths * 4));
	T.waveCap = nPaths;
	return PBR_OK;
}

int ensureShadow(pbr_ctx* ctx, pbr_ctx::WaveSet& T, size_t nPaths) {
	if (nPaths <= T.shadowCap) return PBR_OK;
	cudaFree(T.shadowO); cudaFree(T.shadowD); cudaFree(T.shadowQ);
	T.shadowO = T.shadowD = nullptr;
	T.shadowQ = nullptr;
	T.shadowCap = 0;
	CK(cudaMalloc(&T.shadowO, nPaths * 16));
	CK(cudaMalloc(&T.shadowD, nPaths * 16));
	CK(cudaMalloc(&T.shadowQ, nPaths * 4));
	T.shadowCap = nPaths;
	return PBR_OK;
}

void freeWaveSet(pbr_ctx::WaveSet& T) {
	WaveState& W = T.wave;
	cudaFree(W.rayO); cudaFree(W.rayD); cudaFree(W.colS); cudaFree(W.finF); cudaFree(W.misc); cudaFree(W.dbg); cudaFree(T.hitN);
	cudaFree(T.qctl.queue[0]); cudaFree(T.qctl.queue[1]); cudaFree(T.qctl.ctrl);
	cudaFree(T.shadowO); cudaFree(T.shadowD); cudaFree(T.shadowQ); cudaFree(T.frameOut);
	if (T.stream) { cudaStreamSynchronize(T.stream); cudaStreamDestroy(T.stream); }
	if (T.evTraced) cudaEventDestroy(T.evTraced);
	if (T.evConsumed) cudaEventDestroy(T.evConsumed);
}

/* extendDepth (pt_utils.cl:89-96) and the transparency branch of getNewRay (pt_brdf.cl:352-354) are the
 * only places that set addDepth.  If no material can trigger either, depthAdded stays 0 and a path ends
 * after MAX_DEPTH bounces: the MAX_ADDED_DEPTH extra wavefront iterations would all be empty. */
int updateCanExtendDepth(pbr_ctx* ctx, pbr_mem hMaterials, int brdf) {
	Mem* m = getMem(ctx, hMaterials);
	if (!m) return fail(ctx, PBR_ERR_INVALID, "pathTracing: materials argument is not a live buffer");
	if (ctx->extendCacheMem == hMaterials && ctx->extendCacheEpoch == m->epoch && ctx->extendCacheBrdf == brdf) return PBR_OK;
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
	ctx->extendCacheEpoch = m->epoch;
	ctx->extendCacheBrdf = brdf;
	return PBR_OK;
}

/* pipeline 0: one traverse + one shade launch per bounce (what the measured choice picks on all but small scenes).
 * The traverse stage is the reference-order engine (traverseKernel) or, with P.scene.wide set, the ordered walk
 * (traverseWideKernel): both leave the same t / hitFace in the path state. */
template <int BRDF, bool SHADOW, bool PHONG>
int runWavefront(pbr_ctx* ctx, const FrameParams& P, pbr_ctx::WaveSet& T, cudaStream_t stream, int nPaths) {
	WaveState W = T.wave;
	W.hitN = PHONG ? T.hitN : nullptr;
	const QueueCtl& Q = T.qctl;
	const bool useWide = !PHONG && P.scene.wide != nullptr;
	int occT = 0, occS = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occT, traverseKernel<PHONG>, 128, 0));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occS, shadeKernel<BRDF, SHADOW, PHONG>, 128, 0));
	if (ctx->traverseBlocks > 0 && ctx->traverseBlocks < occT) occT = ctx->traverseBlocks;
	const int gridT = ctx->smCount * (occT > 0 ? occT : 1);
	const int gridS = ctx->smCount * (occS > 0 ? occS : 1);
	int gridW = 0, gridWS = 0;
	size_t sharedW = 0;
	if (useWide) {
		int rc = wideLaunchShape(ctx, traverseWideKernel, &gridW, &sharedW);
		if (rc) return rc;
		if (ctx->wideBlocks > 0 && ctx->wideBlocks * ctx->smCount < gridW) gridW = ctx->wideBlocks * ctx->smCount;
	}

	/* shadow rays through the traversal engine instead of one thread per path inside the shade kernel */
	const bool shadowStage = SHADOW && P.scene.numLights > 0 && ctx->shadowStage != 0;
	int gridG = 0;
	if (shadowStage) {
		int rc = ensureShadow(ctx, T, (size_t) nPaths);
		if (rc) return rc;
		int occG = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occG, shadowGenKernel<BRDF, PHONG>, 128, 0));
		gridG = ctx->smCount * (occG > 0 ? occG : 1);
		if (useWide) {
			rc = wideLaunchShape(ctx, traverseWideShadowKernel, &gridWS, &sharedW);
			if (rc) return rc;
		}
	}
	{
		LaunchScope ls(ctx, K_RAYGEN, stream);
		raygenKernel<<<ctx->smCount * 8, 256, 0, stream>>>(P, W, Q, nPaths);
	}
	const int iterations = P.frameCount * P.samples * (P.maxDepth + (ctx->canExtendDepth ? P.maxAddedDepth : 0));
	static const bool dump = getenv("PBR_PROFILE_DUMP") != nullptr;     /* diagnostics: one line per iteration */
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
	if (dump) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
	for (int it = 0; it < iterations; it++) {
		const int in = it & 1, out = in ^ 1;
		const uint32_t* qIn = (it == 0) ? nullptr : Q.queue[in];
		if (dump) cudaEventRecord(e0, stream);
		{
			LaunchScope ls(ctx, K_TRAVERSE, stream);
			if (useWide) traverseWideKernel<<<gridW, PT_WIDE_BLOCK, sharedW, stream>>>(P.scene, W, qIn, Q.ctrl + in, Q.ctrl + 2, Q.ctrl + out, ctx->stats);
			else traverseKernel<PHONG><<<gridT, 128, 0, stream>>>(P.scene, W, qIn, Q.ctrl + in, Q.ctrl + 2, Q.ctrl + out, ctx->stats);
		}
		if (dump) cudaEventRecord(e1, stream);
		if (shadowStage) {
			{
				LaunchScope ls(ctx, K_SHADE, stream);
				shadowGenKernel<BRDF, PHONG><<<gridG, 128, 0, stream>>>(
					P, W, qIn, Q.ctrl + in, T.shadowO, T.shadowD, T.shadowQ, Q.ctrl + 3);
			}
			{
				LaunchScope ls(ctx, K_TRAVERSE, stream);
				if (useWide) traverseWideShadowKernel<<<gridWS, PT_WIDE_BLOCK, sharedW, stream>>>(
					P.scene, W, T.shadowO, T.shadowD, T.shadowQ, Q.ctrl + 3, Q.ctrl + 4, ctx->stats);
				else traverseShadowKernel<PHONG><<<gridT, 128, 0, stream>>>(
					P.scene, W, T.shadowO, T.shadowD, T.shadowQ, Q.ctrl + 3, Q.ctrl + 4, ctx->stats);
			}
			{
				LaunchScope ls(ctx, K_SHADE, stream);
				shadeKernel<BRDF, SHADOW, PHONG, true><<<gridS, 128, 0, stream>>>(
					P, W, qIn, Q.ctrl + in, Q.queue[out], Q.ctrl + out, Q.ctrl + 2, Q.ctrl + 3, Q.ctrl + 4, T.shadowO);
			}
		}
		else {
			LaunchScope ls(ctx, K_SHADE, stream);
			shadeKernel<BRDF, SHADOW, PHONG><<<gridS, 128, 0, stream>>>(P, W, qIn, Q.ctrl + in, Q.queue[out], Q.ctrl + out, Q.ctrl + 2);
		}
		if (dump) {
			cudaEventRecord(e2, stream);
			uint32_t c[4] = {0, 0, 0, 0};
			cudaMemcpyAsync(c, Q.ctrl, sizeof(c), cudaMemcpyDeviceToHost, stream);
			cudaStreamSynchronize(stream);
			float tMs = 0.0f, sMs = 0.0f;
			cudaEventElapsedTime(&tMs, e0, e1);
			cudaEventElapsedTime(&sMs, e1, e2);
			fprintf(stderr, "[wavefront%s] it %3d  traverse %7.3f ms  shade%s %7.3f ms  alive after %u\n", useWide ? ", ordered walk" : "", it, tMs,
				shadowStage ? " + shadow stage" : "", sMs, c[out]);
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
	pbr_ctx::WaveSet& T = ctx->launchSet < 0 ? ctx->serialSet : ctx->sets[ctx->launchSet];
	int rc = ensureWave(ctx, T, (size_t) nPaths);
	if (rc) return rc;
	return runWavefront<BRDF, SHADOW, PHONG>(ctx, P, T, ctx->launchStream ? ctx->launchStream : ctx->stream, nPaths);
}

#ifndef PBR_DEFAULT_NCCL_MAX_CTAS
#define PBR_DEFAULT_NCCL_MAX_CTAS 4   /* measured: NCCL's own choice 3 080, 2 / 4 / 8 blocks 3 186 / 3 186 / 3 169 Mrays/s on 2 GPUs */
#endif

/* ---- NCCL, loaded on demand ---------------------------------------------------------------------------- */

struct NcclApi {
	void* handle = nullptr;
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&ncclCommInitRank) CommInitRank = nullptr;
	decltype(&ncclCommSplit) CommSplit = nullptr;                       /* optional */
	decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr;
	decltype(&ncclAllGather) AllGather = nullptr;
	decltype(&ncclBroadcast) Broadcast = nullptr;
	decltype(&ncclGroupStart) GroupStart = nullptr;
	decltype(&ncclGroupEnd) GroupEnd = nullptr;
	decltype(&ncclGetErrorString) GetErrorString = nullptr;
	decltype(&ncclGetVersion) GetVersion = nullptr;
	std::string error;
};

NcclApi& nccl() {
	static NcclApi api;
	if (api.handle || !api.error.empty()) return api;
	/* a process that already runs NCCL (torch.distributed) gets that copy: same SONAME */
	for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
		api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
		if (api.handle) break;
	}
	if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
	bool ok = true;
	auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) { ok = false; api.error = std::string("libnccl lacks ") + n; } return p; };
	api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
	api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
	api.CommSplit = (decltype(api.CommSplit)) dlsym(api.handle, "ncclCommSplit");
	api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
	api.AllReduce = (decltype(api.AllReduce)) sym("ncclAllReduce");
	api.AllGather = (decltype(api.AllGather)) sym("ncclAllGather");
	api.Broadcast = (decltype(api.Broadcast)) sym("ncclBroadcast");
	api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
	api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
	api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
	api.GetVersion = (decltype(api.GetVersion)) sym("ncclGetVersion");
	if (!ok) { dlclose(api.handle); api.handle = nullptr; }
	return api;
}

int ncclFail(pbr_ctx* ctx, ncclResult_t r, const char* what) {
	if (r == ncclSuccess) return PBR_OK;
	std::string m = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
	if (ctx) ctx->lastError = m;
	return PBR_ERR_NCCL_BASE + (int) r;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return ncclFail(ctx, r_, #call); } while (0)

/* out = image / world: the one pass before the all-reduce (the sum of the pre-divided images is the mean) */
__global__ void scaleCopyKernel(const float4* __restrict__ in, float4* __restrict__ out, const size_t n, const float s) {
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float4 v = in[i];
		out[i] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
	}
}

/* interleaved stripes (pbr_set_tile_stripes): this rank's stripes, packed / every other rank's stripes, unpacked.
 * rowF4 = float4s per image row; local row r of rank k is image row (r / stripe) * stripe * world + k * stripe + r % stripe */
__global__ void packStripesKernel(const float4* __restrict__ image, float4* __restrict__ send, const int rowF4, const int localRows,
                                  const int stripe, const int world, const int rank) {
	const size_t n = (size_t) rowF4 * localRows, step = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
		const int r = (int) (i / rowF4), x = (int) (i % rowF4);
		const int y = (r / stripe) * stripe * world + rank * stripe + r % stripe;
		send[i] = image[(size_t) y * rowF4 + x];
	}
}

__global__ void unpackStripesKernel(float4* __restrict__ image, const float4* __restrict__ recv, const int rowF4, const int localRows,
                                    const int stripe, const int world, const int rank) {
	const size_t per = (size_t) rowF4 * localRows, n = per * world, step = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
		const int k = (int) (i / per);
		if (k == rank) continue;                   /* its own rows are in place */
		const size_t j = i % per;
		const int r = (int) (j / rowF4), x = (int) (j % rowF4);
		const int y = (r / stripe) * stripe * world + k * stripe + r % stripe;
		image[(size_t) y * rowF4 + x] = recv[i];
	}
}

int markWritten(pbr_ctx* ctx, Mem* m) {
	if (!m) return PBR_OK;
	if (!m->evWritten) CK(cudaEventCreateWithFlags(&m->evWritten, cudaEventDisableTiming));
	CK(cudaEventRecord(m->evWritten, ctx->stream));
	m->writtenValid = true;
	return PBR_OK;
}

/* A launch that overwrites an image waits for the combine that still reads it (frame k + 2 and the combine of frame k
 * share a buffer of PathTracer's ping-pong pair); with depth of field a frame also READS other pixels of imageIn, which
 * a row gather is still filling. */
int waitForCombine(pbr_ctx* ctx, Mem* m, bool ownRowsOnly = false) {
	if (!m) return PBR_OK;
	if (ownRowsOnly && m->ownRowsPending) {
		/* (the collective itself stays pending: whoever reads the whole image still waits for it) */
		CK(cudaStreamWaitEvent(ctx->stream, m->evOwnRowsFree, 0));
		m->ownRowsPending = false;
		return PBR_OK;
	}
	if (m->combinePending) {
		CK(cudaStreamWaitEvent(ctx->stream, m->evCombined, 0));
		m->combinePending = false;
		m->ownRowsPending = false;
	}
	return PBR_OK;
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
	if (cudaMalloc(&ctx->stats, 8 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMalloc(&ctx->cursor64, sizeof(unsigned long long)) != cudaSuccess) {
		delete ctx;
		return PBR_ERR_NO_DEVICE;
	}
	cudaMemset(ctx->stats, 0, 8 * sizeof(unsigned long long));
	if (const char* e = getenv("PBR_NODE_PHASE_MIN")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->nodePhaseMin = v;
	}
	if (const char* e = getenv("PBR_REFILL_MIN")) {
		const int v = atoi(e);
		if (v >= 1 && v <= 32) ctx->refillMin = v;
	}
	if (const char* e = getenv("PBR_PIPELINE")) { const int v = atoi(e); if (v >= 0 && v <= 1) { ctx->pipeline = v; ctx->pipelineAuto = false; } }
	if (const char* e = getenv("PBR_TRAVERSAL")) { const int v = atoi(e); if (v >= -1 && v <= 1) ctx->traversal = v; }
	if (const char* e = getenv("PBR_FRAMES_IN_FLIGHT")) { const int v = atoi(e); if (v >= 1 && v <= pbr_ctx::MAX_IN_FLIGHT) ctx->framesInFlight = v; }
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
	cudaFree(ctx->nodes); cudaFree(ctx->tris); cudaFree(ctx->wide); cudaFree(ctx->faceLeaf);
	for (pbr_ctx::WaveSet& T : ctx->sets) freeWaveSet(T);
	freeWaveSet(ctx->serialSet);
	if (ctx->evPrepared) cudaEventDestroy(ctx->evPrepared);
	cudaFree(ctx->stats); cudaFree(ctx->cursor64);
	cudaEventDestroy(ctx->evStart); cudaEventDestroy(ctx->evStop);
	for (const pbr_ctx::Timed& t : ctx->timedInFlight) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
	for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
	if (ctx->commStream) cudaStreamSynchronize(ctx->commStream);
	if (ctx->commReduce && ctx->commReduce != ctx->comm && nccl().CommDestroy) nccl().CommDestroy(ctx->commReduce);
	if (ctx->comm && nccl().CommDestroy) nccl().CommDestroy(ctx->comm);
	if (ctx->commStream) cudaStreamDestroy(ctx->commStream);
	if (ctx->evRendered) cudaEventDestroy(ctx->evRendered);
	cudaFree(ctx->commSend); cudaFree(ctx->commRecv);
	for (Mem& m : ctx->mems) {
		if (m.evCombined) cudaEventDestroy(m.evCombined);
		if (m.evWritten) cudaEventDestroy(m.evWritten);
		if (m.evOwnRowsFree) cudaEventDestroy(m.evOwnRowsFree);
	}
	if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
	if (ctx->evCopy) cudaEventDestroy(ctx->evCopy);
	for (int i = 0; i < 8; i++) if (ctx->evAuto[i]) cudaEventDestroy(ctx->evAuto[i]);
	cudaStreamDestroy(ctx->ownStream);
	delete ctx;
	return PBR_OK;
}

const char* pbr_last_error(pbr_ctx* ctx) { return ctx ? ctx->lastError.c_str() : "no context"; }

#ifndef PBR_BUILD_ID
#define PBR_BUILD_ID "unknown"
#endif
const char* pbr_build_id(void) { return PBR_BUILD_ID; }

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
	m->epoch = ++ctx->epochCounter;           /* only what was built from THIS buffer is rebuilt */
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
	m->writtenValid = false;
	return PBR_OK;
}

int pbr_image_read(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, float* host) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* m = getMem(ctx, image);
	if (!m || !m->image || width != m->width || height != m->height || !host)
		return fail(ctx, PBR_ERR_INVALID, "pbr_image_read: bad image or size");
	CK(cudaSetDevice(ctx->device));
	int rc = waitForCombine(ctx, m);
	if (rc) return rc;
	CK(cudaMemcpyAsync(host, m->dptr, m->bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return PBR_OK;
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
	if (m->writtenValid) CK(cudaStreamWaitEvent(ctx->copyStream, m->evWritten, 0));
	else if (!m->combinePending) {
		CK(cudaEventRecord(ctx->evCopy, ctx->stream));
		CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evCopy, 0));
	}
	if (m->combinePending) CK(cudaStreamWaitEvent(ctx->copyStream, m->evCombined, 0));
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
	d->writtenValid = false;
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
		m.epoch = ++ctx->epochCounter;
	}
	ctx->cacheBvh = 0;                         /* the repacked scene belongs to buffers that are gone */
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
	/* (MAX_DEPTH 0 would render nothing: the reference's depth loop does not run and every pixel becomes
	 *  mix(0, imageIn, weight); rejected here rather than served by two pipelines that disagree about it) */
	if (d.img_width <= 0 || d.img_height <= 0 || d.samples <= 0 || d.max_depth < 1 || d.max_added_depth < 0)
		return fail(ctx, PBR_ERR_INVALID, "invalid image size / samples / depth (render.max_depth must be at least 1)");
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

/* Everything a launch needs that is built once per scene and cached: the repacked scene, the "can a path be extended"
 * answer, the walk (and with it the 4-wide BVH).  Work goes to the context's stream. */
static int prepareLaunch(pbr_ctx* ctx) {
	KernelArgs& a = ctx->args;
	const pbr_defines& D = ctx->defines;
	const bool phong = (D.phongtess == 1);
	int rc = ensureScene(ctx, a.mem[4], a.mem[5], a.mem[7], D.bvh_num_nodes, phong, a.mem[6], a.mem[8]);
	if (rc) return rc;
	rc = updateCanExtendDepth(ctx, a.mem[9], D.brdf);
	if (rc) return rc;
	/* the reference's visit counters are observable through the debug image only */
	bool useWide = false;
	rc = chooseTraversal(ctx, ctx->debugImage || phong, &useWide);
	if (rc) return rc;
	ctx->launchUseWide = useWide;
	/* frames traced on streams of their own start behind whatever the preparation has put on the context's stream */
	if (!ctx->evPrepared) CK(cudaEventCreateWithFlags(&ctx->evPrepared, cudaEventDisableTiming));
	if (ctx->preparedVersion != ctx->geometryVersion || ctx->preparedWide != (ctx->wideBuilt ? ctx->wideVersion : ~0ull)) {
		CK(cudaEventRecord(ctx->evPrepared, ctx->stream));
		ctx->preparedVersion = ctx->geometryVersion;
		ctx->preparedWide = ctx->wideBuilt ? ctx->wideVersion : ~0ull;
	}
	return PBR_OK;
}

/* n consecutive frames (n <= PT_MAX_BATCH) in one pass over the device: frame f uses seeds[f] / weights[f];
 * frame 0 reads hIn, every later frame reads what the one before wrote to hOut. */
static int launchFrames(pbr_ctx* ctx, int n, const float* seeds, const float* weights, pbr_mem hIn, pbr_mem hOut) {
	KernelArgs& a = ctx->args;
	const pbr_defines& D = ctx->defines;

	const bool phong = (D.phongtess == 1);
	int rc = prepareLaunch(ctx);
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

	const bool useWide = ctx->launchUseWide;
	ctx->lastTraversal = useWide ? 1 : 0;

	FrameParams P;
	fillScene(ctx, P.scene, useWide);
	P.scene.lights = (const pbr_light*) lights->dptr;
	P.scene.numLights = D.num_lights;
	P.scene.phongAlpha = D.phongtess_alpha;
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
	P.frameOut = ctx->launchFrameOut;
	P.stats = ctx->stats;

	const int nPaths = D.img_width * (P.y1 - P.y0);
	const bool shadow = (D.shadow_rays == 1);
	ctx->lastFrameParams = P;
	ctx->lastNumPaths = nPaths;

	const int variant = (D.brdf == 1 ? 4 : 0) | (shadow ? 2 : 0) | (phong ? 1 : 0);

	/* which pipeline: the caller's, or the one that measured faster for this configuration */
	const int callerPipeline = ctx->pipeline;
	int timing = -1;
	if (ctx->pipelineAuto && n == 1) {
		const unsigned long long key = ctx->geometryVersion * 0x9e3779b97f4a7c15ull ^ ((unsigned long long) nPaths << 20) ^
			((unsigned long long) variant << 8) ^ ((unsigned long long) D.max_depth << 12) ^ (unsigned long long) D.samples ^
			((unsigned long long) (useWide ? 1 : 0) << 40);
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
		ctx->pipeline = 0;
	}
	if (ctx->pipeline == 1) ctx->lastTraversal = 0;       /* the megakernel walks in the reference's order */
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

/* Several frames in flight: a frame's rays do not depend on the frame before it -- only the final mix of a pixel does
 * (setColors reads the previous image, pt_rgb.cl:15).  So consecutive frames are traced on streams of their own, each
 * into its own wave state, finished pixels leave their radiance in a per-frame buffer, and mixFrameKernel folds the
 * frames into imageOut in order on the context's stream: the long tail of one frame's traverse launches is filled with
 * the next frame's work.  Same operands, same operations, same bits.  Not with the megakernel (nothing to overlap),
 * the debug image (written by the shade kernels) or depth of field (a frame reads the previous image when it starts). */
static bool overlapEligible(pbr_ctx* ctx) {
	const KernelArgs& a = ctx->args;
	if (ctx->framesInFlight < 2 || ctx->debugImage) return false;
	if (a.cam.focusPoint.x >= 0 && a.cam.focusPoint.y >= 0) return false;
	if (ctx->pipelineAuto) return ctx->autoState >= 6 ? ctx->autoChoice == 0 : false;    /* (while the two pipelines are being timed: no) */
	return ctx->pipeline == 0;
}

static int launchOverlapped(pbr_ctx* ctx, const float* seed, const float* weight, pbr_mem hIn, pbr_mem hOut) {
	Mem* outM = getMem(ctx, hOut);
	if (!outM) return fail(ctx, PBR_ERR_INVALID, "pathTracing: imageOut is not live");
	const size_t outF4 = outM->bytes / 16;
	int rc = prepareLaunch(ctx);                   /* (scene repack, wide BVH: on the context's stream, before the fork) */
	if (rc) return rc;
	pbr_ctx::WaveSet& T = ctx->sets[ctx->overlapCounter % (uint64_t) ctx->framesInFlight];
	ctx->overlapCounter++;
	if (!T.stream) {
		CK(cudaStreamCreateWithFlags(&T.stream, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&T.evTraced, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&T.evConsumed, cudaEventDisableTiming));
	}
	if (outF4 > T.frameOutCap) {
		if (T.consumedPending) { CK(cudaEventSynchronize(T.evConsumed)); T.consumedPending = false; }
		cudaFree(T.frameOut);
		T.frameOut = nullptr;
		T.frameOutCap = 0;
		CK(cudaMalloc(&T.frameOut, outF4 * 16));
		T.frameOutCap = outF4;
	}
	/* this set's stream: behind the scene preparation, and behind the mix that last read its buffer */
	CK(cudaStreamWaitEvent(T.stream, ctx->evPrepared, 0));
	if (T.consumedPending) { CK(cudaStreamWaitEvent(T.stream, T.evConsumed, 0)); T.consumedPending = false; }
	const int savedPipeline = ctx->pipeline;
	const bool savedAuto = ctx->pipelineAuto;
	ctx->pipeline = 0;
	ctx->pipelineAuto = false;
	ctx->launchSet = (int) (&T - ctx->sets);
	ctx->launchStream = T.stream;
	ctx->launchFrameOut = T.frameOut;
	rc = launchFrames(ctx, 1, seed, weight, hIn, hOut);
	ctx->launchSet = -1;
	ctx->launchStream = nullptr;
	ctx->launchFrameOut = nullptr;
	ctx->pipeline = savedPipeline;
	ctx->pipelineAuto = savedAuto;
	if (rc) return rc;
	CK(cudaEventRecord(T.evTraced, T.stream));
	CK(cudaStreamWaitEvent(ctx->stream, T.evTraced, 0));
	rc = waitForCombine(ctx, outM, true);          /* the mix reads and writes this rank's rows only */
	if (rc) return rc;
	{
		LaunchScope ls(ctx, K_SHADE);
		mixFrameKernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(ctx->lastFrameParams, T.frameOut, ctx->lastNumPaths);
	}
	CK(cudaGetLastError());
	CK(cudaEventRecord(T.evConsumed, ctx->stream));
	T.consumedPending = true;
	return PBR_OK;
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
	const bool dof = a.cam.focusPoint.x >= 0 && a.cam.focusPoint.y >= 0;
	rc = waitForCombine(ctx, getMem(ctx, a.mem[12]), !dof);
	if (rc) return rc;
	if (dof) { rc = waitForCombine(ctx, getMem(ctx, a.mem[11])); if (rc) return rc; }
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	if (overlapEligible(ctx)) rc = launchOverlapped(ctx, &a.seed, &a.pixelWeight, a.mem[11], a.mem[12]);
	else rc = launchFrames(ctx, 1, &a.seed, &a.pixelWeight, a.mem[11], a.mem[12]);
	if (rc) return rc;
	rc = markWritten(ctx, getMem(ctx, a.mem[12]));
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
	const bool depthOfField = a.cam.focusPoint.x >= 0 && a.cam.focusPoint.y >= 0;
	rc = waitForCombine(ctx, getMem(ctx, hOut), !depthOfField);
	if (rc) return rc;
	rc = waitForCombine(ctx, getMem(ctx, hIn), !depthOfField);
	if (rc) return rc;
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	Mem* outM = getMem(ctx, hOut);
	if (!outM) return fail(ctx, PBR_ERR_INVALID, "pathTracing: imageOut is not live");
	const bool combine = ctx->batchCombineMode >= 0 && ctx->comm != nullptr;
	if (!depthOfField) {
		/* pixels are independent of each other, so everything after frame 0 accumulates in place in imageOut.  While the
		 * measured pipeline choice is still open the first frames of the batch are its timing rounds; after that the
		 * frames overlap (launchOverlapped) unless the megakernel won, the debug image is on, or frames_in_flight is 1 */
		for (int f = 0; f < n_frames; f++) {
			const pbr_mem in = f == 0 ? hIn : hOut;
			if (overlapEligible(ctx)) rc = launchOverlapped(ctx, seeds + f, pixel_weights + f, in, hOut);
			else {
				if (f > 0) { rc = waitForCombine(ctx, outM, true); if (rc) return rc; }
				rc = launchFrames(ctx, 1, seeds + f, pixel_weights + f, in, hOut);
			}
			if (rc) return rc;
			if (combine) {
				rc = pbr_frame_combine(ctx, hOut, ctx->batchCombineMode, (ctx->batchCombineOuts > 0 ? ctx->batchCombineOut[(ctx->batchCombineFirst + f) % ctx->batchCombineOuts] : 0));
				if (rc) return rc;
				outM = getMem(ctx, hOut);
			}
		}
	}
	else {
		/* depthOfField reads ANOTHER pixel of the previous frame (pathtracing.cl:41-43): frame by frame,
		 * ping-pong between imageOut and a scratch image so that the last frame lands in imageOut */
		Mem* out = getMem(ctx, hOut);
		if (!out) return fail(ctx, PBR_ERR_INVALID, "pathTracing: imageOut is not live");
		const size_t outBytes = out->bytes, outW = out->width, outH = out->height;
		Mem* scratch = getMem(ctx, ctx->scratchImage);
		if (ctx->scratchImage == 0 || !scratch || scratch->bytes < outBytes) {
			Mem* m = nullptr;
			rc = newMem(ctx, outBytes, &ctx->scratchImage, &m);
			if (rc) return rc;
			m->image = true; m->width = outW; m->height = outH;
		}
		pbr_mem prev = hIn;
		for (int f = 0; f < n_frames; f++) {
			const pbr_mem dst = ((n_frames - 1 - f) % 2 == 0) ? hOut : ctx->scratchImage;
			rc = waitForCombine(ctx, getMem(ctx, dst));
			if (rc) return rc;
			rc = launchFrames(ctx, 1, seeds + f, pixel_weights + f, prev, dst);
			if (rc) return rc;
			if (combine) {
				CK(cudaStreamSynchronize(ctx->stream));          /* (depth of field reads what the gather writes: no overlap here) */
				rc = pbr_frame_combine(ctx, dst, ctx->batchCombineMode, (ctx->batchCombineOuts > 0 ? ctx->batchCombineOut[(ctx->batchCombineFirst + f) % ctx->batchCombineOuts] : 0));
				if (rc) return rc;
				rc = pbr_comm_fence(ctx);
				if (rc) return rc;
			}
			prev = dst;
		}
	}
	rc = markWritten(ctx, getMem(ctx, hOut));
	if (rc) return rc;
	if (Mem* scratch = getMem(ctx, ctx->scratchImage)) scratch->writtenValid = false;
	CK(cudaEventRecord(ctx->evStop, ctx->stream));
	ctx->timed = true;
	return PBR_OK;
}

int pbr_finish(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	if (ctx->commStream) CK(cudaStreamSynchronize(ctx->commStream));
	return PBR_OK;
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
	if (!ctx) return PBR_ERR_INVALID;
	if (mode == 2 || mode == 3) return fail(ctx, PBR_ERR_UNSUPPORTED, "pipelines 2 (persistent) and 3 (carry-over) of round 1 measured slower on every scene and were removed");
	if (mode < -1 || mode > 1) return PBR_ERR_INVALID;
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
	else if (k == "traverse_blocks" && value >= 0 && value <= 32) ctx->traverseBlocks = value;
	else if (k == "wide_blocks" && value >= 0 && value <= 32) ctx->wideBlocks = value;
	else if (k == "wide_node_phase_min" && value >= 1 && value <= 32) ctx->wideNodePhaseMin = value;
	else if (k == "wide_refill_min" && value >= 1 && value <= 32) ctx->wideRefillMin = value;
	else if (k == "wide_top" && value >= 1 && value <= 1365) ctx->wideTopBudget = value;      /* rebuilt at the next launch */
	else if (k == "shadow_stage" && (value == 0 || value == 1)) ctx->shadowStage = value;
	else if (k == "frames_in_flight" && value >= 1 && value <= pbr_ctx::MAX_IN_FLIGHT) ctx->framesInFlight = value;
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

int pbr_set_traversal(pbr_ctx* ctx, int32_t mode) {
	if (!ctx || mode < -1 || mode > 1) return PBR_ERR_INVALID;
	ctx->traversal = mode;
	ctx->autoState = 0;                        /* the pipeline choice was measured with the other walk */
	ctx->autoKey = 0;
	return PBR_OK;
}

int pbr_traversal_info(pbr_ctx* ctx, pbr_traversal_info_t* out, int32_t reset) {
	if (!ctx || !out) return PBR_ERR_INVALID;
	CK(cudaSetDevice(ctx->device));
	memset(out, 0, sizeof(*out));
	unsigned long long c[2] = {0, 0};
	CK(cudaMemcpyAsync(c, ctx->stats + 6, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
	if (reset) CK(cudaMemsetAsync(ctx->stats + 6, 0, sizeof(c), ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	out->mode = ctx->traversal;
	out->last_used = ctx->lastTraversal;
	out->wide_available = (ctx->wideBuilt && ctx->wideOk) ? 1 : 0;
	out->wide_nodes = ctx->wideCount;
	out->wide_top = ctx->wideTop;
	out->wide_depth = ctx->wideDepth;
	out->wide_build_ms = ctx->wideBuildMs;
	out->rewalked_rays = c[0];
	out->ordered_rays = c[1];
	strncpy(out->why_not, ctx->wideWhy.c_str(), sizeof(out->why_not) - 1);
	return PBR_OK;
}

/* ---- multi-GPU ------------------------------------------------------------------------------------------- */

int pbr_tile_rows(int32_t height, int32_t rank, int32_t world, int32_t* y0, int32_t* y1) {
	if (height <= 0 || world < 1 || rank < 0 || rank >= world || !y0 || !y1) return PBR_ERR_INVALID;
	const long long align = 4, blocks = (height + align - 1) / align;
	const long long b0 = blocks * rank / world, b1 = blocks * (rank + 1) / world;
	*y0 = (int32_t) std::min<long long>(b0 * align, height);
	*y1 = (int32_t) std::min<long long>(b1 * align, height);
	return PBR_OK;
}

int pbr_comm_unique_id(void* id128) {
	if (!id128) return PBR_ERR_INVALID;
	NcclApi& N = nccl();
	if (!N.handle) return PBR_ERR_UNSUPPORTED;
	ncclUniqueId id;
	const ncclResult_t r = N.GetUniqueId(&id);
	if (r != ncclSuccess) return PBR_ERR_NCCL_BASE + (int) r;
	memcpy(id128, &id, sizeof(id));
	return PBR_OK;
}

int pbr_comm_init(pbr_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
	if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return PBR_ERR_INVALID;
	if (ctx->comm) return fail(ctx, PBR_ERR_INVALID, "pbr_comm_init: this context already has a communicator");
	NcclApi& N = nccl();
	if (!N.handle) return fail(ctx, PBR_ERR_UNSUPPORTED, "pbr_comm_init: " + N.error);
	CK(cudaSetDevice(ctx->device));
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	NK(N.CommInitRank(&ctx->comm, world, id, rank));
	/* The per-frame all-reduce (33 MB at 1080p) shares the SMs with persistent traversal kernels and has a whole frame
	 * to finish in: on a communicator of its own it is held to a few blocks (PBR_NCCL_MAX_CTAS; 0 = NCCL's own choice:
	 * 3 080 instead of 3 186 Mrays/s on 2 GPUs, 12 355 instead of 12 676 on 8).  The row gather (132 MB at 4K, on the
	 * critical path of ONE image) keeps NCCL's choice: held to 4 blocks it costs a quarter of the strong scaling. */
	int maxCtas = PBR_DEFAULT_NCCL_MAX_CTAS;
	if (const char* e = getenv("PBR_NCCL_MAX_CTAS")) maxCtas = atoi(e);
	ctx->commReduce = ctx->comm;
	if (maxCtas > 0 && N.CommSplit) {
		ncclConfig_t config = NCCL_CONFIG_INITIALIZER;
		config.minCTAs = 1;
		config.maxCTAs = maxCtas;
		ncclComm_t split = nullptr;
		NK(N.CommSplit(ctx->comm, 0, rank, &split, &config));
		if (split) ctx->commReduce = split;
	}
	ctx->commRank = rank;
	ctx->commWorld = world;
	/* default (lowest) priority, measured: at the highest priority the collective's kernel takes its blocks early and then
	 * spins on them waiting for its peers -- 11.46 instead of 12.36 Grays/s on 8 GPUs (profiles/r02p_comm_priority.log) */
	CK(cudaStreamCreateWithFlags(&ctx->commStream, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&ctx->evRendered, cudaEventDisableTiming));
	return PBR_OK;
}

int pbr_comm_info(pbr_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
	if (!ctx) return PBR_ERR_INVALID;
	if (rank) *rank = ctx->commRank;
	if (world) *world = ctx->comm ? ctx->commWorld : 1;
	if (nccl_version) {
		int v = 0;
		if (ctx->comm && nccl().GetVersion) nccl().GetVersion(&v);
		*nccl_version = v;
	}
	return PBR_OK;
}

int pbr_comm_destroy(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	if (!ctx->comm) return PBR_OK;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->commStream));
	if (ctx->commReduce && ctx->commReduce != ctx->comm) NK(nccl().CommDestroy(ctx->commReduce));
	ctx->commReduce = nullptr;
	NK(nccl().CommDestroy(ctx->comm));
	ctx->comm = nullptr;
	ctx->commWorld = 1;
	ctx->commRank = 0;
	for (Mem& m : ctx->mems) { m.combinePending = false; m.ownRowsPending = false; }
	return PBR_OK;
}

/* One collective per frame, enqueued on the communicator's own stream behind everything the render
 * stream has been given so far, so that the next frame is traced while this one crosses NVLink. */
int pbr_frame_combine(pbr_ctx* ctx, pbr_mem image, int32_t mode, pbr_mem out) {
	if (!ctx) return PBR_ERR_INVALID;
	Mem* img = getMem(ctx, image);
	if (!img || !img->image) return fail(ctx, PBR_ERR_INVALID, "pbr_frame_combine: not an image");
	if (!ctx->comm) return fail(ctx, PBR_ERR_NOT_READY, "pbr_frame_combine before pbr_comm_init");
	NcclApi& N = nccl();
	CK(cudaSetDevice(ctx->device));
	const int world = ctx->commWorld, rank = ctx->commRank;
	const size_t nF4 = img->width * img->height;
	CK(cudaEventRecord(ctx->evRendered, ctx->stream));
	CK(cudaStreamWaitEvent(ctx->commStream, ctx->evRendered, 0));
	Mem* pending = img;
	if (mode == PBR_COMBINE_SPP) {
		Mem* dst = getMem(ctx, out);
		if (!dst || !dst->image || dst->bytes != img->bytes || dst == img)
			return fail(ctx, PBR_ERR_INVALID, "pbr_frame_combine(SPP): `out` must be another image of the same size");
		if (dst->combinePending) CK(cudaStreamWaitEvent(ctx->commStream, dst->evCombined, 0));
		ctx->prof.launches++; ctx->prof.other_launches++;
		scaleCopyKernel<<<ctx->smCount * 4, 256, 0, ctx->commStream>>>((const float4*) img->dptr, (float4*) dst->dptr, nF4, 1.0f / (float) world);
		CK(cudaGetLastError());
		/* `image` is free again once it has been copied; `out` is busy until the collective is done */
		if (!img->evCombined) CK(cudaEventCreateWithFlags(&img->evCombined, cudaEventDisableTiming));
		CK(cudaEventRecord(img->evCombined, ctx->commStream));
		img->combinePending = true;
		pending = nullptr;
		NK(N.AllReduce(dst->dptr, dst->dptr, nF4 * 4, ncclFloat, ncclSum, ctx->commReduce, ctx->commStream));
		if (!dst->evCombined) CK(cudaEventCreateWithFlags(&dst->evCombined, cudaEventDisableTiming));
		CK(cudaEventRecord(dst->evCombined, ctx->commStream));
		dst->combinePending = true;
	}
	else if (mode == PBR_COMBINE_ROWS) {
		const int H = (int) img->height, rowF4 = (int) img->width;
		if (ctx->stripeRows > 0) {
			if (ctx->stripeWorld != world || ctx->stripeRank != rank || H % (ctx->stripeRows * world) != 0)
				return fail(ctx, PBR_ERR_INVALID, "pbr_frame_combine(ROWS): pbr_set_tile_stripes does not match the communicator");
			const int localRows = H / world;
			const size_t per = (size_t) localRows * rowF4;
			if (per > ctx->commSendCap) { cudaFree(ctx->commSend); ctx->commSend = nullptr; CK(cudaMalloc(&ctx->commSend, per * 16)); ctx->commSendCap = per; }
			if (per * world > ctx->commRecvCap) { cudaFree(ctx->commRecv); ctx->commRecv = nullptr; CK(cudaMalloc(&ctx->commRecv, per * world * 16)); ctx->commRecvCap = per * world; }
			ctx->prof.launches += 2; ctx->prof.other_launches += 2;
			packStripesKernel<<<ctx->smCount * 4, 256, 0, ctx->commStream>>>((const float4*) img->dptr, ctx->commSend, rowF4, localRows, ctx->stripeRows, world, rank);
			if (!img->evOwnRowsFree) CK(cudaEventCreateWithFlags(&img->evOwnRowsFree, cudaEventDisableTiming));
			CK(cudaEventRecord(img->evOwnRowsFree, ctx->commStream));
			img->ownRowsPending = true;
			NK(N.AllGather(ctx->commSend, ctx->commRecv, per * 4, ncclFloat, ctx->comm, ctx->commStream));
			unpackStripesKernel<<<ctx->smCount * 4, 256, 0, ctx->commStream>>>((float4*) img->dptr, ctx->commRecv, rowF4, localRows, ctx->stripeRows, world, rank);
			CK(cudaGetLastError());
		}
		else {
			/* contiguous row blocks, the partition of pbr_tile_rows: gathered in place, no packing */
			std::vector<int32_t> y0((size_t) world), y1((size_t) world);
			bool equal = true;
			for (int r = 0; r < world; r++) {
				pbr_tile_rows(H, r, world, &y0[(size_t) r], &y1[(size_t) r]);
				if (y1[(size_t) r] - y0[(size_t) r] != y1[0] - y0[0]) equal = false;
			}
			const int myY0 = ctx->tileY0 < 0 ? 0 : ctx->tileY0, myY1 = ctx->tileY1 < 0 ? H : ctx->tileY1;
			if (myY0 != y0[(size_t) rank] || myY1 != y1[(size_t) rank])
				return fail(ctx, PBR_ERR_INVALID, "pbr_frame_combine(ROWS): pbr_set_tile is not this rank's block of pbr_tile_rows");
			float* base = (float*) img->dptr;
			if (equal) {
				const size_t cnt = (size_t) (y1[0] - y0[0]) * rowF4 * 4;
				NK(N.AllGather(base + (size_t) rank * cnt, base, cnt, ncclFloat, ctx->comm, ctx->commStream));
			}
			else {
				NK(N.GroupStart());
				for (int r = 0; r < world; r++) {
					float* p = base + (size_t) y0[(size_t) r] * rowF4 * 4;
					const size_t cnt = (size_t) (y1[(size_t) r] - y0[(size_t) r]) * rowF4 * 4;
					const ncclResult_t e = N.Broadcast(p, p, cnt, ncclFloat, r, ctx->comm, ctx->commStream);
					if (e != ncclSuccess) { N.GroupEnd(); return ncclFail(ctx, e, "ncclBroadcast"); }
				}
				NK(N.GroupEnd());
			}
		}
	}
	else return fail(ctx, PBR_ERR_INVALID, "pbr_frame_combine: unknown mode");
	if (pending) {
		if (!pending->evCombined) CK(cudaEventCreateWithFlags(&pending->evCombined, cudaEventDisableTiming));
		CK(cudaEventRecord(pending->evCombined, ctx->commStream));
		pending->combinePending = true;
	}
	ctx->combines++;
	return PBR_OK;
}

int pbr_set_batch_combine(pbr_ctx* ctx, int32_t mode, const pbr_mem* outs, int32_t n_outs, int32_t first) {
	if (!ctx || mode < -1 || mode > PBR_COMBINE_ROWS || n_outs < 0 || n_outs > 8 || (n_outs > 0 && !outs)) return PBR_ERR_INVALID;
	if (mode == PBR_COMBINE_SPP && n_outs < 1) return fail(ctx, PBR_ERR_INVALID, "pbr_set_batch_combine(SPP) needs at least one display image");
	ctx->batchCombineMode = mode;
	ctx->batchCombineOuts = n_outs;
	for (int i = 0; i < n_outs; i++) ctx->batchCombineOut[i] = outs[i];
	ctx->batchCombineFirst = n_outs > 0 ? ((first % n_outs) + n_outs) % n_outs : 0;
	return PBR_OK;
}

/* Block the render stream (not the host) until the combines enqueued so far are done. */
int pbr_comm_fence(pbr_ctx* ctx) {
	if (!ctx) return PBR_ERR_INVALID;
	if (!ctx->comm) return PBR_OK;
	CK(cudaSetDevice(ctx->device));
	for (Mem& m : ctx->mems) {
		if (m.alive && m.combinePending) {
			CK(cudaStreamWaitEvent(ctx->stream, m.evCombined, 0));
			m.combinePending = false;
			m.ownRowsPending = false;
		}
	}
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
	/* pbr_hit carries the visit counters, so the automatic choice is the reference-order walk; the ordered walk on
	 * request (closest hit only: WHICH face ends an any-hit walk depends on the visiting order, see pt_wide.cuh) */
	bool useWide = false;
	if (ctx->traversal == 1 && !any_hit) {
		rc = chooseTraversal(ctx, false, &useWide);
		if (rc) return rc;
	}
	ctx->lastTraversal = useWide ? 1 : 0;
	SceneDev S;
	fillScene(ctx, S, useWide);
	S.numLights = 0;
	S.lights = nullptr;
	S.phongAlpha = 0.0f;
	if (num_lights > 0) {
		Mem* l = getMem(ctx, lights);
		if (!l || (size_t) num_lights * sizeof(pbr_light) > l->bytes) return fail(ctx, PBR_ERR_INVALID, "pbr_trace: bad lights buffer");
		S.lights = (const pbr_light*) l->dptr;
		S.numLights = num_lights;
	}
	if (n <= 0) return PBR_OK;
	CK(cudaMemsetAsync(ctx->cursor64, 0, sizeof(unsigned long long), ctx->stream));
	int occ = 0, grid = 0;
	size_t shared = 0;
	if (useWide) {
		rc = wideLaunchShape(ctx, traceRaysWideKernel, &grid, &shared);
		if (rc) return rc;
		if (ctx->wideBlocks > 0 && ctx->wideBlocks * ctx->smCount < grid) grid = ctx->wideBlocks * ctx->smCount;
	}
	else {
		if (any_hit) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, traceRaysKernel<true>, 128, 0));
		else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, traceRaysKernel<false>, 128, 0));
		grid = ctx->smCount * (occ > 0 ? occ : 1);
	}
	CK(cudaEventRecord(ctx->evStart, ctx->stream));
	{
		LaunchScope ls(ctx, K_TRAVERSE);
		if (useWide) traceRaysWideKernel<<<grid, PT_WIDE_BLOCK, shared, ctx->stream>>>(S, dRays, (long long) n, dHits, ctx->cursor64, ctx->stats);
		else if (any_hit) traceRaysKernel<true><<<grid, 128, 0, ctx->stream>>>(S, dRays, (long long) n, dHits, ctx->cursor64, ctx->stats);
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
	struct Guard { float** p[3]; ~Guard() { for (float** q : p) if (*q) cudaFree(*q); } } guard = {{&dx, &dy, &dout}};
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
	return PBR_OK;
}

} /* extern "C" */
