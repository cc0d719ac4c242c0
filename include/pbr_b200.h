/*
 * pbr_b200.h -- C ABI of libpbr_b200.so, the B200 (sm_100a) replacement for the reference's
 * OpenCL device layer.
 *
 * The reference reaches its device through the C++ class `CL` (source/CL.h:20-83, source/CL.cpp),
 * a thin wrapper over the OpenCL 1.1 C API, and runs exactly one kernel, `pathTracing`
 * (source/opencl/pathtracing.cl:207-334).  Every entry point below names the `CL` method (and the
 * OpenCL call underneath it) that it replaces.  A header-only `CL` shim with the reference's method
 * names (physically-based-rendering_b200/host/CL.h) forwards to these, so the reference's
 * PathTracer code binds to this library without changes to its call sites; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: opaque context pointer, 64-bit handles, raw pointers + sizes.  No torch / CUDA types.
 *   - every function returns 0 on success, else a non-zero code (cudaError_t value, or
 *     PBR_ERR_* below); pbr_last_error(ctx) returns the message.  The `CL` shim maps failures to
 *     the reference's behaviour (log "[OpenCL] Error in function ..." and continue, or
 *     exit(EXIT_FAILURE) where the reference does: CL.cpp:73-79,209-211,347-350,438-448,523-566).
 *   - host pointers are borrowed for the duration of a call; the context owns all device memory
 *     and every handle dies with pbr_destroy (reference: CL::~CL, CL.cpp:30-52).
 *   - one host thread per context; device work is issued on one CUDA stream per context and is
 *     asynchronous until pbr_finish / pbr_image_read / pbr_buffer_read.
 *   - there is NO CPU fallback: without a CUDA device pbr_create fails.
 */
#ifndef PBR_B200_H
#define PBR_B200_H

#include <stddef.h>
#include <stdint.h>

#include "pbr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbr_ctx pbr_ctx;
typedef uint64_t pbr_mem;      /* stands in for cl_mem    (buffer or image) */
typedef uint64_t pbr_kernel;   /* stands in for cl_kernel */

#define PBR_OK 0
#define PBR_ERR_INVALID 10001      /* bad handle / argument          (CL_INVALID_*)            */
#define PBR_ERR_NO_DEVICE 10002    /* no usable CUDA device          (CL_DEVICE_NOT_FOUND)     */
#define PBR_ERR_NOT_READY 10003    /* launch before program/args set (CL_INVALID_KERNEL_ARGS)  */
#define PBR_ERR_UNSUPPORTED 10004  /* define / kernel name not known (CL_INVALID_KERNEL_NAME)  */
#define PBR_ERR_NCCL_BASE 20000    /* + ncclResult_t of a failed NCCL call                     */

/* ---- context ------------------------------------------------------------------------------- */

/* CL::CL() -> getDefaultPlatform / getDefaultDevice / initContext / initCommandQueue
 * (CL.cpp:10-24, 338-473, 513-547).  device = CUDA ordinal, -1 = current device. */
int pbr_create(int device, pbr_ctx** out);
/* CL::~CL() (CL.cpp:30-52): releases every buffer, image, kernel and the stream. */
int pbr_destroy(pbr_ctx* ctx);
const char* pbr_last_error(pbr_ctx* ctx);
/* Device name and SM count (CL::getDefaultDevice debug dump, CL.cpp:377-419). */
/* Twelve hex digits identifying the sources this library was built from (sha1 over the .cu / .cuh / .h files of csrc and the public
 * headers).  Profiles under profiles/ record it; bench.py only quotes counters of the build it is timing. */
const char* pbr_build_id(void);
int pbr_device_info(pbr_ctx* ctx, char* name, size_t name_len, int* sm_count, size_t* total_mem);

/* ---- buffers and images -------------------------------------------------------------------- */

/* CL::createBuffer<T>(vector<T>, bytes) = clCreateBuffer(READ_ONLY|COPY_HOST_PTR) (CL.h:26-33). */
int pbr_buffer_create(pbr_ctx* ctx, const void* host, size_t bytes, pbr_mem* out);
/* CL::createEmptyBuffer(size, flags) (CL.cpp:136-144). */
int pbr_buffer_create_empty(pbr_ctx* ctx, size_t bytes, pbr_mem* out);
/* CL::updateBuffer(buffer, size, data) = clEnqueueWriteBuffer, blocking (CL.cpp:715-728). */
int pbr_buffer_update(pbr_ctx* ctx, pbr_mem buf, size_t bytes, const void* host);
/* Additive: read a buffer back (tests). */
int pbr_buffer_read(pbr_ctx* ctx, pbr_mem buf, size_t bytes, void* host);
/* CL::createImage2DReadOnly(w,h,data) / createImage2DWriteOnly(w,h): CL_RGBA / CL_FLOAT 2-D images
 * (CL.cpp:153-197).  host may be NULL (write-only image, zero-filled). */
int pbr_image_create(pbr_ctx* ctx, size_t width, size_t height, const float* host, pbr_mem* out);
/* CL::updateImageReadOnly = clEnqueueWriteImage, blocking (CL.cpp:738-753). */
int pbr_image_write(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, const float* host);
/* CL::readImageOutput = clEnqueueReadImage, blocking (CL.cpp:581-594). */
int pbr_image_read(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, float* host);
/* Additive: the same read-back split in two, so that work enqueued between the two calls (the next frame)
 * overlaps the copy.  _begin orders the copy after everything enqueued so far and runs it on a stream of
 * its own (host should be pinned: pbr_host_alloc); _end blocks until it has landed.  One read in flight. */
int pbr_image_read_begin(pbr_ctx* ctx, pbr_mem image, size_t width, size_t height, float* host);
int pbr_image_read_end(pbr_ctx* ctx);
/* Additive: device-side copy src -> dst.  Replaces the per-frame host round trip
 * readImageOutput(imageOut) ; updateImageReadOnly(imageIn) of PathTracer::generateImage
 * (PathTracer.cpp:61-66) when the host copy has not been modified in between. */
int pbr_image_copy(pbr_ctx* ctx, pbr_mem dst, pbr_mem src);
/* Additive: raw device pointer of a buffer / image (zero-copy interop with a caller that already
 * owns device memory or wants to run a collective on the accumulation buffer). */
int pbr_mem_device_ptr(pbr_ctx* ctx, pbr_mem mem, void** dev_ptr, size_t* bytes);
/* CL::freeBuffers (CL.cpp:322-331). */
int pbr_free_buffers(pbr_ctx* ctx);
/* Additive: pinned host memory so image reads/writes run at full PCIe rate. */
int pbr_host_alloc(pbr_ctx* ctx, size_t bytes, void** out);
int pbr_host_free(pbr_ctx* ctx, void* ptr);

/* ---- program ------------------------------------------------------------------------------- */

/* CL::setReplacement("#NAME#", value) (CL.cpp:615-617).  Accepted names (with or without the
 * surrounding '#'): BVH_NUM_NODES, NUM_LIGHTS, SKY_LIGHT ("(float4)( r, g, b, 0.0f )"). */
int pbr_set_define(pbr_ctx* ctx, const char* name, const char* value);
/* CL::loadProgram(path) = combineParts + setValues + clCreateProgramWithSource + buildProgram
 * (CL.cpp:554-571, 107-127, 626-705, 58-80).  There is no source to compile: the values that
 * CL::setValues splices into pt_header.cl arrive in `defines` (bvh_num_nodes / num_lights /
 * sky_light are overridden by earlier pbr_set_define calls) and select a precompiled sm_100a
 * kernel specialisation. */
int pbr_program_load(pbr_ctx* ctx, const pbr_defines* defines);
/* CL::createKernel(name) (CL.cpp:205-217).  Only "pathTracing" exists. */
int pbr_kernel_get(pbr_ctx* ctx, const char* name, pbr_kernel* out);
/* CL::setKernelArg(kernel, index, size, data) = clSetKernelArg (CL.cpp:604-607), with the 14-slot
 * signature of the reference kernel (pathtracing.cl:207-234, PathTracer.cpp:88-125):
 *   0 float seed, 1 float pixelWeight, 2 float pxDim, 3 pbr_camera (80 B), 4 bvh, 5 facesV,
 *   6 facesN, 7 vertices, 8 normals, 9 materials, 10 lights, 11 imageIn, 12 imageOut,
 *   13 imageDebug -- slots 4..13 take a pbr_mem (size 8). */
int pbr_kernel_set_arg(pbr_ctx* ctx, pbr_kernel k, uint32_t index, size_t size, const void* data);
/* CL::execute(kernel) = clEnqueueNDRangeKernel over window.width x window.height (CL.cpp:289-306).
 * opencl.localgroupsize has no meaning here and is ignored. */
int pbr_kernel_launch(pbr_ctx* ctx, pbr_kernel k);
/* CL::finish() = clFlush + clFinish (CL.cpp:312-316). */
int pbr_finish(pbr_ctx* ctx);
/* CL::getKernelTimes()[kernel] (CL.cpp:480-506): device time of the last launch in ms. */
int pbr_kernel_time_ms(pbr_ctx* ctx, pbr_kernel k, double* ms);

/* ---- additive entry points --------------------------------------------------- */

/* Restrict launches to image rows [y0, y1) -- tile sharding across GPUs. Default: all. */
int pbr_set_tile(pbr_ctx* ctx, int32_t y0, int32_t y1);
/* Interleaved variant for load balance: this context renders the stripes of `stripe_rows` rows with
 * (row / stripe_rows) % world == rank; IMG_HEIGHT must be a multiple of stripe_rows * world.  stripe_rows <= 0
 * switches back to pbr_set_tile's contiguous rows. */
int pbr_set_tile_stripes(pbr_ctx* ctx, int32_t stripe_rows, int32_t world, int32_t rank);
/* Choose the device pipeline.  -1 (default) = by measurement: after one warm-up frame of a configuration (scene, frame
 * size, kernel variant, walk) two frames run as 0 and two as 1, alternately, all timed, and the faster of the two renders
 * the rest -- the megakernel wins on small scenes and small frames, the wavefront everywhere else.
 * 0 = wavefront, one traverse + one shade launch per bounce;
 * 1 = one-thread-per-pixel megakernel (the reference's launch structure; kept as an on-device cross-check).
 * Both write identical pixels.  (Round 1 also had 2 = persistent kernels and 3 = carry-over wavefront; both measured
 * slower on every scene and were removed: PBR_ERR_UNSUPPORTED.) */
int pbr_set_pipeline(pbr_ctx* ctx, int32_t mode);
/* The pipeline frames are rendered with right now: the mode set explicitly, or with mode -1 the measured choice
 * (0 or 1), -1 while the measurement is still running. */
int pbr_pipeline_in_use(pbr_ctx* ctx, int32_t* mode);
/* n_frames consecutive frames in one call.  Same pixels as the reference's frame loop
 *     for f in 0..n-1: setKernelArg(0, seeds[f]); setKernelArg(1, pixel_weights[f]); execute();
 *                      imageIn <- imageOut                     (PathTracer::generateImage, PathTracer.cpp:59-71)
 * with the camera and every other argument as currently set; the result is in imageOut (slot 12), imageIn
 * (slot 11) is only read.  One call instead of 4 n, no host round trip between frames.  Without depth of field a
 * pixel's frames depend only on that pixel: the frames accumulate in place in imageOut, one after the other.  With a
 * focus point set (camera.focusPoint >= 0) a frame reads another pixel of the previous one, so the frames
 * ping-pong between imageOut and a scratch image. */
int pbr_kernel_launch_batch(pbr_ctx* ctx, pbr_kernel k, int32_t n_frames, const float* seeds, const float* pixel_weights);
/* Scheduling knobs (never change a pixel): "node_phase_min", "refill_min" (reference-order traversal engine),
 * "wide_node_phase_min", "wide_refill_min" (the same two thresholds for the ordered walk),
 * "traverse_blocks" / "wide_blocks" (cap on resident blocks per SM of the reference-order / ordered traversal kernels,
 * 0 = all that fit), "wide_top" (how many nodes of the top of the 4-wide BVH the ordered walk stages in shared memory;
 * the tree is renumbered at the next launch), "shadow_stage" (1: shadow rays are a wavefront stage of their own, walked
 * by the traversal engine; 0: inside the shade kernel), "frames_in_flight" (1..8, default 4: how many consecutive frames of a
 * pbr_kernel_launch_batch are traced concurrently, each on a stream and in a wave state of its own; finished pixels leave
 * their frame's radiance in a per-frame buffer and are mixed into imageOut in frame order -- same bits; 1 = one frame
 * after the other).
 * The environment variables PBR_NODE_PHASE_MIN, PBR_REFILL_MIN, PBR_PIPELINE, PBR_TRAVERSAL, PBR_FRAMES_IN_FLIGHT set the
 * initial values.
 * Stands where opencl.localgroupsize stands in the reference's config.json. */
int pbr_set_tuning(pbr_ctx* ctx, const char* key, int32_t value);
/* Skip the imageDebug write.  The debug image is the only place where the reference's visit counters
 * (writeDebugImage, pathtracing.cl:73-78) can be observed; with it switched off the automatic traversal choice
 * (pbr_set_traversal) is free to take the ordered walk. */
int pbr_set_debug_image(pbr_ctx* ctx, int32_t enabled);
/* Which walk finds the hits.
 *   0  the reference's: stackless, fixed pre-order over the uploaded node array (traverse / traverseShadows,
 *      pt_bvh.cl:82-177) -- hit, t, visit counters and debug image all bit-exact;
 *   1  the ordered walk: a 4-wide BVH collapsed from the same array, nearest child first, four lanes per ray
 *      (csrc/pt_wide.cuh) -- the same hit face, leaf and t bits, hence the same image bits; the visit counters
 *      (pbr_stats [2] [3] [5], pbr_hit.visits, the debug image) then count wide nodes / the ordered walk's tests.
 *      PBR_ERR_UNSUPPORTED at launch when the node array is not one the ordered walk can honour (PHONGTESS, links that
 *      do not nest, a child box outside its parent's, ...: pbr_traversal_info().why_not);
 *  -1  (default) automatic: 1 for frames whose debug image is switched off, 0 otherwise -- and 0 for explicit rays,
 *      whose pbr_hit carries the counters.  Explicit any-hit rays always take 0 (WHICH face ends traverseShadows
 *      depends on its visiting order; inside a frame only "occluded or not" is consumed, which does not). */
int pbr_set_traversal(pbr_ctx* ctx, int32_t mode);
typedef struct {
	int32_t mode;                 /* as set */
	int32_t last_used;            /* 0 / 1: what the last launch walked with */
	int32_t wide_available;       /* the 4-wide BVH exists for the bound scene */
	int32_t wide_nodes, wide_top, wide_depth;
	double wide_build_ms;         /* read-back + host collapse + upload */
	uint64_t ordered_rays;        /* rays walked by the ordered walk since the last reset */
	uint64_t rewalked_rays;       /* of those: ambiguous (or stack overflow), walked again in reference order */
	char why_not[96];             /* when wide_available == 0 after a build attempt */
} pbr_traversal_info_t;
int pbr_traversal_info(pbr_ctx* ctx, pbr_traversal_info_t* out, int32_t reset);
/* Counters accumulated since the last call with reset != 0:
 *   [0] traverse() calls  [1] traverseShadows() calls  [2] BVH nodes visited by traverse()
 *   [3] triangle tests    [4] shaded hits              [5] BVH nodes visited by traverseShadows()
 * These are the inputs of the algorithmic-byte formula of SURVEY.md 8d. */
int pbr_stats(pbr_ctx* ctx, uint64_t out[6], int32_t reset);

/* ---- multi-GPU: one process (one pbr_ctx) per GPU, the scene replicated, ONE NCCL collective per frame on the float
 * accumulation buffer.  Replaces the reference's single-device assumption (CL.cpp:355, 470, 521: first
 * platform, first GPU, one queue).  libnccl.so.2 is loaded by pbr_comm_init; a single GPU needs no NCCL at all.
 *   rank 0:      pbr_comm_unique_id(id)  -> hand the 128 bytes to the other ranks (file, pipe, MPI, torch.distributed ...)
 *   every rank:  pbr_comm_init(ctx, id, rank, world)
 *   per frame:   pbr_kernel_launch(...); pbr_frame_combine(ctx, imageOut, mode, display)
 * pbr_frame_combine returns at once: the collective runs on the communicator's own stream behind the frame just launched,
 * while the next frame is traced.  The library orders what has to be ordered: a later launch that overwrites an image a
 * combine still uses (PathTracer's ping-pong pair: frame k + 2 and the combine of frame k), and pbr_image_read /
 * pbr_image_read_begin of such an image, wait for it on the device.
 *   PBR_COMBINE_SPP   every rank rendered whole frames with its own seeds: out = sum over ranks of image / world (one
 *                     scale-and-copy pass + ncclAllReduce); `image` keeps accumulating untouched.  Not bit-identical to
 *                     one GPU rendering the same frames (summation order): tolerance, see tests.
 *   PBR_COMBINE_ROWS  every rank rendered its rows (pbr_set_tile with the block pbr_tile_rows gives it, or
 *                     pbr_set_tile_stripes): the other ranks' rows are gathered into `image` in place (ncclAllGather; grouped
 *                     ncclBroadcast for unequal blocks; pack + all-gather + unpack for stripes).  Bit-identical to one GPU. */
#define PBR_COMBINE_SPP 0
#define PBR_COMBINE_ROWS 1
int pbr_comm_unique_id(void* id128);
int pbr_comm_init(pbr_ctx* ctx, const void* id128, int32_t rank, int32_t world);
int pbr_comm_info(pbr_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version);
int pbr_comm_destroy(pbr_ctx* ctx);
int pbr_frame_combine(pbr_ctx* ctx, pbr_mem image, int32_t mode, pbr_mem out);
/* The same inside pbr_kernel_launch_batch: after every frame f of a batch, pbr_frame_combine(imageOut, mode,
 * outs[(first + f) % n_outs]) -- a ring of up to 8 display images for PBR_COMBINE_SPP, none needed for PBR_COMBINE_ROWS.
 * mode -1 switches it off.  (With several frames in flight the batch keeps tracing the next frames while a frame is mixed
 * and combined.) */
int pbr_set_batch_combine(pbr_ctx* ctx, int32_t mode, const pbr_mem* outs, int32_t n_outs, int32_t first);
/* Make the render stream wait (on the device) for every combine enqueued so far. */
int pbr_comm_fence(pbr_ctx* ctx);
/* Rows [y0, y1) of rank `rank` of `world`: contiguous blocks of multiples of 4 rows covering [0, height). */
int pbr_tile_rows(int32_t height, int32_t rank, int32_t world, int32_t* y0, int32_t* y1);

/* Issue all device work of this context on a caller-owned CUDA stream (a cudaStream_t passed as
 * void*, e.g. torch.cuda.current_stream().cuda_stream; 0 is the legacy default stream) so that the
 * caller's own copies, collectives and CUDA events are ordered with the kernels.
 * own_stream != 0 ignores the pointer and restores the context's own stream. */
int pbr_set_stream(pbr_ctx* ctx, void* cuda_stream, int32_t own_stream);
/* Per-kernel device timing with CUDA events recorded on the launching stream around every launch
 * (the OpenCL queue of the reference is created with CL_QUEUE_PROFILING_ENABLE, CL.cpp:538). */
typedef struct {
	uint64_t launches;            /* every kernel launched by this library since the last reset */
	uint64_t raygen_launches, traverse_launches, shade_launches, other_launches;
	double raygen_ms, traverse_ms, shade_ms, other_ms;   /* summed event times, 0 unless enabled */
} pbr_profile;
int pbr_profile_enable(pbr_ctx* ctx, int32_t enabled);
int pbr_profile_read(pbr_ctx* ctx, pbr_profile* out, int32_t reset);

/* Explicit rays (BASELINE config 5).  `rays`/`hits` are HOST arrays of n elements; the scene is
 * given by buffer handles in the reference layout.  any_hit = 0: traverse() (closest hit,
 * pt_bvh.cl:82-123); any_hit = 1: traverseShadows() (pt_bvh.cl:133-177).  lights may be 0. */
int pbr_trace(pbr_ctx* ctx, pbr_mem bvh, pbr_mem facesV, pbr_mem vertices, pbr_mem lights, int32_t num_lights,
              const pbr_ray* rays, int64_t n, int32_t any_hit, pbr_hit* hits);
/* Same, rays and hits already resident: device pointers (from pbr_mem_device_ptr). */
int pbr_trace_device(pbr_ctx* ctx, pbr_mem bvh, pbr_mem facesV, pbr_mem vertices, pbr_mem lights, int32_t num_lights,
                     pbr_mem rays, int64_t n, int32_t any_hit, pbr_mem hits);

/* Pinned-math probes: evaluate include/pbr_pinned_math.h ON THE DEVICE (op: 0 sin 1 cos 2 tan
 * 3 acos 4 atan 5 pow(x,y) 6 cbrt 7 rand-seed-step) for n host inputs -- used to prove the
 * device and host agree bit for bit. */
int pbr_pinned_math_eval(pbr_ctx* ctx, int32_t op, const float* x, const float* y, int64_t n, float* out);

#ifdef __cplusplus
}
#endif

#endif /* PBR_B200_H */
