/*
 * pbr_host.h -- flat C view of the host-side mirror (libpbr_host.so) for bindings and tests.
 *
 * libpbr_host.so holds the C++ classes that mirror the reference's host code above the device
 * boundary -- Cfg, ObjParser / MtlParser / LightParser / ModelLoader, BVH, Camera, PathTracer, the
 * `CL` shim and a headless GLWidget (physically-based-rendering_b200/host/).  A C++ caller uses those
 * classes directly, exactly as the reference's qt/GLWidget.cpp does; this header exposes the same
 * operations to C / ctypes.  All device work goes through libpbr_b200.so (include/pbr_b200.h).
 *
 * Return value: 0 = ok, non-zero = failure (message via pbrh_last_error()).
 */
#ifndef PBR_HOST_H
#define PBR_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "pbr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* pbrh_last_error(void);

/* ---- Cfg (source/Cfg.{h,cpp}) ---------------------------------------------------------------- */
int pbrh_config_reset(void);                                   /* the reference's shipped config.json */
int pbrh_config_load_file(const char* path);                   /* Cfg::loadConfigFile */
int pbrh_config_load_string(const char* json);
int pbrh_config_set(const char* key, const char* value);       /* Cfg::value(key, value) */
int pbrh_config_get(const char* key, char* out, size_t out_len);   /* Cfg::value<string>(key) */

/* ---- scene: ModelLoader / ObjParser (source/ModelLoader.cpp, ObjParser.cpp) -------------------- */
typedef struct pbrh_scene pbrh_scene;
/* ModelLoader::loadModel(filepath, filename) */
int pbrh_scene_load(const char* filepath, const char* filename, pbrh_scene** out);
/* A scene already in memory, in the shape ObjParser holds it.  objFaceCounts[i] faces (and
 * objNormalFaceCounts[i] normal-index triples) of the global lists belong to object i, in order.
 * materials: 24 floats each (Ka4 Kd4 Ks4 d Ni Ns illum light rough p nu nv Rs Rd pad), names in one
 * '\n'-separated string; lights: 10 floats each (type pos4 rgb4 radius). */
int pbrh_scene_from_arrays(
	const float* vertices, int64_t numVertexFloats, const float* normals, int64_t numNormalFloats,
	const uint32_t* facesV, int64_t numFaceIdx, const uint32_t* facesVN, int64_t numFaceVNIdx,
	const int32_t* facesMtl, int64_t numFaces,
	const uint32_t* objFaceCounts, const uint32_t* objNormalFaceCounts, int32_t numObjects,
	const float* materials24, int32_t numMaterials, const char* materialNames,
	const float* lights10, int32_t numLights, pbrh_scene** out);
void pbrh_scene_free(pbrh_scene* s);
/* what: 0 vertices(f32) 1 normals(f32) 2 facesV(u32) 3 facesVN(u32) 4 facesMtl(i32) 5 per-object face
 * counts(u32) 6 per-object facesV concatenated(u32) 7 per-object facesVN concatenated(u32) 8 per-object
 * normal-face counts(u32) 9 materials(f32 x24) 10 lights(f32 x10) 12 facesVT(u32) 13 texture coords(f32).
 * Returns the element count; copies when dst != NULL. */
int64_t pbrh_scene_get(pbrh_scene* s, int32_t what, void* dst);
/* kind: 0 object, 1 material, 2 light */
const char* pbrh_scene_name(pbrh_scene* s, int32_t kind, int32_t idx);

/* ---- BVH build + flatten on the host, no device needed (BVH.cpp, PathTracer.cpp:238-347) -------- */
typedef struct pbrh_flat pbrh_flat;
int pbrh_flat_build(pbrh_scene* s, pbrh_flat** out);
/* info: all nodes, leaves, depth, skipped left children, emitted nodes, faces; seconds: build time */
void pbrh_flat_info(pbrh_flat* f, int64_t info[6], double* build_seconds);
void pbrh_flat_get(pbrh_flat* f, pbr_bvh_node* nodes, pbr_uint4* facesV, pbr_uint4* facesN);
void pbrh_flat_free(pbrh_flat* f);

/* ---- renderer: headless GLWidget + PathTracer + Camera ----------------------------------------- */
typedef struct pbrh_renderer pbrh_renderer;
void pbrh_set_device(int device);                              /* CUDA ordinal for renderers created next */
int pbrh_renderer_create(pbrh_renderer** out);
void pbrh_renderer_destroy(pbrh_renderer* r);
/* GLWidget::loadModel: parse, build the BVH, upload, prepare the kernel.  Consumes `s`. */
int pbrh_renderer_load_scene(pbrh_renderer* r, pbrh_scene* s);
int pbrh_renderer_load_model(pbrh_renderer* r, const char* filepath, const char* filename);
int pbrh_renderer_set_deterministic(pbrh_renderer* r, int32_t enabled);
/* frame k uses seed 0.0333f * (k * stride + offset + 1): disjoint seeds for sample-sharded ranks */
int pbrh_renderer_set_seed_schedule(pbrh_renderer* r, uint32_t stride, uint32_t offset);
/* simulated clock: frame with global index g gets the seed the reference derives from its wall clock at
 * t = ms * (g + 1) milliseconds, (ms * (g + 1)) * 0.001f (PathTracer.cpp:78-82); 0 = off */
int pbrh_renderer_set_frame_time_ms(pbrh_renderer* r, uint32_t ms);
/* generate_image starts tracing the next `depth` (0..3) frames before it waits for the copy of this one
 * (PathTracer::setRenderAhead); 0 = off */
int pbrh_renderer_set_render_ahead(pbrh_renderer* r, int32_t depth);
int pbrh_renderer_set_tile(pbrh_renderer* r, int32_t y0, int32_t y1);
/* interleaved stripes of rows for load balance (pbr_set_tile_stripes); stripe_rows <= 0 = off */
int pbrh_renderer_set_tile_stripes(pbrh_renderer* r, int32_t stripe_rows, int32_t world, int32_t rank);
/* Multi-GPU, one process per GPU (PathTracer::setRanks; replaces the single-device assumption of CL.cpp:355,470,521).
 * Rank 0 calls pbrh_comm_unique_id and hands the 128 bytes to the others by any means; every rank then calls
 * pbrh_renderer_set_ranks after loading the scene.  sharding: 0 samples (own seeds per rank, delivered frame = mean over
 * ranks), 1 contiguous row blocks, 2 interleaved stripes of rows.  From then on every frame ends with ONE NCCL collective
 * inside the library, overlapped with the next frame. */
int pbrh_comm_unique_id(void* id128);
int pbrh_renderer_set_ranks(pbrh_renderer* r, int32_t rank, int32_t world, const void* id128, int32_t sharding);
/* another sharding on the same communicator (restarts the accumulation); 3 = none: this rank renders whole frames alone */
int pbrh_renderer_set_sharding(pbrh_renderer* r, int32_t sharding);
/* the render stream waits, on the device, for every collective enqueued so far (for device-side timing) */
int pbrh_renderer_comm_fence(pbrh_renderer* r);
/* -1 automatic, 0 the reference's visiting order, 1 the ordered walk over the 4-wide BVH (pbr_set_traversal) */
int pbrh_renderer_set_traversal(pbrh_renderer* r, int32_t mode);
/* PathTracer::generateImage: one frame, accumulated image into out[W*H*4]; debug may be NULL */
int pbrh_renderer_generate_image(pbrh_renderer* r, float* out, float* debug);
/* n frames resident on the device, nothing read back */
int pbrh_renderer_render_frames(pbrh_renderer* r, int32_t n);
int pbrh_renderer_read_image(pbrh_renderer* r, float* out, float* debug);
int pbrh_renderer_write_image(pbrh_renderer* r, const float* image, uint32_t sample_count);
int pbrh_renderer_finish(pbrh_renderer* r);
int pbrh_renderer_reset_sample_count(pbrh_renderer* r);
int pbrh_renderer_set_focus(pbrh_renderer* r, int32_t x, int32_t y);
int pbrh_renderer_set_eye(pbrh_renderer* r, float x, float y, float z);
int pbrh_renderer_rotate_camera(pbrh_renderer* r, int32_t move_x, int32_t move_y);
/* Camera::cameraMove{Forward,Backward,Left,Right,Up,Down} / cameraReset (Camera.cpp:24-88): direction 0..6 */
int pbrh_renderer_move_camera(pbrh_renderer* r, int32_t direction);
/* info: width, height, sample count, BVH nodes (all), emitted nodes, faces, lights, skipped */
int pbrh_renderer_info(pbrh_renderer* r, int64_t info[8], double* bvh_build_seconds, double* last_kernel_ms);
int pbrh_renderer_stats(pbrh_renderer* r, uint64_t out[6], int32_t reset);
int pbrh_renderer_flat_get(pbrh_renderer* r, pbr_bvh_node* nodes, pbr_uint4* facesV, pbr_uint4* facesN);
int pbrh_renderer_camera(pbrh_renderer* r, pbr_camera* cam, float* px_dim);
/* explicit rays against the loaded scene (pbr_trace); host arrays */
int pbrh_renderer_trace(pbrh_renderer* r, const pbr_ray* rays, int64_t n, int32_t any_hit, pbr_hit* hits);
/* raw handles for callers that drive libpbr_b200.so directly (multi-GPU collectives, resident rays) */
int pbrh_renderer_handles(pbrh_renderer* r, void** pbr_ctx_out, uint64_t handles[6]);   /* bvh facesV vertices lights image kernel */

/* ---- image files (headless driver) ---------------------------------------------- */
int pbrh_write_pfm(const char* path, const float* rgba, int32_t width, int32_t height);
int pbrh_write_checkpoint(const char* path, const float* rgba, int32_t width, int32_t height, uint32_t sample_count);
int pbrh_read_checkpoint(const char* path, float* rgba, int32_t width, int32_t height, uint32_t* sample_count);

#ifdef __cplusplus
}
#endif

#endif /* PBR_HOST_H */
