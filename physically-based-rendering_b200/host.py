"""ctypes binding of libpbr_host.so (include/pbr_host.h): the C++ mirror of the reference's host
classes -- Cfg, ObjParser/MtlParser/LightParser/ModelLoader, BVH, Camera, PathTracer, headless GLWidget.

    cfg = host.Config()                      # Cfg singleton
    cfg.set("window.width", 512)
    r = host.Renderer(device=0)              # GLWidget + PathTracer + Camera on one GPU
    r.load_model("tests/golden/models/", "suzanne.obj")     # GLWidget::loadModel
    img = r.generate_image()                 # PathTracer::generateImage

No CPU fallback: the renderer needs libpbr_b200.so and a CUDA device.  Scene parsing and the BVH build
(Scene, build_flat) are host-only and run anywhere.
"""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libpbr_host.so")
HEADLESS_PATH = os.path.join(_HERE, "host", "pbr_headless")

SYMBOLS = [
    "pbrh_last_error", "pbrh_config_reset", "pbrh_config_load_file", "pbrh_config_load_string", "pbrh_config_set",
    "pbrh_config_get", "pbrh_scene_load", "pbrh_scene_from_arrays", "pbrh_scene_free", "pbrh_scene_get",
    "pbrh_scene_name", "pbrh_flat_build", "pbrh_flat_info", "pbrh_flat_get", "pbrh_flat_free", "pbrh_set_device",
    "pbrh_renderer_create", "pbrh_renderer_destroy", "pbrh_renderer_load_scene", "pbrh_renderer_load_model",
    "pbrh_renderer_set_deterministic", "pbrh_renderer_set_seed_schedule", "pbrh_renderer_set_frame_time_ms", "pbrh_renderer_set_render_ahead", "pbrh_renderer_set_tile", "pbrh_renderer_set_tile_stripes", "pbrh_renderer_generate_image",
    "pbrh_renderer_render_frames", "pbrh_renderer_read_image", "pbrh_renderer_write_image", "pbrh_renderer_finish",
    "pbrh_renderer_reset_sample_count", "pbrh_renderer_set_focus", "pbrh_renderer_set_eye",
    "pbrh_renderer_rotate_camera", "pbrh_renderer_move_camera", "pbrh_renderer_info", "pbrh_renderer_stats", "pbrh_renderer_flat_get",
    "pbrh_renderer_camera", "pbrh_renderer_trace", "pbrh_renderer_handles",
    "pbrh_comm_unique_id", "pbrh_renderer_set_ranks", "pbrh_renderer_set_traversal", "pbrh_renderer_set_sharding",
    "pbrh_renderer_comm_fence",
    "pbrh_write_pfm", "pbrh_write_checkpoint", "pbrh_read_checkpoint",
]

_lib = None


class HostError(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HostError("libpbr_host.so is not built (%s): run __graft_entry__.build()" % LIB_PATH)
    capi.load_library()            # dependency, dlopen'ed first so that the rpath does not matter
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_void_p
    lib.pbrh_last_error.restype = C.c_char_p
    lib.pbrh_config_load_file.argtypes = [C.c_char_p]
    lib.pbrh_config_load_string.argtypes = [C.c_char_p]
    lib.pbrh_config_set.argtypes = [C.c_char_p, C.c_char_p]
    lib.pbrh_config_get.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    lib.pbrh_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    lib.pbrh_scene_from_arrays.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i32, vp, i32,
                                           C.c_char_p, vp, i32, C.POINTER(vp)]
    lib.pbrh_scene_free.argtypes = [vp]
    lib.pbrh_scene_free.restype = None
    lib.pbrh_scene_get.argtypes = [vp, i32, vp]
    lib.pbrh_scene_get.restype = i64
    lib.pbrh_scene_name.argtypes = [vp, i32, i32]
    lib.pbrh_scene_name.restype = C.c_char_p
    lib.pbrh_flat_build.argtypes = [vp, C.POINTER(vp)]
    lib.pbrh_flat_info.argtypes = [vp, vp, C.POINTER(C.c_double)]
    lib.pbrh_flat_info.restype = None
    lib.pbrh_flat_get.argtypes = [vp, vp, vp, vp]
    lib.pbrh_flat_get.restype = None
    lib.pbrh_flat_free.argtypes = [vp]
    lib.pbrh_flat_free.restype = None
    lib.pbrh_set_device.argtypes = [C.c_int]
    lib.pbrh_set_device.restype = None
    lib.pbrh_renderer_create.argtypes = [C.POINTER(vp)]
    lib.pbrh_renderer_destroy.argtypes = [vp]
    lib.pbrh_renderer_destroy.restype = None
    lib.pbrh_renderer_load_scene.argtypes = [vp, vp]
    lib.pbrh_renderer_load_model.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.pbrh_renderer_set_deterministic.argtypes = [vp, i32]
    lib.pbrh_renderer_set_seed_schedule.argtypes = [vp, u32, u32]
    lib.pbrh_renderer_set_frame_time_ms.argtypes = [vp, u32]
    lib.pbrh_renderer_set_render_ahead.argtypes = [vp, i32]
    lib.pbrh_renderer_set_tile_stripes.argtypes = [vp, i32, i32, i32]
    lib.pbrh_renderer_set_tile.argtypes = [vp, i32, i32]
    lib.pbrh_comm_unique_id.argtypes = [vp]
    lib.pbrh_renderer_set_ranks.argtypes = [vp, i32, i32, vp, i32]
    lib.pbrh_renderer_set_traversal.argtypes = [vp, i32]
    lib.pbrh_renderer_set_sharding.argtypes = [vp, i32]
    lib.pbrh_renderer_comm_fence.argtypes = [vp]
    lib.pbrh_renderer_generate_image.argtypes = [vp, f32p, f32p]
    lib.pbrh_renderer_render_frames.argtypes = [vp, i32]
    lib.pbrh_renderer_read_image.argtypes = [vp, f32p, f32p]
    lib.pbrh_renderer_write_image.argtypes = [vp, f32p, u32]
    lib.pbrh_renderer_finish.argtypes = [vp]
    lib.pbrh_renderer_reset_sample_count.argtypes = [vp]
    lib.pbrh_renderer_set_focus.argtypes = [vp, i32, i32]
    lib.pbrh_renderer_set_eye.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    lib.pbrh_renderer_rotate_camera.argtypes = [vp, i32, i32]
    lib.pbrh_renderer_move_camera.argtypes = [vp, i32]
    lib.pbrh_renderer_info.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pbrh_renderer_stats.argtypes = [vp, vp, i32]
    lib.pbrh_renderer_flat_get.argtypes = [vp, vp, vp, vp]
    lib.pbrh_renderer_camera.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.pbrh_renderer_trace.argtypes = [vp, vp, i64, i32, vp]
    lib.pbrh_renderer_handles.argtypes = [vp, C.POINTER(vp), vp]
    lib.pbrh_write_pfm.argtypes = [C.c_char_p, vp, i32, i32]
    lib.pbrh_write_checkpoint.argtypes = [C.c_char_p, vp, i32, i32, u32]
    lib.pbrh_read_checkpoint.argtypes = [C.c_char_p, vp, i32, i32, C.POINTER(u32)]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ck(rc, what):
    if rc != 0:
        raise HostError("%s: %s" % (what, load_library().pbrh_last_error().decode()))


class Config:
    """The Cfg singleton (reference: source/Cfg.h).  Keys are the reference's dotted config.json keys."""

    def __init__(self):
        self.lib = load_library()

    def reset(self):
        _ck(self.lib.pbrh_config_reset(), "Cfg reset")

    def load_file(self, path):
        _ck(self.lib.pbrh_config_load_file(os.fsencode(path)), "Cfg::loadConfigFile")

    def load_string(self, text):
        _ck(self.lib.pbrh_config_load_string(text.encode()), "Cfg load string")

    def set(self, key, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        elif isinstance(value, (float, np.floating)):
            value = repr(float(value))
        _ck(self.lib.pbrh_config_set(key.encode(), str(value).encode()), "Cfg::value(%s)" % key)

    def get(self, key):
        buf = C.create_string_buffer(1024)
        _ck(self.lib.pbrh_config_get(key.encode(), buf, 1024), "Cfg::value<string>(%s)" % key)
        return buf.value.decode()

    def update(self, mapping):
        for k, v in mapping.items():
            self.set(k, v)


_FIELDS = {
    "vertices": (0, np.float32), "normals": (1, np.float32), "facesV": (2, np.uint32),
    "facesVN": (3, np.uint32), "facesMtl": (4, np.int32), "objFaceCounts": (5, np.uint32),
    "objFacesV": (6, np.uint32), "objFacesVN": (7, np.uint32), "objNormalFaceCounts": (8, np.uint32),
    "facesVT": (12, np.uint32), "textures": (13, np.float32),
}


class Scene:
    """A ModelLoader holding a parsed scene (reference: source/ModelLoader.cpp, ObjParser.cpp)."""

    def __init__(self, handle):
        self.lib = load_library()
        self.h = handle

    @classmethod
    def load(cls, filepath, filename):
        """ModelLoader::loadModel(filepath, filename): `filepath` is the directory, with trailing slash."""
        lib = load_library()
        h = C.c_void_p()
        _ck(lib.pbrh_scene_load(os.fsencode(filepath), os.fsencode(filename), C.byref(h)), "ModelLoader::loadModel")
        return cls(h)

    @classmethod
    def from_arrays(cls, d):
        """`d`: dict as produced by pbr_b200.scenes.* (or the oracle's loader)."""
        lib = load_library()
        v = np.ascontiguousarray(d["vertices"], np.float32)
        n = np.ascontiguousarray(d["normals"], np.float32)
        fv = np.ascontiguousarray(d["facesV"], np.uint32)
        fvn = np.ascontiguousarray(d["facesVN"], np.uint32)
        fm = np.ascontiguousarray(d["facesMtl"], np.int32)
        oc = np.ascontiguousarray(d["objFaceCounts"], np.uint32)
        onc = np.ascontiguousarray(d["objNormalFaceCounts"], np.uint32)
        mats = np.ascontiguousarray(d["materials"], np.float32).reshape(-1, 24)
        names = "\n".join(d.get("materialNames", [""] * len(mats))).encode()
        lights = np.ascontiguousarray(d.get("lights", np.zeros((0, 10))), np.float32).reshape(-1, 10)
        h = C.c_void_p()
        _ck(lib.pbrh_scene_from_arrays(_p(v), v.size, _p(n), n.size, _p(fv), fv.size, _p(fvn), fvn.size,
                                       _p(fm), fm.size, _p(oc), _p(onc), len(oc), _p(mats), len(mats), names,
                                       _p(lights), len(lights), C.byref(h)), "pbrh_scene_from_arrays")
        return cls(h)

    def to_dict(self):
        out = {}
        for k, (what, dt) in _FIELDS.items():
            n = self.lib.pbrh_scene_get(self.h, what, None)
            a = np.zeros(n, dt)
            if n:
                self.lib.pbrh_scene_get(self.h, what, _p(a))
            out[k] = a
        n = self.lib.pbrh_scene_get(self.h, 9, None)
        m = np.zeros((n, 24), np.float32)
        if n:
            self.lib.pbrh_scene_get(self.h, 9, _p(m))
        out["materials"] = m
        out["materialNames"] = [self.lib.pbrh_scene_name(self.h, 1, i).decode() for i in range(n)]
        n = self.lib.pbrh_scene_get(self.h, 10, None)
        li = np.zeros((n, 10), np.float32)
        if n:
            self.lib.pbrh_scene_get(self.h, 10, _p(li))
        out["lights"] = li
        out["lightNames"] = [self.lib.pbrh_scene_name(self.h, 2, i).decode() for i in range(n)]
        out["objectNames"] = [self.lib.pbrh_scene_name(self.h, 0, i).decode() for i in range(len(out["objFaceCounts"]))]
        return out

    def build_flat(self):
        """BVH build + flatten on the host (BVH.cpp + PathTracer.cpp:238-347), no device needed.
        Returns dict(nodes [n,8] f32, facesV [m,4] u32, facesN [m,4] u32, info, build_seconds)."""
        f = C.c_void_p()
        _ck(self.lib.pbrh_flat_build(self.h, C.byref(f)), "BVH build")
        try:
            info = np.zeros(6, np.int64)
            sec = C.c_double()
            self.lib.pbrh_flat_info(f, _p(info), C.byref(sec))
            nodes = np.zeros((info[4], 8), np.float32)
            fv = np.zeros((info[5], 4), np.uint32)
            fn = np.zeros((info[5], 4), np.uint32)
            self.lib.pbrh_flat_get(f, _p(nodes), _p(fv), _p(fn))
        finally:
            self.lib.pbrh_flat_free(f)
        return {"nodes": nodes, "facesV": fv, "facesN": fn, "build_seconds": sec.value,
                "info": dict(zip(("allNodes", "leaves", "depth", "skipped", "emitted", "faces"), info.tolist()))}

    def release(self):
        h, self.h = self.h, None
        return h

    def close(self):
        if self.h:
            self.lib.pbrh_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """Headless GLWidget + PathTracer + Camera bound to one GPU (one per process in multi-GPU runs)."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.lib.pbrh_set_device(device)
        self.h = C.c_void_p()
        _ck(self.lib.pbrh_renderer_create(C.byref(self.h)), "renderer create")
        self.W = self.H = 0

    def close(self):
        if self.h:
            self.lib.pbrh_renderer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _after_load(self):
        info = self.info()
        self.W, self.H = info["width"], info["height"]

    def load_model(self, filepath, filename):
        _ck(self.lib.pbrh_renderer_load_model(self.h, os.fsencode(filepath), os.fsencode(filename)), "GLWidget::loadModel")
        self._after_load()

    def load_scene(self, scene):
        if isinstance(scene, dict):
            scene = Scene.from_arrays(scene)
        _ck(self.lib.pbrh_renderer_load_scene(self.h, scene.release()), "GLWidget::loadModel")
        self._after_load()

    def set_deterministic(self, enabled=True):
        _ck(self.lib.pbrh_renderer_set_deterministic(self.h, int(enabled)), "setDeterministicSeeds")

    def set_frame_time_ms(self, ms):
        """Simulated clock: frame k gets the seed of the reference's wall clock at (k + 1) * ms milliseconds."""
        _ck(self.lib.pbrh_renderer_set_frame_time_ms(self.h, int(ms)), "setFrameTimeMs")

    def set_render_ahead(self, depth=1):
        """generate_image() traces the next `depth` (0..3; True = 1) frames while this one is copied to the host
        (PathTracer::setRenderAhead)."""
        _ck(self.lib.pbrh_renderer_set_render_ahead(self.h, int(depth)), "setRenderAhead")

    def set_seed_schedule(self, stride, offset):
        _ck(self.lib.pbrh_renderer_set_seed_schedule(self.h, stride, offset), "setSeedSchedule")

    def set_tile_stripes(self, stripe_rows, world=1, rank=0):
        _ck(self.lib.pbrh_renderer_set_tile_stripes(self.h, stripe_rows, world, rank), "setTileStripes")

    def set_ranks(self, rank, world, nccl_id, sharding="spp"):
        """PathTracer::setRanks: this renderer is rank `rank` of `world` (one process per GPU); `nccl_id` are the 128 bytes
        of comm_unique_id() of rank 0.  sharding: "spp" (own seeds, delivered frame = mean over ranks), "rows", "stripes".
        Every frame then ends with one NCCL collective inside the library."""
        mode = {"spp": 0, "rows": 1, "stripes": 2}[sharding]
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id)) if world > 1 else None
        _ck(self.lib.pbrh_renderer_set_ranks(self.h, rank, world, buf, mode), "PathTracer::setRanks")

    def set_sharding(self, sharding):
        """Another sharding on the communicator set_ranks created: "spp", "rows", "stripes", or "none" (this rank renders
        whole frames on its own, no collective).  Restarts the accumulation."""
        _ck(self.lib.pbrh_renderer_set_sharding(self.h, {"spp": 0, "rows": 1, "stripes": 2, "none": 3}[sharding]), "setSharding")

    def comm_fence(self):
        """The render stream waits, on the device, for the collectives enqueued so far."""
        _ck(self.lib.pbrh_renderer_comm_fence(self.h), "commFence")

    def set_traversal(self, mode):
        """-1 automatic, 0 the reference's visiting order, 1 the ordered walk (pbr_set_traversal)."""
        _ck(self.lib.pbrh_renderer_set_traversal(self.h, int(mode)), "setTraversal")

    def set_tile(self, y0, y1):
        _ck(self.lib.pbrh_renderer_set_tile(self.h, y0, y1), "setTileRows")

    def generate_image(self, out=None, debug=False):
        """PathTracer::generateImage: one more frame; returns the accumulated image [H,W,4]
        (and the debug image when debug=True)."""
        if out is None:
            out = np.zeros((self.H, self.W, 4), np.float32)
        dbg = np.zeros((self.H, self.W, 4), np.float32) if debug is True else (debug if isinstance(debug, np.ndarray) else None)
        _ck(self.lib.pbrh_renderer_generate_image(self.h, _p(out), _p(dbg)), "PathTracer::generateImage")
        return (out, dbg) if dbg is not None else out

    def render_frames(self, n):
        _ck(self.lib.pbrh_renderer_render_frames(self.h, n), "PathTracer::renderFrames")

    def read_image(self, out=None, debug=None):
        if out is None:
            out = np.zeros((self.H, self.W, 4), np.float32)
        _ck(self.lib.pbrh_renderer_read_image(self.h, _p(out), _p(debug)), "PathTracer::readImage")
        return out

    def write_image(self, image, sample_count):
        a = np.ascontiguousarray(image, np.float32)
        _ck(self.lib.pbrh_renderer_write_image(self.h, _p(a), sample_count), "PathTracer::writeImage")

    def finish(self):
        _ck(self.lib.pbrh_renderer_finish(self.h), "CL::finish")

    def reset_sample_count(self):
        _ck(self.lib.pbrh_renderer_reset_sample_count(self.h), "resetSampleCount")

    def set_focus(self, x, y):
        _ck(self.lib.pbrh_renderer_set_focus(self.h, x, y), "setFocus")

    def set_eye(self, x, y, z):
        _ck(self.lib.pbrh_renderer_set_eye(self.h, x, y, z), "Camera::setEye")

    def rotate_camera(self, dx, dy):
        _ck(self.lib.pbrh_renderer_rotate_camera(self.h, dx, dy), "Camera::updateCameraRot")

    def move_camera(self, direction):
        """0 forward, 1 backward, 2 left, 3 right, 4 up, 5 down, 6 reset (Camera::cameraMove*)."""
        _ck(self.lib.pbrh_renderer_move_camera(self.h, int(direction)), "Camera::cameraMove")

    def info(self):
        a = np.zeros(8, np.int64)
        sec, ms = C.c_double(), C.c_double()
        _ck(self.lib.pbrh_renderer_info(self.h, _p(a), C.byref(sec), C.byref(ms)), "renderer info")
        keys = ("width", "height", "sample_count", "bvh_nodes", "emitted_nodes", "faces", "lights", "skipped")
        d = dict(zip(keys, a.tolist()))
        d["bvh_build_seconds"] = sec.value
        d["last_kernel_ms"] = ms.value
        return d

    def stats(self, reset=False):
        out = np.zeros(6, np.uint64)
        _ck(self.lib.pbrh_renderer_stats(self.h, _p(out), int(reset)), "stats")
        return out

    def flat(self):
        info = self.info()
        nodes = np.zeros((info["emitted_nodes"], 8), np.float32)
        fv = np.zeros((info["faces"], 4), np.uint32)
        fn = np.zeros((info["faces"], 4), np.uint32)
        _ck(self.lib.pbrh_renderer_flat_get(self.h, _p(nodes), _p(fv), _p(fn)), "flat get")
        return {"nodes": nodes, "facesV": fv, "facesN": fn}

    def camera(self):
        cam = np.zeros(1, capi.CAMERA_DTYPE)
        px = C.c_float()
        _ck(self.lib.pbrh_renderer_camera(self.h, _p(cam), C.byref(px)), "camera")
        return cam, np.float32(px.value)

    def trace(self, rays, any_hit=False):
        r = np.ascontiguousarray(rays, np.float32)
        out = np.zeros(r.shape[0], capi.HIT_DTYPE)
        _ck(self.lib.pbrh_renderer_trace(self.h, _p(r), r.shape[0], int(any_hit), _p(out)), "pbr_trace")
        return out

    def handles(self):
        """(pbr_ctx*, dict of pbr_mem handles) for driving libpbr_b200.so directly."""
        ctx = C.c_void_p()
        h = np.zeros(6, np.uint64)
        _ck(self.lib.pbrh_renderer_handles(self.h, C.byref(ctx), _p(h)), "handles")
        return ctx, dict(zip(("bvh", "facesV", "vertices", "lights", "image", "kernel"), (int(x) for x in h)))


    def device(self):
        """capi.Device view of this renderer's pbr_ctx (streams, profiling, raw buffers)."""
        ctx, _ = self.handles()
        return capi.Device.from_ctx(ctx)


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (pbr_comm_unique_id): rank 0 creates it, every rank passes it to
    Renderer.set_ranks."""
    buf = (C.c_char * 128)()
    _ck(load_library().pbrh_comm_unique_id(buf), "pbrh_comm_unique_id")
    return bytes(buf)


def write_pfm(path, image):
    a = np.ascontiguousarray(image, np.float32)
    _ck(load_library().pbrh_write_pfm(os.fsencode(path), _p(a), a.shape[1], a.shape[0]), "write_pfm")


def write_checkpoint(path, image, sample_count):
    a = np.ascontiguousarray(image, np.float32)
    _ck(load_library().pbrh_write_checkpoint(os.fsencode(path), _p(a), a.shape[1], a.shape[0], sample_count), "write_checkpoint")


def read_checkpoint(path, width, height):
    a = np.zeros((height, width, 4), np.float32)
    sc = C.c_uint32()
    _ck(load_library().pbrh_read_checkpoint(os.fsencode(path), _p(a), width, height, C.byref(sc)), "read_checkpoint")
    return a, sc.value
