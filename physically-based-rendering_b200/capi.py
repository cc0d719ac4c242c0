"""ctypes binding of libpbr_b200.so (include/pbr_b200.h).

This is the thinnest possible Python view of the C ABI: one method per entry point, numpy arrays in
and out.  It exists so that tests and bench.py can drive the library the way the reference's
`PathTracer` drives its `CL` object (source/PathTracer.cpp:88-125, 43-71).  There is no fallback:
if the shared library is missing, or no CUDA device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpbr_b200.so")

# Every symbol include/pbr_b200.h declares (checked by tests/test_capi_symbols.py).
SYMBOLS = [
    "pbr_create", "pbr_destroy", "pbr_last_error", "pbr_build_id", "pbr_device_info",
    "pbr_buffer_create", "pbr_buffer_create_empty", "pbr_buffer_update", "pbr_buffer_read",
    "pbr_image_create", "pbr_image_write", "pbr_image_read", "pbr_image_copy", "pbr_mem_device_ptr",
    "pbr_free_buffers", "pbr_host_alloc", "pbr_host_free",
    "pbr_set_define", "pbr_program_load", "pbr_kernel_get", "pbr_kernel_set_arg", "pbr_kernel_launch",
    "pbr_finish", "pbr_kernel_time_ms",
    "pbr_image_read_begin", "pbr_image_read_end", "pbr_set_tile", "pbr_set_tile_stripes", "pbr_set_pipeline", "pbr_pipeline_in_use", "pbr_set_tuning", "pbr_kernel_launch_batch", "pbr_set_debug_image", "pbr_stats",
    "pbr_set_traversal", "pbr_traversal_info",
    "pbr_comm_unique_id", "pbr_comm_init", "pbr_comm_info", "pbr_comm_destroy", "pbr_frame_combine", "pbr_set_batch_combine", "pbr_comm_fence", "pbr_tile_rows",
    "pbr_trace", "pbr_trace_device", "pbr_pinned_math_eval",
    "pbr_set_stream", "pbr_profile_enable", "pbr_profile_read",
]

DEFINES_DTYPE = np.dtype([
    ("accel_struct", "<i4"), ("brdf", "<i4"), ("img_width", "<i4"), ("img_height", "<i4"),
    ("shadow_rays", "<i4"), ("max_depth", "<i4"), ("max_added_depth", "<i4"), ("phongtess", "<i4"),
    ("samples", "<i4"), ("anti_aliasing", "<f4"), ("phongtess_alpha", "<f4"),
    ("bvh_num_nodes", "<i4"), ("num_lights", "<i4"), ("_pad", "<i4", (3,)),
    ("sky_light", "<f4", (4,)),
])
CAMERA_DTYPE = np.dtype([
    ("eye", "<f4", (4,)), ("w", "<f4", (4,)), ("u", "<f4", (4,)), ("v", "<f4", (4,)),
    ("focusPoint", "<i4", (2,)), ("lense", "<f4", (2,)),
])
HIT_DTYPE = np.dtype([("t", "<f4"), ("hitFace", "<i4"), ("leaf", "<i4"), ("visits", "<u4")])
PROFILE_DTYPE = np.dtype([("launches", "<u8"), ("raygen_launches", "<u8"), ("traverse_launches", "<u8"),
                          ("shade_launches", "<u8"), ("other_launches", "<u8"), ("raygen_ms", "<f8"),
                          ("traverse_ms", "<f8"), ("shade_ms", "<f8"), ("other_ms", "<f8")])

TRAVERSAL_INFO_DTYPE = np.dtype([("mode", "<i4"), ("last_used", "<i4"), ("wide_available", "<i4"), ("wide_nodes", "<i4"),
                                 ("wide_top", "<i4"), ("wide_depth", "<i4"), ("wide_build_ms", "<f8"), ("ordered_rays", "<u8"),
                                 ("rewalked_rays", "<u8"), ("why_not", "S96")])

_lib = None


class PbrError(RuntimeError):
    pass


def load_library():
    """dlopen libpbr_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PbrError(
            "libpbr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C physically-based-rendering_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, i64, sz = C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, C.c_size_t
    sig = {
        "pbr_create": [C.c_int, C.POINTER(vp)],
        "pbr_destroy": [vp],
        "pbr_device_info": [vp, C.c_char_p, sz, C.POINTER(C.c_int), C.POINTER(sz)],
        "pbr_buffer_create": [vp, vp, sz, C.POINTER(u64)],
        "pbr_buffer_create_empty": [vp, sz, C.POINTER(u64)],
        "pbr_buffer_update": [vp, u64, sz, vp],
        "pbr_buffer_read": [vp, u64, sz, vp],
        "pbr_image_create": [vp, sz, sz, vp, C.POINTER(u64)],
        "pbr_image_write": [vp, u64, sz, sz, vp],
        "pbr_image_read": [vp, u64, sz, sz, vp],
        "pbr_image_copy": [vp, u64, u64],
        "pbr_mem_device_ptr": [vp, u64, C.POINTER(vp), C.POINTER(sz)],
        "pbr_free_buffers": [vp],
        "pbr_host_alloc": [vp, sz, C.POINTER(vp)],
        "pbr_host_free": [vp, vp],
        "pbr_set_define": [vp, C.c_char_p, C.c_char_p],
        "pbr_program_load": [vp, vp],
        "pbr_kernel_get": [vp, C.c_char_p, C.POINTER(u64)],
        "pbr_kernel_set_arg": [vp, u64, C.c_uint32, sz, vp],
        "pbr_kernel_launch": [vp, u64],
        "pbr_finish": [vp],
        "pbr_kernel_time_ms": [vp, u64, C.POINTER(C.c_double)],
        "pbr_set_tile": [vp, i32, i32],
        "pbr_image_read_begin": [vp, u64, sz, sz, vp],
        "pbr_image_read_end": [vp],
        "pbr_set_tile_stripes": [vp, i32, i32, i32],
        "pbr_set_pipeline": [vp, i32],
        "pbr_pipeline_in_use": [vp, C.POINTER(i32)],
        "pbr_set_tuning": [vp, C.c_char_p, i32],
        "pbr_kernel_launch_batch": [vp, u64, i32, vp, vp],
        "pbr_set_debug_image": [vp, i32],
        "pbr_stats": [vp, vp, i32],
        "pbr_set_traversal": [vp, i32],
        "pbr_traversal_info": [vp, vp, i32],
        "pbr_comm_unique_id": [vp],
        "pbr_comm_init": [vp, vp, i32, i32],
        "pbr_comm_info": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "pbr_comm_destroy": [vp],
        "pbr_frame_combine": [vp, u64, i32, u64],
        "pbr_set_batch_combine": [vp, i32, vp, i32, i32],
        "pbr_comm_fence": [vp],
        "pbr_tile_rows": [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)],
        "pbr_trace": [vp, u64, u64, u64, u64, i32, vp, i64, i32, vp],
        "pbr_trace_device": [vp, u64, u64, u64, u64, i32, u64, i64, i32, u64],
        "pbr_pinned_math_eval": [vp, i32, vp, vp, i64, vp],
        "pbr_set_stream": [vp, vp, i32],
        "pbr_profile_enable": [vp, i32],
        "pbr_profile_read": [vp, vp, i32],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    lib.pbr_last_error.argtypes = [vp]
    lib.pbr_last_error.restype = C.c_char_p
    lib.pbr_build_id.argtypes = []
    lib.pbr_build_id.restype = C.c_char_p
    _lib = lib
    return lib


def build_id():
    """pbr_build_id(): which sources the loaded libpbr_b200.so was built from."""
    return load_library().pbr_build_id().decode()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Device:
    """One pbr_ctx.  Method names follow the reference's `CL` class where one exists."""

    def __init__(self, device=-1, _borrowed_ctx=None):
        self.lib = load_library()
        self.owned = _borrowed_ctx is None
        if not self.owned:
            self.ctx = _borrowed_ctx
            return
        self.ctx = C.c_void_p()
        rc = self.lib.pbr_create(device, C.byref(self.ctx))
        if rc != 0:
            raise PbrError("pbr_create failed with code %d: no usable CUDA device (there is no CPU fallback)" % rc)

    @classmethod
    def from_ctx(cls, ctx_pointer):
        """Non-owning view of a pbr_ctx created elsewhere (e.g. by the C++ CL shim)."""
        return cls(_borrowed_ctx=ctx_pointer)

    def close(self):
        if self.ctx and self.owned:
            self.lib.pbr_destroy(self.ctx)
        self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise PbrError("%s failed (%d): %s" % (what, rc, self.lib.pbr_last_error(self.ctx).decode()))

    # ---- CL::createBuffer / createImage2D* / read / update ---------------------------------
    def createBuffer(self, array):
        a = np.ascontiguousarray(array)
        h = C.c_uint64()
        self._ck(self.lib.pbr_buffer_create(self.ctx, _p(a), a.nbytes, C.byref(h)), "pbr_buffer_create")
        return h.value

    def createEmptyBuffer(self, nbytes):
        h = C.c_uint64()
        self._ck(self.lib.pbr_buffer_create_empty(self.ctx, nbytes, C.byref(h)), "pbr_buffer_create_empty")
        return h.value

    def updateBuffer(self, buf, array):
        a = np.ascontiguousarray(array)
        self._ck(self.lib.pbr_buffer_update(self.ctx, buf, a.nbytes, _p(a)), "pbr_buffer_update")

    def readBuffer(self, buf, nbytes, dtype=np.uint8):
        out = np.zeros(nbytes // np.dtype(dtype).itemsize, dtype)
        self._ck(self.lib.pbr_buffer_read(self.ctx, buf, out.nbytes, _p(out)), "pbr_buffer_read")
        return out

    def createImage2DReadOnly(self, width, height, data):
        a = np.ascontiguousarray(data, np.float32)
        assert a.size == width * height * 4
        h = C.c_uint64()
        self._ck(self.lib.pbr_image_create(self.ctx, width, height, _p(a), C.byref(h)), "pbr_image_create")
        return h.value

    def createImage2DWriteOnly(self, width, height):
        h = C.c_uint64()
        self._ck(self.lib.pbr_image_create(self.ctx, width, height, None, C.byref(h)), "pbr_image_create")
        return h.value

    def updateImageReadOnly(self, image, width, height, data):
        a = np.ascontiguousarray(data, np.float32)
        self._ck(self.lib.pbr_image_write(self.ctx, image, width, height, _p(a)), "pbr_image_write")

    def readImageOutput(self, image, width, height, out=None):
        if out is None:
            out = np.zeros((height, width, 4), np.float32)
        self._ck(self.lib.pbr_image_read(self.ctx, image, width, height, _p(out)), "pbr_image_read")
        return out

    def copyImage(self, dst, src):
        self._ck(self.lib.pbr_image_copy(self.ctx, dst, src), "pbr_image_copy")

    def devicePtr(self, mem):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.pbr_mem_device_ptr(self.ctx, mem, C.byref(p), C.byref(n)), "pbr_mem_device_ptr")
        return p.value, n.value

    def freeBuffers(self):
        self._ck(self.lib.pbr_free_buffers(self.ctx), "pbr_free_buffers")

    # ---- program / kernel --------------------------------------------------------------------
    def setReplacement(self, before, after):
        self._ck(self.lib.pbr_set_define(self.ctx, before.encode(), after.encode()), "pbr_set_define")

    def loadProgram(self, defines):
        d = np.ascontiguousarray(defines)
        assert d.dtype.itemsize == 80
        self._ck(self.lib.pbr_program_load(self.ctx, _p(d)), "pbr_program_load")

    def createKernel(self, name="pathTracing"):
        k = C.c_uint64()
        self._ck(self.lib.pbr_kernel_get(self.ctx, name.encode(), C.byref(k)), "pbr_kernel_get")
        return k.value

    def setKernelArg(self, kernel, index, value):
        """value: np.float32 scalar, a 1-element CAMERA_DTYPE array, or an int handle."""
        if isinstance(value, (int, np.integer)) and index >= 4:
            v = np.array([value], np.uint64)
        elif isinstance(value, np.ndarray):
            v = np.ascontiguousarray(value)
        else:
            v = np.array([value], np.float32)
        self._ck(self.lib.pbr_kernel_set_arg(self.ctx, kernel, index, v.nbytes, _p(v)), "pbr_kernel_set_arg")

    def execute(self, kernel):
        self._ck(self.lib.pbr_kernel_launch(self.ctx, kernel), "pbr_kernel_launch")

    def finish(self):
        self._ck(self.lib.pbr_finish(self.ctx), "pbr_finish")

    def kernelTimeMs(self, kernel):
        ms = C.c_double()
        self._ck(self.lib.pbr_kernel_time_ms(self.ctx, kernel, C.byref(ms)), "pbr_kernel_time_ms")
        return ms.value

    # ---- additive ------------------------------------------------------------------------------
    def setTile(self, y0, y1):
        self._ck(self.lib.pbr_set_tile(self.ctx, y0, y1), "pbr_set_tile")

    def setTileStripes(self, stripe_rows, world=1, rank=0):
        self._ck(self.lib.pbr_set_tile_stripes(self.ctx, stripe_rows, world, rank), "pbr_set_tile_stripes")

    def setPipeline(self, mode):
        self._ck(self.lib.pbr_set_pipeline(self.ctx, mode), "pbr_set_pipeline")

    def pipelineInUse(self):
        m = C.c_int32(-2)
        self._ck(self.lib.pbr_pipeline_in_use(self.ctx, C.byref(m)), "pbr_pipeline_in_use")
        return int(m.value)

    def executeBatch(self, kernel, seeds, pixel_weights):
        """pbr_kernel_launch_batch: len(seeds) consecutive frames, result in imageOut."""
        sd = np.ascontiguousarray(seeds, np.float32)
        pw = np.ascontiguousarray(pixel_weights, np.float32)
        assert sd.shape == pw.shape and sd.ndim == 1
        self._ck(self.lib.pbr_kernel_launch_batch(self.ctx, kernel, len(sd), sd.ctypes.data, pw.ctypes.data),
                 "pbr_kernel_launch_batch")

    def setTuning(self, key, value):
        self._ck(self.lib.pbr_set_tuning(self.ctx, key.encode(), int(value)), "pbr_set_tuning")

    def setDebugImage(self, enabled):
        self._ck(self.lib.pbr_set_debug_image(self.ctx, int(enabled)), "pbr_set_debug_image")

    def stats(self, reset=False):
        out = np.zeros(6, np.uint64)
        self._ck(self.lib.pbr_stats(self.ctx, _p(out), int(reset)), "pbr_stats")
        return out

    # ---- multi-GPU (one pbr_ctx per process and GPU) -------------------------------------------
    @staticmethod
    def commUniqueId():
        buf = (C.c_char * 128)()
        rc = load_library().pbr_comm_unique_id(buf)
        if rc != 0:
            raise PbrError("pbr_comm_unique_id failed (%d): libnccl.so.2 not loadable?" % rc)
        return bytes(buf)

    def commInit(self, nccl_id, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        self._ck(self.lib.pbr_comm_init(self.ctx, buf, rank, world), "pbr_comm_init")

    def commInfo(self):
        r, w, v = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.pbr_comm_info(self.ctx, C.byref(r), C.byref(w), C.byref(v)), "pbr_comm_info")
        return r.value, w.value, v.value

    def commDestroy(self):
        self._ck(self.lib.pbr_comm_destroy(self.ctx), "pbr_comm_destroy")

    def frameCombine(self, image, mode, out=0):
        """mode: 0 samples (out = mean over ranks of image), 1 rows (the other ranks' rows gathered into image)."""
        self._ck(self.lib.pbr_frame_combine(self.ctx, image, mode, out), "pbr_frame_combine")

    def setBatchCombine(self, mode, outs=(), first=0):
        a = np.asarray(list(outs), np.uint64)
        self._ck(self.lib.pbr_set_batch_combine(self.ctx, mode, _p(a) if len(a) else None, len(a), first), "pbr_set_batch_combine")

    def commFence(self):
        self._ck(self.lib.pbr_comm_fence(self.ctx), "pbr_comm_fence")

    @staticmethod
    def tileRows(height, rank, world):
        y0, y1 = C.c_int32(), C.c_int32()
        rc = load_library().pbr_tile_rows(height, rank, world, C.byref(y0), C.byref(y1))
        if rc != 0:
            raise PbrError("pbr_tile_rows: bad arguments")
        return y0.value, y1.value

    def setTraversal(self, mode):
        """-1 automatic (default), 0 the reference's visiting order, 1 the ordered walk over the 4-wide BVH."""
        self._ck(self.lib.pbr_set_traversal(self.ctx, int(mode)), "pbr_set_traversal")

    def traversalInfo(self, reset=False):
        out = np.zeros(1, TRAVERSAL_INFO_DTYPE)
        self._ck(self.lib.pbr_traversal_info(self.ctx, _p(out), int(reset)), "pbr_traversal_info")
        d = {k: out[k][0].item() for k in TRAVERSAL_INFO_DTYPE.names}
        d["why_not"] = d["why_not"].decode(errors="replace")
        return d

    def deviceInfo(self):
        name = C.create_string_buffer(256)
        sm, mem = C.c_int(), C.c_size_t()
        self._ck(self.lib.pbr_device_info(self.ctx, name, 256, C.byref(sm), C.byref(mem)), "pbr_device_info")
        return name.value.decode(), sm.value, mem.value

    def trace(self, bvh, facesV, vertices, rays, any_hit=False, lights=0, num_lights=0):
        r = np.ascontiguousarray(rays, np.float32)
        n = r.shape[0]
        out = np.zeros(n, HIT_DTYPE)
        self._ck(self.lib.pbr_trace(self.ctx, bvh, facesV, vertices, lights, num_lights, _p(r), n, int(any_hit), _p(out)),
                 "pbr_trace")
        return out

    def traceDevice(self, bvh, facesV, vertices, rays_mem, n, hits_mem, any_hit=False, lights=0, num_lights=0):
        self._ck(self.lib.pbr_trace_device(self.ctx, bvh, facesV, vertices, lights, num_lights, rays_mem, n,
                                           int(any_hit), hits_mem), "pbr_trace_device")

    def setStream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (0 = legacy default stream); None = the context's own."""
        own = cuda_stream is None
        self._ck(self.lib.pbr_set_stream(self.ctx, None if own else C.c_void_p(cuda_stream), int(own)), "pbr_set_stream")

    def profileEnable(self, enabled=True):
        self._ck(self.lib.pbr_profile_enable(self.ctx, int(enabled)), "pbr_profile_enable")

    def profileRead(self, reset=False):
        out = np.zeros(1, PROFILE_DTYPE)
        self._ck(self.lib.pbr_profile_read(self.ctx, _p(out), int(reset)), "pbr_profile_read")
        return {k: out[k][0].item() for k in PROFILE_DTYPE.names}

    def pinnedMath(self, op, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = None if y is None else np.ascontiguousarray(y, np.float32)
        out = np.zeros_like(x)
        self._ck(self.lib.pbr_pinned_math_eval(self.ctx, op, _p(x), _p(y), x.size, _p(out)), "pbr_pinned_math_eval")
        return out
