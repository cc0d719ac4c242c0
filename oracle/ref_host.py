"""ctypes front of oracle/_ref/libref_host.so -- the reference's own ObjParser / MtlParser / LightParser /
ModelLoader / BVH classes compiled for the tests by oracle/build_ref_host.py.  TEST INFRASTRUCTURE: the
yardstick for obj_oracle.cpp and bvh_oracle.cpp.  load_obj() and build_bvh() return what oracle.load_obj() and
oracle.build_bvh() return, so a test compares dict with dict."""
import ctypes as C
import json
import os
import tempfile

import numpy as np

from . import build_ref_host
from .oracle import _OBJ_FIELDS

_LIB = None


def available():
    return build_ref_host.reference_available() or os.path.isfile(build_ref_host.SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_ref_host.build())
        L.refhost_load.restype = C.c_void_p
        L.refhost_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.refhost_obj_get.restype = C.c_int64
        L.refhost_obj_get.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.refhost_obj_name.restype = C.c_char_p
        L.refhost_obj_name.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.refhost_bvh_info.argtypes = [C.c_void_p, C.c_void_p]
        L.refhost_bvh_get.argtypes = [C.c_void_p] * 4
        L.refhost_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _config_json(shadow_rays=0, max_faces=2, sah_faces_limit=100000, skip_ahead=True, skip_ahead_compare=0.7,
                 phong_tess=0.0):
    """The keys the compiled classes read (BVH.cpp:57,158,349,771; MathHelp.cpp:264,330; ObjParser.cpp:133;
    Logger.cpp), in the shape of the reference's config.json."""
    return {
        "accel_struct": 0,
        "bvh": {"max_faces": int(max_faces), "sah_faces_limit": int(sah_faces_limit),
                "skip_ahead": bool(skip_ahead), "skip_ahead_compare": float(skip_ahead_compare)},
        "logging": {"level": 0},
        "render": {"phong_tessellation": float(phong_tess), "shadow_rays": int(shadow_rays)},
    }


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load(path, build_bvh=True, **cfg):
    """Parse `path` (an .obj with its .mtl / .lights next to it) and, optionally, build + flatten the BVH with
    the reference's classes.  Returns (scene dict like oracle.load_obj, bvh dict like oracle.build_bvh or None)."""
    L = lib()
    d, f = os.path.split(os.path.abspath(path))
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as tmp:
        json.dump(_config_json(**cfg), tmp)
    try:
        h = L.refhost_load(os.fsencode(d + "/"), os.fsencode(f), os.fsencode(tmp.name), int(build_bvh))
    finally:
        os.unlink(tmp.name)
    try:
        out = {}
        for k, (what, dt) in _OBJ_FIELDS.items():
            n = L.refhost_obj_get(h, what, None)
            a = np.zeros(n, dtype=dt)
            if n:
                L.refhost_obj_get(h, what, _p(a))
            out[k] = a
        n = L.refhost_obj_get(h, 9, None)
        m = np.zeros((n, 24), np.float32)
        if n:
            L.refhost_obj_get(h, 9, _p(m))
        out["materials"] = m
        out["materialNames"] = [L.refhost_obj_name(h, 1, i).decode() for i in range(n)]
        n = L.refhost_obj_get(h, 10, None)
        li = np.zeros((n, 10), np.float32)
        if n:
            L.refhost_obj_get(h, 10, _p(li))
        out["lights"] = li
        out["lightNames"] = [L.refhost_obj_name(h, 2, i).decode() for i in range(n)]
        out["objectNames"] = [L.refhost_obj_name(h, 0, i).decode() for i in range(len(out["objFaceCounts"]))]
        bvh = None
        if build_bvh:
            info = np.zeros(6, np.int64)
            L.refhost_bvh_info(h, _p(info))
            nodes = np.zeros((info[4], 8), np.float32)
            fv = np.zeros((info[5], 4), np.uint32)
            fn = np.zeros((info[5], 4), np.uint32)
            L.refhost_bvh_get(h, _p(nodes), _p(fv), _p(fn))
            bvh = {"nodes": nodes, "facesV": fv, "facesN": fn,
                   "info": dict(zip(("allNodes", "leaves", "depth", "skipped", "emitted", "faces"), info.tolist()))}
        return out, bvh
    finally:
        L.refhost_free(h)


# ------------------------------------------------------------------------------------------------------
# The reference's renderer core: PathTracer.cpp + Camera.cpp on top of ref_shim/fake_cl.cpp

def full_config(width, height, brdf=1, samples=1, max_depth=3, max_added_depth=5, shadow_rays=0, antialiasing=0.7,
                phong_tess=0.0, eye=(0.0, 1.0, 3.0), center=(0.0, 0.0, 1.0), fov=45.0, focal_length=0.035,
                aperture=1.8, speed=0.2, **bvh):
    """A config.json with every key PathTracer.cpp / Camera.cpp / CL read, shaped like the reference's."""
    cfg = _config_json(shadow_rays=shadow_rays, phong_tess=phong_tess, **bvh)
    cfg["camera"] = {
        "eye": {"x": float(eye[0]), "y": float(eye[1]), "z": float(eye[2])},
        "center": {"x": float(center[0]), "y": float(center[1]), "z": float(center[2])},
        "perspective": {"fov": float(fov), "zfar": 1000.0, "znear": 0.1},
        "thin_lense": {"aperture": float(aperture), "focal_length": float(focal_length)},
        "speed": float(speed),
    }
    cfg["render"].update({"antialiasing": float(antialiasing), "brdf": int(brdf), "interval": 16.666,
                          "max_added_depth": int(max_added_depth), "max_depth": int(max_depth), "samples": int(samples)})
    cfg["window"] = {"width": int(width), "height": int(height)}
    cfg["opencl"] = {"build_options": "", "check_errors": True, "localgroupsize": 8, "program": "pathtracing.cl"}
    cfg["info"] = {"kernel_times": 0.0}
    return cfg


class Renderer:
    """GLWidget's constructor + loadModel + paintGL sequence with the reference's PathTracer and Camera
    (qt/GLWidget.cpp:28-32, 339-387, 504-517), the kernel being the reference kernel built for the host."""

    def __init__(self, path, nthreads=4, **cfg):
        from . import build_ref
        L = lib()
        L.refhost_renderer_create.restype = C.c_void_p
        L.refhost_renderer_create.argtypes = [C.c_char_p] * 3
        L.refhost_renderer_generate.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.refhost_renderer_command.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refhost_renderer_free.argtypes = [C.c_void_p]
        L.fakecl_program_values.restype = C.c_char_p
        L.fakecl_set_kernel_library.argtypes = [C.c_char_p, C.c_int]
        L.fakecl_kernel_arg.restype = C.c_longlong
        L.fakecl_kernel_arg.argtypes = [C.c_int, C.c_void_p]
        self.L = L
        self.config = full_config(**cfg)
        self.W, self.H = self.config["window"]["width"], self.config["window"]["height"]
        d, f = os.path.split(os.path.abspath(path))
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as tmp:
            json.dump(self.config, tmp)
        try:
            self.h = L.refhost_renderer_create(os.fsencode(tmp.name), os.fsencode(d + "/"), os.fsencode(f))
        finally:
            os.unlink(tmp.name)
        # the program text values the reference's CL::setValues would splice in -> the matching kernel library
        self.values = dict(line.split("=", 1) for line in L.fakecl_program_values().decode().splitlines() if line)
        so = build_ref.build(self.values)
        rc = L.fakecl_set_kernel_library(os.fsencode(so), nthreads)
        assert rc == 0, "cannot load %s" % so

    def generate_image(self, ms):
        """PathTracer::generateImage() `ms` milliseconds after start (seed = ms * 0.001f)."""
        img = np.zeros((self.H, self.W, 4), np.float32)
        dbg = np.zeros((self.H, self.W, 4), np.float32)
        n = self.L.refhost_renderer_generate(self.h, int(ms), _p(img), _p(dbg))
        assert n == img.size
        return img, dbg

    def kernel_arg(self, slot, dtype=np.uint8):
        n = self.L.fakecl_kernel_arg(slot, None)
        assert n >= 0
        a = np.zeros(n, np.uint8)
        if n:
            self.L.fakecl_kernel_arg(slot, _p(a))
        return a.view(dtype)

    def command(self, what, a=0, b=0):
        self.L.refhost_renderer_command(self.h, what, a, b)

    def close(self):
        if self.h:
            self.L.refhost_renderer_free(self.h)
            self.h = None
