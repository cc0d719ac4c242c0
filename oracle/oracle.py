"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# numpy views of include/pbr_types.h
DEFINES_DTYPE = np.dtype([
    ("accel_struct", "<i4"), ("brdf", "<i4"), ("img_width", "<i4"), ("img_height", "<i4"),
    ("shadow_rays", "<i4"), ("max_depth", "<i4"), ("max_added_depth", "<i4"), ("phongtess", "<i4"),
    ("samples", "<i4"), ("anti_aliasing", "<f4"), ("phongtess_alpha", "<f4"),
    ("bvh_num_nodes", "<i4"), ("num_lights", "<i4"), ("_pad", "<i4", (3,)),
    ("sky_light", "<f4", (4,)),
], align=False)
assert DEFINES_DTYPE.itemsize == 80

CAMERA_DTYPE = np.dtype([
    ("eye", "<f4", (4,)), ("w", "<f4", (4,)), ("u", "<f4", (4,)), ("v", "<f4", (4,)),
    ("focusPoint", "<i4", (2,)), ("lense", "<f4", (2,)),
])
assert CAMERA_DTYPE.itemsize == 80

HIT_DTYPE = np.dtype([("t", "<f4"), ("hitFace", "<i4"), ("leaf", "<i4"), ("visits", "<u4")])


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pt_oracle.cpp", "bvh_oracle.cpp", "obj_oracle.cpp")]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("pbr_pinned_math.h", "pbr_types.h")]
    stale = force or not os.path.exists(so) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("sin", "cos", "tan", "acos", "atan", "cbrt"):
            f = getattr(_LIB, "oracle_pm_" + name)
            f.restype = C.c_float
            f.argtypes = [C.c_float]
        _LIB.oracle_pm_pow.restype = C.c_float
        _LIB.oracle_pm_pow.argtypes = [C.c_float, C.c_float]
        _LIB.oracle_rand.restype = C.c_float
        _LIB.oracle_rand.argtypes = [C.POINTER(C.c_float)]
        _LIB.oracle_obj_load.restype = C.c_void_p
        _LIB.oracle_obj_load.argtypes = [C.c_char_p, C.c_int32]
        _LIB.oracle_obj_get.restype = C.c_int64
        _LIB.oracle_obj_get.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _LIB.oracle_obj_name.restype = C.c_char_p
        _LIB.oracle_obj_name.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        _LIB.oracle_obj_free.argtypes = [C.c_void_p]
        _LIB.oracle_bvh_build.restype = C.c_void_p
        _LIB.oracle_bvh_build.argtypes = (
            [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int32] +
            [C.c_void_p, C.c_int64] * 3 + [C.c_uint32, C.c_uint32, C.c_int32, C.c_float, C.c_float])
        _LIB.oracle_bvh_info.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.oracle_bvh_get.argtypes = [C.c_void_p] * 4
        _LIB.oracle_bvh_free.argtypes = [C.c_void_p]
        _LIB.oracle_path_tracing.argtypes = (
            [C.c_void_p, C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 7 + [C.c_int32] +
            [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p])
        _LIB.oracle_trace.argtypes = [C.c_void_p] * 8 + [C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        _LIB.oracle_trace_bruteforce.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_OBJ_FIELDS = {
    "vertices": (0, np.float32), "normals": (1, np.float32), "facesV": (2, np.uint32),
    "facesVN": (3, np.uint32), "facesMtl": (4, np.int32), "objFaceCounts": (5, np.uint32),
    "objFacesV": (6, np.uint32), "objFacesVN": (7, np.uint32), "objNormalFaceCounts": (8, np.uint32),
    "facesVT": (12, np.uint32), "textures": (13, np.float32),
}


def load_obj(path, shadow_rays=0):
    """Reference-faithful OBJ/MTL/LIGHTS load -> dict of numpy arrays (see obj_oracle.cpp)."""
    L = lib()
    h = L.oracle_obj_load(os.fsencode(path), shadow_rays)
    try:
        out = {}
        for k, (what, dt) in _OBJ_FIELDS.items():
            n = L.oracle_obj_get(h, what, None)
            a = np.zeros(n, dtype=dt)
            if n:
                L.oracle_obj_get(h, what, _p(a))
            out[k] = a
        n = L.oracle_obj_get(h, 9, None)
        m = np.zeros((n, 24), np.float32)
        if n:
            L.oracle_obj_get(h, 9, _p(m))
        out["materials"] = m
        out["materialNames"] = [L.oracle_obj_name(h, 1, i).decode() for i in range(n)]
        n = L.oracle_obj_get(h, 10, None)
        li = np.zeros((n, 10), np.float32)
        if n:
            L.oracle_obj_get(h, 10, _p(li))
        out["lights"] = li
        out["lightNames"] = [L.oracle_obj_name(h, 2, i).decode() for i in range(n)]
        out["objectNames"] = [L.oracle_obj_name(h, 0, i).decode()
                              for i in range(len(out["objFaceCounts"]))]
        out["shadowRaysForcedOff"] = bool(L.oracle_obj_get(h, 11, None))
        return out
    finally:
        L.oracle_obj_free(h)


def build_bvh(scene, max_faces=2, sah_faces_limit=100000, skip_ahead=True, skip_ahead_compare=0.7,
              phong_tess=0.0):
    """Reference-faithful BVH build + flatten.  `scene` is a dict as returned by load_obj().
    Returns dict(nodes[n,8] f32, facesV[m,4] u32, facesN[m,4] u32, info)."""
    L = lib()
    v = np.ascontiguousarray(scene["vertices"], np.float32)
    nrm = np.ascontiguousarray(scene["normals"], np.float32)
    ofv = np.ascontiguousarray(scene["objFacesV"], np.uint32)
    ofn = np.ascontiguousarray(scene["objFacesVN"], np.uint32)
    ofc = np.ascontiguousarray(scene["objFaceCounts"], np.uint32)
    onc = np.ascontiguousarray(scene["objNormalFaceCounts"], np.uint32)
    f = np.ascontiguousarray(scene["facesV"], np.uint32)
    fvn = np.ascontiguousarray(scene["facesVN"], np.uint32)
    fm = np.ascontiguousarray(scene["facesMtl"], np.int32)
    h = L.oracle_bvh_build(_p(v), v.size, _p(nrm), nrm.size, _p(ofv), _p(ofn), _p(ofc), _p(onc), len(ofc),
                           _p(f), f.size, _p(fvn), fvn.size, _p(fm), fm.size,
                           max_faces, sah_faces_limit, int(skip_ahead), skip_ahead_compare, phong_tess)
    try:
        info = np.zeros(6, np.int64)
        L.oracle_bvh_info(h, _p(info))
        nodes = np.zeros((info[4], 8), np.float32)
        fv = np.zeros((info[5], 4), np.uint32)
        fn = np.zeros((info[5], 4), np.uint32)
        L.oracle_bvh_get(h, _p(nodes), _p(fv), _p(fn))
        return {"nodes": nodes, "facesV": fv, "facesN": fn,
                "info": dict(zip(("allNodes", "leaves", "depth", "skipped", "emitted", "faces"), info.tolist()))}
    finally:
        L.oracle_bvh_free(h)


def make_defines(**kw):
    d = np.zeros(1, DEFINES_DTYPE)
    d["brdf"] = 1
    d["samples"] = 1
    d["max_depth"] = 3
    d["max_added_depth"] = 5
    d["anti_aliasing"] = 0.7
    d["sky_light"] = (1.0, 1.0, 1.0, 0.0)
    for k, val in kw.items():
        d[k] = val
    return d


def path_tracing(defines, seed, pixel_weight, px_dim, camera, nodes, facesV, facesN, vertices4, normals4,
                 materials, lights, image_in, y0=0, y1=None, nthreads=1, debug=True):
    """One launch of the kernel over rows [y0,y1).  Returns (imageOut, imageDebug, stats[6])."""
    L = lib()
    W, H = int(defines["img_width"][0]), int(defines["img_height"][0])
    y1 = H if y1 is None else y1
    out = np.zeros((H, W, 4), np.float32)
    dbg = np.zeros((H, W, 4), np.float32) if debug else None
    stats = np.zeros(6, np.uint64)
    mats = np.ascontiguousarray(materials, np.float32)
    nmat = mats.shape[0] if mats.ndim == 2 else 0
    if lights is None or len(lights) == 0:
        lights = np.zeros((1, 12), np.float32)
    lights = np.ascontiguousarray(lights, np.float32)
    image_in = np.ascontiguousarray(image_in, np.float32)
    if normals4 is None or len(normals4) == 0:
        normals4 = np.zeros((1, 4), np.float32)
    L.oracle_path_tracing(_p(defines), seed, pixel_weight, px_dim, _p(camera), _p(nodes), _p(facesV), _p(facesN),
                          _p(vertices4), _p(normals4), _p(mats), nmat, _p(lights), _p(image_in), _p(out), _p(dbg),
                          y0, y1, nthreads, _p(stats))
    return out, dbg, stats


def trace(defines, nodes, facesV, facesN, vertices4, normals4, lights, rays, any_hit=False, nthreads=1):
    """Explicit rays [n,8] f32 (origin.xyz_, dir.xyz, t0) -> structured HIT array + stats[6]."""
    L = lib()
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]
    out = np.zeros(n, HIT_DTYPE)
    stats = np.zeros(6, np.uint64)
    if lights is None or len(lights) == 0:
        lights = np.zeros((1, 12), np.float32)
    lights = np.ascontiguousarray(lights, np.float32)
    if normals4 is None or len(normals4) == 0:
        normals4 = np.zeros((1, 4), np.float32)
    L.oracle_trace(_p(defines), _p(nodes), _p(facesV), _p(facesN), _p(vertices4), _p(normals4), _p(lights),
                   _p(rays), n, int(any_hit), _p(out), nthreads, _p(stats))
    return out, stats


def trace_bruteforce(facesV, vertices4, rays):
    L = lib()
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros(rays.shape[0], HIT_DTYPE)
    L.oracle_trace_bruteforce(_p(facesV), facesV.shape[0], _p(vertices4), _p(rays), rays.shape[0], _p(out))
    return out
