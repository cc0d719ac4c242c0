"""ctypes front of oracle/_ref/pt_ref_<key>.so -- the reference's own kernel compiled for the host CPU by
oracle/build_ref.py.  TEST INFRASTRUCTURE: the yardstick for the restatement in pt_oracle.cpp.

    path_tracing(defines, seed, pixel_weight, px_dim, camera, nodes, facesV, ...)   same arguments as
    oracle.path_tracing, so a test can hand both the same arrays.
"""
import ctypes as C

import numpy as np

from . import build_ref

_LIBS = {}


def values_from_defines(defines):
    """pbr_defines (oracle.DEFINES_DTYPE) -> the placeholder texts CL::setValues would splice in."""
    d = defines
    return build_ref.program_values(
        int(d["img_width"][0]), int(d["img_height"][0]), int(d["bvh_num_nodes"][0]), int(d["num_lights"][0]),
        sky_light=tuple(float(c) for c in d["sky_light"][0][:3]), brdf=int(d["brdf"][0]), samples=int(d["samples"][0]),
        max_depth=int(d["max_depth"][0]), max_added_depth=int(d["max_added_depth"][0]),
        shadow_rays=int(d["shadow_rays"][0]), antialiasing=float(d["anti_aliasing"][0]),
        phong_tessellation=float(d["phongtess_alpha"][0]))


def available(defines):
    """Is the reference kernel for this configuration built, or buildable here?"""
    import os
    return build_ref.reference_available() or os.path.isfile(build_ref.library_path(values_from_defines(defines)))


def lib(defines):
    values = values_from_defines(defines)
    so = build_ref.build(values)
    if so not in _LIBS:
        L = C.CDLL(so)
        L.ref_defines.restype = C.c_char_p
        L.ref_path_tracing.restype = C.c_int
        vp, ll, f32, i32 = C.c_void_p, C.c_longlong, C.c_float, C.c_int
        L.ref_path_tracing.argtypes = [f32, f32, f32, vp, vp, ll, vp, vp, ll, vp, ll, vp, ll, vp, ll, vp, ll,
                                       vp, vp, vp, i32, i32, i32, i32, i32]
        L.ref_trace.restype = C.c_int
        L.ref_trace.argtypes = [vp, ll, vp, vp, ll, vp, ll, vp, ll, vp, ll, vp, ll, i32, vp, i32]
        brdf = int(defines["brdf"][0])
        assert L.ref_sizeof(0) == 80 and L.ref_sizeof(1) == 32 and L.ref_sizeof(3) == 48
        assert L.ref_sizeof(2) == (48 if brdf == 0 else 64)
        _LIBS[so] = L
    return _LIBS[so]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def path_tracing(defines, seed, pixel_weight, px_dim, camera, nodes, facesV, facesN, vertices4, normals4,
                 materials, lights, image_in, y0=0, y1=None, nthreads=1):
    """One launch of the reference kernel over rows [y0, y1).  Returns (imageOut, imageDebug)."""
    L = lib(defines)
    W, H = int(defines["img_width"][0]), int(defines["img_height"][0])
    y1 = H if y1 is None else y1
    f32 = np.float32
    out = np.zeros((H, W, 4), f32)
    dbg = np.zeros((H, W, 4), f32)
    nodes = np.ascontiguousarray(nodes, f32)
    facesV = np.ascontiguousarray(facesV, np.uint32)
    facesN = np.ascontiguousarray(facesN, np.uint32)
    vertices4 = np.ascontiguousarray(vertices4, f32)
    normals4 = np.ascontiguousarray(normals4 if normals4 is not None and len(normals4) else np.zeros((1, 4), f32), f32)
    mats = np.ascontiguousarray(materials, f32)
    nmat = mats.shape[0] if mats.ndim == 2 else 0
    lights = np.ascontiguousarray(lights if lights is not None and len(lights) else np.zeros((1, 12), f32), f32)
    image_in = np.ascontiguousarray(image_in, f32)
    cam = np.ascontiguousarray(camera)
    rc = L.ref_path_tracing(
        f32(seed), f32(pixel_weight), f32(px_dim), _p(cam), _p(nodes), nodes.shape[0], _p(facesV), _p(facesN),
        facesV.shape[0], _p(vertices4), vertices4.shape[0], _p(normals4), normals4.shape[0], _p(mats), nmat,
        _p(lights), lights.shape[0], _p(image_in), _p(out), _p(dbg), W, H, y0, y1, nthreads)
    assert rc == 0
    return out, dbg


def trace(defines, nodes, facesV, facesN, vertices4, normals4, lights, rays, any_hit=False, nthreads=1):
    """Explicit rays [n,8] through the reference's traverse() (any_hit: traverseShadows()).
    Returns (t f32[n], hitFace i32[n], node visits i32[n], face tests i32[n])."""
    L = lib(defines)
    f32 = np.float32
    nodes = np.ascontiguousarray(nodes, f32)
    facesV = np.ascontiguousarray(facesV, np.uint32)
    facesN = np.ascontiguousarray(facesN, np.uint32)
    vertices4 = np.ascontiguousarray(vertices4, f32)
    normals4 = np.ascontiguousarray(normals4 if normals4 is not None and len(normals4) else np.zeros((1, 4), f32), f32)
    lights = np.ascontiguousarray(lights if lights is not None and len(lights) else np.zeros((1, 12), f32), f32)
    rays = np.ascontiguousarray(rays, f32)
    n = rays.shape[0]
    hits = np.zeros((n, 4), np.int32)
    rc = L.ref_trace(_p(nodes), nodes.shape[0], _p(facesV), _p(facesN), facesV.shape[0], _p(vertices4),
                     vertices4.shape[0], _p(normals4), normals4.shape[0], _p(lights), lights.shape[0], _p(rays), n,
                     int(any_hit), _p(hits), nthreads)
    assert rc == 0
    return hits[:, 0].copy().view(f32), hits[:, 1].copy(), hits[:, 2].copy(), hits[:, 3].copy()
