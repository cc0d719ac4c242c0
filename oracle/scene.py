"""Host-side half of the CPU oracle (TEST INFRASTRUCTURE): numpy restatement of what the
reference's PathTracer / Camera compute on the host before a launch.

    camera()         PathTracer::updateEyeBuffer (PathTracer.cpp:628-652) on top of
                     Camera::cameraReset / getAdjustedCenter (Camera.cpp:80-107)
    px_dim()         PathTracer::initKernelArgs (PathTracer.cpp:88-91), MathHelp::degToRad (MathHelp.cpp:9-11)
    pack_float4()    PathTracer::initOpenCLBuffers_Faces (PathTracer.cpp:357-380)
    pack_materials() PathTracer::initOpenCLBuffers_MaterialsRGB (PathTracer.cpp:435-519)
    pack_lights()    PathTracer::initOpenCLBuffers_Lights (PathTracer.cpp:387-428)
    defines()        CL::setValues (CL.cpp:626-705) + the three setReplacement() strings

All arithmetic is carried out in float32 step by step, like the C++ it restates.
"""
import math
import zlib
import struct

import numpy as np

from . import oracle as O

f32 = np.float32


def _normalize(v):
    # glm::normalize(v) = v * inversesqrt(dot(v, v)); inversesqrt(x) = 1 / sqrt(x)
    d = f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))
    s = f32(f32(1.0) / np.sqrt(d, dtype=f32))
    return np.array([v[0] * s, v[1] * s, v[2] * s], f32)


def _cross(x, y):
    # glm::cross
    return np.array([
        f32(x[1] * y[2]) - f32(y[1] * x[2]),
        f32(x[2] * y[0]) - f32(y[2] * x[0]),
        f32(x[0] * y[1]) - f32(y[0] * x[1]),
    ], f32)


def camera(eye=(0.0, 1.0, 3.0), center=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0),
           focus_point=(-1, -1), focal_length=0.035, aperture=1.8):
    """camera_cl for the config.json camera block.  `center` is the config value: the reference
    normalises it and then looks at (eye.x + c.x, eye.y - c.y, eye.z - c.z)."""
    eye = np.asarray(eye, f32)
    c = _normalize(np.asarray(center, f32))
    adj = np.array([eye[0] + c[0], eye[1] - c[1], eye[2] - c[2]], f32)
    up = np.asarray(up, f32)
    w = _normalize(adj - eye)
    u = _normalize(_cross(w, up))
    v = _normalize(_cross(u, w))
    cam = np.zeros(1, O.CAMERA_DTYPE)
    cam["eye"][0, :3] = eye
    cam["w"][0, :3] = w
    cam["u"][0, :3] = u
    cam["v"][0, :3] = v
    cam["focusPoint"][0] = focus_point
    cam["lense"][0] = (focal_length, aperture)
    return cam


def px_dim(width, height, fov_deg=45.0):
    aspect = f32(f32(width) / f32(height))
    rad = f32(float(f32(fov_deg)) * 3.14159265359 / float(f32(180.0)))
    f = f32(float(f32(aspect * f32(2.0))) * math.tan(float(f32(rad / f32(2.0)))))
    return f32(f / f32(width))


def pack_float4(flat):
    a = np.asarray(flat, f32).reshape(-1, 3)
    out = np.zeros((a.shape[0], 4), f32)
    out[:, :3] = a
    return out


def pack_materials(materials24, names, brdf):
    """materials24: [n,24] as returned by oracle.load_obj().  Returns (buffer [n,12|16] f32, sky_light)."""
    m = np.asarray(materials24, f32).reshape(-1, 24)
    n = m.shape[0]
    sky = np.array([1.0, 1.0, 1.0, 0.0], f32)
    if brdf == 0:
        buf = np.zeros((n, 12), f32)
        buf[:, 0] = m[:, 12]   # d
        buf[:, 1] = m[:, 13]   # Ni
        buf[:, 2] = m[:, 18]   # p
        buf[:, 3] = m[:, 17]   # rough
        buf[:, 4:8] = m[:, 4:8]    # Kd
        buf[:, 8:12] = m[:, 8:12]  # Ks
    else:
        buf = np.zeros((n, 16), f32)
        buf[:, 0] = m[:, 12]   # d
        buf[:, 1] = m[:, 13]   # Ni
        buf[:, 2] = m[:, 19]   # nu
        buf[:, 3] = m[:, 20]   # nv
        buf[:, 4] = m[:, 21]   # Rs
        buf[:, 5] = m[:, 22]   # Rd
        buf[:, 8:12] = m[:, 4:8]
        buf[:, 12:16] = m[:, 8:12]
    for i, name in enumerate(names):
        if name == "sky_light":
            # snprintf("(float4)( %f, %f, %f, 0.0f )") -> six decimals, re-read as float literals
            sky = np.array([float("%f" % m[i, 4]), float("%f" % m[i, 5]), float("%f" % m[i, 6]), 0.0], f32)
    return buf, sky


def pack_lights(lights10):
    li = np.asarray(lights10, f32).reshape(-1, 10)
    out = np.zeros((li.shape[0], 12), f32)
    out[:, 0:4] = li[:, 1:5]
    out[:, 4:8] = li[:, 5:9]
    out[:, 8] = li[:, 0]
    orb = li[:, 0] == 2
    out[orb, 9] = li[orb, 9]
    return out


def defines(width, height, bvh_num_nodes, num_lights, sky_light=(1.0, 1.0, 1.0, 0.0), brdf=1, samples=1,
            max_depth=3, max_added_depth=5, shadow_rays=0, antialiasing=0.7, phong_tessellation=0.0):
    return O.make_defines(
        accel_struct=0, brdf=brdf, img_width=width, img_height=height, shadow_rays=shadow_rays,
        max_depth=max_depth, max_added_depth=max_added_depth,
        phongtess=1 if phong_tessellation > 0.0 else 0, samples=samples,
        anti_aliasing=f32(float("%f" % f32(antialiasing))),
        phongtess_alpha=f32(float("%f" % f32(phong_tessellation))),
        bvh_num_nodes=bvh_num_nodes, num_lights=num_lights, sky_light=np.asarray(sky_light, f32))


def frame_seed(k):
    """Deterministic stand-in for the wall-clock seed: seed_k = 0.0333f * (k + 1)."""
    return f32(f32(0.0333) * f32(k + 1))


def pixel_weight(sample_count):
    # PathTracer.cpp:44
    return f32(f32(sample_count) / f32(sample_count + 1))


def write_png(path, rgb):
    """rgb: [H,W,3] float in [0,1], row 0 = bottom (the kernel's convention)."""
    img = (np.clip(np.nan_to_num(rgb[::-1]), 0.0, 1.0) ** (1 / 2.2) * 255.0 + 0.5).astype(np.uint8)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                 chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
