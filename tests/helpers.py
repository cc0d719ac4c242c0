"""Shared plumbing for the parity tests: prepare one scene the way the reference's
PathTracer::initOpenCLBuffers does (PathTracer.cpp:136-230), run it through the oracle and through
the C ABI with identical inputs."""
import os

import numpy as np

from oracle import oracle as O
from oracle import scene as S

f32 = np.float32


class Prepared:
    """Flattened, packed scene + launch parameters (host side)."""

    def __init__(self, scene, width, height, brdf=1, samples=1, max_depth=3, max_added_depth=5,
                 shadow_rays=0, antialiasing=0.7, eye=(0.0, 1.0, 3.0), center=(0.0, 0.0, 1.0), fov=45.0,
                 focus_point=(-1, -1), bvh=None, bvh_kwargs=None, phong_tessellation=0.0):
        self.scene = scene
        self.W, self.H = width, height
        self.brdf = brdf
        kw = dict(bvh_kwargs or {})
        if phong_tessellation > 0.0:
            kw.setdefault("phong_tess", phong_tessellation)     # MathHelp::triCalcAABB grows the face boxes
        self.bvh = bvh if bvh is not None else O.build_bvh(scene, **kw)
        self.nodes = self.bvh["nodes"]
        self.facesV = self.bvh["facesV"]
        self.facesN = self.bvh["facesN"]
        self.vertices4 = S.pack_float4(scene["vertices"])
        self.normals4 = S.pack_float4(scene["normals"]) if len(scene["normals"]) else np.zeros((1, 4), f32)
        self.materials, sky = S.pack_materials(scene["materials"], scene["materialNames"], brdf)
        lights = scene["lights"] if shadow_rays > 0 else np.zeros((0, 10), f32)
        self.num_lights = len(lights)
        self.lights = S.pack_lights(lights) if len(lights) else np.zeros((1, 12), f32)
        if shadow_rays > 0 and self.num_lights == 0:
            shadow_rays = 0          # LightParser.cpp:119-121
        self.defines = S.defines(width, height, self.nodes.shape[0], self.num_lights, sky, brdf=brdf,
                                 samples=samples, max_depth=max_depth, max_added_depth=max_added_depth,
                                 shadow_rays=shadow_rays, antialiasing=antialiasing,
                                 phong_tessellation=phong_tessellation)
        self.camera = S.camera(eye=eye, center=center, focus_point=focus_point)
        self.px_dim = S.px_dim(width, height, fov)

    # ---- oracle ---------------------------------------------------------------------------
    def oracle_frames(self, nframes, nthreads=8, y0=0, y1=None, image=None, first=0):
        img = np.zeros((self.H, self.W, 4), f32) if image is None else image
        dbg = None
        total = np.zeros(6, np.uint64)
        for k in range(first, first + nframes):
            img, dbg, st = O.path_tracing(
                self.defines, S.frame_seed(k), S.pixel_weight(k), self.px_dim, self.camera, self.nodes,
                self.facesV, self.facesN, self.vertices4, self.normals4, self.materials, self.lights, img,
                y0=y0, y1=y1, nthreads=nthreads)
            total += st
        return img, dbg, total

    def oracle_trace(self, rays, any_hit=False, nthreads=8):
        return O.trace(self.defines, self.nodes, self.facesV, self.facesN, self.vertices4, self.normals4,
                       self.lights, rays, any_hit=any_hit, nthreads=nthreads)


class DeviceScene:
    """The same scene bound to a pbr_ctx exactly like PathTracer::initOpenCLBuffers + initKernelArgs."""

    def __init__(self, dev, prep):
        self.dev, self.prep = dev, prep
        p = prep
        self.bufBVH = dev.createBuffer(p.nodes)
        self.bufFacesV = dev.createBuffer(p.facesV)
        self.bufFacesN = dev.createBuffer(p.facesN)
        self.bufVertices = dev.createBuffer(p.vertices4)
        self.bufNormals = dev.createBuffer(p.normals4)
        mats = p.materials if len(p.materials) else np.zeros((1, 16), f32)
        self.bufMaterials = dev.createBuffer(mats) if len(p.materials) else dev.createEmptyBuffer(0)
        self.bufLights = dev.createBuffer(p.lights)
        dev.setReplacement("#BVH_NUM_NODES#", "%d" % p.nodes.shape[0])
        dev.setReplacement("#NUM_LIGHTS#", "%d" % p.num_lights)
        sky = p.defines["sky_light"][0]
        dev.setReplacement("#SKY_LIGHT#", "(float4)( %f, %f, %f, 0.0f )" % (sky[0], sky[1], sky[2]))
        zero = np.zeros((p.H, p.W, 4), f32)
        self.texIn = dev.createImage2DReadOnly(p.W, p.H, zero)
        self.texOut = dev.createImage2DWriteOnly(p.W, p.H)
        self.texDebug = dev.createImage2DWriteOnly(p.W, p.H)
        dev.loadProgram(p.defines)
        self.kernel = dev.createKernel("pathTracing")
        k = self.kernel
        dev.setKernelArg(k, 2, f32(p.px_dim))
        dev.setKernelArg(k, 3, p.camera)
        for i, h in enumerate([self.bufBVH, self.bufFacesV, self.bufFacesN, self.bufVertices, self.bufNormals,
                               self.bufMaterials, self.bufLights, self.texIn, self.texOut, self.texDebug]):
            dev.setKernelArg(k, 4 + i, h)

    def frames(self, nframes, image=None, first=0, host_roundtrip=True):
        """PathTracer::generateImage nframes times (PathTracer.cpp:59-71) with the deterministic
        seed schedule.  Returns (imageOut, imageDebug)."""
        dev, p, k = self.dev, self.prep, self.kernel
        img = np.zeros((p.H, p.W, 4), f32) if image is None else image
        dbg = None
        for n in range(first, first + nframes):
            if host_roundtrip or n == first:
                dev.updateImageReadOnly(self.texIn, p.W, p.H, img)
            else:
                dev.copyImage(self.texIn, self.texOut)
            dev.setKernelArg(k, 0, S.frame_seed(n))
            dev.setKernelArg(k, 1, S.pixel_weight(n))
            dev.setKernelArg(k, 3, p.camera)
            dev.execute(k)
            dev.finish()
            if host_roundtrip or n == first + nframes - 1:
                img = dev.readImageOutput(self.texOut, p.W, p.H)
                dbg = dev.readImageOutput(self.texDebug, p.W, p.H)
        return img, dbg

    def frames_batch(self, nframes, image=None, first=0):
        """The same frames through pbr_kernel_launch_batch (one call).  Returns (imageOut, imageDebug)."""
        dev, p, k = self.dev, self.prep, self.kernel
        img = np.zeros((p.H, p.W, 4), f32) if image is None else image
        dev.updateImageReadOnly(self.texIn, p.W, p.H, img)
        dev.setKernelArg(k, 3, p.camera)
        ks = range(first, first + nframes)
        dev.executeBatch(k, [S.frame_seed(n) for n in ks], [S.pixel_weight(n) for n in ks])
        dev.finish()
        return dev.readImageOutput(self.texOut, p.W, p.H), dev.readImageOutput(self.texDebug, p.W, p.H)

    def trace(self, rays, any_hit=False):
        return self.dev.trace(self.bufBVH, self.bufFacesV, self.bufVertices, rays, any_hit=any_hit,
                              lights=self.bufLights, num_lights=self.prep.num_lights)


def primary_rays(prep, width, height, t0=np.inf):
    """Pinhole rays through pixel centres (no jitter): initRay without antiAliasing
    (pathtracing.cl:30-38), evaluated in float32 in the kernel's operation order."""
    cam = prep.camera
    w, u, v = (cam[k][0, :3].astype(f32) for k in ("w", "u", "v"))
    px_dim = S.px_dim(width, height)
    xs, ys = np.meshgrid(np.arange(width, dtype=f32), np.arange(height, dtype=f32))
    xs, ys = xs.reshape(-1, 1), ys.reshape(-1, 1)
    W, H = f32(width), f32(height)
    inner = (u[None] - W * u[None] + (f32(2.0) * xs) * u[None] + v[None] - H * v[None] + (f32(2.0) * ys) * v[None]).astype(f32)
    d = (w[None] + (f32(px_dim) * f32(0.5)) * inner).astype(f32)
    s = (f32(1.0) / np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(f32))).astype(f32)
    d = (d * s[:, None]).astype(f32)
    rays = np.zeros((width * height, 8), f32)
    rays[:, 0:3] = cam["eye"][0, :3]
    rays[:, 4:7] = d
    rays[:, 7] = t0
    return rays


def random_rays(n, seed, lo=-1.5, hi=1.5):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3)).astype(f32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(f32)
    rays = np.zeros((n, 8), f32)
    rays[:, 0:3] = o
    rays[:, 4:7] = d.astype(f32)
    rays[:, 7] = np.inf
    return rays


def shadow_rays_from_hits(rays, hits, light_pos):
    """C5 shadow rays: from each primary hit point towards a point light, t = distance."""
    t = hits["t"]
    ok = np.isfinite(t)
    o = (rays[ok, 0:3] + rays[ok, 4:7] * t[ok, None]).astype(f32)
    to = (np.asarray(light_pos, f32)[None] - o).astype(f32)
    dist = np.sqrt((to * to).sum(1, dtype=f32)).astype(f32)
    out = np.zeros((o.shape[0], 8), f32)
    out[:, 0:3] = o
    out[:, 4:7] = to / dist[:, None]
    out[:, 7] = dist
    return out


def images_equal(a, b):
    """Bit-level equality (+0 and -0 differ) with NaN == NaN (the two sides may carry different NaN payloads)."""
    a, b = np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32)
    if a.shape != b.shape:
        return False
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def count_identical_pixels(a, b):
    """Number of pixels whose four channels are bit-identical (NaN == NaN)."""
    a, b = np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return int(same.reshape(-1, a.shape[-1]).all(axis=1).sum())


def mean_relative_error(a, b):
    """mean(|a-b| / (max(a,b) + 1e-3)) over RGB of pixels finite on both sides."""
    a, b = np.asarray(a, np.float64)[..., :3], np.asarray(b, np.float64)[..., :3]
    ok = np.isfinite(a) & np.isfinite(b)
    return float((np.abs(a - b)[ok] / (np.maximum(a, b)[ok] + 1e-3)).mean())


def model_path(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models", name)
