"""Randomised parity (seeded): the geometry of tests/material_scene.py with random materials, lights and cameras
-- the values the program text does not depend on, so one build of the reference kernel serves every trial.
CPU: restatement vs the reference kernel.  GPU: CUDA vs the reference kernel.  Bit for bit, NaN pixels included."""
import os
import tempfile

import numpy as np
import pytest

import helpers as Hh
import material_scene
from oracle import oracle as O
from oracle import ref as R
from oracle import scene as S

W, H = 64, 48
TRIALS = 8


def _random_mtl(rng, transparent):
    names = ["floor", "wall", "frosted", "brushed", "mirrorish", "glass"]
    out = []
    for n in names:
        d = 1.0
        if transparent and n in ("frosted", "glass"):
            d = float(rng.choice([0.0, 0.25, 0.6, 0.9]))
        nu, nv = (float(10 ** rng.uniform(0, 4.5)) for _ in range(2))
        if rng.random() < 0.3:
            nv = nu
        out.append("newmtl %s\nKd %.3f %.3f %.3f\nKs %.3f %.3f %.3f\nd %.2f\nNi %.3f\nrough %.3f\np %.3f\nnu %.3f\nnv %.3f\nRs %.3f\nRd %.3f\n" % (
            n, *rng.uniform(0.1, 1.0, 3), *rng.uniform(0.0, 1.0, 3), d, rng.uniform(1.0, 2.0),
            float(rng.choice([0.0, 1.0, rng.uniform(0.01, 0.99)])), float(rng.choice([1.0, rng.uniform(0.05, 0.95)])),
            nu, nv, rng.uniform(0.0, 1.0), rng.uniform(0.0, 1.0)))
    out.append("newmtl sky_light\nKd %.3f %.3f %.3f\n" % tuple(rng.uniform(0.3, 1.0, 3)))
    return "\n".join(out)


def _trial(seed, brdf):
    rng = np.random.default_rng(1000 * brdf + seed)
    transparent = brdf == 0 or seed % 3 == 0
    saved_m, saved_l = material_scene.MATERIALS, material_scene.LIGHTS
    try:
        material_scene.MATERIALS = _random_mtl(rng, transparent)
        material_scene.LIGHTS = "newlight l\ntype 2\npos %.3f %.3f %.3f\nrgb %.3f %.3f %.3f\nradius %.3f\n" % (
            *rng.uniform(-0.5, 0.8, 1), *rng.uniform(1.2, 1.9, 1), *rng.uniform(0.5, 1.8, 1), *rng.uniform(0.5, 1.0, 3), rng.uniform(0.05, 0.4))
        with tempfile.TemporaryDirectory() as d:
            scene = O.load_obj(material_scene.write(d), 1)
    finally:
        material_scene.MATERIALS, material_scene.LIGHTS = saved_m, saved_l
    eye = (float(rng.uniform(-1.2, 1.2)), float(rng.uniform(0.4, 1.6)), float(rng.uniform(2.2, 3.6)))
    center = (float(rng.uniform(-0.3, 0.3)), float(rng.uniform(0.0, 0.4)), 1.0)
    return Hh.Prepared(scene, W, H, brdf=brdf, shadow_rays=int(seed % 2), samples=1 + int(seed % 3 == 2), max_depth=4 + seed % 3,
                       max_added_depth=2 + seed % 4, antialiasing=float(rng.uniform(0.0, 1.0)), eye=eye, center=center,
                       fov=float(rng.uniform(35.0, 70.0)),
                       focus_point=(int(rng.integers(0, W)), int(rng.integers(0, H))) if seed % 5 == 4 else (-1, -1))


def _reference_frames(p, frames=2):
    img = np.zeros((p.H, p.W, 4), np.float32)
    for k in range(frames):
        img, dbg = R.path_tracing(p.defines, S.frame_seed(k), S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV, p.facesN,
                                  p.vertices4, p.normals4, p.materials, p.lights, img, nthreads=4)
    return img, dbg


@pytest.mark.parametrize("brdf", [1, 0])
@pytest.mark.parametrize("seed", range(TRIALS))
def test_random_materials_oracle_vs_reference_kernel(seed, brdf):
    p = _trial(seed, brdf)
    if not R.available(p.defines):
        pytest.skip("reference kernel for this configuration neither prebuilt nor buildable")
    want, wdbg = _reference_frames(p)
    got, gdbg, _ = p.oracle_frames(2, nthreads=4)
    assert Hh.images_equal(got, want) and Hh.images_equal(gdbg, wdbg)
    assert np.isfinite(want[..., :3]).mean() > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("brdf", [1, 0])
@pytest.mark.parametrize("seed", range(TRIALS))
def test_random_materials_device_vs_reference_kernel(device, seed, brdf):
    p = _trial(seed, brdf)
    if not R.available(p.defines):
        pytest.skip("oracle/_ref not built for this configuration")
    want, wdbg = _reference_frames(p)
    got, gdbg = Hh.DeviceScene(device, p).frames(2)
    assert Hh.images_equal(got, want) and Hh.images_equal(gdbg, wdbg)
    # the same frames without the debug image: the ordered walk (where the scene allows it: no depth of field needed,
    # it only changes how frames overlap), several frames in flight -- the same bits
    ds = Hh.DeviceScene(device, p)
    device.setDebugImage(False)
    device.setPipeline(0)
    try:
        fast, _ = ds.frames_batch(2)
        info = device.traversalInfo()
    finally:
        device.setPipeline(-1)
        device.setDebugImage(True)
    assert info["last_used"] == 1, info
    assert Hh.images_equal(fast, want)
