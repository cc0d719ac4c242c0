"""CPU: every model the reference ships (resources/models/testing/*.obj, SURVEY.md 8c "soft pins") through
  * the reference's own host classes (oracle/_ref/libref_host.so),
  * the oracle's restatement, and
  * the product's host library,
same arrays bit for bit -- and then a small frame of each through the reference's own kernel and the restatement,
same bits again.  Only suzanne and pillars are committed as fixtures (tests/golden/models); the others are read
where they lie under /root/reference, so these cases exist in the build container only and skip elsewhere."""
import os

import numpy as np
import pytest

import helpers as Hh
from oracle import oracle as O
from oracle import ref as R
from oracle import ref_host as RH
from oracle import scene as S
from test_host_parity import _same_flat, _same_scene

REF_MODELS = "/root/reference/resources/models/testing"
# name -> (vertices, normals, faces, objects): counted from the files
BUNDLED = {
    "suzanne.obj": (603, 549, 1082, 10),
    "pillars.obj": (40, 6, 56, 5),
    "spheres.obj": (410, 134, 800, 4),
    "squirrel-mirror.obj": (520, 1009, 1020, 3),
    "squirrels.obj": (716, 712, 1408, 3),
    "applejack2.obj": (4140, 4090, 8180, 6),
    "applejack3.obj": (4076, 4074, 8068, 2),
}

pytestmark = pytest.mark.skipif(not (os.path.isdir(REF_MODELS) and RH.available()),
                                reason="the reference's bundled models / host classes are only in the build container")


@pytest.fixture()
def cfg():
    from pbr_b200 import host
    c = host.Config()
    c.reset()
    yield c
    c.reset()


@pytest.mark.parametrize("name", sorted(BUNDLED))
def test_bundled_model_three_way(cfg, name):
    from pbr_b200 import host
    path = os.path.join(REF_MODELS, name)
    ref_scene, ref_flat = RH.load(path, shadow_rays=1)
    nv, nn, nf, no = BUNDLED[name]
    assert (ref_scene["vertices"].size // 3, ref_scene["normals"].size // 3, ref_scene["facesV"].size // 3,
            len(ref_scene["objectNames"])) == (nv, nn, nf, no)
    ora_scene = O.load_obj(path, 1)
    _same_scene(ora_scene, ref_scene)
    _same_flat(O.build_bvh(ora_scene), ref_flat)
    cfg.set("render.shadow_rays", 1)
    prod = host.Scene.load(REF_MODELS + "/", name)
    _same_scene(prod.to_dict(), ref_scene)
    cfg.update({"bvh.max_faces": 2, "bvh.sah_faces_limit": 100000, "bvh.skip_ahead": True,
                "bvh.skip_ahead_compare": 0.7, "render.phong_tessellation": 0.0})
    _same_flat(prod.build_flat(), ref_flat)
    # every face sits in exactly one leaf slot
    assert ref_flat["info"]["faces"] == nf == ref_flat["facesV"].shape[0]


@pytest.mark.parametrize("name,brdf", [("spheres.obj", 1), ("squirrel-mirror.obj", 0), ("squirrels.obj", 1),
                                       ("applejack2.obj", 0), ("applejack3.obj", 1)])
def test_bundled_model_frames_equal_reference_kernel(name, brdf):
    """Two accumulated 64x40 frames: the reference's kernel source (built for these values on demand) against the
    restatement -- image and visit counters."""
    scene = O.load_obj(os.path.join(REF_MODELS, name), 0)
    p = Hh.Prepared(scene, 64, 40, brdf=brdf, max_depth=4, max_added_depth=3)
    if not R.available(p.defines):
        pytest.skip("reference kernel not buildable here")
    img_o = np.zeros((p.H, p.W, 4), np.float32)
    img_r = img_o.copy()
    for k in range(2):
        args = (p.defines, S.frame_seed(k), S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV, p.facesN,
                p.vertices4, p.normals4, p.materials, p.lights)
        img_o, dbg_o, _ = O.path_tracing(*args, img_o, nthreads=4)
        img_r, dbg_r = R.path_tracing(*args, img_r, nthreads=4)
        assert Hh.images_equal(img_o, img_r), "frame %d: radiance" % k
        assert Hh.images_equal(dbg_o, dbg_r), "frame %d: visit counters" % k
    assert dbg_r[..., 0].max() > 0, "no triangle was ever tested: the camera does not see the model"


def test_shipped_config_json_is_read_as_is(cfg):
    """The reference's own config.json (// comments, tabs, a path with spaces) through the product's Cfg: every
    key of Cfg.cpp:4-39 with the value the file states -- and those are also the product's built-in defaults."""
    shipped = {
        "camera.eye.x": 0.0, "camera.eye.y": 1.0, "camera.eye.z": 3.0,
        "camera.center.x": 0.0, "camera.center.y": 0.0, "camera.center.z": 1.0,
        "camera.perspective.fov": 45.0, "camera.perspective.zfar": 1000.0, "camera.perspective.znear": 0.1,
        "camera.thin_lense.aperture": 1.8, "camera.thin_lense.focal_length": 0.035, "camera.speed": 0.2,
        "info.kernel_times": 250.0, "accel_struct": 0, "bvh.max_faces": 2, "bvh.sah_faces_limit": 100000,
        "bvh.skip_ahead": "true", "bvh.skip_ahead_compare": 0.7, "logging.level": 4,
        "opencl.build_options": "", "opencl.check_errors": "true", "opencl.program": "source/opencl/pathtracing.cl",
        "opencl.localgroupsize": 8, "render.antialiasing": 0.7, "render.brdf": 1, "render.interval": 33.3,
        "render.max_added_depth": 5, "render.max_depth": 3, "render.phong_tessellation": 0.0, "render.samples": 1,
        "render.shadow_rays": 0, "shader.name": "pathtracing", "shader.path": "source/shader/",
        "window.height": 600, "window.width": 800,
        "import_path": "/home/seba/programming/Physically-based Rendering/resources/models/",
    }
    defaults = {k: cfg.get(k) for k in shipped if k not in ("logging.level", "import_path")}
    cfg.load_file("/root/reference/config.json")
    for key, want in shipped.items():
        got = cfg.get(key)
        if isinstance(want, str):
            assert got == want, key
        else:
            assert float(got) == float(want), key
    for key, was in defaults.items():
        got = cfg.get(key)
        assert got == was or float(got) == float(was), "built-in default of %s differs from the shipped file" % key
