"""GPU: the host mirror end to end -- Cfg -> ObjParser -> BVH -> PathTracer -> CL shim -> C ABI -> CUDA --
against the oracle, plus the behaviours PathTracer adds around the kernel (accumulation, reset on
camera change, checkpoint / resume, headless driver)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as Hh
from conftest import MODELS

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cfg():
    from pbr_b200 import host
    c = host.Config()
    c.reset()
    yield c
    c.reset()


def _oracle_for(oracle, cfg, scene_dict, frames, **kw):
    W, H = int(cfg.get("window.width")), int(cfg.get("window.height"))
    p = Hh.Prepared(scene_dict, W, H, brdf=int(cfg.get("render.brdf")), samples=int(cfg.get("render.samples")),
                    max_depth=int(cfg.get("render.max_depth")), max_added_depth=int(cfg.get("render.max_added_depth")),
                    shadow_rays=int(cfg.get("render.shadow_rays")), antialiasing=float(cfg.get("render.antialiasing")),
                    eye=tuple(float(cfg.get("camera.eye." + a)) for a in "xyz"),
                    center=tuple(float(cfg.get("camera.center." + a)) for a in "xyz"),
                    fov=float(cfg.get("camera.perspective.fov")), **kw)
    return p, p.oracle_frames(frames)


@pytest.mark.parametrize("brdf,shadow", [(1, 0), (0, 0), (1, 1)])
def test_pathtracer_matches_oracle_on_suzanne(oracle, cfg, brdf, shadow):
    """BASELINE config 1 (reduced frame size): bundled Cornell box + Suzanne through the whole host path."""
    from pbr_b200 import host
    cfg.update({"window.width": 160, "window.height": 120, "render.brdf": brdf, "render.shadow_rays": shadow,
                "render.max_depth": 4})
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    img = None
    for _ in range(3):
        img, dbg = r.generate_image(debug=True)
    scene = oracle.load_obj(os.path.join(MODELS, "suzanne.obj"), shadow)
    p, (want, wdbg, wstats) = _oracle_for(oracle, cfg, scene, 3)
    # host-side products: flattened BVH, camera, pixel size
    flat = r.flat()
    assert np.array_equal(flat["nodes"].view(np.uint32), p.nodes.view(np.uint32))
    assert np.array_equal(flat["facesV"], p.facesV)
    cam, px = r.camera()
    assert cam.tobytes() == p.camera.tobytes() and px == p.px_dim
    # the frame
    assert Hh.mean_relative_error(img, want) <= 0.02
    assert Hh.images_equal(img, want)
    assert Hh.images_equal(dbg, wdbg)
    assert np.array_equal(r.stats(reset=True), wstats)
    assert r.info()["sample_count"] == 3 and r.info()["lights"] == (1 if shadow else 0)
    r.close()


def test_pathtracer_phong_tessellation(oracle, cfg):
    from pbr_b200 import host
    cfg.update({"window.width": 96, "window.height": 72, "render.phong_tessellation": 0.6})
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    img = None
    for _ in range(2):
        img = r.generate_image()
    scene = oracle.load_obj(os.path.join(MODELS, "suzanne.obj"), 0)
    p, (want, _, wstats) = _oracle_for(oracle, cfg, scene, 2, phong_tessellation=0.6)
    assert np.array_equal(r.flat()["nodes"].view(np.uint32), p.nodes.view(np.uint32))
    assert Hh.images_equal(img, want)
    assert np.array_equal(r.stats(reset=True), wstats)
    r.close()


def test_resident_frames_equal_per_frame_readback(cfg):
    from pbr_b200 import host
    cfg.update({"window.width": 96, "window.height": 64})
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    for _ in range(6):
        a = r.generate_image()
    r.reset_sample_count()
    r.render_frames(6)
    b = r.read_image()
    assert Hh.images_equal(a, b)
    # resetSampleCount (what a camera change triggers, qt/GLWidget.cpp:80-84) restarts the accumulation
    r.reset_sample_count()
    c = r.generate_image()
    r2 = host.Renderer(0)
    r2.set_deterministic(True)
    r2.load_model(MODELS + "/", "suzanne.obj")
    assert Hh.images_equal(c, r2.generate_image())
    # moving the camera changes the picture and resets the count
    r2.set_eye(0.3, 1.0, 3.0)
    assert r2.info()["sample_count"] == 0
    moved = r2.generate_image()
    assert not Hh.images_equal(moved, c)
    r.close()
    r2.close()


def test_checkpoint_resume(cfg, tmp_path):
    from pbr_b200 import host
    cfg.update({"window.width": 64, "window.height": 64})
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    r.render_frames(4)
    half = r.read_image()
    host.write_checkpoint(str(tmp_path / "acc.bin"), half, r.info()["sample_count"])
    r.render_frames(4)
    full = r.read_image()
    img, sc = host.read_checkpoint(str(tmp_path / "acc.bin"), 64, 64)
    assert sc == 4 and Hh.images_equal(img, half)
    r2 = host.Renderer(0)
    r2.set_deterministic(True)
    r2.load_model(MODELS + "/", "suzanne.obj")
    r2.write_image(img, sc)
    r2.render_frames(4)
    assert Hh.images_equal(r2.read_image(), full)
    r.close()
    r2.close()


def test_synthetic_scene_and_explicit_rays(oracle, cfg):
    """Soup through Scene.from_arrays: product BVH builder + kernels vs oracle builder + oracle kernels."""
    import pbr_b200
    from pbr_b200 import host
    cfg.update({"window.width": 128, "window.height": 72, "camera.eye.x": 0.0, "camera.eye.y": 0.0, "camera.eye.z": 3.5})
    sc = pbr_b200.scenes.soup(30000, seed=21)
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_scene(sc)
    img = r.generate_image()
    p, (want, _, _) = _oracle_for(oracle, cfg, sc, 1)
    assert Hh.images_equal(img, want)
    rays = Hh.primary_rays(p, 300, 200)
    got = r.trace(rays)
    exp, _ = p.oracle_trace(rays)
    assert np.array_equal(got["hitFace"], exp["hitFace"]) and np.array_equal(got["leaf"], exp["leaf"])
    assert np.array_equal(got["t"].view(np.uint32), exp["t"].view(np.uint32))
    r.close()


def test_headless_driver(cfg, tmp_path):
    from pbr_b200 import host
    out = tmp_path / "img.pfm"
    ck = tmp_path / "acc.bin"
    cmd = [host.HEADLESS_PATH, "--model", MODELS + "/", "suzanne.obj", "--frames", "3", "--deterministic",
           "--set", "window.width=64", "--set", "window.height=48", "--out", str(out), "--checkpoint", str(ck)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert "Mrays/s" in res.stderr
    data = open(out, "rb").read()
    assert data.startswith(b"PF\n64 48\n-1.0\n") and len(data) == len(b"PF\n64 48\n-1.0\n") + 64 * 48 * 12
    img, sc = host.read_checkpoint(str(ck), 64, 48)
    assert sc == 3
    cfg.update({"window.width": 64, "window.height": 48})
    r = host.Renderer(0)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    r.render_frames(3)
    assert Hh.images_equal(r.read_image(), img)
    r.close()


# ---------------------------------------------------------------------------------------------------------
# Product versus the reference itself: libpbr_host.so + libpbr_b200.so on the GPU against the reference's own
# ModelLoader / BVH / PathTracer.cpp / Camera.cpp / kernel source running on the host
# (oracle/_ref/libref_host.so + oracle/_ref/pt_ref_*.so, built by __graft_entry__.build() where
# /root/reference exists).

REF_RENDER = {
    "sa": dict(brdf=1),
    "schlick_shadow_ms": dict(brdf=0, shadow_rays=1, samples=2),
    "phong_camera_fov": dict(brdf=1, phong_tess=0.7, eye=(0.3, 0.9, 2.5), center=(0.1, 0.2, 1.0), fov=60.0, antialiasing=0.4),
}
REF_W, REF_H = 88, 56


def _cam_fields_equal(a, b):
    return all(np.array_equal(a[f][0, :3], b[f][0, :3]) for f in ("eye", "w", "u", "v")) and \
        np.array_equal(a["focusPoint"], b["focusPoint"]) and np.array_equal(a["lense"], b["lense"])


@pytest.mark.parametrize("name", sorted(REF_RENDER))
def test_product_equals_reference_renderer(cfg, name):
    from oracle import oracle as O
    from oracle import ref_host as RH
    from pbr_b200 import host
    kw = dict(REF_RENDER[name], max_depth=4)
    path = os.path.join(MODELS, "suzanne.obj")
    if not RH.available():
        pytest.skip("oracle/_ref/libref_host.so not built")
    try:
        ref = RH.Renderer(path, width=REF_W, height=REF_H, nthreads=8, **kw)
    except FileNotFoundError:
        pytest.skip("reference kernel for this configuration not prebuilt")
    cfg.update({"window.width": REF_W, "window.height": REF_H, "render.brdf": kw["brdf"], "render.max_depth": 4,
                "render.shadow_rays": kw.get("shadow_rays", 0), "render.samples": kw.get("samples", 1),
                "render.phong_tessellation": kw.get("phong_tess", 0.0), "render.antialiasing": kw.get("antialiasing", 0.7),
                "camera.perspective.fov": kw.get("fov", 45.0)})
    for axis, e, c in zip("xyz", kw.get("eye", (0.0, 1.0, 3.0)), kw.get("center", (0.0, 0.0, 1.0))):
        cfg.update({"camera.eye." + axis: e, "camera.center." + axis: c})
    r = host.Renderer(0)
    r.set_frame_time_ms(33)                      # frame k is "rendered" 33 (k + 1) ms after start, on both sides
    r.load_model(MODELS + "/", "suzanne.obj")
    try:
        for k in range(3):
            got, gdbg = r.generate_image(debug=True)
            want, wdbg = ref.generate_image(33 * (k + 1))
            assert Hh.images_equal(got, want) and Hh.images_equal(gdbg, wdbg), "frame %d" % k
        cam, px = r.camera()
        assert px == ref.kernel_arg(2, np.float32)[0]
        assert _cam_fields_equal(cam, ref.kernel_arg(3).view(O.CAMERA_DTYPE))
        # Camera.cpp: rotate, move, focus -- every change resets the accumulation on both sides
        for what, args, ours in [(2, (37, -12), lambda: r.rotate_camera(37, -12)), (3, (), lambda: r.move_camera(0)),
                                 (5, (), lambda: r.move_camera(2)), (7, (), lambda: r.move_camera(4)),
                                 (0, (40, 25), lambda: r.set_focus(40, 25)), (4, (), lambda: r.move_camera(1)),
                                 (9, (), lambda: r.move_camera(6))]:
            ref.command(what, *args)
            ours()
            for k in range(2):
                got = r.generate_image()
                want, _ = ref.generate_image(33 * (k + 1))
                assert Hh.images_equal(got, want), "after camera command %d, frame %d" % (what, k)
            cam, _ = r.camera()
            assert _cam_fields_equal(cam, ref.kernel_arg(3).view(O.CAMERA_DTYPE)), "camera after command %d" % what
    finally:
        r.close()
        ref.close()


def test_render_ahead_returns_the_same_frames(cfg):
    """PathTracer::setRenderAhead: the next 1..3 frames are traced while this one is copied out.  Same frames, also
    across everything that invalidates the frame traced ahead (camera, focus, reset, renderFrames, readImage)."""
    from pbr_b200 import host
    cfg.update({"window.width": 120, "window.height": 72, "render.max_depth": 4})

    def session(ahead):
        r = host.Renderer(0)
        r.set_deterministic(True)
        r.load_model(MODELS + "/", "suzanne.obj")
        r.set_render_ahead(ahead)
        out = []
        for _ in range(4):
            out.append(r.generate_image().copy())
        assert r.info()["sample_count"] == 4
        r.rotate_camera(25, -8)                      # drops the frame traced ahead, restarts the accumulation
        for _ in range(2):
            out.append(r.generate_image().copy())
        r.set_focus(60, 30)
        out.append(r.generate_image().copy())
        img, dbg = r.generate_image(debug=True)      # debug image: traced inside the call
        out += [img.copy(), dbg.copy()]
        out.append(r.generate_image().copy())
        r.render_frames(3)                           # the frame traced ahead is the first of the three
        out.append(r.read_image().copy())
        assert r.info()["sample_count"] == 6
        out.append(r.generate_image().copy())
        out.append(r.read_image().copy())            # the frame returned last, not the one traced ahead
        out.append(r.generate_image().copy())
        r.close()
        return out

    plain = session(0)
    for depth in (1, 2, 3):
        ahead = session(depth)
        assert len(plain) == len(ahead)
        for i, (a, b) in enumerate(zip(plain, ahead)):
            assert Hh.images_equal(a, b), "image %d differs with render-ahead %d" % (i, depth)
