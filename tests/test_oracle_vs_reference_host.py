"""Pins the host side: the reference's OWN ObjParser / MtlParser / LightParser / ModelLoader / MathHelp / BVH
classes, compiled from /root/reference/source by oracle/build_ref_host.py (oracle/_ref/libref_host.so), against
  * the oracle's restatement (oracle/obj_oracle.cpp, oracle/bvh_oracle.cpp), and
  * the product's host library (libpbr_host.so: parsers, index-range BVH builder, flatten).
Same files, same config values -> identical arrays, bit for bit: every parsed array, the flattened bvhNode_cl[]
(boxes, face slots, miss links), the leaf-ordered facesV / facesN, node / leaf / depth / skip counts."""
import os

import numpy as np
import pytest

from conftest import MODELS
from oracle import oracle as O
from oracle import ref_host as RH
from test_host_parity import _same_flat, _same_scene

pytestmark = pytest.mark.skipif(not RH.available(), reason="oracle/_ref/libref_host.so neither prebuilt nor buildable")

BVH_CONFIGS = [dict(), dict(max_faces=1), dict(skip_ahead=False), dict(sah_faces_limit=50),
               dict(phong_tess=0.7), dict(skip_ahead_compare=0.3), dict(max_faces=1, skip_ahead_compare=0.95)]


@pytest.fixture()
def cfg():
    from pbr_b200 import host
    c = host.Config()
    c.reset()
    yield c
    c.reset()


def _product_flat(cfg, scene_loader, kw):
    cfg.update({"bvh.max_faces": kw.get("max_faces", 2), "bvh.sah_faces_limit": kw.get("sah_faces_limit", 100000),
                "bvh.skip_ahead": kw.get("skip_ahead", True), "bvh.skip_ahead_compare": kw.get("skip_ahead_compare", 0.7),
                "render.phong_tessellation": kw.get("phong_tess", 0.0)})
    return scene_loader().build_flat()


@pytest.mark.parametrize("name", ["suzanne.obj", "pillars.obj"])
@pytest.mark.parametrize("kw", BVH_CONFIGS, ids=lambda k: ",".join("%s=%s" % i for i in k.items()) or "default")
def test_bundled_models_parse_and_bvh(cfg, name, kw):
    from pbr_b200 import host
    path = os.path.join(MODELS, name)
    ref_scene, ref_flat = RH.load(path, shadow_rays=1, **kw)
    ora_scene = O.load_obj(path, 1)
    _same_scene(ora_scene, ref_scene)
    _same_flat(O.build_bvh(ora_scene, **kw), ref_flat)
    cfg.set("render.shadow_rays", 1)
    loader = lambda: host.Scene.load(MODELS + "/", name)          # noqa: E731
    _same_scene(loader().to_dict(), ref_scene)
    _same_flat(_product_flat(cfg, loader, kw), ref_flat)
    assert ref_flat["info"]["faces"] == {"suzanne.obj": 1082, "pillars.obj": 56}[name]   # pathtracing.cl:75


def test_parser_quirks(cfg):
    """tests/golden/models/quirks.obj: empty tokens, v/vt read as v//vn, faces before any `o` / `usemtl`, ragged
    face lists ...  The reference's BVH constructor does not survive this file (it indexes past its arrays),
    so only the parsers are compared."""
    from pbr_b200 import host
    path = os.path.join(MODELS, "quirks.obj")
    for shadow_rays in (0, 1):
        ref_scene, _ = RH.load(path, build_bvh=False, shadow_rays=shadow_rays)
        _same_scene(O.load_obj(path, shadow_rays), ref_scene)
        cfg.set("render.shadow_rays", shadow_rays)
        _same_scene(host.Scene.load(MODELS + "/", "quirks.obj").to_dict(), ref_scene)
    assert ref_scene["facesV"].size == 25 and ref_scene["materialNames"] == ["red", "glass", "later", "sky_light"]


def _with_face_normals(scene):
    """The reference's BVH constructor indexes the per-object normal faces unconditionally (BVH.cpp:227-236): a
    file without `vn` crashes it.  Give every face one geometric normal so that `f v//vn` is written."""
    if scene["facesVN"].size == scene["facesV"].size:
        return scene
    v = scene["vertices"].reshape(-1, 3)
    f = scene["facesV"].reshape(-1, 3)
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]).astype(np.float32)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20).astype(np.float32)
    out = dict(scene)
    out["normals"] = n.astype(np.float32).ravel()
    out["facesVN"] = np.repeat(np.arange(len(f), dtype=np.uint32), 3)
    return out


@pytest.mark.parametrize("gen,kw", [
    ("soup", dict()), ("soup", dict(sah_faces_limit=700)), ("soup", dict(max_faces=1, skip_ahead=False)),
    ("grid", dict()), ("grid", dict(sah_faces_limit=300)), ("interior", dict()), ("interior", dict(phong_tess=0.5)),
], ids=lambda v: v if isinstance(v, str) else ",".join("%s=%s" % i for i in v.items()) or "default")
def test_generated_scenes_parse_and_bvh(cfg, tmp_path, gen, kw):
    """Scenes written to .obj/.mtl and read back by all three: SAH sweep and mean-split paths, 16 objects with
    per-object trees and exactly tied centres, several materials, v//vn faces."""
    import pbr_b200
    from pbr_b200 import host
    S = pbr_b200.scenes
    scene = {"soup": lambda: S.soup(6000, seed=11), "grid": lambda: S.displaced_grid(40, 32, patches=4, seed=5),
             "interior": lambda: S.interior(detail=0.12)}[gen]()
    scene = _with_face_normals(scene)
    path = str(tmp_path / "scene.obj")
    S.write_obj(scene, path)
    ref_scene, ref_flat = RH.load(path, shadow_rays=0, **kw)
    ora_scene = O.load_obj(path, 0)
    _same_scene(ora_scene, ref_scene)
    _same_flat(O.build_bvh(ora_scene, **kw), ref_flat)
    loader = lambda: host.Scene.load(str(tmp_path) + "/", "scene.obj")      # noqa: E731
    _same_scene(loader().to_dict(), ref_scene)
    _same_flat(_product_flat(cfg, loader, kw), ref_flat)
    assert ref_flat["info"]["faces"] == scene["facesV"].size // 3 > 1000


# ---------------------------------------------------------------------------------------------------------
# The reference's renderer core end to end: ModelLoader -> BVH -> PathTracer.cpp (buffers, kernel arguments,
# camera, per-frame seed and weight) -> the reference kernel, all of it the reference's own code.

RENDER_CONFIGS = {
    "sa": dict(brdf=1),
    "schlick_shadow_ms": dict(brdf=0, shadow_rays=1, samples=2),
    "phong_camera_fov": dict(brdf=1, phong_tess=0.7, eye=(0.3, 0.9, 2.5), center=(0.1, 0.2, 1.0), fov=60.0, antialiasing=0.4),
    "sa_bvh1": dict(brdf=1, max_faces=1, skip_ahead_compare=0.5, max_depth=5, max_added_depth=2),
}


def _prepared_for(path, W, H, kw):
    import helpers as Hh
    kw = dict(kw)
    bvh_kw = {k: kw.pop(k) for k in ("max_faces", "skip_ahead", "skip_ahead_compare", "sah_faces_limit") if k in kw}
    pt = kw.pop("phong_tess", 0.0)
    kw.setdefault("max_depth", 4)
    scene = O.load_obj(path, kw.get("shadow_rays", 0))
    return Hh.Prepared(scene, W, H, phong_tessellation=pt, bvh_kwargs=bvh_kw, **kw)


@pytest.mark.parametrize("name", sorted(RENDER_CONFIGS))
def test_reference_renderer_equals_oracle_pipeline(name):
    """PathTracer.cpp's host arithmetic (camera basis, pixel size with its binary64 tangent, material / light
    packing, SKY_LIGHT / NUM_LIGHTS / BVH_NUM_NODES texts, seed and mixing weight per frame) against the
    restatement in oracle/scene.py + tests/helpers.py, through the image it produces."""
    import helpers as Hh
    from oracle import scene as S
    kw = dict(RENDER_CONFIGS[name])
    kw.setdefault("max_depth", 4)
    path = os.path.join(MODELS, "suzanne.obj")
    W, H = 88, 56
    r = RH.Renderer(path, width=W, height=H, **kw)
    try:
        p = _prepared_for(path, W, H, kw)
        assert int(r.values["BVH_NUM_NODES"]) == p.nodes.shape[0] and int(r.values["NUM_LIGHTS"]) == p.num_lights
        img_o = np.zeros((H, W, 4), np.float32)
        for k, ms in enumerate((33, 67, 100)):
            img_r, dbg_r = r.generate_image(ms)
            seed = np.float32(ms) * np.float32(0.001)
            img_o, dbg_o, _ = O.path_tracing(p.defines, seed, S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV,
                                             p.facesN, p.vertices4, p.normals4, p.materials, p.lights, img_o, nthreads=4)
            assert Hh.images_equal(img_r, img_o) and Hh.images_equal(dbg_r, dbg_o), "frame %d" % k
        # the kernel arguments themselves
        assert r.kernel_arg(2, np.float32)[0] == p.px_dim
        cam = r.kernel_arg(3).view(O.CAMERA_DTYPE)
        for f in ("eye", "w", "u", "v"):
            assert np.array_equal(cam[f][0, :3], p.camera[f][0, :3]), f
        assert np.array_equal(cam["focusPoint"], p.camera["focusPoint"]) and np.array_equal(cam["lense"], p.camera["lense"])
        assert np.array_equal(r.kernel_arg(4, np.float32).view(np.uint32), p.nodes.ravel().view(np.uint32))
        assert np.array_equal(r.kernel_arg(5, np.uint32), p.facesV.ravel())
        assert np.array_equal(r.kernel_arg(9, np.float32), np.asarray(p.materials, np.float32).ravel())
        if p.num_lights:
            assert np.array_equal(r.kernel_arg(10, np.float32), p.lights.ravel())
    finally:
        r.close()


def test_number_spellings_parse_like_the_reference(cfg, tmp_path):
    """The reference reads coordinates with atof(); the product takes a shortcut (std::from_chars) for tokens that are
    one plain decimal number and keeps atof for the rest.  Every spelling below must give the reference's bits."""
    from pbr_b200 import host
    spellings = ["1", "-1", "0.1", "-0.3333333333333333333333", "1e-3", "1E+2", "+0.5", ".5", "-.5", "1.", "1e", "1e+",
                 "0x10", "-0x1p-2", "1e999", "-1e999", "1e-999", "4.9e-324", "1.17549435e-38", "3.4028235e38", "3.5e38",
                 "-", "+", ".", "e5", "1.5abc", "12,5", "nan", "inf", "-infinity", "0", "-0", "00012.5000", "1_000",
                 "0.1234567890123456789012345678901234567890123456789012345678901234567890", "16777217", "1e23", "8.5e-46"]
    lines = ["o numbers", "usemtl none"]
    for k in range(0, len(spellings) - 2):
        lines.append("v %s %s %s" % (spellings[k], spellings[k + 1], spellings[k + 2]))
        lines.append("vn %s %s %s" % (spellings[k + 2], spellings[k], spellings[k + 1]))
    lines += ["f 1//1 2//2 3//3", "f 2//2 3//3 4//4"]
    (tmp_path / "numbers.obj").write_text("\n".join(lines) + "\n")
    (tmp_path / "numbers.mtl").write_text("newmtl none\nKd 1e-1 +0.5 .25\nd 0x1p-1\nNi 1.5abc\n")
    path = str(tmp_path / "numbers.obj")
    ref_scene, _ = RH.load(path, build_bvh=False, shadow_rays=0)
    assert ref_scene["vertices"].size == 3 * (len(spellings) - 2)
    _same_scene(O.load_obj(path, 0), ref_scene)
    _same_scene(host.Scene.load(str(tmp_path) + "/", "numbers.obj").to_dict(), ref_scene)


@pytest.mark.parametrize("seed", range(24))
def test_random_obj_text_parses_like_the_reference(cfg, tmp_path, seed):
    """tests/obj_fuzz.py: reference classes == restatement == product on random OBJ text (parsers only)."""
    import obj_fuzz
    from pbr_b200 import host
    rng = np.random.default_rng(seed)
    with open(tmp_path / "f.obj", "w", newline="") as f:
        f.write(obj_fuzz.gen(rng, int(rng.integers(20, 400))))
    (tmp_path / "f.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl glass\nd 0.5\nNi 1.5\nnewmtl sky_light\nKd 0.9 0.9 1\n")
    path = str(tmp_path / "f.obj")
    ref_scene, _ = RH.load(path, build_bvh=False, shadow_rays=0)
    _same_scene(O.load_obj(path, 0), ref_scene)
    _same_scene(host.Scene.load(str(tmp_path) + "/", "f.obj").to_dict(), ref_scene)


@pytest.mark.parametrize("seed", range(24))
def test_random_mtl_and_lights_text_parse_like_the_reference(cfg, tmp_path, seed):
    """Random .mtl and .lights next to a fixed .obj; with render.shadow_rays = 1 the .lights file is read, and an
    empty one switches shadow rays off (LightParser.cpp:119-121)."""
    import obj_fuzz
    from pbr_b200 import host
    rng = np.random.default_rng(1000 + seed)
    (tmp_path / "f.obj").write_text("o a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nusemtl red\nf 1//1 2//1 3//1\n"
                                    "usemtl glass\nf 3//1 2//1 1//1\nusemtl m3\nf 1//1 3//1 2//1\n")
    with open(tmp_path / "f.mtl", "w", newline="") as f:
        f.write(obj_fuzz.gen_mtl(rng, int(rng.integers(5, 120))))
    with open(tmp_path / "f.lights", "w", newline="") as f:
        f.write(obj_fuzz.gen_lights(rng, int(rng.integers(0, 40))))
    path = str(tmp_path / "f.obj")
    for shadow_rays in (1, 0):
        ref_scene, _ = RH.load(path, build_bvh=False, shadow_rays=shadow_rays)
        ora = O.load_obj(path, shadow_rays)
        _same_scene(ora, ref_scene)
        cfg.set("render.shadow_rays", shadow_rays)
        _same_scene(host.Scene.load(str(tmp_path) + "/", "f.obj").to_dict(), ref_scene)
        forced_off = int(cfg.get("render.shadow_rays")) == 0 and shadow_rays == 1
        assert forced_off == ora["shadowRaysForcedOff"] == (shadow_rays == 1 and ref_scene["lights"].shape[0] == 0)


_REF_CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, sys.argv[4]); sys.path.insert(0, sys.argv[4] + "/tests")
from oracle import ref_host as RH
scene, flat = RH.load(sys.argv[1], shadow_rays=0, **json.loads(sys.argv[2]))
np.savez(sys.argv[3], nodes=flat["nodes"], facesV=flat["facesV"], facesN=flat["facesN"], info=json.dumps(flat["info"]))
"""


@pytest.mark.parametrize("seed", range(16))
def test_random_scenes_build_the_reference_tree(cfg, tmp_path, seed):
    """tests/obj_fuzz.py::gen_bvh_scene through the three builders.  The reference's BVH class runs in a child process:
    it does not survive every input (a scene of one single face crashes it) -- there the restatement and the product
    still have to agree with each other."""
    import json
    import subprocess
    import sys
    import obj_fuzz
    from conftest import ROOT
    from pbr_b200 import host
    text, kw, nfaces = obj_fuzz.gen_bvh_scene(np.random.default_rng(seed))
    (tmp_path / "s.obj").write_text(text)
    (tmp_path / "s.mtl").write_text("newmtl m\nKd 1 1 1\n")
    path = str(tmp_path / "s.obj")
    child = subprocess.run([sys.executable, "-c", _REF_CHILD, path, json.dumps(kw), str(tmp_path / "ref.npz"), ROOT],
                           capture_output=True)
    ora_flat = O.build_bvh(O.load_obj(path, 0), **kw)
    _same_flat(_product_flat(cfg, lambda: host.Scene.load(str(tmp_path) + "/", "s.obj"), kw), ora_flat)
    assert ora_flat["info"]["faces"] == nfaces
    if child.returncode != 0:
        assert nfaces == 1, child.stderr.decode()[-400:]        # the only input known to crash the reference
        return
    z = np.load(str(tmp_path / "ref.npz"))
    _same_flat(ora_flat, dict(nodes=z["nodes"], facesV=z["facesV"], facesN=z["facesN"], info=json.loads(str(z["info"]))))
