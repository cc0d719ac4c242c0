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
