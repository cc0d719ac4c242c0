"""CPU: the product's host mirror (libpbr_host.so: ObjParser / MtlParser / LightParser / Cfg / BVH /
flatten) against the oracle's literal restatement of the reference (oracle/obj_oracle.cpp,
bvh_oracle.cpp).  Bar: identical arrays, bit for bit."""
import os

import numpy as np
import pytest

import helpers as Hh
from conftest import MODELS

ARRAY_KEYS = ["vertices", "normals", "facesV", "facesVN", "facesVT", "textures", "facesMtl", "objFaceCounts",
              "objFacesV", "objFacesVN", "objNormalFaceCounts", "materials", "lights"]


@pytest.fixture()
def cfg():
    from pbr_b200 import host
    c = host.Config()
    c.reset()
    yield c
    c.reset()


def _same_scene(a, b):
    for k in ARRAY_KEYS:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k]), k
    for k in ("materialNames", "objectNames", "lightNames"):
        assert a[k] == b[k], k


def _same_flat(a, b):
    assert a["info"] == b["info"]
    assert np.array_equal(a["nodes"].view(np.uint32), b["nodes"].view(np.uint32))
    assert np.array_equal(a["facesV"], b["facesV"])
    assert np.array_equal(a["facesN"], b["facesN"])


@pytest.mark.parametrize("name", ["suzanne.obj", "pillars.obj", "quirks.obj"])
@pytest.mark.parametrize("shadow_rays", [0, 1])
def test_parsers_match_reference_restatement(oracle, cfg, name, shadow_rays):
    from pbr_b200 import host
    cfg.set("render.shadow_rays", shadow_rays)
    got = host.Scene.load(MODELS + "/", name).to_dict()
    want = oracle.load_obj(os.path.join(MODELS, name), shadow_rays)
    _same_scene(got, want)
    # LightParser switches shadow rays off when the .lights file holds no light (LightParser.cpp:119-121)
    forced_off = int(cfg.get("render.shadow_rays")) == 0 and shadow_rays == 1
    assert forced_off == want["shadowRaysForcedOff"]


def test_parser_quirks_are_the_reference_ones(cfg):
    """Spot checks that the quirks are really there (not just equal on both sides)."""
    from pbr_b200 import host
    cfg.set("render.shadow_rays", 1)
    d = host.Scene.load(MODELS + "/", "quirks.obj").to_dict()
    v = d["vertices"].reshape(-1, 3)
    assert np.array_equal(v[3], (0, 2, 0))                      # "v  2 0 0": the empty token reads as 0
    # 8 `f` lines; the last one, "f 2  3 5", has an empty group that reads as index 0 - 1 = 0xFFFFFFFF
    # and so contributes FOUR indices: everything after it would be misaligned (reference behaviour)
    assert d["facesV"].size == 7 * 3 + 4 and len(d["facesMtl"]) == 8
    assert d["facesV"][-4:].tolist() == [1, 0xFFFFFFFF, 2, 4]
    assert d["objFaceCounts"].tolist() == [4, 3]                # face before the first `o` is in no object
    fv = d["facesV"][:21].reshape(-1, 3)
    assert fv[6].tolist() == [2, 1, 0]                          # leading/trailing blanks are trimmed
    assert d["facesMtl"].tolist() == [-1, 0, 0, 0, -1, 1, 1, 1]   # unknown usemtl -> -1; kept across `o`
    # "v/vt" is read as "v//vn" (is_any_of("//")): face 2 put its vt indices into facesVN
    assert d["facesVN"][3:6].tolist() == [1, 0, 1]
    m = dict(zip(d["materialNames"], d["materials"]))
    assert m["red"][15] == 2.0                                   # illum 99 is rejected -> 2
    assert np.isclose(m["glass"][12], 0.5)                       # d wins over the earlier Tr
    assert np.isclose(m["later"][12], 1.0)                       # Tr ignored once any d was seen
    assert np.array_equal(m["glass"][0:3], (1, 1, 1))            # "Ka 0.1 0.2": too few values, ignored
    assert np.allclose(m["later"][4:7], (0, 0.3, 0.3))           # "Kd  0.3 0.3 0.3": the empty token reads as 0
    li = d["lights"]
    assert li.shape == (2, 10) and li[1, 0] == 2 and np.isclose(li[1, 9], 0.2)
    assert np.array_equal(li[1, 5:8], (1, 1, 1))                 # "rgb 0.5 0.5": too few values, ignored


def test_config_reader(cfg, tmp_path):
    from pbr_b200 import host
    # the reference's config.json carries // comments (boost::property_tree tolerates them)
    p = tmp_path / "config.json"
    p.write_text('{\n // comment\n "window": { "width": 512, /* c */ "height": 256 },\n'
                 ' "render": { "brdf": 0, "antialiasing": 0.25 }, "bvh": { "skip_ahead": false },\n'
                 ' "import_path": "/a path/with spaces/" }\n')
    cfg.load_file(str(p))
    assert cfg.get("window.width") == "512" and cfg.get("window.height") == "256"
    assert cfg.get("render.brdf") == "0" and cfg.get("bvh.skip_ahead") == "false"
    assert cfg.get("import_path") == "/a path/with spaces/"
    assert cfg.get("render.max_depth") == "3"                    # untouched keys keep the shipped default
    with pytest.raises(host.HostError):
        cfg.get("no.such.key")
    # every key of the reference (Cfg.cpp:4-39) exists in the defaults
    cfg.reset()
    for key in ("accel_struct", "bvh.max_faces", "bvh.sah_faces_limit", "bvh.skip_ahead", "bvh.skip_ahead_compare",
                "camera.center.x", "camera.center.y", "camera.center.z", "camera.eye.x", "camera.eye.y", "camera.eye.z",
                "camera.thin_lense.aperture", "camera.thin_lense.focal_length", "camera.speed", "import_path",
                "info.kernel_times", "logging.level", "opencl.build_options", "opencl.check_errors",
                "opencl.localgroupsize", "opencl.program", "camera.perspective.fov", "camera.perspective.zfar",
                "camera.perspective.znear", "render.antialiasing", "render.brdf", "render.interval",
                "render.max_added_depth", "render.max_depth", "render.phong_tessellation", "render.samples",
                "render.shadow_rays", "shader.name", "shader.path", "window.height", "window.width"):
        cfg.get(key)


@pytest.mark.parametrize("name", ["suzanne.obj", "pillars.obj"])
@pytest.mark.parametrize("max_faces,skip_ahead,cmp_", [(2, True, 0.7), (1, True, 0.5), (2, False, 0.7), (4, True, 0.9)])
def test_bvh_flat_identical_bundled(oracle, cfg, name, max_faces, skip_ahead, cmp_):
    from pbr_b200 import host
    cfg.update({"bvh.max_faces": max_faces, "bvh.skip_ahead": skip_ahead, "bvh.skip_ahead_compare": cmp_})
    got = host.Scene.load(MODELS + "/", name).build_flat()
    want = oracle.build_bvh(oracle.load_obj(os.path.join(MODELS, name), 0), max_faces=max_faces,
                            skip_ahead=skip_ahead, skip_ahead_compare=cmp_)
    _same_flat(got, want)


@pytest.mark.parametrize("ntri,sah_limit", [(2, 100000), (3, 100000), (5000, 100000), (5000, 700), (60000, 20000)])
def test_bvh_flat_identical_soup(oracle, cfg, ntri, sah_limit):
    """SAH sweep path, mean-split path (faces > sah_faces_limit) and their mix; 20k+ faces also
    exercise the worker threads of the product builder."""
    import pbr_b200
    from pbr_b200 import host
    sc = pbr_b200.scenes.soup(ntri, seed=11)
    cfg.set("bvh.sah_faces_limit", sah_limit)
    got = host.Scene.from_arrays(sc).build_flat()
    want = oracle.build_bvh(sc, sah_faces_limit=sah_limit)
    _same_flat(got, want)


def test_bvh_flat_identical_grid_objects_and_ties(oracle, cfg):
    """Displaced grid split into 16 `o` objects: per-object trees grouped into a top tree, and many
    exactly equal centre coordinates (ties in the SAH sort must fall as in the reference)."""
    import pbr_b200
    from pbr_b200 import host
    sc = pbr_b200.scenes.displaced_grid(48, 40, patches=4, seed=5)
    assert len(sc["objFaceCounts"]) == 16 and sc["facesV"].size // 3 == 48 * 40 * 2
    got = host.Scene.from_arrays(sc).build_flat()
    want = oracle.build_bvh(sc)
    _same_flat(got, want)


def test_phong_tessellation_grows_boxes_identically(oracle, cfg):
    from pbr_b200 import host
    cfg.set("render.phong_tessellation", 0.8)
    got = host.Scene.load(MODELS + "/", "suzanne.obj").build_flat()
    o = oracle.load_obj(os.path.join(MODELS, "suzanne.obj"), 0)
    want = oracle.build_bvh(o, phong_tess=0.8)
    _same_flat(got, want)
    flat = oracle.build_bvh(o, phong_tess=0.0)
    assert not np.array_equal(flat["nodes"], want["nodes"])


def test_obj_roundtrip_through_writer(oracle, cfg, tmp_path):
    """scenes.write_obj -> the parser gives back the generated arrays (the 1M / 10M configurations can
    therefore be fed through the reference's own OBJ path)."""
    import pbr_b200
    from pbr_b200 import host
    sc = pbr_b200.scenes.soup(500, seed=2)
    path = pbr_b200.scenes.write_obj(sc, str(tmp_path / "soup.obj"))
    d = host.Scene.load(str(tmp_path) + "/", os.path.basename(path)).to_dict()
    assert np.array_equal(d["vertices"], sc["vertices"]) and np.array_equal(d["facesV"], sc["facesV"])
    assert d["objFaceCounts"].tolist() == [500] and np.array_equal(d["facesMtl"], sc["facesMtl"])
    assert np.allclose(d["materials"][0, 4:7], 0.8)


def test_host_library_exports_declared_symbols():
    import ctypes
    import re
    from pbr_b200 import host
    from conftest import ROOT
    text = open(os.path.join(ROOT, "include", "pbr_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pbrh_[a-z0-9_]+)\s*\(", text)))
    lib = ctypes.CDLL(host.LIB_PATH)
    assert not [n for n in declared if not hasattr(lib, n)]
    assert declared == sorted(host.SYMBOLS)
