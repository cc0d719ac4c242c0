"""tests/golden/reference_outputs.npz holds outputs of THE REFERENCE ITSELF (its kernel source and host classes
compiled for the host, tests/golden/make_golden.py): frames, debug images, the flattened BVH, explicit-ray hits.
It needs neither /root/reference nor oracle/_ref at test time, so these tests run everywhere.
CPU: the oracle restatement and the product's host library against it.  GPU: the CUDA path against it."""
import os
import sys

import numpy as np
import pytest

import helpers as Hh
from conftest import MODELS

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402   (only its constants: W, H, FRAME_MS, CASES)


@pytest.fixture(scope="module")
def gold():
    return np.load(G.OUT)


def _prepared(oracle, kw):
    kw = dict(kw)
    pt = kw.pop("phong_tess", 0.0)
    scene = oracle.load_obj(os.path.join(MODELS, "suzanne.obj"), kw.get("shadow_rays", 0))
    return Hh.Prepared(scene, G.W, G.H, phong_tessellation=pt, **kw)


def _seeds():
    return [np.float32(ms) * np.float32(0.001) for ms in G.FRAME_MS]


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_frames_equal_reference_outputs(oracle, gold, name):
    from oracle import scene as S
    p = _prepared(oracle, G.CASES[name])
    values = dict(v.split("=", 1) for v in gold[name + "/values"])
    assert int(values["BVH_NUM_NODES"]) == p.nodes.shape[0] and int(values["NUM_LIGHTS"]) == p.num_lights
    img = np.zeros((G.H, G.W, 4), np.float32)
    for k, seed in enumerate(_seeds()):
        img, dbg, _ = oracle.path_tracing(p.defines, seed, S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV,
                                          p.facesN, p.vertices4, p.normals4, p.materials, p.lights, img, nthreads=4)
    assert Hh.images_equal(img, gold[name + "/image"])
    assert Hh.images_equal(dbg, gold[name + "/debug"])


def test_bvh_equals_reference_outputs(oracle, gold):
    from pbr_b200 import host
    want = {"nodes": gold["bvh/nodes"], "facesV": gold["bvh/facesV"], "facesN": gold["bvh/facesN"]}
    path = os.path.join(MODELS, "suzanne.obj")
    got = oracle.build_bvh(oracle.load_obj(path, 1))
    c = host.Config()
    c.reset()
    c.set("render.shadow_rays", 1)
    prod = host.Scene.load(MODELS + "/", "suzanne.obj").build_flat()
    c.reset()
    for flat in (got, prod):
        assert np.array_equal(flat["nodes"].view(np.uint32), want["nodes"].view(np.uint32))
        assert np.array_equal(flat["facesV"], want["facesV"]) and np.array_equal(flat["facesN"], want["facesN"])
        assert [flat["info"][k] for k in ("allNodes", "leaves", "depth", "skipped", "emitted", "faces")] == gold["bvh/info"].tolist()
    assert gold["bvh/info"][5] == 1082                       # pathtracing.cl:75


def test_oracle_hits_equal_reference_outputs(oracle, gold):
    p = _prepared(oracle, dict(brdf=1, shadow_rays=1))
    got, _ = p.oracle_trace(gold["rays"], nthreads=4)
    assert np.array_equal(got["t"].view(np.uint32), gold["hits/t"].view(np.uint32))
    assert np.array_equal(got["hitFace"], gold["hits/face"])
    assert np.array_equal(got["visits"] & 0xfffff, gold["hits/nodes"].astype(np.uint32))
    assert np.array_equal(got["visits"] >> 20, gold["hits/tris"].astype(np.uint32))
    got, _ = p.oracle_trace(gold["shadow/rays"], any_hit=True, nthreads=4)
    assert np.array_equal(got["t"].view(np.uint32), gold["shadow/t"].view(np.uint32))
    assert np.array_equal(got["hitFace"], gold["shadow/face"])
    assert np.array_equal(got["visits"] >> 20, gold["shadow/tris"].astype(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [0, 1])
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_device_frames_equal_reference_outputs(device, oracle, gold, name, pipeline):
    from oracle import scene as S
    p = _prepared(oracle, G.CASES[name])
    ds = Hh.DeviceScene(device, p)
    device.setPipeline(pipeline)
    try:
        img = np.zeros((G.H, G.W, 4), np.float32)
        k = ds.kernel
        for n, seed in enumerate(_seeds()):
            device.updateImageReadOnly(ds.texIn, p.W, p.H, img)
            device.setKernelArg(k, 0, seed)
            device.setKernelArg(k, 1, S.pixel_weight(n))
            device.setKernelArg(k, 3, p.camera)
            device.execute(k)
            device.finish()
            img = device.readImageOutput(ds.texOut, p.W, p.H)
        dbg = device.readImageOutput(ds.texDebug, p.W, p.H)
    finally:
        device.setPipeline(-1)
    assert Hh.images_equal(img, gold[name + "/image"])
    assert Hh.images_equal(dbg, gold[name + "/debug"])


@pytest.mark.gpu
def test_device_hits_equal_reference_outputs(device, oracle, gold):
    p = _prepared(oracle, dict(brdf=1, shadow_rays=1))
    ds = Hh.DeviceScene(device, p)
    got = ds.trace(gold["rays"])
    assert np.array_equal(got["t"].view(np.uint32), gold["hits/t"].view(np.uint32))
    assert np.array_equal(got["hitFace"], gold["hits/face"])
    assert np.array_equal(got["visits"] & 0xfffff, gold["hits/nodes"].astype(np.uint32))
    assert np.array_equal(got["visits"] >> 20, gold["hits/tris"].astype(np.uint32))
    got = ds.trace(gold["shadow/rays"], any_hit=True)
    assert np.array_equal(got["t"].view(np.uint32), gold["shadow/t"].view(np.uint32))
    assert np.array_equal(got["hitFace"], gold["shadow/face"])
    assert np.array_equal(got["visits"] >> 20, gold["shadow/tris"].astype(np.uint32))
