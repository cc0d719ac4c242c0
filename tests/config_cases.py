"""BASELINE.json's configurations as the tests use them: one place for scene, camera and
kernel values, shared by the GPU parity tests (tests/test_gpu_configs.py) and by tests/ref_configs.py, which
prebuilds the reference's own kernel (oracle/_ref) for the configurations that compare whole frames.

C1  bundled Cornell box + Suzanne (tests/golden/models/suzanne.obj), 512x512, 1 spp, max_depth 4
C2  1 000 000-triangle soup, 1920x1080 (bench.py; tests/test_gpu_fullsize.py)
C3  procedural interior (~254 k triangles, 8 materials), 1920x1080, BRDF 1 and BRDF 0
C4  displaced grid, 10 003 864 triangles in 64 objects, 3840x2160
C5  explicit primary + shadow rays on the C2 scene, 1 M and 5 M rays
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, _HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

C1 = dict(width=512, height=512, brdf=1, samples=1, max_depth=4, max_added_depth=5, shadow_rays=0, antialiasing=0.7,
          eye=(0.0, 1.0, 3.0), center=(0.0, 0.0, 1.0))
C3 = dict(width=1920, height=1080, samples=1, max_depth=3, max_added_depth=5, shadow_rays=0, antialiasing=0.7,
          eye=(0.0, 1.4, 5.2), center=(0.0, 0.1, 1.0))
C4 = dict(width=3840, height=2160, brdf=1, samples=1, max_depth=3, max_added_depth=5, shadow_rays=0, antialiasing=0.7,
          eye=(0.0, 1.2, 1.8), center=(0.0, 0.55, 1.0))
C4_CELLS = (2237, 2236)
C5_LIGHT = (0.0, 3.0, 0.0)

_CACHE = {}


def c1_prepared():
    import helpers as Hh
    from oracle import oracle as O
    if "c1" not in _CACHE:
        _CACHE["c1"] = Hh.Prepared(O.load_obj(Hh.model_path("suzanne.obj"), 0), **C1)
    return _CACHE["c1"]


def c3_scene():
    import pbr_b200
    if "c3s" not in _CACHE:
        _CACHE["c3s"] = pbr_b200.scenes.interior()
    return _CACHE["c3s"]


def c3_prepared(brdf, width=None, height=None):
    """The oracle-side view of C3; the BVH is the oracle's own (the product's builder gives the same tree bit for
    bit: tests/test_host_parity.py, test_oracle_vs_reference_host.py)."""
    import helpers as Hh
    from oracle import oracle as O
    key = ("c3", brdf, width, height)
    if key not in _CACHE:
        if "c3bvh" not in _CACHE:
            _CACHE["c3bvh"] = O.build_bvh(c3_scene())
        kw = dict(C3)
        if width:
            kw["width"], kw["height"] = width, height
        _CACHE[key] = Hh.Prepared(c3_scene(), brdf=brdf, bvh=_CACHE["c3bvh"], **kw)
    return _CACHE[key]


def host_config(cfg, case, **extra):
    """The same case as Cfg keys for the product's host layer."""
    cfg.reset()
    cfg.update({
        "window.width": case["width"], "window.height": case["height"],
        "render.samples": case["samples"], "render.max_depth": case["max_depth"],
        "render.max_added_depth": case["max_added_depth"], "render.shadow_rays": case["shadow_rays"],
        "render.antialiasing": case["antialiasing"], "logging.level": 1,
    })
    if "brdf" in case:
        cfg.set("render.brdf", case["brdf"])
    for axis, e, c in zip("xyz", case["eye"], case["center"]):
        cfg.update({"camera.eye." + axis: e, "camera.center." + axis: c})
    cfg.update(extra)


def checker_frames(prep, nframes, nthreads=None, image=None):
    """`nframes` accumulated frames of `prep` on the host CPU: the reference's own kernel when it was prebuilt for
    this configuration (oracle/_ref), else the restatement (pinned to it bit for bit by the CPU tests).
    Returns (image, debug image, kind)."""
    from oracle import ref as R
    from oracle import scene as S
    nthreads = nthreads or (os.cpu_count() or 1)
    if R.available(prep.defines):
        img = np.zeros((prep.H, prep.W, 4), np.float32) if image is None else image
        dbg = None
        for k in range(nframes):
            img, dbg = R.path_tracing(prep.defines, S.frame_seed(k), S.pixel_weight(k), prep.px_dim, prep.camera,
                                      prep.nodes, prep.facesV, prep.facesN, prep.vertices4, prep.normals4,
                                      prep.materials, prep.lights, img, nthreads=nthreads)
        return img, dbg, "reference"
    img, dbg, _ = prep.oracle_frames(nframes, nthreads=nthreads, image=image)
    return img, dbg, "port"


def program_values():
    """What oracle/build_ref.py prebuilds for these configurations."""
    from oracle import ref as R
    yield R.values_from_defines(c1_prepared().defines)
    for brdf in (1, 0):
        yield R.values_from_defines(c3_prepared(brdf).defines)
