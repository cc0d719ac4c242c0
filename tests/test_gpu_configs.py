"""GPU: BASELINE.json's configurations at their full sizes against the reference kernel (oracle/_ref, prebuilt by
__graft_entry__.build(); the restatement where it is missing), through the host layer and the C ABI.

  C1  suzanne.obj 512x512, max_depth 4: whole frame + debug image bit for bit
  C3  interior, 1920x1080, one frame each for BRDF 1 and BRDF 0: whole frame + debug image bit for bit
  C4  10 003 864-triangle displaced grid: 90 k explicit rays (face, leaf, t, visits) bit for bit; the 3840x2160 frame
      rendered in 4 interleaved stripes == the frame rendered whole
  C5  explicit primary + shadow rays on the 1 M-triangle soup, 1 M and 5 M rays: face, leaf, t bit for bit on every ray
(C2 at full size: tests/test_gpu_fullsize.py and bench.py's `verify`.)
"""
import os

import numpy as np
import pytest

import config_cases as CC
import helpers as Hh
from conftest import MODELS

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


@pytest.fixture()
def cfg():
    from pbr_b200 import host
    c = host.Config()
    c.reset()
    yield c
    c.reset()


def _set_traversal(r, mode):
    """-1 automatic, 0 the reference's visiting order, 1 the ordered wide-BVH walk (pbr_set_traversal)."""
    r.device().setTraversal(mode)


def _assert_same_tree(flat, prep):
    assert np.array_equal(flat["nodes"].view(np.uint32), prep.nodes.view(np.uint32))
    assert np.array_equal(flat["facesV"], prep.facesV)


def test_c1_suzanne_512(cfg):
    from pbr_b200 import host
    CC.host_config(cfg, CC.C1)
    r = host.Renderer(0)
    try:
        r.set_deterministic(True)
        r.load_model(MODELS + "/", "suzanne.obj")
        got, gdbg = r.generate_image(debug=True)
        prep = CC.c1_prepared()
        _assert_same_tree(r.flat(), prep)
        want, wdbg, kind = CC.checker_frames(prep, 1, THREADS)
        assert Hh.mean_relative_error(got, want) <= 0.02
        assert Hh.images_equal(got, want), "C1 frame differs from the %s kernel" % kind
        assert Hh.images_equal(gdbg, wdbg)
        # without the debug image (the walk may then take the ordered wide-BVH route): the same picture
        r.reset_sample_count()
        assert Hh.images_equal(r.generate_image(), want)
    finally:
        r.close()


@pytest.mark.parametrize("brdf", [1, 0])
def test_c3_interior_1080p(cfg, brdf):
    from pbr_b200 import host
    CC.host_config(cfg, dict(CC.C3, brdf=brdf))
    r = host.Renderer(0)
    try:
        r.set_deterministic(True)
        r.load_scene(CC.c3_scene())
        got, gdbg = r.generate_image(debug=True)
        prep = CC.c3_prepared(brdf)
        _assert_same_tree(r.flat(), prep)
        want, wdbg, kind = CC.checker_frames(prep, 1, THREADS)
        assert Hh.mean_relative_error(got, want) <= 0.02
        assert Hh.count_identical_pixels(got, want) == prep.W * prep.H, "C3 BRDF %d vs the %s kernel" % (brdf, kind)
        assert Hh.images_equal(gdbg, wdbg)
        r.reset_sample_count()
        assert Hh.images_equal(r.generate_image(), want)
    finally:
        r.close()


def test_c4_grid_10m(cfg):
    import pbr_b200
    from pbr_b200 import host
    CC.host_config(cfg, CC.C4)
    scene = pbr_b200.scenes.displaced_grid(CC.C4_CELLS[0], CC.C4_CELLS[1], patches=8)
    r = host.Renderer(0)
    try:
        r.set_deterministic(True)
        r.load_scene(scene)
        info = r.info()
        assert info["faces"] == 10_003_864
        flat = r.flat()
        prep = Hh.Prepared(scene, 64, 64, bvh=flat, eye=CC.C4["eye"], center=CC.C4["center"])
        cam, _ = r.camera()

        class P:
            camera = cam
        rng = np.random.default_rng(11)
        prim = Hh.primary_rays(P, 400, 225)                       # 90 000 rays through the 16:9 frame
        rnd = Hh.random_rays(10_000, 4, -1.0, 1.0)
        rnd[:, 1] = np.abs(rnd[:, 1]) * 0.5 + 0.3                # origins above the height field
        rays = np.concatenate([prim, rnd])
        want, _ = prep.oracle_trace(rays, nthreads=THREADS)
        for forced in (0, 1):                                     # reference-order walk, ordered wide-BVH walk
            _set_traversal(r, forced)
            got = r.trace(rays)
            assert np.array_equal(got["hitFace"], want["hitFace"]), "traversal %d" % forced
            assert np.array_equal(got["leaf"], want["leaf"]), "traversal %d" % forced
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), "traversal %d" % forced
            if forced == 0:
                assert np.array_equal(got["visits"], want["visits"])
        _set_traversal(r, -1)
        assert 0.2 < np.isfinite(want["t"]).mean() <= 1.0
        sh = Hh.shadow_rays_from_hits(rays, want, (0.0, 3.0, 0.0))[:: max(1, int(rng.integers(1, 3)))]
        want_s, _ = prep.oracle_trace(sh, any_hit=True, nthreads=THREADS)
        got_s = r.trace(sh, any_hit=True)
        assert np.array_equal(got_s["hitFace"], want_s["hitFace"])
        assert np.array_equal(got_s["t"].view(np.uint32), want_s["t"].view(np.uint32))

        # the 4K frame: whole, and stitched from 4 interleaved stripe sets (what 4 ranks would render)
        W, H = CC.C4["width"], CC.C4["height"]
        r.render_frames(2)
        whole = r.read_image()
        assert np.isfinite(whole[..., :3]).all() and whole[..., :3].mean() > 0.05
        stitched = np.zeros_like(whole)
        from pbr_b200 import multigpu
        world = 4
        stripe = multigpu.stripe_rows_for(H, world, want=8)      # IMG_HEIGHT must be a multiple of stripe * world
        assert stripe > 0
        rows = np.arange(H)
        for rank in range(world):
            r.set_tile_stripes(stripe, world, rank)
            r.reset_sample_count()
            r.render_frames(2)
            part = r.read_image()
            mine = (rows // stripe) % world == rank
            stitched[mine] = part[mine]
        r.set_tile_stripes(0)
        assert Hh.count_identical_pixels(stitched, whole) == W * H
    finally:
        r.close()


@pytest.mark.parametrize("mega", [1, 5])
def test_c5_explicit_rays(cfg, mega):
    import pbr_b200
    from pbr_b200 import host
    CC.host_config(cfg, dict(CC.C1, width=1920, height=1080, eye=(0.0, 0.0, 3.5), max_depth=3))
    scene = pbr_b200.scenes.soup(1_000_000, seed=12345)
    r = host.Renderer(0)
    try:
        r.set_deterministic(True)
        r.load_scene(scene)
        flat = r.flat()
        prep = Hh.Prepared(scene, 64, 64, bvh=flat, eye=(0.0, 0.0, 3.5))
        cam, _ = r.camera()

        class P:
            camera = cam
        n_total = mega * 1_000_000
        h = int(round((n_total * 9 / 16) ** 0.5))
        w = n_total // h
        rays = Hh.primary_rays(P, w, h)                           # pinhole grid oversampling the 1080p image
        want, _ = prep.oracle_trace(rays, nthreads=THREADS)
        for forced in (0, 1):
            _set_traversal(r, forced)
            got = r.trace(rays)
            for f in ("hitFace", "leaf"):
                assert np.array_equal(got[f], want[f]), "%s, traversal %d" % (f, forced)
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), "t, traversal %d" % forced
        _set_traversal(r, -1)
        sh = Hh.shadow_rays_from_hits(rays, want, CC.C5_LIGHT)
        want_s, _ = prep.oracle_trace(sh, any_hit=True, nthreads=THREADS)
        got_s = r.trace(sh, any_hit=True)
        assert np.array_equal(got_s["hitFace"], want_s["hitFace"]) and np.array_equal(got_s["leaf"], want_s["leaf"])
        assert np.array_equal(got_s["t"].view(np.uint32), want_s["t"].view(np.uint32))
    finally:
        r.close()
