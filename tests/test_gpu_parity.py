"""GPU: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star / SURVEY.md 8d):
  - nearest-hit face index, leaf index: bit-exact.  t and the visit counters: bit-exact too, because
    both sides follow include/pbr_pinned_math.h.
  - radiance: the stated tolerance is a mean relative error <= 2 %; under the pinned arithmetic the
    frame is in fact bit-identical, which is what is asserted (NaN == NaN), with the MRE reported.
"""
import numpy as np
import pytest

import helpers as Hh

pytestmark = pytest.mark.gpu

RADIANCE_MRE_TOLERANCE = 0.02


@pytest.fixture(scope="module")
def suzanne(oracle):
    return oracle.load_obj(Hh.model_path("suzanne.obj"), 1)


def test_pinned_math_device_equals_host(device, oracle):
    L = oracle.lib()
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-100, 100, 20000), rng.uniform(-1, 1, 20000),
                         [0.0, 1.0, -1.0, 0.0333, 1e-20, 1e20, np.inf, -np.inf, np.nan]]).astype(np.float32)
    for op, name in enumerate(["sin", "cos", "tan", "acos", "atan"]):
        host = np.array([getattr(L, "oracle_pm_" + name)(float(x)) for x in xs], np.float32)
        dev = device.pinnedMath(op, xs)
        assert Hh.images_equal(host, dev), name
    ys = np.concatenate([rng.uniform(0, 300, 20000), rng.uniform(0, 200000, 20000), [0, 1, 2, 0.5, 3, -1, 1e5, 0, 2]]).astype(np.float32)
    xb = np.abs(xs)
    xb[:20000] = rng.uniform(0, 1, 20000)
    host = np.array([L.oracle_pm_pow(float(a), float(b)) for a, b in zip(xb, ys)], np.float32)
    assert Hh.images_equal(host, device.pinnedMath(5, xb, ys))
    host = np.array([L.oracle_pm_cbrt(float(x)) for x in xs], np.float32)
    assert Hh.images_equal(host, device.pinnedMath(6, xs))
    # one rand() step from a given seed
    import ctypes
    seeds = rng.uniform(0, 5000, 5000).astype(np.float32)
    host = np.array([L.oracle_rand(ctypes.byref(ctypes.c_float(float(s)))) for s in seeds], np.float32)
    assert np.array_equal(host, device.pinnedMath(7, seeds))


def _assert_hits_equal(a, b):
    assert np.array_equal(a["hitFace"], b["hitFace"])
    assert np.array_equal(a["leaf"], b["leaf"])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    assert np.array_equal(a["visits"], b["visits"])


def test_trace_parity_suzanne(device, suzanne):
    p = Hh.Prepared(suzanne, 64, 64, shadow_rays=1)
    ds = Hh.DeviceScene(device, p)
    rays = np.concatenate([Hh.primary_rays(p, 200, 200), Hh.random_rays(50000, 3, -1.0, 2.5)])
    want, _ = p.oracle_trace(rays)
    got = ds.trace(rays)
    _assert_hits_equal(got, want)
    srays = Hh.shadow_rays_from_hits(rays, want, (0.1, 1.3, 1.2))
    want_s, _ = p.oracle_trace(srays, any_hit=True)
    got_s = ds.trace(srays, any_hit=True)
    _assert_hits_equal(got_s, want_s)
    assert (np.isfinite(want["t"])).mean() > 0.5


@pytest.mark.parametrize("ntri,skip_ahead,max_faces", [(20000, True, 2), (20000, False, 1), (200000, True, 2)])
def test_trace_parity_soup(device, oracle, ntri, skip_ahead, max_faces):
    import pbr_b200
    s = pbr_b200.scenes.soup(ntri, seed=99)
    p = Hh.Prepared(s, 64, 64, eye=(0.0, 0.0, 3.5), bvh_kwargs=dict(skip_ahead=skip_ahead, max_faces=max_faces))
    ds = Hh.DeviceScene(device, p)
    rays = np.concatenate([Hh.primary_rays(p, 256, 256), Hh.random_rays(100000, 4, -1.0, 1.0)])
    want, _ = p.oracle_trace(rays)
    got = ds.trace(rays)
    _assert_hits_equal(got, want)
    srays = Hh.shadow_rays_from_hits(rays, want, (0.0, 3.0, 0.0))
    want_s, _ = p.oracle_trace(srays, any_hit=True)
    _assert_hits_equal(ds.trace(srays, any_hit=True), want_s)


def test_trace_edge_cases(device, suzanne):
    p = Hh.Prepared(suzanne, 64, 64)
    ds = Hh.DeviceScene(device, p)
    # empty batch
    assert ds.trace(np.zeros((0, 8), np.float32)).shape == (0,)
    # one ray, axis-aligned direction (zero components -> inf in invDir, NaN dropped by fmin/fmax),
    # and rays that start on geometry
    rays = np.zeros((5, 8), np.float32)
    rays[:, 7] = np.inf
    rays[0, 0:3] = (0, 1, 3); rays[0, 4:7] = (0, 0, -1)
    rays[1, 0:3] = (0, 1, 3); rays[1, 4:7] = (0, -1, 0)
    rays[2, 0:3] = (0, 1, 3); rays[2, 4:7] = (1, 0, 0)
    rays[3, 0:3] = (0, 5, 0); rays[3, 4:7] = (0, 1, 0)       # leaves the scene
    rays[4, 0:3] = (0, 0, 0); rays[4, 4:7] = (0, 1, 0)       # starts on the floor plane
    want, _ = p.oracle_trace(rays, nthreads=1)
    _assert_hits_equal(ds.trace(rays), want)
    # ragged sizes around the 32-ray warp batches
    big = Hh.random_rays(1000, 8, -1.0, 2.5)
    for n in (1, 31, 32, 33, 127, 999):
        want, _ = p.oracle_trace(big[:n], nthreads=1)
        _assert_hits_equal(ds.trace(big[:n]), want)


def _render_both(device, prep, frames, pipeline=0):
    ds = Hh.DeviceScene(device, prep)
    device.setPipeline(pipeline)
    device.stats(reset=True)
    try:
        got, gdbg = ds.frames(frames)
        gstats = device.stats(reset=True)
    finally:
        device.setPipeline(-1)
    want, wdbg, wstats = prep.oracle_frames(frames)
    return got, gdbg, gstats, want, wdbg, wstats


@pytest.mark.parametrize("brdf", [1, 0])
@pytest.mark.parametrize("pipeline", [0, 1])
def test_render_parity_suzanne(device, suzanne, brdf, pipeline):
    p = Hh.Prepared(suzanne, 128, 96, brdf=brdf, max_depth=4)
    got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 4, pipeline)
    mre = Hh.mean_relative_error(got, want)
    assert mre <= RADIANCE_MRE_TOLERANCE
    assert Hh.images_equal(got, want), "radiance not bit-identical (MRE %.3g)" % mre
    assert Hh.images_equal(gdbg, wdbg)
    assert np.array_equal(gstats, wstats)
    assert np.isfinite(got[..., :3]).mean() > 0.99 and got[..., :3][np.isfinite(got[..., :3])].mean() > 0.05


@pytest.mark.parametrize("brdf", [1, 0])
@pytest.mark.parametrize("shadow_stage", [1, 0])
def test_render_parity_shadow_rays(device, suzanne, brdf, shadow_stage):
    """Shadow rays walked by the traversal engine as their own wavefront stage (default) or inside the shade
    kernel (shadow_stage 0): same pixels, same counters."""
    p = Hh.Prepared(suzanne, 96, 96, brdf=brdf, shadow_rays=1, max_depth=3)
    assert p.num_lights == 1
    device.setTuning("shadow_stage", shadow_stage)
    try:
        got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 3)
    finally:
        device.setTuning("shadow_stage", 1)
    assert Hh.mean_relative_error(got, want) <= RADIANCE_MRE_TOLERANCE
    assert Hh.images_equal(got, want)
    assert Hh.images_equal(gdbg, wdbg)
    assert np.array_equal(gstats, wstats)
    assert int(gstats[1]) > 0


def test_render_parity_multisample_and_dof(device, suzanne):
    p = Hh.Prepared(suzanne, 64, 64, samples=4, max_depth=3, focus_point=(32, 20))
    got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 3)
    assert Hh.images_equal(got, want)
    assert np.array_equal(gstats, wstats)


@pytest.mark.parametrize("pipeline", [0, 1])
@pytest.mark.parametrize("brdf,shadow", [(1, 0), (0, 1)])
def test_render_parity_phong_tessellation(device, suzanne, pipeline, brdf, shadow):
    """render.phong_tessellation > 0: Ogaki-Tokuyoshi direct ray tracing of Phong tessellation
    (pt_phongtess.cl) for faces with differing vertex normals, flat test for the others."""
    p = Hh.Prepared(suzanne, 96, 80, brdf=brdf, shadow_rays=shadow, max_depth=3, phong_tessellation=0.7)
    assert int(p.defines["phongtess"][0]) == 1
    got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 2, pipeline)
    assert Hh.mean_relative_error(got, want) <= RADIANCE_MRE_TOLERANCE
    assert Hh.images_equal(got, want)
    assert Hh.images_equal(gdbg, wdbg)
    assert np.array_equal(gstats, wstats)
    # the tessellated picture differs from the flat one (Suzanne has smooth normals)
    flat = Hh.Prepared(suzanne, 96, 80, brdf=brdf, shadow_rays=shadow, max_depth=3)
    fimg, _, _ = flat.oracle_frames(2)
    assert not Hh.images_equal(fimg, want)


@pytest.mark.parametrize("name", sorted(__import__("ref_configs").CASES))
def test_device_equals_reference_kernel(device, name):
    """The CUDA path against the reference's OWN kernel source compiled for the host (oracle/_ref, built by
    oracle/build_ref.py where /root/reference exists): image and debug image, bit for bit."""
    import ref_configs
    from oracle import ref as R
    from oracle import scene as S
    p = ref_configs.prepared(name)
    if not R.available(p.defines):
        pytest.skip("oracle/_ref not built for this configuration")
    ds = Hh.DeviceScene(device, p)
    got, gdbg = ds.frames(3)
    img = np.zeros((p.H, p.W, 4), np.float32)
    for k in range(3):
        img, dbg = R.path_tracing(p.defines, S.frame_seed(k), S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV,
                                  p.facesN, p.vertices4, p.normals4, p.materials, p.lights, img, nthreads=8)
    assert Hh.images_equal(got, img)
    assert Hh.images_equal(gdbg, dbg)


def _batch_both(device, prep, frames, pipeline):
    ds = Hh.DeviceScene(device, prep)
    device.setPipeline(pipeline)
    device.stats(reset=True)
    try:
        got, gdbg = ds.frames_batch(frames)
        gstats = device.stats(reset=True)
    finally:
        device.setPipeline(-1)
    want, wdbg, wstats = prep.oracle_frames(frames)
    return got, gdbg, gstats, want, wdbg, wstats


@pytest.mark.parametrize("pipeline", [0, 1])
@pytest.mark.parametrize("brdf,shadow,samples", [(1, 0, 1), (0, 1, 2)])
def test_render_parity_batched_frames(device, suzanne, pipeline, brdf, shadow, samples):
    """pbr_kernel_launch_batch: the pixels stay the reference's."""
    p = Hh.Prepared(suzanne, 112, 80, brdf=brdf, shadow_rays=shadow, samples=samples, max_depth=4)
    got, gdbg, gstats, want, wdbg, wstats = _batch_both(device, p, 6, pipeline)
    assert Hh.mean_relative_error(got, want) <= RADIANCE_MRE_TOLERANCE
    assert Hh.images_equal(got, want)
    assert Hh.images_equal(gdbg, wdbg)
    assert np.array_equal(gstats, wstats)


def test_render_batched_frames_depth_of_field_and_long_batches(device, suzanne):
    # depth of field couples pixels across frames: the batch call must fall back to frame-by-frame
    p = Hh.Prepared(suzanne, 64, 64, samples=2, max_depth=3, focus_point=(32, 20))
    got, gdbg, gstats, want, wdbg, wstats = _batch_both(device, p, 5, 0)
    assert Hh.images_equal(got, want)
    assert np.array_equal(gstats, wstats)
    # a long batch, continuing an accumulated image
    p = Hh.Prepared(suzanne, 48, 32, max_depth=3)
    ds = Hh.DeviceScene(device, p)
    start, _ = ds.frames(2)
    seq, _ = ds.frames(35, image=start.copy(), first=2, host_roundtrip=False)
    bat, _ = ds.frames_batch(35, image=start.copy(), first=2)
    assert Hh.images_equal(seq, bat)


def test_render_parity_soup(device, oracle):
    import pbr_b200
    s = pbr_b200.scenes.soup(50000, seed=5)
    p = Hh.Prepared(s, 160, 96, eye=(0.0, 0.0, 3.5))
    got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 2)
    assert Hh.images_equal(got, want)
    assert Hh.images_equal(gdbg, wdbg)
    assert np.array_equal(gstats, wstats)


def test_non_multiple_image_size_and_tiles(device, suzanne):
    """Sizes that are not multiples of the 8x4 block; tile rows stitched == full frame."""
    p = Hh.Prepared(suzanne, 50, 37, max_depth=3)
    got, gdbg, gstats, want, wdbg, wstats = _render_both(device, p, 2)
    assert Hh.images_equal(got, want)
    p2 = Hh.Prepared(suzanne, 64, 48, max_depth=3)
    ds = Hh.DeviceScene(device, p2)
    full, _ = ds.frames(1)
    stitched = np.zeros_like(full)
    for (y0, y1) in ((0, 12), (12, 20), (20, 48)):
        device.setTile(y0, y1)
        part, _ = ds.frames(1)
        stitched[y0:y1] = part[y0:y1]
    device.setTile(-1, -1)
    assert Hh.images_equal(stitched, full)
    # interleaved stripes (pbr_set_tile_stripes): 3 "ranks", stripes of 4 and of 2 rows, every pipeline
    for stripe, pipeline in ((4, 0), (2, 0), (8, 1)):
        world = 48 // (stripe * (4 if stripe < 8 else 2))
        world = max(world, 2)
        if 48 % (stripe * world):
            continue
        stitched = np.zeros_like(full)
        device.setPipeline(pipeline)
        try:
            for rank in range(world):
                device.setTileStripes(stripe, world, rank)
                part, _ = ds.frames(1)
                mine = ((np.arange(48) // stripe) % world) == rank
                stitched[mine] = part[mine]
        finally:
            device.setTileStripes(0)
            device.setPipeline(-1)
        assert Hh.images_equal(stitched, full), "stripes of %d rows, pipeline %d" % (stripe, pipeline)
    device.setTileStripes(5, 2, 0)
    with pytest.raises(Exception):
        ds.frames(1)                                   # 48 is not a multiple of 5 * 2
    device.setTileStripes(0)


def test_device_resident_accumulation_equals_host_roundtrip(device, suzanne):
    """imageIn <- imageOut on the device (pbr_image_copy) gives the same frames as the reference's
    read-back / re-upload (PathTracer.cpp:61-66)."""
    p = Hh.Prepared(suzanne, 64, 64)
    ds = Hh.DeviceScene(device, p)
    a, _ = ds.frames(5, host_roundtrip=True)
    b, _ = ds.frames(5, host_roundtrip=False)
    assert Hh.images_equal(a, b)


def test_error_reporting(device, suzanne):
    import pbr_b200
    p = Hh.Prepared(suzanne, 32, 32)
    with pytest.raises(pbr_b200.PbrError):
        device.createKernel("noise_filtering")           # CL_INVALID_KERNEL_NAME
    with pytest.raises(pbr_b200.PbrError):
        device.setReplacement("#NOT_A_DEFINE#", "1")
    with pytest.raises(pbr_b200.PbrError):
        device.readImageOutput(12345678, 32, 32)          # dead handle
    bad = p.defines.copy()
    bad["brdf"] = 7
    with pytest.raises(pbr_b200.PbrError):
        device.loadProgram(bad)


# ---------------------------------------------------------------------------------------------------------
# The ordered walk (pbr_set_traversal(1), csrc/pt_wide.cuh): same hit face, leaf and t bits, same image bits.
# The visit counters and the debug image are the ordered walk's own and are not compared.

def _ordered(device):
    class _Ctx:
        def __enter__(self_):
            device.setTraversal(1)
            device.traversalInfo(reset=True)

        def __exit__(self_, *a):
            device.setTraversal(-1)
    return _Ctx()


@pytest.mark.parametrize("brdf,shadow,samples", [(1, 0, 1), (0, 0, 1), (1, 1, 1), (0, 1, 2)])
def test_ordered_walk_frames_suzanne(device, suzanne, brdf, shadow, samples):
    p = Hh.Prepared(suzanne, 128, 96, brdf=brdf, shadow_rays=shadow, samples=samples, max_depth=4)
    ds = Hh.DeviceScene(device, p)
    device.setPipeline(0)
    try:
        with _ordered(device):
            got, _ = ds.frames(3)
            info = device.traversalInfo()
    finally:
        device.setPipeline(-1)
    want, _, wstats = p.oracle_frames(3)
    assert info["last_used"] == 1 and info["wide_available"] == 1
    assert info["ordered_rays"] == int(wstats[0]) + int(wstats[1])          # every ray of the frames took the ordered walk
    assert Hh.mean_relative_error(got, want) <= RADIANCE_MRE_TOLERANCE
    assert Hh.images_equal(got, want)


def test_ordered_walk_is_automatic_without_debug_image(device, suzanne):
    p = Hh.Prepared(suzanne, 96, 64, max_depth=4)
    ds = Hh.DeviceScene(device, p)
    want, wdbg, _ = p.oracle_frames(2)
    device.setPipeline(0)
    try:
        got, gdbg = ds.frames(2)                                    # debug image on (the default of a pbr_ctx)
        assert device.traversalInfo()["last_used"] == 0
        assert Hh.images_equal(got, want) and Hh.images_equal(gdbg, wdbg)
        device.setDebugImage(False)
        got, _ = ds.frames(2)
        assert device.traversalInfo()["last_used"] == 1
        assert Hh.images_equal(got, want)
        device.setTraversal(0)                                      # forced reference order
        got, _ = ds.frames(2)
        assert device.traversalInfo()["last_used"] == 0
        assert Hh.images_equal(got, want)
    finally:
        device.setTraversal(-1)
        device.setDebugImage(True)
        device.setPipeline(-1)


@pytest.mark.parametrize("kw", [dict(), dict(skip_ahead=False), dict(max_faces=1)])
def test_ordered_walk_explicit_rays(device, oracle, kw):
    import pbr_b200
    s = pbr_b200.scenes.soup(60000, seed=31)
    p = Hh.Prepared(s, 64, 64, eye=(0.0, 0.0, 3.5), bvh_kwargs=kw)
    ds = Hh.DeviceScene(device, p)
    rays = np.concatenate([Hh.primary_rays(p, 200, 120), Hh.random_rays(30001, 6, -1.0, 1.0)])
    rng = np.random.default_rng(3)
    rays[-5000:, 7] = rng.uniform(0.0, 2.0, 5000).astype(np.float32)          # finite initial ray.t
    want, _ = p.oracle_trace(rays)
    with _ordered(device):
        got = ds.trace(rays)
        info = device.traversalInfo()
    assert info["last_used"] == 1 and info["ordered_rays"] == len(rays)
    assert np.array_equal(got["hitFace"], want["hitFace"]) and np.array_equal(got["leaf"], want["leaf"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    # any-hit rays keep the reference-order engine even when the ordered walk is forced
    hits = want
    sh = Hh.shadow_rays_from_hits(rays, hits, (0.0, 3.0, 0.0))
    want_s, _ = p.oracle_trace(sh, any_hit=True)
    with _ordered(device):
        got_s = ds.trace(sh, any_hit=True)
        assert device.traversalInfo()["last_used"] == 0
    assert np.array_equal(got_s["hitFace"], want_s["hitFace"]) and np.array_equal(got_s["visits"], want_s["visits"])


def test_ordered_walk_flat_coplanar_geometry_rewalks(device, oracle):
    """Zero-thickness leaf boxes, coplanar overlapping layers: hits in front of their own leaf box, ties between leaves.
    The ambiguity rule sends those rays through the reference-order walk; every hit is still the reference's."""
    import test_wide_walk as T
    p = Hh.Prepared(T.flat_floor_scene(), 160, 96, eye=(0.3, 1.5, 3.0), center=(0.0, 0.2, 1.0), max_depth=4)
    ds = Hh.DeviceScene(device, p)
    rays = T.rays_for(p, 60000, 11, -1.9, 1.9, grid=(320, 180))
    rays[-30000:, 1] = np.abs(rays[-30000:, 1]) + 0.01
    want, _ = p.oracle_trace(rays)
    with _ordered(device):
        got = ds.trace(rays)
        info = device.traversalInfo()
    assert info["rewalked_rays"] > 0
    assert np.array_equal(got["hitFace"], want["hitFace"]) and np.array_equal(got["leaf"], want["leaf"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    device.setPipeline(0)
    try:
        with _ordered(device):
            img, _ = ds.frames(2)
    finally:
        device.setPipeline(-1)
    ref, _, _ = p.oracle_frames(2)
    assert Hh.images_equal(img, ref)


def test_ordered_walk_refused_for_foreign_node_arrays(device, suzanne):
    """A node array the ordered walk cannot honour: automatic mode keeps the reference-order walk, forcing it fails loudly."""
    import pbr_b200
    p = Hh.Prepared(suzanne, 64, 48, max_depth=3)
    leaf = np.where(p.nodes[1:, 3] >= 0.0)[0] + 1
    p.nodes = p.nodes.copy()
    p.nodes[leaf[5], 0] -= 10.0                    # a leaf box sticking out of its parent's
    ds = Hh.DeviceScene(device, p)
    want, _, _ = p.oracle_frames(1)
    device.setPipeline(0)
    device.setDebugImage(False)
    try:
        got, _ = ds.frames(1)
        info = device.traversalInfo()
        assert info["last_used"] == 0 and info["wide_available"] == 0 and "not inside" in info["why_not"]
        assert Hh.images_equal(got, want)
        device.setTraversal(1)
        with pytest.raises(pbr_b200.capi.PbrError):
            ds.frames(1)
    finally:
        device.setTraversal(-1)
        device.setDebugImage(True)
        device.setPipeline(-1)


def test_ordered_walk_top_of_tree_budgets(device, oracle):
    """The number of nodes staged in shared memory renumbers the tree and changes nothing else."""
    import pbr_b200
    s = pbr_b200.scenes.soup(40000, seed=8)
    p = Hh.Prepared(s, 64, 64, eye=(0.0, 0.0, 3.5))
    ds = Hh.DeviceScene(device, p)
    rays = np.concatenate([Hh.primary_rays(p, 160, 90), Hh.random_rays(20000, 2, -1.0, 1.0)])
    want, _ = p.oracle_trace(rays)
    try:
        for top in (1, 5, 21, 85, 341, 1365):
            device.setTuning("wide_top", top)
            with _ordered(device):
                got = ds.trace(rays)
                assert device.traversalInfo()["wide_top"] == min(top, device.traversalInfo()["wide_nodes"])
            assert np.array_equal(got["hitFace"], want["hitFace"]) and np.array_equal(got["leaf"], want["leaf"]), top
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), top
    finally:
        device.setTuning("wide_top", 21)


# ---------------------------------------------------------------------------------------------------------
# Several frames in flight (pbr_kernel_launch_batch, tuning "frames_in_flight"): consecutive frames traced on
# streams of their own, mixed into the image in order -- the same bits as one frame after the other.

@pytest.mark.parametrize("brdf,shadow,samples", [(1, 0, 1), (0, 1, 2)])
@pytest.mark.parametrize("traversal", [0, 1])
def test_frames_in_flight_same_bits(device, suzanne, brdf, shadow, samples, traversal):
    p = Hh.Prepared(suzanne, 136, 88, brdf=brdf, shadow_rays=shadow, samples=samples, max_depth=4)
    ds = Hh.DeviceScene(device, p)
    want, _, wstats = p.oracle_frames(7)
    device.setDebugImage(False)
    device.setTraversal(traversal)
    try:
        for n in (1, 2, 3, 4):
            device.setTuning("frames_in_flight", n)
            device.stats(reset=True)
            got, _ = ds.frames_batch(7)
            st = device.stats(reset=True)
            assert Hh.images_equal(got, want), "frames in flight: %d" % n
            assert int(st[0]) + int(st[1]) == int(wstats[0]) + int(wstats[1])
            # ... and continuing an accumulated image, with a tile set
            start, _ = ds.frames_batch(3)
            device.setTileStripes(4, 2, 1)
            part, _ = ds.frames_batch(4, image=start.copy(), first=3)
            device.setTileStripes(0)
            rows = [y for y in range(p.H) if (y // 4) % 2 == 1]
            assert Hh.images_equal(part[rows], want[rows]), "frames in flight: %d, stripes" % n
    finally:
        device.setTuning("frames_in_flight", 4)
        device.setTraversal(-1)
        device.setDebugImage(True)


# ---------------------------------------------------------------------------------------------------------
# Node arrays the reference's host never writes but its kernel gives a meaning to, and the
# bookkeeping around scene buffers.

def test_foreign_node_words_follow_the_reference_kernel(device, suzanne):
    """bbMin.w == -2 (traverseShadows' "skip the next left child", pt_bvh.cl:157-159) and bbMin.w strictly between -1
    and 0 (neither inner nor leaf for the reference: the walk steps over the node) behave as in the reference kernel."""
    p = Hh.Prepared(suzanne, 64, 48, max_depth=3, shadow_rays=1, brdf=0)
    inner = np.where(p.nodes[1:, 3] <= -1.0)[0] + 1
    leaf = np.where(p.nodes[1:, 3] >= 0.0)[0] + 1
    p.nodes = p.nodes.copy()
    p.nodes[inner[5::7], 3] = -2.0
    p.nodes[leaf[3::11], 3] = -0.5
    ds = Hh.DeviceScene(device, p)
    rays = np.concatenate([Hh.primary_rays(p, 96, 64), Hh.random_rays(20000, 12, -2.0, 2.0)])
    want, _ = p.oracle_trace(rays)
    got = ds.trace(rays)
    for f in ("hitFace", "leaf", "visits"):
        assert np.array_equal(got[f], want[f]), f
    sh = Hh.shadow_rays_from_hits(rays, want, (0.1, 1.3, 1.2))
    want_s, _ = p.oracle_trace(sh, any_hit=True)
    got_s = ds.trace(sh, any_hit=True)
    assert np.array_equal(got_s["hitFace"], want_s["hitFace"]) and np.array_equal(got_s["visits"], want_s["visits"])
    assert np.array_equal(got_s["t"].view(np.uint32), want_s["t"].view(np.uint32))
    img, dbg = ds.frames(2)
    wimg, wdbg, _ = p.oracle_frames(2)
    assert Hh.images_equal(img, wimg) and Hh.images_equal(dbg, wdbg)
    # such an array is not one the ordered walk takes: the automatic choice stays with the reference order
    device.setDebugImage(False)
    try:
        img2, _ = ds.frames(2)
        info = device.traversalInfo()
    finally:
        device.setDebugImage(True)
    assert info["last_used"] == 0 and "skip flag" in info["why_not"]
    assert Hh.images_equal(img2, wimg)


def test_max_depth_zero_is_rejected(device, suzanne):
    import pbr_b200
    p = Hh.Prepared(suzanne, 32, 32, max_depth=3)
    p.defines = p.defines.copy()
    p.defines["max_depth"] = 0
    with pytest.raises(pbr_b200.capi.PbrError):
        Hh.DeviceScene(device, p)


def test_updating_the_lights_does_not_rebuild_the_scene(device, oracle):
    """Per-buffer epochs: only what was built from the buffer that changed is rebuilt."""
    scene = oracle.load_obj(Hh.model_path("suzanne.obj"), 1)
    p = Hh.Prepared(scene, 64, 48, max_depth=3, shadow_rays=1)
    ds = Hh.DeviceScene(device, p)
    device.setPipeline(0)
    try:
        ds.frames(1)
        device.profileRead(reset=True)
        ds.frames(1)
        steady = device.profileRead(reset=True)["other_launches"]
        lights = p.lights.copy()
        lights[0, 0] += 0.25                                        # move the light
        device.updateBuffer(ds.bufLights, lights)
        got, _ = ds.frames(1)
        after = device.profileRead(reset=True)["other_launches"]
        assert after == steady, "a lights-only update re-ran the scene repack"
        p.lights = lights
        want, _, _ = p.oracle_frames(1)
        assert Hh.images_equal(got, want)
        nodes = p.nodes.copy()
        device.updateBuffer(ds.bufBVH, nodes)                       # the same bytes, but the library cannot know: rebuilt
        ds.frames(1)
        assert device.profileRead(reset=True)["other_launches"] > steady
    finally:
        device.setPipeline(-1)
