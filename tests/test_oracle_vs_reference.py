"""Pins the oracle: oracle/pt_oracle.cpp (the restatement every GPU parity test is checked against) versus
the reference's OWN kernel source, compiled for the host by oracle/build_ref.py (oracle/_ref/).  Same scene,
same seeds, same arithmetic contract -> the two must agree bit for bit: image, debug image, and for explicit
rays t / hitFace / node visits / face tests."""
import numpy as np
import pytest

import helpers as Hh
import ref_configs
from oracle import oracle as O
from oracle import ref as R
from oracle import scene as S


def _need_ref(prep):
    if not R.available(prep.defines):
        pytest.skip("reference kernel for this configuration neither prebuilt nor buildable (no /root/reference)")


@pytest.mark.parametrize("name", sorted(ref_configs.CASES))
def test_restatement_equals_reference_kernel(name):
    p = ref_configs.prepared(name)
    _need_ref(p)
    img_o = np.zeros((p.H, p.W, 4), np.float32)
    img_r = img_o.copy()
    for k in range(3):
        args = (p.defines, S.frame_seed(k), S.pixel_weight(k), p.px_dim, p.camera, p.nodes, p.facesV, p.facesN,
                p.vertices4, p.normals4, p.materials, p.lights)
        img_o, dbg_o, _ = O.path_tracing(*args, img_o, nthreads=4)
        img_r, dbg_r = R.path_tracing(*args, img_r, nthreads=4)
        assert Hh.images_equal(img_o, img_r), "frame %d: radiance differs from the reference kernel" % k
        assert Hh.images_equal(dbg_o, dbg_r), "frame %d: visit counters differ from the reference kernel" % k
    rgb = img_r[..., :3]
    # a picture, not zeros (NaN pixels are the reference's BRDF arithmetic: they must match too, and do)
    assert np.isfinite(rgb).mean() > 0.6 and rgb[np.isfinite(rgb)].mean() > 0.02


def test_reference_kernel_rows_are_independent():
    """The host driver around the reference kernel may run any block of rows (tile sharding)."""
    p = ref_configs.prepared("suzanne_sa")
    _need_ref(p)
    args = (p.defines, S.frame_seed(0), S.pixel_weight(0), p.px_dim, p.camera, p.nodes, p.facesV, p.facesN,
            p.vertices4, p.normals4, p.materials, p.lights)
    zero = np.zeros((p.H, p.W, 4), np.float32)
    full, _ = R.path_tracing(*args, zero, nthreads=2)
    part, _ = R.path_tracing(*args, zero, y0=16, y1=40, nthreads=2)
    assert Hh.images_equal(full[16:40], part[16:40])
    assert not part[:16].any() and not part[40:].any()


@pytest.mark.parametrize("name", ["suzanne_sa", "suzanne_sa_shadow", "soup_sa", "pillars_sa"])
def test_explicit_rays_equal_reference_traverse(name):
    p = ref_configs.prepared(name)
    _need_ref(p)
    rays = np.concatenate([Hh.primary_rays(p, 48, 32), Hh.random_rays(3000, 11, -1.5, 1.5)])
    want_t, want_face, want_nodes, want_tris = R.trace(p.defines, p.nodes, p.facesV, p.facesN, p.vertices4, p.normals4,
                                                       p.lights, rays, nthreads=4)
    got, _ = p.oracle_trace(rays, nthreads=4)
    assert np.array_equal(got["t"].view(np.uint32), want_t.view(np.uint32))
    assert np.array_equal(got["hitFace"], want_face)
    assert np.array_equal(got["visits"] & 0xfffff, want_nodes.astype(np.uint32))
    assert np.array_equal(got["visits"] >> 20, want_tris.astype(np.uint32))
    assert (want_face > 0).sum() > 100

    # shadow rays: from the primary hits towards a point above the scene, t = distance (traverseShadows)
    sh = Hh.shadow_rays_from_hits(rays, got, (0.5, 4.0, 1.0))
    want_t, want_face, want_nodes, want_tris = R.trace(p.defines, p.nodes, p.facesV, p.facesN, p.vertices4, p.normals4,
                                                       p.lights, sh, any_hit=True, nthreads=4)
    got, _ = p.oracle_trace(sh, any_hit=True, nthreads=4)
    assert np.array_equal(got["t"].view(np.uint32), want_t.view(np.uint32))
    assert np.array_equal(got["hitFace"], want_face)
    assert not want_nodes.any()        # traverseShadows does not count node visits (pt_bvh.cl:133-177); ours does
    assert np.array_equal(got["visits"] >> 20, want_tris.astype(np.uint32))
    assert (want_t < sh[:, 7]).sum() > 50 and (want_t >= sh[:, 7]).sum() > 50      # occluded and unoccluded
