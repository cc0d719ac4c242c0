"""CPU: the product's wide-BVH builder (csrc/wide_bvh.h) and the exactness argument of the ordered walk, through the
CPU model in tests/wide_walk_model.cpp (test infrastructure): for every ray the ordered walk -- nearest child first,
margin pruning, ambiguous rays re-walked in reference order -- returns the (t bits, face, leaf) of the reference-order
walk, which is itself checked against the oracle here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as Hh

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "wide_walk_model.cpp")
HDR = os.path.join(ROOT, "physically-based-rendering_b200", "csrc", "wide_bvh.h")
SO = os.path.join(HERE, "_wide_walk_model.so")

STAT_NAMES = ("wide_nodes", "wide_depth", "top_count", "leaf_refs", "inner_refs", "strict_nodes", "strict_tris",
              "wide_visits", "fast_tris", "fallbacks", "overflows", "insane_winners", "max_stack", "mismatches")


def model():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", SRC, "-o", SO])
    lib = C.CDLL(SO)
    vp, ll, i32 = C.c_void_p, C.c_longlong, C.c_int
    lib.wide_model_run.argtypes = [vp, i32, vp, i32, vp, vp, ll, i32, i32, vp, vp, vp, C.c_char_p, i32]
    lib.wide_model_run.restype = i32
    return lib


def run_model(prep, rays, stack_cap=32, top_budget=85, nodes=None):
    lib = model()
    nodes = np.ascontiguousarray(prep.nodes if nodes is None else nodes, np.float32)
    fv = np.ascontiguousarray(prep.facesV, np.uint32)
    v4 = np.ascontiguousarray(prep.vertices4, np.float32)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]
    strict = np.zeros((n, 4), np.int32)
    fast = np.zeros((n, 4), np.int32)
    stats = np.zeros(len(STAT_NAMES), np.int64)
    msg = C.create_string_buffer(256)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.wide_model_run(p(nodes), nodes.shape[0], p(fv), fv.shape[0], p(v4), p(rays), n, stack_cap, top_budget,
                            p(strict), p(fast), p(stats), msg, 256)
    return rc, msg.value.decode(), strict, fast, dict(zip(STAT_NAMES, stats.tolist()))


def check_scene(prep, rays, **kw):
    rc, why, strict, fast, st = run_model(prep, rays, **kw)
    assert rc == 0, why
    want, _ = prep.oracle_trace(rays, nthreads=os.cpu_count() or 1)
    # the model's reference-order walk is the oracle's
    assert np.array_equal(strict[:, 0].view(np.uint32), want["t"].view(np.uint32))
    assert np.array_equal(strict[:, 1], want["hitFace"]) and np.array_equal(strict[:, 2], want["leaf"])
    # and the ordered walk returns the same hits
    assert st["mismatches"] == 0 and np.array_equal(strict[:, :3], fast[:, :3])
    return st


def flat_floor_scene(n=24, overlap=True):
    """Axis-aligned quads with zero-thickness leaf boxes (tNear == tFar == the hit distance) and, with `overlap`, a
    second coplanar layer and a slab exactly on top of it: hits within an ulp of a box's tNear, and ties between faces
    of different leaves -- the cases the ambiguity rule of the ordered walk exists for."""
    import pbr_b200
    from pbr_b200 import scenes
    M = scenes._Mesh()
    M.grid("floor", (-2.0, 0.0, -2.0), (4.0, 0, 0), (0, 0, 4.0), n, n, (0, 1, 0), 0)
    M.grid("wall", (-2.0, 0.0, -2.0), (4.0, 0, 0), (0, 3.0, 0), n, n, (0, 0, 1), 0)
    if overlap:
        M.grid("floor2", (-1.0, 0.0, -1.0), (2.0, 0, 0), (0, 0, 2.0), n // 2 + 1, n // 2 + 1, (0, 1, 0), 0)
        M.grid("slab_top", (-0.5, 0.25, -0.5), (1.0, 0, 0), (0, 0, 1.0), 5, 5, (0, 1, 0), 0)
        M.grid("slab_top_again", (-0.5, 0.25, -0.5), (1.0, 0, 0), (0, 0, 1.0), 3, 3, (0, 1, 0), 0)
    return M.finish([scenes.default_material("m")], None)


def rays_for(prep, n_random, seed, lo=-1.5, hi=1.5, grid=(160, 90)):
    prim = Hh.primary_rays(prep, *grid)
    return np.concatenate([prim, Hh.random_rays(n_random, seed, lo, hi)])


def test_model_suzanne(oracle):
    prep = Hh.Prepared(oracle.load_obj(Hh.model_path("suzanne.obj"), 0), 64, 64)
    st = check_scene(prep, rays_for(prep, 20000, 3, -2.0, 2.0))
    assert st["wide_nodes"] < prep.nodes.shape[0] and st["wide_visits"] < st["strict_nodes"]


@pytest.mark.parametrize("model_name", ["pillars.obj"])
def test_model_bundled(oracle, model_name):
    prep = Hh.Prepared(oracle.load_obj(Hh.model_path(model_name), 0), 64, 64)
    check_scene(prep, rays_for(prep, 20000, 5, -3.0, 3.0))


@pytest.mark.parametrize("kw", [dict(), dict(skip_ahead=False), dict(max_faces=1), dict(skip_ahead_compare=0.3)])
def test_model_soup_builder_settings(oracle, kw):
    import pbr_b200
    scene = pbr_b200.scenes.soup(30000, seed=17)
    prep = Hh.Prepared(scene, 64, 64, eye=(0.0, 0.0, 3.5), bvh_kwargs=kw)
    st = check_scene(prep, rays_for(prep, 30000, 9, -1.0, 1.0))
    assert st["wide_visits"] * 3 < st["strict_nodes"]          # the point of the exercise


def test_model_flat_and_coplanar_geometry(oracle):
    prep = Hh.Prepared(flat_floor_scene(), 64, 64, eye=(0.3, 1.5, 3.0), center=(0.0, 0.2, 1.0))
    rays = rays_for(prep, 60000, 11, -1.9, 1.9, grid=(320, 180))
    rays[-30000:, 1] = np.abs(rays[-30000:, 1]) + 0.01            # origins above the floor
    st = check_scene(prep, rays)
    assert st["insane_winners"] > 0, "the scene was meant to produce hits in front of their own leaf box"


def test_model_small_stack_falls_back(oracle):
    import pbr_b200
    scene = pbr_b200.scenes.soup(20000, seed=23)
    prep = Hh.Prepared(scene, 64, 64, eye=(0.0, 0.0, 3.5))
    st = check_scene(prep, rays_for(prep, 5000, 2, -1.0, 1.0), stack_cap=2)
    assert st["overflows"] > 0 and st["fallbacks"] >= st["overflows"]


def test_model_initial_t(oracle):
    """Rays that start with a finite ray.t (what pbr_trace accepts in dir.w)."""
    import pbr_b200
    scene = pbr_b200.scenes.soup(20000, seed=29)
    prep = Hh.Prepared(scene, 64, 64, eye=(0.0, 0.0, 3.5))
    rays = rays_for(prep, 20000, 4, -1.0, 1.0)
    rng = np.random.default_rng(1)
    rays[:, 7] = rng.uniform(0.0, 3.0, len(rays)).astype(np.float32)
    check_scene(prep, rays)


def test_builder_refuses_what_the_ordered_walk_cannot_honour(oracle):
    prep = Hh.Prepared(oracle.load_obj(Hh.model_path("suzanne.obj"), 0), 64, 64)
    rays = Hh.random_rays(10, 1)
    inner = np.where(prep.nodes[1:, 3] <= -1.0)[0] + 1
    leaf = np.where(prep.nodes[1:, 3] >= 0.0)[0] + 1

    def refused(mutate, expect):
        nodes = prep.nodes.copy()
        mutate(nodes)
        rc, why, *_ = run_model(prep, rays, nodes=nodes)
        assert rc == 1 and expect in why, (rc, why)

    refused(lambda n: n.__setitem__((inner[3], 3), -2.0), "skip flag")
    refused(lambda n: n.__setitem__((leaf[5], 0), n[leaf[5], 0] - 10.0), "not inside")
    refused(lambda n: n.__setitem__((inner[2], 7), 1.0), "backwards")
    refused(lambda n: n.__setitem__((leaf[0], 3), 1.0e7), "out of range")
    two = leaf[prep.nodes[leaf, 7] != -1.0]
    refused(lambda n: n.__setitem__((two[0], 7), n[two[0], 3] + 2.0), "first + 1")
    refused(lambda n: n.__setitem__((leaf[7], 1), np.nan), "NaN")


def test_model_many_objects_and_tied_centres(oracle):
    """A grid in 64 objects (one tree per object under an object-level tree, centres tied everywhere) and the interior
    scene (flat walls and floors, 34 objects): the collapse of per-object trees and of skip-ahead chains."""
    import pbr_b200
    grid = pbr_b200.scenes.displaced_grid(60, 58, patches=8)
    prep = Hh.Prepared(grid, 64, 64, eye=(0.0, 1.2, 1.8), center=(0.0, 0.55, 1.0))
    rays = rays_for(prep, 20000, 21, -1.0, 1.0, grid=(200, 120))
    rays[-20000:, 1] = np.abs(rays[-20000:, 1]) * 0.5 + 0.3
    st = check_scene(prep, rays)
    assert st["wide_visits"] < st["strict_nodes"]
    interior = pbr_b200.scenes.interior(detail=0.15)
    prep = Hh.Prepared(interior, 64, 64, eye=(0.0, 1.4, 5.2), center=(0.0, 0.1, 1.0), bvh_kwargs=dict(skip_ahead_compare=0.2))
    rays = rays_for(prep, 30000, 22, -3.5, 3.5, grid=(200, 120))
    rays[-30000:, 1] = np.abs(rays[-30000:, 1]) * 0.8 + 0.05
    st = check_scene(prep, rays)
    assert st["insane_winners"] > 0            # flat floor / wall boxes: hits in front of their own leaf box do occur here


def test_model_quantised_child_boxes_keep_the_result(oracle):
    """Premise of the layout DESIGN.md §6 names as the next step (64-byte nodes: child boxes quantised conservatively,
    the exact leaf box tested at the leaf): the hits stay those of the reference-order walk; only visits are added.
    Flat, coplanar geometry included (hits within an ulp of their leaf box's tNear)."""
    import pbr_b200
    lib = model()
    vp, ll, i32 = C.c_void_p, C.c_longlong, C.c_int
    lib.wide_model_quant.argtypes = [vp, i32, vp, i32, vp, vp, ll, i32, i32, i32, vp]
    lib.wide_model_quant.restype = i32
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    soup = Hh.Prepared(pbr_b200.scenes.soup(20000, seed=31), 64, 64, eye=(0.0, 0.0, 3.5))
    flat = Hh.Prepared(flat_floor_scene(), 64, 64, eye=(0.3, 1.5, 3.0), center=(0.0, 0.2, 1.0))
    for prep, rays in ((soup, rays_for(soup, 10000, 6, -1.0, 1.0)), (flat, rays_for(flat, 20000, 12, -1.9, 1.9))):
        nodes = np.ascontiguousarray(prep.nodes, np.float32)
        fv = np.ascontiguousarray(prep.facesV, np.uint32)
        v4 = np.ascontiguousarray(prep.vertices4, np.float32)
        rays = np.ascontiguousarray(rays, np.float32)
        visits = []
        for bits, pow2 in ((0, 0), (8, 0), (8, 1), (4, 1)):
            st = np.zeros(6, np.int64)
            assert lib.wide_model_quant(p(nodes), nodes.shape[0], p(fv), fv.shape[0], p(v4), p(rays), rays.shape[0], 21, bits, pow2, p(st)) == 0
            assert st[5] == 0, (bits, pow2, st.tolist())
            visits.append(int(st[0]))
        assert visits[0] <= visits[1] <= visits[3]          # coarser boxes only add visits
        assert visits[1] < visits[0] * 1.1                  # and 8 bits add few
