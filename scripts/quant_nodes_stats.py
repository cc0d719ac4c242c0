"""CPU experiment for DESIGN.md §6 ("where the next factor is"): what the ordered walk would visit with 64-byte nodes --
child boxes quantised to 8 (or fewer) bits per plane relative to the node, the exact leaf box moved next to the leaf's
faces -- counted by the CPU model of tests/wide_walk_model.cpp, results checked against the reference-order walk.
    python scripts/quant_nodes_stats.py [c2] [c3] [c4small]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402
from pbr_b200 import host, scenes  # noqa: E402
import helpers as Hh  # noqa: E402
import test_wide_walk as T  # noqa: E402

which = sys.argv[1:] or ["c2", "c3", "c4small"]
cfg = host.Config()
cfg.reset()
lib = T.model()
vp, ll, i32 = C.c_void_p, C.c_longlong, C.c_int
lib.wide_model_quant.argtypes = [vp, i32, vp, i32, vp, vp, ll, i32, i32, i32, vp]
lib.wide_model_quant.restype = i32


def quant(prep, rays, bits, pow2):
    nodes = np.ascontiguousarray(prep.nodes, np.float32)
    fv = np.ascontiguousarray(prep.facesV, np.uint32)
    v4 = np.ascontiguousarray(prep.vertices4, np.float32)
    rays = np.ascontiguousarray(rays, np.float32)
    st = np.zeros(6, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.wide_model_quant(p(nodes), nodes.shape[0], p(fv), fv.shape[0], p(v4), p(rays), rays.shape[0], 21, bits, pow2, p(st))
    assert rc == 0
    return st.tolist()


def report(name, prep, sets):
    for label, rays in sets:
        n = float(len(rays))
        for bits, pow2, what in ((0, 0, "exact boxes"), (8, 0, "8 bit"), (8, 1, "8 bit, 2^k steps"), (6, 1, "6 bit, 2^k steps")):
            inner, leaves, rej, tris, fb, mism = quant(prep, rays, bits, pow2)
            assert mism == 0, (name, label, what, mism)
            # lines fetched: a 128-byte line per inner visit now (leaf boxes ride in the parent), 36 B per triangle test;
            # quantised: 64 B per inner visit, 32 B of leaf box per leaf visit, 36 B per triangle test
            now = inner / n * 128 + tris / n * 36
            then = inner / n * 64 + leaves / n * 32 + tris / n * 36
            print("%-6s %-10s %-17s inner %6.2f  leaves %6.2f (%5.2f rejected by the exact box)  tests %6.2f  re-walks %d | bytes/ray %s" % (
                name, label, what, inner / n, leaves / n, rej / n, tris / n, fb,
                ("%.0f (128-B nodes)" % now) if bits == 0 else ("%.0f (64-B nodes + leaf boxes)" % then)), flush=True)


if "c2" in which:
    sc = scenes.soup(1_000_000, seed=12345)
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 0.0, 3.5))
    report("C2", prep, [("primary", Hh.primary_rays(prep, 320, 180)), ("incoherent", Hh.random_rays(50_000, 9, -1.0, 1.0))])
if "c3" in which:
    sc = scenes.interior()
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 1.4, 5.2), center=(0.0, 0.1, 1.0))
    rnd = Hh.random_rays(50_000, 5, -3.5, 3.5)
    rnd[:, 1] = np.abs(rnd[:, 1]) * 0.8 + 0.05
    report("C3", prep, [("primary", Hh.primary_rays(prep, 320, 180)), ("incoherent", rnd)])
if "c4small" in which:
    sc = scenes.displaced_grid(700, 700, patches=8)
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 1.2, 1.8), center=(0.0, 0.55, 1.0))
    rnd = Hh.random_rays(50_000, 4, -1.0, 1.0)
    rnd[:, 1] = np.abs(rnd[:, 1]) * 0.5 + 0.3
    report("C4/10", prep, [("primary", Hh.primary_rays(prep, 320, 180)), ("incoherent", rnd)])
