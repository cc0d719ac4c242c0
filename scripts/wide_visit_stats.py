"""CPU: what the ordered walk saves, counted by the CPU model of tests/wide_walk_model.cpp (which includes the product's
builder header): node visits, triangle tests, stack depth and re-walks per ray for both walks on the BASELINE scenes.
    python scripts/wide_visit_stats.py [c2] [c3] [c4small]"""
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


def report(name, prep, sets):
    for label, rays in sets:
        rc, why, s, f, st = T.run_model(prep, rays, stack_cap=64, top_budget=21)
        assert rc == 0 and st["mismatches"] == 0, (rc, why, st)
        n = float(len(rays))
        print("%-8s %-10s rays %7d | reference order: %6.1f nodes %5.1f tests | ordered: %5.1f wide nodes %5.1f tests | "
              "max stack %2d  hits in front of their leaf box %6d  re-walked %5d | wide nodes %d (depth %d)" % (
                  name, label, len(rays), st["strict_nodes"] / n, st["strict_tris"] / n, st["wide_visits"] / n, st["fast_tris"] / n,
                  st["max_stack"], st["insane_winners"], st["fallbacks"], st["wide_nodes"], st["wide_depth"]), flush=True)


if "c2" in which:
    sc = scenes.soup(1_000_000, seed=12345)
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 0.0, 3.5))
    report("C2", prep, [("primary", Hh.primary_rays(prep, 480, 270)), ("incoherent", Hh.random_rays(100_000, 9, -1.0, 1.0))])
if "c3" in which:
    sc = scenes.interior()
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 1.4, 5.2), center=(0.0, 0.1, 1.0))
    rnd = Hh.random_rays(100_000, 5, -3.5, 3.5)
    rnd[:, 1] = np.abs(rnd[:, 1]) * 0.8 + 0.05
    report("C3", prep, [("primary", Hh.primary_rays(prep, 480, 270)), ("incoherent", rnd)])
if "c4small" in which:
    sc = scenes.displaced_grid(700, 700, patches=8)          # 980 000 triangles of the C4 kind (the full grid takes minutes here)
    prep = Hh.Prepared(sc, 64, 64, bvh=host.Scene.from_arrays(sc).build_flat(), eye=(0.0, 1.2, 1.8), center=(0.0, 0.55, 1.0))
    rnd = Hh.random_rays(100_000, 4, -1.0, 1.0)
    rnd[:, 1] = np.abs(rnd[:, 1]) * 0.5 + 0.3
    report("C4/10", prep, [("primary", Hh.primary_rays(prep, 480, 270)), ("incoherent", rnd)])
