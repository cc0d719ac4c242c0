"""CPU analysis for the layout discussion in DESIGN.md section 6: where do the node visits of the walk go?
Walks primary and random rays through the C2 tree with oracle_visit_histogram_pruned (the walk traverse() really
does, t pruning and face tests included) and splits the visits into inner nodes / one-face leaves / two-face leaves, and by how
much of the array the hottest nodes cover.      python scripts/visit_breakdown.py [tris]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402
from pbr_b200 import host, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import scene as S  # noqa: E402
import helpers as Hh  # noqa: E402

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = scenes.soup(tris, seed=12345)
flat = host.Scene.from_arrays(sc).build_flat()
nodes = np.ascontiguousarray(flat["nodes"], np.float32).reshape(-1, 8)
N = nodes.shape[0]
leaf = nodes[:, 3] >= 0
two = leaf & (nodes[:, 7] >= 0)
facesV = np.ascontiguousarray(flat["facesV"], np.uint32)
facesN = np.ascontiguousarray(flat["facesN"], np.uint32)
vertices4 = S.pack_float4(sc["vertices"])
normals4 = np.zeros((1, 4), np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


print("tree: %d nodes = %d inner + %d leaves (%d with two faces), %.1f MB as 32-byte nodes" % (
    N, (~leaf).sum() - 1, leaf.sum(), two.sum(), N * 32 / 1e6))


class P:
    camera = S.camera(eye=(0.0, 0.0, 3.5), center=(0.0, 0.0, 1.0))


L = O.lib()
D = O.make_defines(img_width=1920, img_height=1080, bvh_num_nodes=N)
for name, rays in (("primary", Hh.primary_rays(P, 320, 180)), ("random", Hh.random_rays(40000, 1, -1.0, 1.0))):
    rays = np.ascontiguousarray(rays, np.float32)
    counts = np.zeros(N, np.uint32)
    L.oracle_visit_histogram_pruned(D.ctypes.data_as(C.c_void_p), nodes.ctypes.data_as(C.c_void_p), _p(facesV), _p(facesN),
                                    _p(vertices4), _p(normals4), rays.ctypes.data_as(C.c_void_p), C.c_int64(len(rays)),
                                    counts.ctypes.data_as(C.c_void_p))
    leaf_hits = int(counts[0])
    counts[0] = 0
    tot = float(counts.sum())
    v_in, v_l1, v_l2 = counts[~leaf].sum() / tot, counts[leaf & ~two].sum() / tot, counts[two].sum() / tot
    order = np.argsort(counts)[::-1]
    cum = np.cumsum(counts[order]) / tot
    hot = {mb: float(cum[min(N - 1, int(mb * 1e6 / 32) - 1)]) for mb in (0.064, 1, 4, 8, 16)}
    n_leaf = float(counts[leaf].sum())
    print("%-8s %6.1f visits/ray: inner %.1f %%, one-face leaves %.1f %%, two-face leaves %.1f %%; box hit at %.1f %% of the leaf "
          "visits (%.1f per ray); nodes touched %.0f %% of the array" % (
              name, tot / len(rays), 100 * v_in, 100 * v_l1, 100 * v_l2, 100 * leaf_hits / n_leaf, leaf_hits / len(rays),
              100 * (counts > 0).mean()))
    print("         share of visits served by the hottest 64 KB / 1 / 4 / 8 / 16 MB of nodes: " +
          " / ".join("%.0f %%" % (100 * hot[k]) for k in (0.064, 1, 4, 8, 16)))
