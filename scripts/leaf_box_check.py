"""Premise of the next layout step: is every leaf box exactly the min / max of the vertices of
the (one or two) faces the leaf holds?  CPU only.   python scripts/leaf_box_check.py"""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pbr_b200
from pbr_b200 import scenes, host
from oracle import oracle as O
import helpers as Hh
def check(name, scene, **kw):
    bvh = O.build_bvh(scene, **kw)
    nodes = bvh["nodes"].reshape(-1, 8); fv = bvh["facesV"].reshape(-1, 4)
    v = np.asarray(scene["vertices"], np.float32).reshape(-1, 3)
    leaf = nodes[:, 3] >= 0
    f0 = nodes[leaf, 3].astype(np.int64); f1 = nodes[leaf, 7].astype(np.int64)
    tri = v[fv[:, :3].astype(np.int64)]            # [F,3,3]
    lo = tri.min(1); hi = tri.max(1)
    two = f1 >= 0
    blo = lo[f0].copy(); bhi = hi[f0].copy()
    blo[two] = np.minimum(blo[two], lo[f1[two]]); bhi[two] = np.maximum(bhi[two], hi[f1[two]])
    ok = (blo == nodes[leaf, 0:3]).all(1) & (bhi == nodes[leaf, 4:7]).all(1)
    print("%-28s nodes %8d leaves %8d (2 faces: %5.1f %%)  box derivable: %d of %d" % (name, len(nodes), leaf.sum(), 100*two.mean(), ok.sum(), len(ok)))
check("suzanne", O.load_obj(Hh.model_path("suzanne.obj"), 0))
check("pillars", O.load_obj(Hh.model_path("pillars.obj"), 0))
check("soup 200k", scenes.soup(200000, seed=12345))
check("soup 200k max_faces=1", scenes.soup(200000, seed=12345), max_faces=1)
check("grid", scenes.displaced_grid(120, 100, patches=4, seed=5))
check("interior", scenes.interior(detail=0.2))
