"""CPU: scene load (OBJ parse + BVH build + flatten) of the product's host library against the reference's own
ObjParser / BVH classes (oracle/_ref/libref_host.so, build container only) on soups written to .obj -- same tree,
bit for bit (asserted), how much faster?      python scripts/bvh_build_bench.py [tris ...]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402
from pbr_b200 import host, scenes  # noqa: E402
from oracle import ref_host as RH  # noqa: E402
from test_oracle_vs_reference_host import _with_face_normals  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [20000, 100000, 300000]
cfg = host.Config()
cfg.reset()
for n in sizes:
    with tempfile.TemporaryDirectory() as d:
        sc = _with_face_normals(scenes.soup(n, seed=12345))
        path = os.path.join(d, "scene.obj")
        scenes.write_obj(sc, path)
        t0 = time.perf_counter()
        s = host.Scene.load(d + "/", "scene.obj")
        t1 = time.perf_counter()
        flat = s.build_flat()
        t2 = time.perf_counter()
        line = "%8d triangles: product parse %.2f s + BVH %.2f s (%d nodes)" % (n, t1 - t0, t2 - t1, flat["nodes"].shape[0])
        if RH.available():
            t3 = time.perf_counter()
            _, ref_flat = RH.load(path, shadow_rays=0)
            t4 = time.perf_counter()
            same = (np.array_equal(ref_flat["nodes"].view(np.uint32), flat["nodes"].view(np.uint32)) and
                    np.array_equal(ref_flat["facesV"], flat["facesV"]))
            line += "; reference classes parse + BVH %.2f s (%.0fx), same tree: %s" % (t4 - t3, (t4 - t3) / (t2 - t0), same)
        print(line, flush=True)
