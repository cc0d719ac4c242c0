"""Generate tests/golden/reference_outputs.npz from THE REFERENCE ITSELF: its kernel source and its host classes
compiled for the host by oracle/build_ref.py / oracle/build_ref_host.py (needs /root/reference; run in the build
container, the result is committed).  Nothing of the oracle restatement or of the CUDA path goes into this file.

    python tests/golden/make_golden.py

Contents (all for tests/golden/models/suzanne.obj, the reference's own test model):
  <case>/image, <case>/debug   the accumulated frame and the debug image after 3 frames of PathTracer::generateImage
                               at t = 33, 67, 100 ms (seed = ms * 0.001f), 72 x 48 pixels, for the cases in CASES
  <case>/values                the program text values CL::setValues would splice in
  bvh/nodes, bvh/facesV, bvh/facesN, bvh/info     BVH( objects, vertices, normals ) flattened as PathTracer.cpp:238-347
  rays, hits/t, hits/face, hits/nodes, hits/tris  48 x 32 pinhole rays + 500 random rays through traverse()
  shadow/rays, shadow/t, shadow/face, shadow/tris the corresponding shadow rays through traverseShadows()
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H = 72, 48
FRAME_MS = (33, 67, 100)
CASES = {
    "sa": dict(brdf=1, max_depth=4),
    "schlick": dict(brdf=0, max_depth=4),
    "sa_shadow_ms": dict(brdf=1, shadow_rays=1, samples=2, max_depth=3),
    "schlick_shadow": dict(brdf=0, shadow_rays=1, max_depth=3),
    "sa_phong": dict(brdf=1, phong_tess=0.7, max_depth=3),
}
OUT = os.path.join(HERE, "reference_outputs.npz")


def main():
    import helpers as Hh
    from oracle import ref as R
    from oracle import ref_host as RH
    path = Hh.model_path("suzanne.obj")
    out = {}
    for name, kw in CASES.items():
        r = RH.Renderer(path, width=W, height=H, nthreads=4, **kw)
        for ms in FRAME_MS:
            img, dbg = r.generate_image(ms)
        out[name + "/image"] = img
        out[name + "/debug"] = dbg
        out[name + "/values"] = np.array(sorted("%s=%s" % kv for kv in r.values.items()))
        r.close()
    scene, flat = RH.load(path, shadow_rays=1)
    out["bvh/nodes"], out["bvh/facesV"], out["bvh/facesN"] = flat["nodes"], flat["facesV"], flat["facesN"]
    out["bvh/info"] = np.array([flat["info"][k] for k in ("allNodes", "leaves", "depth", "skipped", "emitted", "faces")], np.int64)

    # explicit rays through the reference's traverse() / traverseShadows(); inputs packed from the reference's arrays
    from oracle import scene as S
    v4 = S.pack_float4(scene["vertices"])
    n4 = S.pack_float4(scene["normals"])
    lights = S.pack_lights(scene["lights"])
    D = S.defines(W, H, flat["nodes"].shape[0], len(scene["lights"]), (1.0, 1.0, 1.0, 0.0), brdf=1, shadow_rays=1)

    class P:
        camera = S.camera()
    rays = np.concatenate([Hh.primary_rays(P, 48, 32), Hh.random_rays(500, 11, -1.5, 1.5)])
    t, face, nodes, tris = R.trace(D, flat["nodes"], flat["facesV"], flat["facesN"], v4, n4, lights, rays, nthreads=4)
    out["rays"], out["hits/t"], out["hits/face"], out["hits/nodes"], out["hits/tris"] = rays, t, face, nodes, tris
    hits = np.zeros(len(rays), [("t", "<f4")])
    hits["t"] = t
    sh = Hh.shadow_rays_from_hits(rays, hits, (0.5, 4.0, 1.0))
    t, face, nodes, tris = R.trace(D, flat["nodes"], flat["facesV"], flat["facesN"], v4, n4, lights, sh, any_hit=True, nthreads=4)
    out["shadow/rays"], out["shadow/t"], out["shadow/face"], out["shadow/tris"] = sh, t, face, tris
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.0f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
