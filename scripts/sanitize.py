"""What `compute-sanitizer` wraps: every kernel of the library once, at sizes a sanitizer run finishes in seconds.

    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py
    compute-sanitizer --tool synccheck python scripts/sanitize.py
    compute-sanitizer --tool initcheck python scripts/sanitize.py

Covers: scene repack, raygen / traverse / shade (BRDF 0 and 1), the shadow-ray stage (as a wavefront stage and
inline), depth of field, SAMPLES > 1, Phong tessellation, both pipelines (wavefront, megakernel), both walks (the
reference's visiting order, the ordered walk over the 4-wide BVH -- shared-memory stack and top-of-tree staging
included), batched frames, tile rows and stripes, explicit closest-hit and any-hit rays under both walks, the
pinned-math probe.  Every frame is also compared with the reference-order wavefront's (same bits)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402
import helpers as Hh  # noqa: E402
from oracle import oracle as O  # noqa: E402   (scene loading only; the check against it is tests/)

W, H = 48, 32
SKIP = set(a[5:] for a in sys.argv[1:] if a.startswith("--no-"))      # e.g. --no-megakernel
CASES = [
    ("sa", dict(brdf=1, max_depth=4)),
    ("schlick_shadow_ms", dict(brdf=0, shadow_rays=1, samples=2, max_depth=3)),
    ("sa_shadow_dof", dict(brdf=1, shadow_rays=1, max_depth=3, focus_point=(20, 12))),
    ("phong_sa", dict(brdf=1, max_depth=3, phong_tessellation=0.7)),
]


def main():
    dev = pbr_b200.Device(0)
    bad = 0
    try:
        for name, kw in CASES:
            scene = O.load_obj(Hh.model_path("suzanne.obj"), kw.get("shadow_rays", 0))
            prep = Hh.Prepared(scene, W, H, **kw)
            ds = Hh.DeviceScene(dev, prep)
            dev.setPipeline(0)
            want, wdbg = ds.frames(2)
            for label, setup in [
                ("megakernel", lambda: dev.setPipeline(1)),
                ("shadow inline", lambda: (dev.setPipeline(0), dev.setTuning("shadow_stage", 0))),
                ("measured choice", lambda: dev.setPipeline(-1)),
                ("ordered walk", lambda: (dev.setPipeline(0), dev.setTraversal(1))),
                ("ordered, top 1", lambda: (dev.setPipeline(0), dev.setTraversal(1), dev.setTuning("wide_top", 1))),
            ]:
                if label in SKIP:
                    continue
                setup()
                ordered = label.startswith("ordered") and "phong_tessellation" not in kw
                if label.startswith("ordered") and not ordered:
                    dev.setTraversal(-1)                      # (PHONGTESS keeps the reference-order walk)
                got, gdbg = ds.frames(2)
                ok = Hh.images_equal(got, want) and (ordered or Hh.images_equal(gdbg, wdbg))
                bad += not ok
                print("%-18s %-16s %s" % (name, label, "same bits" if ok else "DIFFERENT"), flush=True)
                dev.setTuning("shadow_stage", 1)
                dev.setTuning("wide_top", 21)
                dev.setTraversal(-1)
            dev.setPipeline(0)
            got, _ = ds.frames_batch(2)
            ok = Hh.images_equal(got, want)
            bad += not ok
            print("%-18s %-16s %s" % (name, "batch", "same bits" if ok else "DIFFERENT"), flush=True)
            # no debug image: frames in flight (four streams, deferred mix), single launches and a batch
            dev.setDebugImage(False)
            for label, run in (("in flight, single", lambda: ds.frames(2, host_roundtrip=False)), ("in flight, batch", lambda: ds.frames_batch(2))):
                got, _ = run()
                ok = Hh.images_equal(got, want)
                bad += not ok
                print("%-18s %-16s %s" % (name, label, "same bits" if ok else "DIFFERENT"), flush=True)
            dev.setDebugImage(True)
            # rows [8, 24) only, then stripes of 4 rows for rank 1 of 2
            dev.setTile(8, 24)
            got, _ = ds.frames(2)
            ok = Hh.images_equal(got[8:24], want[8:24])
            dev.setTileStripes(4, 2, 1)
            got, _ = ds.frames(2)
            rows = [y for y in range(H) if (y // 4) % 2 == 1]
            ok = ok and Hh.images_equal(got[rows], want[rows])
            dev.setTileStripes(0)
            dev.setTile(0, H)
            bad += not ok
            print("%-18s %-16s %s" % (name, "tiles / stripes", "same bits" if ok else "DIFFERENT"), flush=True)
            if name == "sa":
                rays = np.concatenate([Hh.primary_rays(prep, W, H), Hh.random_rays(1000, 3)])
                hits = ds.trace(rays)
                sh = Hh.shadow_rays_from_hits(rays, hits, (0.5, 4.0, 1.0))
                occl = ds.trace(sh, any_hit=True)
                dev.setTraversal(1)
                hits2 = ds.trace(rays)
                dev.setTraversal(-1)
                same = np.array_equal(hits["hitFace"], hits2["hitFace"]) and np.array_equal(hits["t"].view(np.uint32), hits2["t"].view(np.uint32))
                bad += not same
                print("explicit rays: %d closest hits, %d occluded shadow rays, ordered walk %s" % (
                    int((hits["hitFace"] >= 0).sum()), int((occl["t"] < sh[:, 7]).sum()), "same hits" if same else "DIFFERENT"), flush=True)
        x = np.linspace(-3, 3, 257, dtype=np.float32)
        for op in range(5):
            dev.pinnedMath(op, x if op != 3 else np.clip(x, -1, 1))
        print("pinned math probe ok", flush=True)
    finally:
        dev.close()
    print("sanitize.py: %s" % ("all frames identical" if not bad else "%d MISMATCHES" % bad), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
