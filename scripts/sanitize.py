"""What `compute-sanitizer` wraps: every kernel of the library once, at sizes a sanitizer run finishes in seconds.

    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py
    compute-sanitizer --tool synccheck python scripts/sanitize.py
    compute-sanitizer --tool initcheck python scripts/sanitize.py

Covers: scene repack, raygen / traverse / shade (BRDF 0 and 1), the shadow-ray stage (as a wavefront stage and
inline), depth of field, SAMPLES > 1, Phong tessellation, the four pipelines (wavefront, megakernel, persistent
rings, carry-over), batched frames with and without interleaving, tile rows and stripes, explicit closest-hit
and any-hit rays, the pinned-math probe.  Every frame is also compared with the wavefront's (same bits)."""
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
SKIP = set(a[5:] for a in sys.argv[1:] if a.startswith("--no-"))      # e.g. --no-persistent
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
                ("persistent", lambda: dev.setPipeline(2)),
                ("carry-over", lambda: dev.setPipeline(3)),
                ("shadow inline", lambda: (dev.setPipeline(0), dev.setTuning("shadow_stage", 0))),
                ("measured choice", lambda: dev.setPipeline(-1)),
            ]:
                if label in SKIP:
                    continue
                setup()
                got, gdbg = ds.frames(2)
                ok = Hh.images_equal(got, want) and Hh.images_equal(gdbg, wdbg)
                bad += not ok
                print("%-18s %-16s %s" % (name, label, "same bits" if ok else "DIFFERENT"), flush=True)
                dev.setTuning("shadow_stage", 1)
            dev.setPipeline(0)
            for inter in (0, 1):
                dev.setTuning("batch_interleave", inter)
                got, _ = ds.frames_batch(2)
                ok = Hh.images_equal(got, want)
                bad += not ok
                print("%-18s %-16s %s" % (name, "batch interleave=%d" % inter, "same bits" if ok else "DIFFERENT"), flush=True)
            dev.setTuning("batch_interleave", 0)
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
                print("explicit rays: %d closest hits, %d occluded shadow rays" % (
                    int((hits["hitFace"] >= 0).sum()), int((occl["t"] < sh[:, 7]).sum())), flush=True)
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
