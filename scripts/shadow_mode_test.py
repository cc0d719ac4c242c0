"""Throughput of the interior scene (C3) with and without render.shadow_rays (the shadow ray to lights[0] is
walked inside the shade kernel, one thread per path)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402

scene = scenes.interior(detail=1.0, with_light=True)
for brdf in (1, 0):
    for shadow, stage in ((0, 1), (1, 0), (1, 1)):
        cfg = host.Config()
        cfg.reset()
        cfg.update({"window.width": 1920, "window.height": 1080, "render.brdf": brdf, "render.shadow_rays": shadow,
                    "render.max_depth": 3, "camera.eye.x": 0.0, "camera.eye.y": 1.4, "camera.eye.z": 5.5,
                    "camera.center.x": 0.0, "camera.center.y": 0.1, "camera.center.z": 1.0, "logging.level": 0})
        r = host.Renderer(0)
        r.set_deterministic(True)
        r.load_scene(scene)
        dev = r.device()
        dev.setTuning("shadow_stage", stage)
        dev.profileEnable(True)
        r.render_frames(4)
        r.finish()
        dev.stats(reset=True)
        dev.profileRead(reset=True)
        t = time.perf_counter()
        r.render_frames(32)
        r.finish()
        sec = time.perf_counter() - t
        st = dev.stats(reset=True)
        pr = dev.profileRead(reset=True)
        print("brdf %d shadow_rays %d stage %d: %.3f ms/frame  %7.1f Mrays/s (closest %d + shadow %d)  traverse %.1f ms  shade %.1f ms  lights %d" % (
            brdf, shadow, stage, sec * 1e3 / 32, (st[0] + st[1]) / sec / 1e6, st[0], st[1], pr["traverse_ms"], pr["shade_ms"], r.info()["lights"]), flush=True)
        r.close()
