"""Wavefront (0) against the megakernel (1) as the scene grows: where is the crossover?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "models")


def run(label, loader, w, h, eye, center=(0.0, 0.0, 1.0), depth=4, frames=32):
    cfg = host.Config()
    cfg.reset()
    cfg.update({"window.width": w, "window.height": h, "render.max_depth": depth, "logging.level": 0,
                "camera.eye.x": eye[0], "camera.eye.y": eye[1], "camera.eye.z": eye[2],
                "camera.center.x": center[0], "camera.center.y": center[1], "camera.center.z": center[2]})
    r = host.Renderer(0)
    r.set_deterministic(True)
    loader(r)
    dev = r.device()
    out = []
    for pipeline in (0, 1, -1):
        dev.setPipeline(pipeline)
        r.reset_sample_count()
        r.render_frames(4)
        r.finish()
        dev.stats(reset=True)
        r.reset_sample_count()
        t = time.perf_counter()
        r.render_frames(frames)
        r.finish()
        sec = time.perf_counter() - t
        st = dev.stats(reset=True)
        out.append("%s %.3f ms/frame %6.0f Mrays/s" % ({0: "wavefront", 1: "megakernel", -1: "auto"}[pipeline], sec * 1e3 / frames, (st[0] + st[1]) / sec / 1e6))
    print("%-34s nodes %8d | %s | %s | %s" % (label, r.info()["emitted_nodes"], out[0], out[1], out[2]), flush=True)
    r.close()


run("suzanne 512x512", lambda r: r.load_model(MODELS + "/", "suzanne.obj"), 512, 512, (0.0, 1.0, 3.0))
run("suzanne 1920x1080", lambda r: r.load_model(MODELS + "/", "suzanne.obj"), 1920, 1080, (0.0, 1.0, 3.0))
for tris in (2_000, 10_000, 50_000, 200_000, 1_000_000):
    sc = scenes.soup(tris, seed=12345)
    run("soup %d tris 1920x1080" % tris, lambda r, sc=sc: r.load_scene(sc), 1920, 1080, (0.0, 0.0, 3.5), depth=3, frames=16)
sc = scenes.interior(detail=1.0, with_light=False)
run("interior 254k tris 1920x1080", lambda r: r.load_scene(sc), 1920, 1080, (0.0, 1.4, 5.5), center=(0.0, 0.1, 1.0), depth=3, frames=16)
