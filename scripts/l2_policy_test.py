"""L2 access-policy window over the scene during traversal (pbr_set_tuning l2_window / l2_hit_pct / l2_persist_mb)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402
import helpers as Hh  # noqa: E402

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = dict(bench.WORKLOADS["c2"])
w["tris"] = tris
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(tris, seed=12345))
dev = r.device()
ref = None
for window, hit, mb in [(0, 100, 0), (1, 100, 0), (1, 100, 40), (2, 100, 0), (2, 60, 0), (2, 40, 0), (2, 80, 0), (2, 60, 64), (0, 100, 0)]:
    dev.setTuning("l2_window", window)
    dev.setTuning("l2_hit_pct", hit)
    dev.setTuning("l2_persist_mb", mb)
    best = 1e9
    for _ in range(3):
        r.reset_sample_count()
        dev.stats(reset=True)
        t = time.perf_counter()
        r.render_frames(16)
        r.finish()
        best = min(best, time.perf_counter() - t)
    st = dev.stats(reset=True)
    img = r.read_image()
    if ref is None:
        ref = img
    print("window %d hit %3d%% persist %3d MB: %.3f ms/frame  %.1f Mrays/s  same image %s" % (
        window, hit, mb, best * 1e3 / 16, (st[0] + st[1]) / best / 1e6, Hh.images_equal(ref, img)), flush=True)
