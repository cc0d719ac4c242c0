"""How sensitive is the wavefront to the number of resident traverse warps?  (pbr_set_tuning "traverse_blocks")"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402

w = dict(bench.WORKLOADS["c2"])
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(w["tris"], seed=12345))
dev = r.device()
for blocks in (10, 9, 8, 7, 6, 5, 4):
    dev.setTuning("traverse_blocks", blocks)
    best = 1e9
    for _ in range(3):
        r.reset_sample_count()
        dev.stats(reset=True)
        t = time.perf_counter()
        r.render_frames(16)
        r.finish()
        best = min(best, time.perf_counter() - t)
    st = dev.stats(reset=True)
    print("traverse blocks/SM %2d (%2d warps): %.3f ms/frame  %.1f Mrays/s" % (blocks, blocks * 4, best * 1e3 / 16, (st[0] + st[1]) / best / 1e6), flush=True)
