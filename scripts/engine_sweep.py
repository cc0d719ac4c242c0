"""Sweep of the traversal engine's two thresholds on C2 (pbr_set_tuning node_phase_min / refill_min)."""
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
for npm in (8, 12, 16, 20, 24, 28):
    row = []
    for rf in (1, 2, 4, 8, 16):
        dev.setTuning("node_phase_min", npm)
        dev.setTuning("refill_min", rf)
        best = 1e9
        for _ in range(3):
            r.reset_sample_count()
            dev.stats(reset=True)
            t = time.perf_counter()
            r.render_frames(16)
            r.finish()
            best = min(best, time.perf_counter() - t)
        st = dev.stats(reset=True)
        row.append("%.0f" % ((st[0] + st[1]) / best / 1e6))
    print("node_phase_min %2d | refill_min 1 2 4 8 16: %s" % (npm, "  ".join(row)), flush=True)
