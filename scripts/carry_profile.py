"""Where does the time go?  Per-kernel totals (event-timed) for the wavefront pipelines on C2."""
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
dev.profileEnable(True)
for pipe, batch in [(0, False), (0, True), (3, False), (3, True)]:
    dev.setPipeline(pipe)
    for rep in range(2):
        r.reset_sample_count()
        dev.profileRead(reset=True)
        t = time.perf_counter()
        if batch:
            r.render_frames(16)
        else:
            for _ in range(16):
                r.render_frames(1)
        r.finish()
        sec = time.perf_counter() - t
        pr = dev.profileRead(reset=True)
    print("pipeline %d %-8s wall %.2f ms/frame | %s" % (pipe, "batched" if batch else "per-frame", sec * 1e3 / 16, pr), flush=True)
