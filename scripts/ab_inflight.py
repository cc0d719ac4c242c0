"""C2 frames (16 per step) with 1 .. 4 frames in flight, both walks: best / median of 5 and the image checksum.
    python scripts/ab_inflight.py [workload tris]"""
import os
import sys
import time

import numpy as np

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
for walk in (1, 0):
    r.set_traversal(walk)
    for n in (1, 2, 3, 4):
        dev.setTuning("frames_in_flight", n)
        r.reset_sample_count()
        r.render_frames(16)
        r.finish()
        times = []
        for _ in range(5):
            r.reset_sample_count()
            dev.stats(reset=True)
            t0 = time.perf_counter()
            r.render_frames(16)
            r.finish()
            times.append(time.perf_counter() - t0)
            st = dev.stats(reset=True)
        img = r.read_image()
        rays = int(st[0]) + int(st[1])
        print("walk %d  frames in flight %d : best %.3f median %.3f ms/frame -> %.1f Mrays/s   image %08x" % (
            walk, n, min(times) / 16 * 1e3, float(np.median(times)) / 16 * 1e3, rays / min(times) / 1e6,
            int(np.bitwise_xor.reduce(img[..., :3].view(np.uint32).ravel()))), flush=True)
r.close()
