import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bench, pbr_b200, helpers as Hh
from pbr_b200 import host, scenes
w = dict(bench.WORKLOADS["c2"]); cfg = host.Config(); bench.host_config(cfg, w)
r = host.Renderer(0); r.set_deterministic(True); r.load_scene(scenes.soup(1_000_000, seed=12345))
dev = r.device(); r.render_frames(1); r.finish(); ctx, hd = r.handles()
for n in (250_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000, 16_000_000):
    rays = Hh.random_rays(n, 1, -1.0, 1.0)
    rb = dev.createBuffer(rays); hb = dev.createEmptyBuffer(n * 16)
    best = 1e9
    for _ in range(3):
        dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb); dev.finish()
        best = min(best, dev.kernelTimeMs(hd["kernel"]))
    print("random rays %9d: %8.3f ms %8.1f Mrays/s" % (n, best, n / best / 1e3), flush=True)
