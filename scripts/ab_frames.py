"""A/B helper: the C2 workload (16 frames, wavefront) and 2 M incoherent explicit rays, best and median of 5
repetitions, for whichever libpbr_b200.so is in place.  Run it once per library build on the SAME box:
    python scripts/ab_frames.py [label]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402
import helpers as Hh  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "lib"
w = dict(bench.WORKLOADS["c2"])
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(w["tris"], seed=12345))
dev = r.device()
dev.setPipeline(0)
r.render_frames(3)
r.finish()
ctx, hd = r.handles()

times = []
for _ in range(5):
    r.reset_sample_count()
    dev.stats(reset=True)
    t0 = time.perf_counter()
    r.render_frames(16)
    r.finish()
    times.append(time.perf_counter() - t0)
    st = dev.stats(reset=True)
rays = int(st[0]) + int(st[1])
img = r.read_image()
print("%-6s frames : best %.3f  median %.3f ms/frame  -> %.1f / %.1f Mrays/s   image checksum %08x" % (
    label, min(times) / 16 * 1e3, float(np.median(times)) / 16 * 1e3, rays / min(times) / 1e6,
    rays / float(np.median(times)) / 1e6, int(np.bitwise_xor.reduce(img.view(np.uint32).ravel()))), flush=True)

rnd = Hh.random_rays(2_000_000, 1, -1.0, 1.0)
n = len(rnd)
rb = dev.createBuffer(rnd)
hb = dev.createEmptyBuffer(n * 16)
ms = []
for _ in range(5):
    dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb)
    dev.finish()
    ms.append(dev.kernelTimeMs(hd["kernel"]))
hits = dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)
print("%-6s random : best %.3f  median %.3f ms  -> %.1f Mrays/s   hit checksum %08x" % (
    label, min(ms), float(np.median(ms)), n / min(ms) / 1e3,
    int(np.bitwise_xor.reduce(hits.view(np.uint32).ravel()))), flush=True)
r.close()
