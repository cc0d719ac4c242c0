"""The C2 workload (16 frames) and 2 M incoherent + 2 M primary explicit rays with the reference-order walk and with
the ordered walk, on one box: best and median of 5, image / hit checksums (they must agree between the two walks).
    python scripts/ab_traversal.py [wide_top ...]"""
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

# arguments: settings of the ordered walk to time, each "top,node_phase_min,refill_min" (default 21,20,8)
settings = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(21, 20, 8)]
w = dict(bench.WORKLOADS["c2"])
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(w["tris"], seed=12345))
dev = r.device()
dev.setPipeline(0)
dev.profileEnable(True)
ctx, hd = r.handles()
cam, _ = r.camera()


class P:
    camera = cam


rnd = Hh.random_rays(2_000_000, 1, -1.0, 1.0)
prim = Hh.primary_rays(P, 1920, 1080)
bufs = {}
for name, rays in (("random", rnd), ("primary", prim)):
    bufs[name] = (dev.createBuffer(rays), dev.createEmptyBuffer(len(rays) * 16), len(rays))


def run(label):
    r.reset_sample_count()
    r.render_frames(3)
    r.finish()
    times = []
    for _ in range(5):
        r.reset_sample_count()
        dev.stats(reset=True)
        dev.profileRead(reset=True)
        t0 = time.perf_counter()
        r.render_frames(16)
        r.finish()
        times.append(time.perf_counter() - t0)
        st = dev.stats(reset=True)
        prof = dev.profileRead(reset=True)
    rays = int(st[0]) + int(st[1])
    img = r.read_image()
    info = dev.traversalInfo(reset=True)
    print("%-14s frames : best %.3f  median %.3f ms/frame -> %.1f Mrays/s  traverse %.3f shade %.3f raygen %.3f ms/frame  nodes/ray %.1f tris/ray %.1f  "
          "walk %d rewalked %d  image %08x" % (
              label, min(times) / 16 * 1e3, float(np.median(times)) / 16 * 1e3, rays / min(times) / 1e6,
              prof["traverse_ms"] / 16, prof["shade_ms"] / 16, prof["raygen_ms"] / 16, st[2] / max(1, st[0]), st[3] / max(1, st[0]),
              info["last_used"], info["rewalked_rays"], int(np.bitwise_xor.reduce(img[..., :3].view(np.uint32).ravel()))), flush=True)
    for name, (rb, hb, n) in bufs.items():
        ms = []
        for _ in range(5):
            dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb)
            dev.finish()
            ms.append(dev.kernelTimeMs(hd["kernel"]))
        hits = dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)
        ck = int(np.bitwise_xor.reduce(hits["t"].view(np.uint32))) ^ int(np.bitwise_xor.reduce(hits["hitFace"].view(np.uint32))) ^ \
            int(np.bitwise_xor.reduce(hits["leaf"].view(np.uint32)))
        print("%-14s %-7s: best %.3f  median %.3f ms -> %.1f Mrays/s   hit checksum %08x" % (
            label, name, min(ms), float(np.median(ms)), n / min(ms) / 1e3, ck), flush=True)


dev.setTraversal(0)
run("reference")
for top, npm, rm in settings:
    dev.setTuning("wide_top", top)
    dev.setTuning("wide_node_phase_min", npm)
    dev.setTuning("wide_refill_min", rm)
    dev.setTraversal(1)
    run("ord %d/%d/%d" % (top, npm, rm))
print(dev.traversalInfo())
r.close()
