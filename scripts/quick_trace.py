"""Quick device-side numbers on the C2 scene (1M-triangle soup): explicit primary rays, incoherent
random rays, and whole frames, with the visit counters.   python scripts/quick_trace.py [tris]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pbr_b200  # noqa: E402
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
r.render_frames(1)          # the camera struct is filled by the first frame
r.finish()
ctx, hd = r.handles()
cam, px = r.camera()


class P:  # what helpers.primary_rays needs
    camera = cam


def run(name, rays, any_hit=False, reps=3):
    n = len(rays)
    rb = dev.createBuffer(rays)
    hb = dev.createEmptyBuffer(n * 16)
    for _ in range(reps):
        dev.stats(reset=True)
        dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb, any_hit=any_hit)
        dev.finish()
        ms = dev.kernelTimeMs(hd["kernel"])
        st = dev.stats(reset=True)
        nodes = st[5] if any_hit else st[2]
    print("%-14s %8.3f ms  %8.1f Mrays/s  nodes/ray %6.1f  tris/ray %5.2f  alg %7.1f GB/s" % (
        name, ms, n / ms / 1e3, nodes / n, st[3] / n, (32 * nodes + 64 * st[3]) / ms / 1e6), flush=True)
    return dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)


prim = Hh.primary_rays(P, 1920, 1080)
print("rays", prim.shape, flush=True)
hits = run("primary", prim)
run("random", Hh.random_rays(2_000_000, 1, -1.0, 1.0))
run("shadow(any)", Hh.shadow_rays_from_hits(prim, hits, (0.0, 3.0, 0.0)), any_hit=True)
dev.profileEnable(True)
r.render_frames(3)
r.finish()
dev.profileRead(reset=True)
dev.stats(reset=True)
t0 = time.perf_counter()
r.render_frames(8)
r.finish()
wall = (time.perf_counter() - t0) / 8
pr = dev.profileRead(reset=True)
st = dev.stats(reset=True)
print("frame: %.3f ms wall, traverse %.3f ms, shade %.3f ms, raygen %.3f ms  -> %.1f Mrays/s, nodes/ray %.1f" % (
    wall * 1e3, pr["traverse_ms"] / 8, pr["shade_ms"] / 8, pr["raygen_ms"] / 8, st[0] / 8 / wall / 1e6, st[2] / st[0]))
r.close()
