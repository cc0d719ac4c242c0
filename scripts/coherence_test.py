"""How much does ray order matter to the traversal engine?  The same rays in three orders: as generated
(pixel order / path order), randomly permuted, and sorted by (origin cell, direction octant).
   python scripts/coherence_test.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pbr_b200  # noqa: E402
from pbr_b200 import host, scenes  # noqa: E402
import helpers as Hh  # noqa: E402

w = dict(bench.WORKLOADS["c2"])
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(w["tris"], seed=12345))
dev = r.device()
r.render_frames(1)
r.finish()
ctx, hd = r.handles()
cam, px = r.camera()


class P:
    camera = cam


def run(name, rays, reps=3):
    n = len(rays)
    rb = dev.createBuffer(rays)
    hb = dev.createEmptyBuffer(n * 16)
    for _ in range(reps):
        dev.stats(reset=True)
        dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb)
        dev.finish()
        ms = dev.kernelTimeMs(hd["kernel"])
        st = dev.stats(reset=True)
    print("%-34s %8.3f ms  %8.1f Mrays/s  nodes/ray %6.1f" % (name, ms, n / ms / 1e3, st[2] / n), flush=True)
    return dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)


def sort_key(rays, cells):
    o = rays[:, 0:3]
    lo, hi = o.min(0), o.max(0)
    c = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * cells).astype(np.int64), 0, cells - 1)
    morton = np.zeros(len(rays), np.int64)
    bits = int(np.log2(cells))
    for b in range(bits):
        for a in range(3):
            morton |= ((c[:, a] >> b) & 1) << (3 * b + a)
    octant = (rays[:, 4] < 0).astype(np.int64) | ((rays[:, 5] < 0).astype(np.int64) << 1) | ((rays[:, 6] < 0).astype(np.int64) << 2)
    return morton * 8 + octant


rng = np.random.default_rng(3)
prim = Hh.primary_rays(P, 1920, 1080)
hits = run("primary, pixel order", prim)
run("primary, shuffled", prim[rng.permutation(len(prim))])

# secondary rays: cosine-ish bounce off the primary hits (what iteration 1 of a frame traces)
ok = np.isfinite(hits["t"])
o = (prim[ok, 0:3] + prim[ok, 4:7] * hits["t"][ok, None]).astype(np.float32)
d = rng.normal(size=o.shape).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros((len(o), 8), np.float32)
sec[:, 0:3] = o + d * np.float32(1e-3)
sec[:, 4:7] = d
sec[:, 7] = np.inf
run("secondary, path order", sec)
run("secondary, shuffled", sec[rng.permutation(len(sec))])
for cells in (8, 32, 128):
    k = sort_key(sec, cells)
    run("secondary, sorted cell%d+octant" % cells, sec[np.argsort(k, kind="stable")])
k = (sec[:, 4] < 0).astype(np.int64) | ((sec[:, 5] < 0).astype(np.int64) << 1) | ((sec[:, 6] < 0).astype(np.int64) << 2)
run("secondary, sorted octant only", sec[np.argsort(k, kind="stable")])
