"""Persistent pipeline (pbr_set_pipeline 2) against the wavefront on the C2 scene: same bits? how fast?
python scripts/persist_test.py [tris]          run under `timeout`: a protocol bug would spin, not crash."""
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

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = dict(bench.WORKLOADS["c2"])
w["tris"] = tris
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(tris, seed=12345))
dev = r.device()
FR = 16


def run(label, reps=3, batch=True):
    best = None
    for _ in range(reps):
        r.reset_sample_count()
        dev.stats(reset=True)
        t = time.perf_counter()
        if batch:
            r.render_frames(FR)
        else:
            for _ in range(FR):
                r.render_frames(1)
        r.finish()
        sec = time.perf_counter() - t
        st = dev.stats(reset=True)
        best = sec if best is None else min(best, sec)
    rays = int(st[0]) + int(st[1])
    img = r.read_image()
    print("%-34s %7.3f ms/frame  %8.1f Mrays/s  rays %d nodes %d" % (label, best * 1e3 / FR, rays / best / 1e6, rays, st[2]), flush=True)
    return img, st


dev.setPipeline(0)
ref, st0 = run("wavefront, frame by frame", batch=False)
img, st = run("wavefront, batched")
print("  identical:", Hh.images_equal(ref, img), " stats equal:", [int(a) for a in st] == [int(a) for a in st0], flush=True)
for bulk, flush, grp in [(64, 256, 4), (16, 256, 4), (256, 256, 4), (64, 64, 4), (64, 1024, 4), (64, 256, 2), (64, 256, 8)]:
    dev.setPipeline(3)
    dev.setTuning("tail_steps_bulk", bulk)
    dev.setTuning("tail_steps_flush", flush)
    dev.setTuning("flush_group", grp)
    img, st = run("carry bulk=%d flush=%d grp=%d, frame by frame" % (bulk, flush, grp), batch=False)
    ok1 = Hh.images_equal(ref, img) and [int(a) for a in st] == [int(a) for a in st0]
    img, st = run("carry bulk=%d flush=%d grp=%d, batched" % (bulk, flush, grp))
    ok2 = Hh.images_equal(ref, img) and [int(a) for a in st] == [int(a) for a in st0]
    print("  identical + stats equal:", ok1, ok2, flush=True)
if len(sys.argv) > 2:
    sys.exit(0)
dev.setPipeline(2)
img, st = run("persistent S=2, frame by frame", batch=False)
print("  identical:", Hh.images_equal(ref, img), " stats equal:", [int(a) for a in st] == [int(a) for a in st0], flush=True)
first = True
for s_blocks, t_blocks, fill in [(2, 0, 4), (1, 0, 4), (3, 0, 4), (1, 0, 0), (1, 0, 16)]:
    dev.setTuning("persist_s", s_blocks)
    dev.setTuning("persist_t", t_blocks)
    dev.setTuning("persist_fill", fill)
    img, st = run("persistent S=%d T=%d fill=%d" % (s_blocks, t_blocks, fill))
    if first:
        print("  identical to wavefront:", Hh.images_equal(ref, img), " stats equal:", [int(a) for a in st] == [int(a) for a in st0], flush=True)
        first = False
    elif not Hh.images_equal(ref, img):
        print("  MISMATCH", flush=True)
