"""Experiment: T independent renderers (own context + stream), each owning a block of rows, all
enqueueing 16 frames without synchronising -- does overlapping the tile pipelines hide the traversal
tails?   python scripts/tile_streams_test.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, bench, pbr_b200
from pbr_b200 import host, scenes, multigpu
w = dict(bench.WORKLOADS["c2"]); cfg = host.Config(); bench.host_config(cfg, w)
scene = scenes.soup(1_000_000, seed=12345)
H = 1080
for T in (1, 2, 3, 4, 6, 8):
    rs = []
    for t in range(T):
        r = host.Renderer(0); r.set_deterministic(True); r.load_scene(scene)
        y0, y1 = multigpu.tile_rows(H, t, T)
        r.set_tile(y0, y1)
        rs.append(r)
    for r in rs: r.render_frames(2)
    for r in rs: r.finish(); r.stats(reset=True)
    t0 = time.perf_counter()
    for f in range(16):
        for r in rs: r.render_frames(1)
    for r in rs: r.finish()
    sec = time.perf_counter() - t0
    rays = sum(float(r.stats(reset=True)[0]) for r in rs)
    print("T=%d tile pipelines: %.3f ms/frame  %.1f Mrays/s" % (T, sec * 1e3 / 16, rays / sec / 1e6), flush=True)
    for r in rs: r.close()
