"""Render a few frames of the C2 workload (1M-triangle soup, 1920x1080) -- the command ncu wraps.
    python scripts/profile_frame.py [frames] [tris]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pbr_b200  # noqa: E402
from pbr_b200 import host, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tris = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
w = dict(bench.WORKLOADS["c2"])
w["tris"] = tris
cfg = host.Config()
bench.host_config(cfg, w)
r = host.Renderer(0)
r.set_deterministic(True)
r.load_scene(scenes.soup(tris, seed=12345))
r.device().setPipeline(0)        # the wavefront (what the measured choice picks on this scene), no megakernel timing rounds
t0 = time.perf_counter()
r.render_frames(frames)
r.finish()
dt = time.perf_counter() - t0
st = r.stats()
print("frames %d: %.2f ms/frame, %.1f Mrays/s, nodes/ray %.1f" % (
    frames, dt * 1e3 / frames, (st[0] + st[1]) / dt / 1e6, st[2] / max(1, st[0])))
r.close()
