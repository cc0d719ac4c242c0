"""Per-iteration times of the wavefront on C2 (PBR_PROFILE_DUMP=1): one frame, then a 16-frame batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PBR_PROFILE_DUMP"] = "1"
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
dev.setPipeline(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for k, v in [a.split("=") for a in sys.argv[2:]]:
    dev.setTuning(k, int(v))
for k in range(2):
    print("--- single frame", k, file=sys.stderr, flush=True)
    r.render_frames(1)
    r.finish()
print("--- batch of 16", file=sys.stderr, flush=True)
r.render_frames(16)
r.finish()
