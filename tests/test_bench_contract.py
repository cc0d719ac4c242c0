"""CPU: the JSON line of bench.py's reference arm (the one arm that runs without a GPU) carries the keys the driver
reads, on a reduced workload; a non-zero rank of a multi-rank launch exits quietly."""
import json
import os
import subprocess
import sys

from conftest import ROOT

ARGS = ["--impl", "reference", "--tris", "20000", "--width", "96", "--height", "64", "--steps", "2", "--warmup", "1"]


def _run(extra_env=None, extra_args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS + list(extra_args), cwd=ROOT, env=env,
                          capture_output=True, timeout=600)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr.decode()[-800:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["data"] == "synthetic" and d["config"]["workload"].startswith("REDUCED")       # overrides are recorded
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"},
             ["--gpus", "2"])
    assert r.returncode == 0, r.stderr.decode()[-800:]
    assert not [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
