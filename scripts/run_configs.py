"""Run the BASELINE.json configurations on one rank (or, under torchrun, shard C4 tiles / C5 rays over
the ranks) and write one JSON document with throughput, parity and CPU-oracle numbers.

    python scripts/run_configs.py [--only c1,c3,c4,c5] [--out gpurun_out/configs.json] [--quick]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/run_configs.py --only c4,c5

C1  bundled Cornell box + Suzanne (tests/golden/models/suzanne.obj), 512x512, 1 spp, max_depth 4
C3  procedural interior (~254 k triangles), 1920x1080, 64 spp, BRDF 1 and BRDF 0
C4  displaced grid, 10 003 864 triangles in 64 objects, 3840x2160, 32 spp, rows sharded over the ranks
C5  explicit primary + shadow rays on the C2 scene (1 M-triangle soup), 1 M .. 50 M rays, ray array
    split contiguously over the ranks, hit index / leaf / t checked bit for bit against the oracle
(C2 is bench.py.)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pbr_b200  # noqa: E402
from pbr_b200 import host, multigpu, scenes  # noqa: E402
import helpers as Hh  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker + CPU baseline only)
from oracle import scene as S  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "models")
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
THREADS = os.cpu_count() or 1


def base_config(cfg, width, height, **kw):
    cfg.reset()
    cfg.update({"window.width": width, "window.height": height, "logging.level": 1})
    cfg.update(kw)


def prepared_like(cfg, scene, flat, width=None, height=None):
    """Oracle-side view of the scene with the current Cfg and the product's flattened BVH."""
    W = width or int(cfg.get("window.width"))
    H = height or int(cfg.get("window.height"))
    return Hh.Prepared(
        scene, W, H, brdf=int(cfg.get("render.brdf")), samples=int(cfg.get("render.samples")),
        max_depth=int(cfg.get("render.max_depth")), max_added_depth=int(cfg.get("render.max_added_depth")),
        shadow_rays=int(cfg.get("render.shadow_rays")), antialiasing=float(cfg.get("render.antialiasing")),
        eye=tuple(float(cfg.get("camera.eye." + a)) for a in "xyz"),
        center=tuple(float(cfg.get("camera.center." + a)) for a in "xyz"), bvh=flat)


def time_frames(r, frames, warm=2):
    r.reset_sample_count()
    r.render_frames(warm)
    r.finish()
    r.stats(reset=True)
    r.reset_sample_count()
    t0 = time.perf_counter()
    r.render_frames(frames)
    r.finish()
    sec = time.perf_counter() - t0
    st = r.stats(reset=True).astype(np.float64)
    return sec, st


def oracle_rate(prep, frames=1, budget_s=8.0):
    img = np.zeros((prep.H, prep.W, 4), np.float32)
    rays, n, t0 = 0, 0, time.perf_counter()
    while n < frames or ((time.perf_counter() - t0) * (n + 1) / max(n, 1) < budget_s and n < 64):
        img, _, st = O.path_tracing(prep.defines, S.frame_seed(n), S.pixel_weight(n), prep.px_dim, prep.camera,
                                    prep.nodes, prep.facesV, prep.facesN, prep.vertices4, prep.normals4,
                                    prep.materials, prep.lights, img, nthreads=THREADS, debug=False)
        rays += int(st[0]) + int(st[1])
        n += 1
    sec = time.perf_counter() - t0
    return {"mrays_per_s": rays / sec / 1e6, "frames": n, "seconds": sec, "cores": THREADS, "image": img}


def run_c1(cfg, quick):
    base_config(cfg, 512, 512, **{"render.max_depth": 4, "render.samples": 1})
    r = host.Renderer(LOCAL)
    r.set_deterministic(True)
    r.load_model(MODELS + "/", "suzanne.obj")
    got = r.generate_image()
    scene = O.load_obj(os.path.join(MODELS, "suzanne.obj"), 0)
    prep = prepared_like(cfg, scene, r.flat())
    orc = oracle_rate(prep, frames=1, budget_s=0.0)
    sec, st = time_frames(r, 64)
    out = {
        "config": "C1 suzanne.obj 512x512 1spp max_depth 4 (BRDF 1, no shadow rays)",
        "gpu_mrays_per_s": (st[0] + st[1]) / sec / 1e6, "gpu_ms_per_frame": sec * 1e3 / 64,
        "gpu_samples_per_s": 64 * 512 * 512 / sec,
        "cpu_oracle_mrays_per_s": orc["mrays_per_s"], "cpu_cores": THREADS,
        "frame_bit_identical_to_oracle": Hh.images_equal(got, orc["image"]),
        "mean_relative_error": Hh.mean_relative_error(got, orc["image"]),
        "nodes_per_ray": st[2] / max(1.0, st[0]),
    }
    r.close()
    return out


def run_c3(cfg, quick):
    res = {}
    scene = scenes.interior()
    W, H, SPP = (1920, 1080, 64) if not quick else (640, 360, 8)
    for brdf in (1, 0):
        base_config(cfg, W, H, **{
            "render.brdf": brdf, "render.max_depth": 3, "camera.eye.x": 0.0, "camera.eye.y": 1.4, "camera.eye.z": 5.2,
            "camera.center.x": 0.0, "camera.center.y": 0.1, "camera.center.z": 1.0})
        r = host.Renderer(LOCAL)
        r.set_deterministic(True)
        r.load_scene(scene)
        info = r.info()
        sec, st = time_frames(r, SPP)
        flat = r.flat()
        # parity on a reduced frame (same scene, same BVH): whole image bit for bit
        r.close()
        pw, ph = 320, 180
        base_config(cfg, pw, ph, **{
            "render.brdf": brdf, "render.max_depth": 3, "camera.eye.x": 0.0, "camera.eye.y": 1.4, "camera.eye.z": 5.2,
            "camera.center.x": 0.0, "camera.center.y": 0.1, "camera.center.z": 1.0})
        r2 = host.Renderer(LOCAL)
        r2.set_deterministic(True)
        r2.load_scene(scene)
        got = None
        for _ in range(2):
            got = r2.generate_image()
        prep = prepared_like(cfg, scene, flat, pw, ph)
        want, _, _ = prep.oracle_frames(2, nthreads=THREADS)
        r2.close()
        base_config(cfg, W, H, **{
            "render.brdf": brdf, "render.max_depth": 3, "camera.eye.x": 0.0, "camera.eye.y": 1.4, "camera.eye.z": 5.2,
            "camera.center.x": 0.0, "camera.center.y": 0.1, "camera.center.z": 1.0})
        orc = oracle_rate(prepared_like(cfg, scene, flat), frames=1, budget_s=6.0)
        res["brdf%d" % brdf] = {
            "config": "C3 interior %d tris %dx%d %dspp BRDF %d" % (info["faces"], W, H, SPP, brdf),
            "gpu_mrays_per_s": (st[0] + st[1]) / sec / 1e6, "gpu_ms_per_frame": sec * 1e3 / SPP,
            "gpu_samples_per_s": SPP * W * H / sec, "nodes_per_ray": st[2] / max(1.0, st[0]),
            "bvh_build_s": info["bvh_build_seconds"],
            "cpu_oracle_mrays_per_s": orc["mrays_per_s"], "cpu_cores": THREADS, "cpu_sample_frames": orc["frames"],
            "parity_frame_320x180_bit_identical": Hh.images_equal(got, want),
            "parity_mean_relative_error": Hh.mean_relative_error(got, want),
            "nan_pixel_fraction": float(np.isnan(got[..., :3]).any(-1).mean()),
        }
    return res


def init_dist():
    if WORLD == 1:
        return None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(LOCAL)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=RANK, world_size=WORLD, device_id=torch.device("cuda", LOCAL))
    return dist


def run_c4(cfg, quick, dist):
    import torch
    W, H, SPP = (3840, 2160, 32) if not quick else (960, 540, 4)
    cells = (2237, 2236) if not quick else (500, 500)
    base_config(cfg, W, H, **{
        "render.max_depth": 3, "camera.eye.x": 0.0, "camera.eye.y": 1.2, "camera.eye.z": 1.8,
        "camera.center.x": 0.0, "camera.center.y": 0.55, "camera.center.z": 1.0})
    t0 = time.perf_counter()
    scene = scenes.displaced_grid(cells[0], cells[1], patches=8)
    gen_s = time.perf_counter() - t0
    r = host.Renderer(LOCAL)
    r.set_deterministic(True)
    t0 = time.perf_counter()
    r.load_scene(scene)
    load_s = time.perf_counter() - t0
    info = r.info()
    # rows interleaved in stripes: a height field seen from above its horizon has cheap sky rows and expensive
    # ground rows, so contiguous blocks would leave most ranks waiting for the one with the ground
    stripe = multigpu.stripe_rows_for(H, WORLD, want=8) if WORLD > 1 else 0
    if stripe > 0:
        r.set_tile_stripes(stripe, WORLD, RANK)
    else:
        y0, y1 = multigpu.tile_rows(H, RANK, WORLD)
        r.set_tile(y0, y1)
    dev = r.device()
    img_t = None
    if dist is not None:
        dev.setStream(torch.cuda.current_stream().cuda_stream)

    # the gather of frame k runs on a side stream while frame k + 1 is traced (frame k + 1 reads only this rank's own
    # rows of image k; frame k + 2, which overwrites that image, waits for the gather of frame k) -- as in bench.py
    render_stream = torch.cuda.current_stream() if dist is not None else None
    gather_stream = torch.cuda.Stream() if dist is not None else None
    gathered = [None, None]
    views = {}
    count = [0]

    def frame():
        j = count[0]
        count[0] += 1
        if dist is not None and gathered[j & 1] is not None:
            render_stream.wait_event(gathered[j & 1])
        r.render_frames(1)
        if dist is not None:
            _, hd = r.handles()
            ptr, _ = dev.devicePtr(hd["image"])
            if ptr not in views:
                views[ptr] = multigpu.DeviceImage(ptr, H, W, torch.device("cuda", LOCAL)).tensor
            t = views[ptr]
            rendered = torch.cuda.Event()
            rendered.record(render_stream)
            with torch.cuda.stream(gather_stream):
                gather_stream.wait_event(rendered)
                if stripe > 0:
                    multigpu.combine_stripes(t, RANK, WORLD, stripe)
                else:
                    multigpu.combine_tiles(t, RANK, WORLD)
                gathered[j & 1] = torch.cuda.Event()
                gathered[j & 1].record(gather_stream)

    for _ in range(2):
        frame()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    r.stats(reset=True)
    r.reset_sample_count()
    t0 = time.perf_counter()
    for _ in range(SPP):
        frame()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sec = time.perf_counter() - t0
    st = r.stats(reset=True).astype(np.float64)
    if dist is not None:
        t = torch.tensor(st, device="cuda")
        dist.all_reduce(t)
        st = t.cpu().numpy()
        ts = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sec = float(ts.item())
    out = None
    if RANK == 0:
        # parity: explicit primary rays of a coarse grid against the oracle on the product's BVH
        cam, px = r.camera()

        class P:
            camera = cam
        rays = Hh.primary_rays(P, 400, 225)
        got = r.trace(rays)
        flat = r.flat()
        prep = prepared_like(cfg, scene, flat, 64, 64)
        want, _ = prep.oracle_trace(rays, nthreads=THREADS)
        out = {
            "config": "C4 displaced grid %d tris (%d objects) %dx%d %dspp, rows sharded over %d rank(s)%s" % (
                info["faces"], len(scene["objFaceCounts"]), W, H, SPP, WORLD,
                " in interleaved stripes of %d rows" % stripe if stripe > 0 else ""),
            "gpu_mrays_per_s": (st[0] + st[1]) / sec / 1e6, "gpu_ms_per_frame": sec * 1e3 / SPP,
            "gpu_samples_per_s": SPP * W * H / sec, "nodes_per_ray": st[2] / max(1.0, st[0]),
            "scene_gen_s": gen_s, "load_s": load_s, "bvh_build_s": info["bvh_build_seconds"], "bvh_nodes": info["emitted_nodes"],
            "parity_rays": len(rays),
            "parity_hit_face_leaf_t_bit_exact": bool(
                np.array_equal(got["hitFace"], want["hitFace"]) and np.array_equal(got["leaf"], want["leaf"]) and
                np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))),
            "primary_hit_fraction": float(np.isfinite(want["t"]).mean()),
        }
    r.close()
    return out


def run_c5(cfg, quick, dist):
    import torch
    base_config(cfg, 1920, 1080, **{"camera.eye.x": 0.0, "camera.eye.y": 0.0, "camera.eye.z": 3.5})
    scene = scenes.soup(1_000_000 if not quick else 100_000, seed=12345)
    r = host.Renderer(LOCAL)
    r.set_deterministic(True)
    r.load_scene(scene)
    r.render_frames(1)
    r.finish()
    cam, px = r.camera()
    dev = r.device()
    _, hd = r.handles()
    flat = r.flat() if RANK == 0 else None
    prep = prepared_like(cfg, scene, flat, 64, 64) if RANK == 0 else None

    class P:
        camera = cam
    sizes = [1, 2, 5, 10, 20, 50] if not quick else [1, 2]
    rows = []
    for mega in sizes:
        n_total = mega * 1_000_000
        # pinhole grid oversampling the 1080p image, 16:9, no jitter
        h = int(round((n_total * 9 / 16) ** 0.5))
        w = n_total // h
        n_total = w * h
        lo, hi = (n_total * RANK) // WORLD, (n_total * (RANK + 1)) // WORLD
        rays_all = Hh.primary_rays(P, w, h)
        rays = np.ascontiguousarray(rays_all[lo:hi])
        n = len(rays)
        rb = dev.createBuffer(rays)
        hb = dev.createEmptyBuffer(n * 16)
        res = {}
        for kind in ("primary", "shadow"):
            if kind == "shadow":
                hits = dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)
                rays = Hh.shadow_rays_from_hits(rays, hits, (0.0, 3.0, 0.0))
                n = len(rays)
                rb = dev.createBuffer(rays)
                hb = dev.createEmptyBuffer(max(n, 1) * 16)
            best = None
            for rep in range(3):
                dev.stats(reset=True)
                if dist is not None:
                    dist.barrier()
                dev.traceDevice(hd["bvh"], hd["facesV"], hd["vertices"], rb, n, hb, any_hit=(kind == "shadow"))
                dev.finish()
                ms = dev.kernelTimeMs(hd["kernel"])
                best = ms if best is None else min(best, ms)
            st = dev.stats(reset=True).astype(np.float64)
            got = dev.readBuffer(hb, n * 16, np.uint8).view(pbr_b200.capi.HIT_DTYPE)
            n_all, ms_all = float(n), best
            if dist is not None:
                t = torch.tensor([n_all], dtype=torch.float64, device="cuda")
                dist.all_reduce(t)
                n_all = float(t.item())
                t = torch.tensor([best], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_all = float(t.item())
            entry = {"rays": int(n_all), "ms": ms_all, "mrays_per_s": n_all / ms_all / 1e3}
            if RANK == 0:
                stride = 1 if n_all <= 5_000_000 else 16
                sub = np.ascontiguousarray(rays[::stride][:400_000])
                want, _ = prep.oracle_trace(sub, any_hit=(kind == "shadow"), nthreads=THREADS)
                g = got[::stride][:400_000]
                entry["checked_rays"] = len(sub)
                entry["hit_face_leaf_t_bit_exact"] = bool(
                    np.array_equal(g["hitFace"], want["hitFace"]) and np.array_equal(g["leaf"], want["leaf"]) and
                    np.array_equal(g["t"].view(np.uint32), want["t"].view(np.uint32)))
                nodes = st[5] if kind == "shadow" else st[2]
                entry["nodes_per_ray"] = nodes / max(1.0, n)
            res[kind] = entry
        dev.freeBuffers() if False else None
        rows.append({"requested_mrays": mega, **res})
    r.close()
    return {"config": "C5 explicit rays on the 1M-triangle soup, %d rank(s), contiguous split" % WORLD, "sweep": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,c5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    only = set(args.only.split(","))
    dist = init_dist()
    cfg = host.Config()
    doc = {"world_size": WORLD, "host_cores": THREADS, "quick": args.quick}
    if "c1" in only and RANK == 0:
        doc["c1"] = run_c1(cfg, args.quick)
        print("c1", json.dumps(doc["c1"]), flush=True)
    if "c3" in only and RANK == 0:
        doc["c3"] = run_c3(cfg, args.quick)
        print("c3", json.dumps(doc["c3"]), flush=True)
    if "c4" in only:
        doc["c4"] = run_c4(cfg, args.quick, dist)
        if RANK == 0:
            print("c4", json.dumps(doc["c4"]), flush=True)
    if "c5" in only:
        doc["c5"] = run_c5(cfg, args.quick, dist)
        if RANK == 0:
            print("c5", json.dumps(doc["c5"]), flush=True)
    if RANK == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(doc, fh, indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
