#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 path-tracing core.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): synthetic 1 000 000-triangle random soup, 1920 x 1080,
16 spp rendered as 16 frames x render.samples = 1 (config.json semantics), max_depth 3 (+<= 5 added),
BRDF 1 (Shirley-Ashikhmin), no shadow rays, anti-aliasing 0.7, camera (0, 0, 3.5) looking down -z.
One STEP = one 16-spp job on every rank.

Metric: Mrays/s = (calls of traverse() + traverseShadows()) / time, whole job over all ranks.
  value : device-resident -- scene, path state and accumulation buffer stay in HBM; timed with CUDA
          events on the stream the kernels run on, barrier + synchronize on both sides, max over ranks.
  e2e   : the same job through the host API a user of the reference calls (PathTracer::generateImage
          via libpbr_host.so): every frame the accumulated image is read back into pinned host memory,
          kernel arguments (seed, weight, camera) go host -> device.
  N > 1 : samples are sharded -- every rank renders whole frames with its own seeds (weak scaling, N x
          the samples per second) and ONE NCCL all-reduce per frame forms the displayed frame
          (--shard tiles: rows are sharded instead, one all-gather per frame, strong scaling).
Extra objects: roofline (dominant kernel = traverse: algorithmic bytes / its summed CUDA-event time,
against the measured HBM peak), cpu_baseline (the CPU oracle on the host cores, bounded sample),
clocks (sampled during the timed region).

--impl reference times the reference's own algorithm on the host CPU (the oracle port -- the
reference itself cannot be built here: no OpenCL ICD, Boost, GLM, Qt) on the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (triangles, width, height, frames per step, description)
    "c2": dict(tris=1_000_000, width=1920, height=1080, spp=16, eye=(0.0, 0.0, 3.5),
               name="C2 soup-1M-tris 1920x1080 16spp"),
}
ALGO_BYTES_PER_NODE = 32      # one bvhNode: 2 x float4              (pt_bvh.cl:90)
ALGO_BYTES_PER_TRI = 64       # facesV entry + 3 vertices            (pt_intersect.cl:146-149)
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="spp", choices=["spp", "tiles"])
    # overrides for quick checks; any override is recorded in `config` and makes the run non-headline
    ap.add_argument("--tris", type=int)
    ap.add_argument("--width", type=int)
    ap.add_argument("--height", type=int)
    ap.add_argument("--spp", type=int)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify", action="store_true",
                    help="N > 1, spp sharding: rank 0 re-renders every rank's frames alone and compares the mean with "
                         "the combined image of the pipelined step")
    return ap.parse_args()


def workload(args):
    w = dict(WORKLOADS[args.workload])
    reduced = False
    for k in ("tris", "width", "height", "spp"):
        v = getattr(args, k)
        if v is not None and v != w[k]:
            w[k] = v
            reduced = True
    if reduced:
        w["name"] = "REDUCED soup-%d-tris %dx%d %dspp" % (w["tris"], w["width"], w["height"], w["spp"])
    w["reduced"] = reduced
    return w


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per traverse launch from the committed ncu --set full summary, if there is one."""
    path = os.path.join(ROOT, "profiles", "traverse_ncu_summary.json")
    try:
        with open(path) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU while the timed region runs (NVML, 100 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except Exception:
                    idx = self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop_flag.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = get_reasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report it, do not fail the run
            self.error = repr(e)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.error or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_config(cfg, w):
    cfg.reset()
    cfg.update({
        "window.width": w["width"], "window.height": w["height"],
        "camera.eye.x": w["eye"][0], "camera.eye.y": w["eye"][1], "camera.eye.z": w["eye"][2],
        "camera.center.x": 0.0, "camera.center.y": 0.0, "camera.center.z": 1.0,
        "render.samples": 1, "render.max_depth": 3, "render.max_added_depth": 5, "render.brdf": 1,
        "render.shadow_rays": 0, "render.antialiasing": 0.7, "render.phong_tessellation": 0.0,
        "bvh.max_faces": 2, "bvh.sah_faces_limit": 100000, "bvh.skip_ahead": True, "bvh.skip_ahead_compare": 0.7,
        "logging.level": 1,
    })


# ------------------------------------------------------------------------------------------------ ours

def run_ours(args):
    import torch
    import torch.distributed as dist
    import pbr_b200
    from pbr_b200 import host, multigpu, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the path tracer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev_t)
    w = workload(args)
    W, H, SPP = w["width"], w["height"], w["spp"]

    cfg = host.Config()
    host_config(cfg, w)
    scene = scenes.soup(w["tris"], seed=12345)
    r = host.Renderer(local_rank)
    t0 = time.perf_counter()
    r.load_scene(scene)
    load_s = time.perf_counter() - t0
    info = r.info()
    r.set_deterministic(True)
    dev = r.device()
    dev.setStream(torch.cuda.current_stream().cuda_stream)
    dev.profileEnable(True)

    tiles = args.shard == "tiles" and world > 1
    if world > 1 and not tiles:
        r.set_seed_schedule(world, rank)
    if tiles:
        y0, y1 = multigpu.tile_rows(H, rank, world)
        r.set_tile(y0, y1)

    views = {}

    def image_tensor():
        _, hd = r.handles()
        ptr, _ = dev.devicePtr(hd["image"])
        if ptr not in views:
            views[ptr] = multigpu.DeviceImage(ptr, H, W, dev_t).tensor
        return views[ptr]

    display = torch.empty((H, W, 4), dtype=torch.float32, device=dev_t) if world > 1 else None
    # N > 1: the cross-rank combine of frame k (a copy + one collective) runs on its own stream while frame k+1
    # is traced.  Frame k+1 only READS the image frame k wrote; frame k+2 overwrites it, so it waits for the
    # combine of frame k (two events, by parity).
    render_stream = torch.cuda.current_stream()
    display_stream = torch.cuda.Stream() if world > 1 else None
    combined = [None, None]
    frame_no = [0]

    def frame():
        j = frame_no[0]
        frame_no[0] += 1
        if world > 1 and combined[j & 1] is not None:
            render_stream.wait_event(combined[j & 1])
        r.render_frames(1)
        if world > 1:
            img = image_tensor()
            rendered = torch.cuda.Event()
            rendered.record(render_stream)
            with torch.cuda.stream(display_stream):
                display_stream.wait_event(rendered)
                if tiles:
                    multigpu.combine_tiles(img, rank, world)
                else:
                    multigpu.combine_spp(img, world, out=display)
                combined[j & 1] = torch.cuda.Event()
                combined[j & 1].record(display_stream)
            # (rows: the next frame only needs this rank's own rows of the previous image, which it wrote
            #  itself; the gathered rows are for display)

    def step_resident():
        r.reset_sample_count()
        if world == 1:
            r.render_frames(SPP)         # one pbr_kernel_launch_batch call, the frames run back to back
        else:
            for _ in range(SPP):
                frame()                  # every frame is combined across ranks (progressive display)
            render_stream.wait_stream(display_stream)

    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    pinned_np = pinned.numpy()
    pinned2 = [pinned, torch.empty((H, W, 4), dtype=torch.float32).pin_memory()] if world > 1 and rank == 0 else None
    landed = [None, None]
    delivered = [0]

    def step_e2e(last=False):
        if world > 1:
            r.reset_sample_count()
        for i in range(SPP):
            if world == 1:
                # one generateImage() per frame, every frame read back into pinned memory; with render-ahead the
                # next frame is traced while this one is copied.  The accumulation simply continues from step to
                # step (no reset: a reset would discard the frame traced ahead), and the very last call of the
                # run switches render-ahead off first, so no frame is traced that is not delivered.
                if last and i == SPP - 1:
                    r.set_render_ahead(False)
                r.generate_image(out=pinned_np)
            else:
                frame()
                if rank == 0:
                    # one consumer of the displayed frame: rank 0's host.  The copy of frame j is enqueued behind its
                    # combine; the host then waits for the copy of frame j - 1, so that it never stalls the launches
                    # of the frame being traced (every frame is delivered, one frame later; two pinned buffers).
                    j = delivered[0]
                    delivered[0] += 1
                    with torch.cuda.stream(display_stream):
                        pinned2[j & 1].copy_(display if not tiles else image_tensor(), non_blocking=True)
                        landed[j & 1] = torch.cuda.Event()
                        landed[j & 1].record(display_stream)
                    if tiles:
                        combined[(frame_no[0] - 1) & 1] = landed[j & 1]   # the copy reads the image frame + 2 overwrites
                    if landed[(j - 1) & 1] is not None:
                        landed[(j - 1) & 1].synchronize()
        if world > 1:
            render_stream.wait_stream(display_stream)
            if rank == 0:
                display_stream.synchronize()                 # the last frame of the step has landed too

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(a):
        if world == 1:
            return a
        t = torch.tensor(np.asarray(a, np.float64), device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    dev.stats(reset=True)
    dev.profileRead(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    start.record()
    for _ in range(args.steps):
        step_resident()
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_local = start.elapsed_time(end)
    clocks = sampler.result()
    ms_total = max_over_ranks(ms_local)
    stats_local = dev.stats(reset=True).astype(np.float64)
    prof = dev.profileRead(reset=True)
    stats_all = sum_over_ranks(stats_local)
    launches_all = float(sum_over_ranks(np.array([prof["launches"]], np.float64))[0])
    rays_all = stats_all[0] + stats_all[1]
    value = rays_all / (ms_total * 1e-3) / 1e6
    frames_all = args.steps * SPP * (1 if tiles else world)
    samples_per_s = frames_all * W * H / (ms_total * 1e-3)

    # roofline of the dominant kernel (this rank's traverse launches)
    peak, peak_src = measured_peak()
    algo_bytes = ALGO_BYTES_PER_NODE * stats_local[2] + ALGO_BYTES_PER_TRI * stats_local[3]
    trav_ms = prof["traverse_ms"]
    achieved = algo_bytes / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": "traverseKernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": ncu_traffic(), "peak_source": peak_src,
        "launches": int(prof["traverse_launches"]),
        "avg_launch_ms": round(trav_ms / max(1, prof["traverse_launches"]), 4),
        "algorithmic_bytes_per_launch": round(algo_bytes / max(1, prof["traverse_launches"])),
        "kernel_share_of_step": round(trav_ms / ms_local, 4),
        "shade_share_of_step": round(prof["shade_ms"] / ms_local, 4),
        "nodes_per_ray": round(stats_local[2] / max(1.0, stats_local[0]), 2),
        "tri_tests_per_ray": round(stats_local[3] / max(1.0, stats_local[0]), 2),
    }

    # ---- end to end through the host API ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            r.reset_sample_count()
        for i in range(max(1, args.warmup // 2)):
            step_e2e(last=True)
        sync_all()
        dev.stats(reset=True)
        if world == 1:
            r.set_render_ahead(True)
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(last=(i == args.steps - 1))
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e_stats = sum_over_ranks(dev.stats(reset=True).astype(np.float64))
        e2e = {
            "value": round((e_stats[0] + e_stats[1]) / e2e_s / 1e6, 2), "unit": "Mrays/s",
            "h2d_bytes_per_step": SPP * (4 + 4 + 80),          # seed, pixelWeight, camera per frame
            "d2h_bytes_per_step": SPP * W * H * 16,            # accumulated frame per frame (N > 1: rank 0 reads it)
            "ms_per_step": round(e2e_s * 1e3 / args.steps, 3),
            "api": "PathTracer::generateImage (libpbr_host.so) per frame, every frame into a pinned host image, "
                   "setRenderAhead(true): the next frame is traced while this one is copied" if world == 1 else
                   "PathTracer::renderFrames(1) per rank + combine on a side stream; rank 0 reads every combined "
                   "frame into pinned host memory (double-buffered: it waits for frame j - 1 while frame j + 1 is traced)",
        }

    # ---- optional: is the pipelined multi-GPU image the right one? ---------------------------------
    verify = None
    if args.verify and world > 1 and tiles:
        step_resident()
        sync_all()
        if rank == 0:
            shown = image_tensor().clone()
            r.set_tile(0, H)
            r.reset_sample_count()
            r.render_frames(SPP)
            torch.cuda.synchronize()
            want = image_tensor()
            same = (shown == want) | (torch.isnan(shown) & torch.isnan(want))
            verify = {"bit_identical_pixels": int(same.all(dim=2).sum()), "pixels": W * H,
                      "what": "all-gathered image of one pipelined step vs the same frames rendered whole on rank 0"}
            y0, y1 = multigpu.tile_rows(H, rank, world)
            r.set_tile(y0, y1)
        sync_all()
    if args.verify and world > 1 and not tiles:
        step_resident()
        sync_all()
        if rank == 0:
            shown = display.clone()
            acc = torch.zeros_like(shown, dtype=torch.float64)
            for rr in range(world):
                r.set_seed_schedule(world, rr)
                r.reset_sample_count()
                r.render_frames(SPP)
                torch.cuda.synchronize()
                acc += image_tensor().double()
            r.set_seed_schedule(world, rank)
            want = (acc / world)[..., :3]
            got = shown[..., :3].double()
            ok = torch.isfinite(want) & torch.isfinite(got)
            diff = (want - got).abs()[ok]
            verify = {"max_abs_diff": float(diff.max()), "mean": float(want[ok].mean()),
                      "pixels_compared": int(ok.all(dim=2).sum()), "what": "combined image of one pipelined step vs "
                      "the mean of all ranks' accumulations re-rendered on rank 0"}
        sync_all()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, cpu_frames, cpu_image = cpu_baseline_sample(w, r.flat(), scene)
        # parity at the bench configuration itself: the frames the CPU arm has just rendered with the reference
        # kernel (same seeds, same accumulation from a black image) against the same frames on the GPU, bit for bit
        r.set_render_ahead(False)
        r.reset_sample_count()
        r.render_frames(cpu_frames)
        gpu_image = r.read_image()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers as Hh
        verify = {
            "frames": cpu_frames, "pixels": W * H,
            "bit_identical_pixels": Hh.count_identical_pixels(gpu_image, cpu_image),
            "mre": Hh.mean_relative_error(gpu_image, cpu_image),
            "what": "accumulated image after %d frame(s): GPU (%s) vs the CPU arm's %s kernel, all four channels" % (
                cpu_frames, "reference-order walk", cpu_baseline["kind"]),
        }

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": w["name"], "triangles": w["tris"], "width": W, "height": H, "spp_per_step": SPP,
                "frames_per_step_per_rank": SPP, "max_depth": 3, "max_added_depth": 5, "brdf": 1,
                "sharding": ("tiles+allgather" if tiles else "spp+allreduce") if world > 1 else "none",
                "bvh_nodes": info["emitted_nodes"], "bvh_build_s": round(info["bvh_build_seconds"], 2),
                "scene_load_s": round(load_s, 2),
                "pipeline": {0: "wavefront", 1: "megakernel", 2: "persistent", 3: "carry-over", -1: "undecided"}[dev.pipelineInUse()] +
                       " (chosen by measurement on the first frames)",
                       "l2_policy": "working set > L2: nodes+tris %.0f MB, path state %.0f MB, images %.0f MB" % (
                    (info["emitted_nodes"] * 32 + info["faces"] * 36) / 1e6, W * H * 104 / 1e6, W * H * 48 / 1e6),
            },
            "samples_per_s": round(samples_per_s), "rays_per_step": round(rays_all / args.steps),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_all), "verify": verify,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def _c2_defines(w, num_nodes, sky):
    from oracle import scene as S
    return S.defines(w["width"], w["height"], num_nodes, 0, sky, brdf=1, samples=1, max_depth=3, max_added_depth=5,
                     shadow_rays=0, antialiasing=0.7)


def reference_program_values(w=None):
    """The program text values (CL::setValues) of the workload the reference arm runs: lets
    oracle/build_ref.py prebuild the reference kernel for it where /root/reference exists."""
    from oracle import oracle as O
    from oracle import ref as R
    from oracle import scene as S
    import pbr_b200
    w = dict(WORKLOADS["c2"]) if w is None else w
    scene = pbr_b200.scenes.soup(w["tris"], seed=12345)
    bvh = O.build_bvh(scene)
    _, sky = S.pack_materials(scene["materials"], scene["materialNames"], 1)
    return R.values_from_defines(_c2_defines(w, bvh["nodes"].shape[0], sky))


def cpu_frame_runner(D, px, cam, nodes, facesV, facesN, v4, mats, threads):
    """One frame of the path on the host CPU.  Prefers the reference's own kernel (oracle/_ref, built from
    the reference's sources by oracle/build_ref.py: kind "reference"); without it, the restatement
    (oracle/pt_oracle.cpp: kind "port").  Returns (frame(k, image) -> image, count_rays(k, image) -> int, kind).
    The ray count of a frame comes from the restatement's counters (same pixels, same rays) and is taken
    outside the timed region."""
    from oracle import oracle as O
    from oracle import ref as R
    from oracle import scene as S

    def port(k, img):
        out, _, st = O.path_tracing(D, S.frame_seed(k), S.pixel_weight(k), px, cam, nodes, facesV, facesN, v4, None,
                                    mats, None, img, nthreads=threads, debug=False)
        return out, int(st[0]) + int(st[1])

    if R.available(D):
        R.lib(D)                                                 # build / load outside the timed region

        def frame(k, img):
            out, _ = R.path_tracing(D, S.frame_seed(k), S.pixel_weight(k), px, cam, nodes, facesV, facesN, v4, None,
                                    mats, None, img, nthreads=threads)
            return out
        return frame, (lambda k, img: port(k, img)[1]), "reference"
    return (lambda k, img: port(k, img)[0]), (lambda k, img: port(k, img)[1]), "port"


def cpu_baseline_sample(w, flat, scene, budget_s=12.0):
    """The reference kernel on all host cores (see cpu_frame_runner) for a bounded sample of the same
    workload: whole frames of the full image at 1 spp for about budget_s."""
    from oracle import scene as S
    threads = os.cpu_count() or 1
    W, H = w["width"], w["height"]
    v4 = S.pack_float4(scene["vertices"])
    mats, sky = S.pack_materials(scene["materials"], scene["materialNames"], 1)
    D = _c2_defines(w, flat["nodes"].shape[0], sky)
    cam = S.camera(eye=w["eye"])
    px = S.px_dim(W, H)
    frame, count_rays, kind = cpu_frame_runner(D, px, cam, flat["nodes"], flat["facesV"], flat["facesN"], v4, mats, threads)
    img = np.zeros((H, W, 4), np.float32)
    inputs = []
    frames = 0
    t0 = time.perf_counter()
    while frames < 1 or (time.perf_counter() - t0) * (frames + 1) / frames < budget_s:
        inputs.append(img)
        img = frame(frames, img)
        frames += 1
    sec = time.perf_counter() - t0
    rays = sum(count_rays(k, inputs[k]) for k in range(frames))
    return {"value": round(rays / sec / 1e6, 3), "unit": "Mrays/s", "cores": threads, "kind": kind,
            "sample": "%d frame(s) of the full %dx%d image at 1 spp (%d rays) in %.1f s" % (frames, W, H, rays, sec)}, \
        frames, img


# ------------------------------------------------------------------------------------------- reference

def run_reference(args):
    """The reference's kernel on the host CPU (cpu_frame_runner), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from oracle import scene as S
    import pbr_b200
    w = workload(args)
    W, H = w["width"], w["height"]
    threads = os.cpu_count() or 1
    scene = pbr_b200.scenes.soup(w["tris"], seed=12345)         # input data only
    bvh = O.build_bvh(scene)                                    # the oracle's own (literal) BVH builder
    v4 = S.pack_float4(scene["vertices"])
    mats, sky = S.pack_materials(scene["materials"], scene["materialNames"], 1)
    D = _c2_defines(w, bvh["nodes"].shape[0], sky)
    cam = S.camera(eye=w["eye"])
    px = S.px_dim(W, H)
    frame, count_rays, kind = cpu_frame_runner(D, px, cam, bvh["nodes"], bvh["facesV"], bvh["facesN"], v4, mats, threads)
    img = np.zeros((H, W, 4), np.float32)

    k = 0
    for _ in range(args.warmup):
        img = frame(k, img)
        k += 1
    inputs = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        inputs.append((k, img))
        img = frame(k, img)
        k += 1
    sec = time.perf_counter() - t0
    rays = sum(count_rays(kk, im) for kk, im in inputs)
    value = round(rays / sec / 1e6, 3)
    sample = "each step = 1 frame of the full %dx%d image at 1 spp (1/%d of the GPU arm's step)" % (W, H, w["spp"])
    note = ("the reference's own kernel source (source/opencl/pathtracing.cl + pt_*.cl) compiled for the host by "
            "oracle/build_ref.py, one work-item after the other on all host threads" if kind == "reference" else
            "CPU restatement of the reference kernels (oracle port); oracle/_ref for this configuration was not "
            "prebuilt and /root/reference is not present")
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3 / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "triangles": w["tris"], "width": W, "height": H,
                   "spp_per_step": 1, "max_depth": 3, "max_added_depth": 5, "brdf": 1, "note": note},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
