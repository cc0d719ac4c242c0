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
  N > 1 : PathTracer::setRanks -- samples are sharded, every rank renders whole frames with its own seeds
          (weak scaling, N x the samples per second) and ONE NCCL all-reduce per frame, issued by the
          library itself, forms the delivered frame.  `strong` adds the strong-scaling view: ONE image,
          rows sharded in stripes, one all-gather per frame, against rank 0 rendering it alone.
Extra objects: roofline (dominant kernel = traverse: the bytes it asks for / its summed CUDA-event time,
against the measured HBM peak; ncu counters of the SAME build when committed), reference_walk (the step
with the reference's visiting order), verify (N = 1: the GPU's frames against the CPU arm's, bit for bit),
cpu_baseline (the reference's own kernel on the host cores, bounded sample), clocks.

--impl reference times the reference's own kernel source (oracle/_ref: source/opencl/*.cl compiled for
the host by oracle/build_ref.py; the restatement in oracle/ when that is missing) on all host cores.
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
    ap.add_argument("--traversal", type=int, default=-1, choices=[-1, 0, 1],
                    help="pbr_set_traversal: -1 automatic (the ordered walk, since no debug image is requested), 0 the "
                         "reference's visiting order, 1 the ordered walk")
    ap.add_argument("--no-reference-walk", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-strong-c4", action="store_true")
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


def count_identical_pixels(a, b):
    """Pixels whose four channels are bit-identical (NaN == NaN)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return int(same.reshape(-1, a.shape[-1]).all(axis=1).sum())


def mean_relative_error(a, b):
    """mean(|a-b| / (max(a,b) + 1e-3)) over RGB of pixels finite on both sides."""
    a, b = np.asarray(a, np.float64)[..., :3], np.asarray(b, np.float64)[..., :3]
    ok = np.isfinite(a) & np.isfinite(b)
    return float((np.abs(a - b)[ok] / (np.maximum(a, b)[ok] + 1e-3)).mean())


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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

WALK_NAMES = {0: "reference-order walk (stackless pre-order, pt_bvh.cl:82-123)",
              1: "ordered walk over the 4-wide BVH collapsed from the uploaded node array (same hits, same image bits)"}


def ncu_summary(build_id):
    """Counters of the dominant kernel from the committed `ncu --set full` summary -- only when it was taken of the
    very build that is being timed (pbr_build_id)."""
    path = os.path.join(ROOT, "profiles", "traverse_ncu_summary.json")
    try:
        with open(path) as fh:
            doc = json.load(fh)
    except Exception:
        return None, "no profiles/traverse_ncu_summary.json"
    if doc.get("build_id") != build_id:
        return None, "profiles/traverse_ncu_summary.json is of build %s, this library is %s" % (doc.get("build_id"), build_id)
    return doc, None


def time_steps(torch, dist, world, step, steps, fence):
    """K steps on the device: CUDA events on the stream the kernels run on, barrier + synchronize on both sides."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        step()
    fence()
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return start.elapsed_time(end)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pbr_b200
    from pbr_b200 import capi, host, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the path tracer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    if world > 1:
        # torch.distributed is plumbing here: barrier, max over ranks, handing the NCCL id around.  The per-frame
        # collective on the image is the library's own (pbr_frame_combine inside PathTracer::setRanks).
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev_t)
    w = workload(args)
    W, H, SPP = w["width"], w["height"], w["spp"]
    tiles = args.shard == "tiles" and world > 1

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(a):
        if world == 1:
            return np.asarray(a, np.float64)
        t = torch.tensor(np.asarray(a, np.float64), device=dev_t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def new_comm_id():
        box = [host.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def make_renderer(scene, sharding):
        r = host.Renderer(local_rank)
        t0 = time.perf_counter()
        r.load_scene(scene)
        load_s = time.perf_counter() - t0
        r.set_deterministic(True)
        r.set_traversal(args.traversal)
        d = r.device()
        d.setStream(torch.cuda.current_stream().cuda_stream)
        if world > 1:
            r.set_ranks(rank, world, new_comm_id(), sharding)
        return r, d, load_s

    cfg = host.Config()
    host_config(cfg, w)
    scene = scenes.soup(w["tris"], seed=12345)
    r, dev, load_s = make_renderer(scene, "stripes" if tiles else "spp")
    info = r.info()

    def step_resident():
        r.reset_sample_count()
        r.render_frames(SPP)             # N = 1: one pbr_kernel_launch_batch; N > 1: every frame ends with the collective

    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    pinned_np = pinned.numpy()

    FRAMES_IN_FLIGHT = 4                 # the library's default: consecutive frames of a batch traced concurrently
    AHEAD = 3                            # frames traced beyond the one being copied out

    def run_e2e(steps):
        # one generateImage() per frame, every frame read back into pinned host memory; with render-ahead the next
        # frames are traced while this one is copied.  The accumulation simply continues from step to step (a reset
        # would discard the frames traced ahead), and render-ahead is switched off AHEAD calls before the end, so that
        # exactly steps * SPP frames are traced and every one of them is delivered.  N > 1: rank 0's host is the one
        # consumer of the combined frame.
        total = steps * SPP
        r.set_render_ahead(AHEAD)
        for j in range(total):
            if rank == 0:
                if j == total - AHEAD:
                    r.set_render_ahead(0)
                r.generate_image(out=pinned_np)
            else:
                r.render_frames(1)
        r.set_render_ahead(0)

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    r.finish()
    dev.stats(reset=True)
    dev.profileRead(reset=True)
    dev.traversalInfo(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_local = time_steps(torch, dist, world, step_resident, args.steps, r.comm_fence)
    clocks = sampler.result()
    ms_total = max_over_ranks(ms_local)
    stats_timed = dev.stats(reset=True).astype(np.float64)
    prof_timed = dev.profileRead(reset=True)
    tinfo = dev.traversalInfo(reset=True)
    stats_all = sum_over_ranks(stats_timed)
    launches_all = float(sum_over_ranks(np.array([prof_timed["launches"]], np.float64))[0])
    rays_all = stats_all[0] + stats_all[1]
    value = rays_all / (ms_total * 1e-3) / 1e6
    frames_all = args.steps * SPP * (1 if tiles else world)
    samples_per_s = frames_all * W * H / (ms_total * 1e-3)
    walk = int(tinfo["last_used"])
    pipeline = dev.pipelineInUse()

    # ---- roofline of the dominant kernel (this rank's traverse launches) -------------------------
    # The timed region above keeps several frames in flight: kernels of consecutive frames share the device, so the
    # duration of ONE launch cannot be read off there.  The same step is therefore run once more with one frame after
    # the other (frames_in_flight = 1) and every launch bracketed by CUDA events on its stream.
    dev.setTuning("frames_in_flight", 1)
    dev.profileEnable(True)
    step_resident()
    r.finish()
    dev.stats(reset=True)
    dev.profileRead(reset=True)
    k_serial = max(3, args.steps // 3)
    ms_serial = time_steps(torch, dist, world, step_resident, k_serial, r.comm_fence)
    stats_local = dev.stats(reset=True).astype(np.float64)
    prof = dev.profileRead(reset=True)
    dev.profileEnable(False)
    dev.setTuning("frames_in_flight", FRAMES_IN_FLIGHT)
    peak, peak_src = measured_peak()
    if walk == 1:
        kernel = "traverseWideKernel"
        bytes_node, bytes_tri = 128, 36        # what the ordered walk asks for: one 128-byte wide node per visit, 36 B per face
    else:
        kernel = "traverseKernel"
        bytes_node, bytes_tri = ALGO_BYTES_PER_NODE, ALGO_BYTES_PER_TRI
    algo_bytes = bytes_node * stats_local[2] + bytes_tri * stats_local[3]
    trav_ms = prof["traverse_ms"]
    achieved = algo_bytes / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
    build = capi.build_id()
    ncu, ncu_note = ncu_summary(build)
    per_kernel = (ncu or {}).get("kernels", {}).get(kernel)
    roofline = {
        "bound": "hbm", "bound_note": "nominally HBM (SURVEY.md 8d); in fact the walk is served from L1 / L2 and limited by the L1 "
                 "load pipe (LSU wavefronts) and issue slots -- see traffic, l2_hit, lsu_wavefront_pct",
        "kernel": kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "peak_source": peak_src,
        "traffic": per_kernel.get("dram_bytes_per_launch") if per_kernel else None,
        "l1_hit": per_kernel.get("l1_hit_pct") if per_kernel else None,
        "l2_hit": per_kernel.get("l2_hit_pct") if per_kernel else None,
        "lsu_wavefront_pct": per_kernel.get("lsu_wavefront_pct") if per_kernel else None,
        "issue_active_pct": per_kernel.get("issue_active_pct") if per_kernel else None,
        "threads_per_inst": per_kernel.get("threads_per_inst") if per_kernel else None,
        "ncu_build_id": (ncu or {}).get("build_id"), "ncu_note": ncu_note, "build_id": build,
        "algorithmic_bytes": "%d B x node visits + %d B x triangle tests, both counted by the kernel" % (bytes_node, bytes_tri),
        "launches": int(prof["traverse_launches"]),
        "avg_launch_ms": round(trav_ms / max(1, prof["traverse_launches"]), 4),
        "algorithmic_bytes_per_launch": round(algo_bytes / max(1, prof["traverse_launches"])),
        "measured_on": "%d steps with one frame after the other (frames_in_flight = 1): %.3f ms per step, %.1f Mrays/s" % (
            k_serial, ms_serial / k_serial, (stats_local[0] + stats_local[1]) / (ms_serial * 1e-3) / 1e6),
        "kernel_share_of_step": round(trav_ms / ms_serial, 4),
        "shade_share_of_step": round(prof["shade_ms"] / ms_serial, 4),
        "nodes_per_ray": round(stats_local[2] / max(1.0, stats_local[0]), 2),
        "tri_tests_per_ray": round(stats_local[3] / max(1.0, stats_local[0]), 2),
        "rewalked_rays": int(tinfo["rewalked_rays"]),
    }

    # ---- the same step with the reference-order walk (N = 1): what the ordered walk replaces ------
    reference_walk = None
    if world == 1 and walk == 1 and not args.no_reference_walk:
        r.set_traversal(0)
        for _ in range(2):
            step_resident()
        r.finish()
        dev.stats(reset=True)
        k = max(3, args.steps // 4)
        ms_ref = time_steps(torch, dist, world, step_resident, k, r.comm_fence)
        st = dev.stats(reset=True).astype(np.float64)
        dev.setTuning("frames_in_flight", 1)
        dev.profileEnable(True)
        step_resident()
        r.finish()
        dev.stats(reset=True)
        dev.profileRead(reset=True)
        time_steps(torch, dist, world, step_resident, k, r.comm_fence)
        st1 = dev.stats(reset=True).astype(np.float64)
        pr = dev.profileRead(reset=True)
        dev.profileEnable(False)
        dev.setTuning("frames_in_flight", FRAMES_IN_FLIGHT)
        ref_bytes = ALGO_BYTES_PER_NODE * st1[2] + ALGO_BYTES_PER_TRI * st1[3]
        reference_walk = {
            "value": round((st[0] + st[1]) / (ms_ref * 1e-3) / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(ms_ref / k, 3),
            "steps": k, "kernel": "traverseKernel", "nodes_per_ray": round(st[2] / max(1.0, st[0]), 2),
            "tri_tests_per_ray": round(st[3] / max(1.0, st[0]), 2),
            "algorithmic_gbs": round(ref_bytes / (pr["traverse_ms"] * 1e-3) / 1e9, 1),
            "frac_of_peak": round(ref_bytes / (pr["traverse_ms"] * 1e-3) / 1e9 / peak, 4),
            "what": "pbr_set_traversal(0): visit counters and debug image bit-exact as well; algorithmic_gbs = 32 B x nodes + 64 B x "
                    "triangle tests (SURVEY.md 8d) over the summed traverseKernel time of steps run one frame after the other",
        }
        r.set_traversal(args.traversal)

    # ---- end to end through the host API ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        r.reset_sample_count()
        run_e2e(max(1, args.warmup // 2))
        r.finish()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dev.stats(reset=True)
        t0 = time.perf_counter()
        run_e2e(args.steps)
        r.finish()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e_stats = sum_over_ranks(dev.stats(reset=True).astype(np.float64))
        e2e = {
            "value": round((e_stats[0] + e_stats[1]) / e2e_s / 1e6, 2), "unit": "Mrays/s",
            "h2d_bytes_per_step": SPP * (4 + 4 + 80),          # seed, pixelWeight, camera per frame
            "d2h_bytes_per_step": SPP * W * H * 16,            # accumulated frame per frame (N > 1: rank 0 reads it)
            "ms_per_step": round(e2e_s * 1e3 / args.steps, 3),
            "api": "PathTracer::generateImage (libpbr_host.so) per frame, every frame into a pinned host image, "
                   "setRenderAhead(%d): the next frames are traced while this one is copied" % AHEAD + (
                       "; N > 1: PathTracer::setRanks, every frame ends with the library's own collective, rank 0's host "
                       "receives every combined frame" if world > 1 else ""),
        }

    # ---- strong scaling of ONE image (N > 1): rows of every frame sharded in interleaved stripes ----
    strong = None
    if world > 1 and not args.no_strong:
        strong = {}
        try:
            strong["c2"] = strong_scaling(torch, dist, world, rank, r, dev, W, H, SPP, max(3, args.steps // 2), max_over_ranks,
                                          sum_over_ranks, w["name"], already_stripes=tiles)
            if not args.no_strong_c4 and not w["reduced"]:
                r.close()
                r = None
                c4 = dict(width=3840, height=2160, spp=8)
                cfg.reset()
                cfg.update({"window.width": c4["width"], "window.height": c4["height"], "render.max_depth": 3, "logging.level": 1,
                            "camera.eye.x": 0.0, "camera.eye.y": 1.2, "camera.eye.z": 1.8,
                            "camera.center.x": 0.0, "camera.center.y": 0.55, "camera.center.z": 1.0})
                grid = scenes.displaced_grid(2237, 2236, patches=8)
                r, dev, _ = make_renderer(grid, "stripes")
                strong["c4"] = strong_scaling(torch, dist, world, rank, r, dev, c4["width"], c4["height"], c4["spp"], 3, max_over_ranks,
                                              sum_over_ranks, "C4 displaced-grid-10M-tris 3840x2160 %dspp" % c4["spp"], already_stripes=True)
        except Exception as e:          # the headline line must not die with an optional measurement
            strong["error"] = repr(e)

    # ---- CPU baseline + parity at the bench configuration (rank 0, N = 1 only) ---------------------
    cpu_baseline = None
    verify = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, cpu_frames, cpu_image = cpu_baseline_sample(w, r.flat(), scene)
        # parity at the bench configuration itself: the frames the CPU arm has just rendered with the reference
        # kernel (same seeds, same accumulation from a black image) against the same frames on the GPU, bit for bit
        verify = {"frames": cpu_frames, "pixels": W * H, "against": "the CPU arm's %s kernel, all four channels" % cpu_baseline["kind"]}
        for mode, name in ((args.traversal, "as_timed"), (0, "reference_order_walk")):
            r.set_traversal(mode)
            r.reset_sample_count()
            r.render_frames(cpu_frames)
            gpu_image = r.read_image()
            verify[name] = {"bit_identical_pixels": count_identical_pixels(gpu_image, cpu_image),
                            "mre": mean_relative_error(gpu_image, cpu_image),
                            "walk": WALK_NAMES[int(dev.traversalInfo()["last_used"])]}
        r.set_traversal(args.traversal)

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": w["name"], "triangles": w["tris"], "width": W, "height": H, "spp_per_step": SPP,
                "frames_per_step_per_rank": SPP, "max_depth": 3, "max_added_depth": 5, "brdf": 1,
                "sharding": ("stripes+allgather" if tiles else "spp+allreduce") + " inside libpbr_b200.so (PathTracer::setRanks)"
                if world > 1 else "none",
                "bvh_nodes": info["emitted_nodes"], "bvh_build_s": round(info["bvh_build_seconds"], 2),
                "scene_load_s": round(load_s, 2),
                "traversal": WALK_NAMES[walk] + ("; chosen automatically: no debug image is requested" if args.traversal < 0 else "; forced"),
                "wide_bvh": {"nodes": tinfo["wide_nodes"], "depth": tinfo["wide_depth"], "staged_in_shared_memory": tinfo["wide_top"],
                             "build_ms": round(tinfo["wide_build_ms"], 1)} if walk == 1 else None,
                "frames_in_flight": FRAMES_IN_FLIGHT,
                "pipeline": {0: "wavefront", 1: "megakernel", -1: "wavefront (batched frames)"}[pipeline] +
                            " (chosen by measurement on the first frames)",
                "l2_policy": "working set > L2: scene %.0f MB (wide nodes + triangles%s), path state %.0f MB, images %.0f MB" % (
                    (tinfo["wide_nodes"] * 128 + info["faces"] * 36) / 1e6 if walk == 1 else (info["emitted_nodes"] * 32 + info["faces"] * 36) / 1e6,
                    "" if walk == 1 else "; reference nodes", W * H * 104 / 1e6, W * H * 48 / 1e6),
            },
            "samples_per_s": round(samples_per_s), "rays_per_step": round(rays_all / args.steps),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_all), "verify": verify,
            "roofline": roofline, "reference_walk": reference_walk, "strong": strong, "cpu_baseline": cpu_baseline,
        }
        emit(line)
    if r is not None:
        r.close()
    if world > 1:
        dist.destroy_process_group()


def strong_scaling(torch, dist, world, rank, r, dev, W, H, spp, steps, max_over_ranks, sum_over_ranks, name, already_stripes):
    """ONE image rendered by all ranks together: every frame's rows sharded in interleaved stripes, completed with one
    all-gather per frame (bit-identical to one GPU) -- against the same frames rendered by rank 0 alone."""
    sharding = "stripes" if H % world == 0 else "rows"
    if not already_stripes:
        r.set_sharding(sharding)

    def step():
        r.reset_sample_count()
        r.render_frames(spp)

    for _ in range(2):
        step()
    r.finish()
    dev.stats(reset=True)
    ms = max_over_ranks(time_steps(torch, dist, world, step, steps, r.comm_fence))
    st = sum_over_ranks(dev.stats(reset=True).astype(np.float64))
    r.finish()
    shown = r.read_image() if rank == 0 else None
    dist.barrier()
    # the same job on one GPU (rank 0 alone; the others wait)
    out = {"workload": name, "sharding": sharding + " + one all-gather per frame (pbr_frame_combine)", "steps": steps,
           "ms_per_step": round(ms / steps, 3), "value": round((st[0] + st[1]) / (ms * 1e-3) / 1e6, 2), "unit": "Mrays/s"}
    r.set_sharding("none")
    if rank == 0:
        for _ in range(2):
            step()
        r.finish()
        ms1 = time_steps(torch, None, 1, step, steps, lambda: None)
        whole = r.read_image()
        out.update({
            "one_gpu_ms_per_step": round(ms1 / steps, 3), "speedup": round(ms1 / ms, 3), "efficiency": round(ms1 / ms / world, 4),
            "bit_identical_pixels": count_identical_pixels(shown, whole), "pixels": W * H,
        })
    dist.barrier()
    r.set_sharding(sharding)
    return out


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
    emit(line)


_STDOUT_FD = None


def emit(line):
    """The ONE JSON line, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    # Libraries talk on fd 1 (NCCL prints its version there when a communicator is created): everything but the
    # result line goes to stderr.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
