"""Turn the ncu captures in gpurun_out/ into the small, committed summaries under profiles/.
    python scripts/summarize_ncu.py r02
Reads  gpurun_out/<tag>_build_id.txt            (pbr_build_id() of the library that was profiled)
       gpurun_out/<tag>_launches.csv            (ncu --metrics gpu__time_duration.sum launch list, ordered walk)
       gpurun_out/<tag>_ref_launches.csv        (the same with PBR_TRAVERSAL=0: reference-order walk)
       gpurun_out/<tag>_traverse.ncu-rep        (ncu --set full, traverseWideKernel)
       gpurun_out/<tag>_ref_traverse.ncu-rep    (ncu --set full, traverseKernel)
       gpurun_out/<tag>_shade.ncu-rep           (ncu --set full, shadeKernel)  [optional]
Writes profiles/<tag>_*launches.csv, profiles/<tag>_*_shares.json, profiles/<tag>_*_ncu.json,
       profiles/traverse_ncu_summary.json       (what bench.py reads for roofline.traffic & co, keyed by build id)"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value) * scale


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[header.index("Kernel Name")][:80]}
        for m in METRICS:
            if m in header:
                i = header.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if m.startswith("dram__bytes"):
                    v = to_bytes(v, units[i])
                elif m == "gpu__time_duration.sum":
                    v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1.0)
                    m = "gpu__time_duration.ms"
                d[m] = v
        launches.append(d)
    return launches


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    total, per = 0.0, {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").strip()
        ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
        per.setdefault(name, [0, 0.0])
        per[name][0] += 1
        per[name][1] += ns
        total += ns
    return {"note": "ncu launch list: cold-cache, serialised; compare SHARES, not absolutes",
            "total_ms": total * 1e-6,
            "kernels": {k: {"launches": n, "ms": t * 1e-6, "share": t / total} for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])}}


build_id = None
bid = os.path.join(G, tag + "_build_id.txt")
if os.path.exists(bid):
    build_id = open(bid).read().strip()

for name in ("launches", "ref_launches"):
    src = os.path.join(G, "%s_%s.csv" % (tag, name))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_%s.csv" % (tag, name)))
        doc = launch_shares(src)
        doc["build_id"] = build_id
        json.dump(doc, open(os.path.join(P, "%s_%s_shares.json" % (tag, name)), "w"), indent=1)


def mean(launches, key):
    v = [l[key] for l in launches if key in l]
    return sum(v) / len(v) if v else None


summary = {"tag": tag, "build_id": build_id,
           "what": "ncu --set full --clock-control none, three consecutive launches (primary, bounce 1, bounce 2 of one frame of the "
                   "C2 workload, scripts/profile_frame.py); means over the three; dram = dram__bytes_read.sum + dram__bytes_write.sum",
           "kernels": {}}
for kind, kernel in (("traverse", "traverseWideKernel"), ("ref_traverse", "traverseKernel"), ("shade", "shadeKernel")):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, kind))
    if not os.path.exists(rep):
        continue
    launches = raw_page(rep)
    doc = {"source": "ncu --set full --clock-control none --import-source on -k regex:%s (scripts/profile_frame.py, "
                     "C2 workload: 1M-triangle soup, 1920x1080)" % kernel, "build_id": build_id, "launches": launches}
    json.dump(doc, open(os.path.join(P, "%s_%s_ncu.json" % (tag, kind)), "w"), indent=1)
    dram = [l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in launches]
    summary["kernels"][kernel] = {
        "dram_bytes_per_launch": sum(dram) / max(1, len(dram)), "dram_bytes_each": dram,
        "duration_ms_each_under_ncu": [l.get("gpu__time_duration.ms") for l in launches],
        "registers": mean(launches, "launch__registers_per_thread"),
        "warps_active_per_sm": mean(launches, "sm__warps_active.avg.per_cycle_active"),
        "threads_per_inst": mean(launches, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": mean(launches, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1_hit_pct": mean(launches, "l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": mean(launches, "lts__t_sector_hit_rate.pct"),
        "lsu_wavefront_pct": mean(launches, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "l2_throughput_pct": mean(launches, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": mean(launches, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "stall_long_scoreboard": mean(launches, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "inst_executed": mean(launches, "smsp__inst_executed.sum"),
    }
if summary["kernels"]:
    json.dump(summary, open(os.path.join(P, "traverse_ncu_summary.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(P)))
