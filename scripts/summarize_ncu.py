"""Turn the ncu captures in gpurun_out/ into the small, committed summaries under profiles/.
    python scripts/summarize_ncu.py r01
Reads  gpurun_out/<tag>_launches.csv        (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/<tag>_traverse.ncu-rep    (ncu --set full, traverseKernel)
       gpurun_out/<tag>_shade.ncu-rep       (ncu --set full, shadeKernel)  [optional]
Writes profiles/<tag>_launches.csv, profiles/<tag>_launch_shares.json,
       profiles/<tag>_traverse_ncu.json, profiles/<tag>_shade_ncu.json,
       profiles/traverse_ncu_summary.json   (what bench.py reads for roofline.traffic)"""
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


src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_launches.csv"))
    json.dump(launch_shares(src), open(os.path.join(P, tag + "_launch_shares.json"), "w"), indent=1)

for kind in ("traverse", "shade"):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, kind))
    if not os.path.exists(rep):
        continue
    launches = raw_page(rep)
    doc = {"source": "ncu --set full --clock-control none --import-source on -k regex:%sKernel (scripts/profile_frame.py, "
                     "C2 workload: 1M-triangle soup, 1920x1080)" % kind,
           "launches": launches}
    json.dump(doc, open(os.path.join(P, "%s_%s_ncu.json" % (tag, kind)), "w"), indent=1)
    if kind == "traverse":
        dram = [l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in launches]
        summary = {
            "tag": tag, "kernel": "traverseKernel",
            "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches "
                    "(primary, bounce 1, bounce 2 of one frame)",
            "dram_bytes_per_launch": sum(dram) / max(1, len(dram)),
            "dram_bytes_each": dram,
        }
        json.dump(summary, open(os.path.join(P, "traverse_ncu_summary.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(P)))
