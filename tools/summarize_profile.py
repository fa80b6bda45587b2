"""Turns one round of gpurun_out/ evidence into the tracked summaries under profiles/.

  python tools/summarize_profile.py <tag>      e.g. r1i
reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list of bench.py)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full of tile_kernel)
writes profiles/<tag>_launches.csv, <tag>_launch_shares.json, <tag>_tile_kernel_ncu_full.json,
       <tag>_tile_kernel_top_source_lines.txt
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

launches = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ik, iv, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0, ""])
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        a = agg[r[ik][:90]]
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
        a[2] = r[ig]
    ours = {k: v for k, v in agg.items() if "md2::" in k}
    tot = sum(v[1] for v in ours.values())
    summ = [{"kernel": k, "launches": v[0], "avg_us": v[1] / v[0] / 1e3, "share_of_our_launches": v[1] / tot, "grid": v[2]}
            for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1])]
    for s in summ:
        print(s)
    json.dump(summ, open(os.path.join(P, f"{tag}_launch_shares.json"), "w"), indent=1)
    shutil.copy(launches, os.path.join(P, f"{tag}_launches.csv"))

rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
    keep += [x for x in hdr if "issue_stalled" in x and "per_issue_active" in x]
    d = {k: (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip() for k in keep if k in hdr}
    json.dump(d, open(os.path.join(P, f"{tag}_tile_kernel_ncu_full.json"), "w"), indent=1)
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread"):
        print(k, d.get(k))
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_tile_kernel_top_source_lines.txt"), "w").write(lines)
