"""Turns the captures of tools/capture_profiles.sh into the tracked files under profiles/.
usage: python tools/profile_summary.py [tag]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
summary = {}
for k in ("tile", "geom"):
    raw = os.path.join(G, f"prof_{k}_c3_{tag}_raw.csv")
    rep = os.path.join(G, f"prof_{k}_c3_{tag}.ncu-rep")
    if not os.path.exists(raw):
        with open(raw, "w") as f:
            f.write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            d[h] = f"{v} {u}".strip()
    summary[k] = d
    shutil.copy(raw, os.path.join(P, f"{tag}_{k}_kernel_c3_raw.csv"))
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "60"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_{k}_kernel_c3_lines.txt"), "w").write(lines)
json.dump(summary, open(os.path.join(P, f"{tag}_summary.json"), "w"), indent=1, sort_keys=True)
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"{tag}_launches_bench_c3.csv"))
print(json.dumps(summary, indent=1))
