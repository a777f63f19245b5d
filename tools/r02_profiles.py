"""Turns the captures of tools/r02_capture.sh / r02_benchall.sh (gpurun_out/r02_*) into the tracked files under
profiles/: r02_summary.json (the counters quoted in DESIGN.md), r02_traffic.json (what bench.py reads for
roofline.traffic), the raw pages, the per-line summaries, the launch list and the bench lines.
usage: python tools/r02_profiles.py"""
import csv
import glob
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
TILES = {"c3": 64, "c2": 32, "c5": 64, "fill": 32, "c0_4k": 32}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

summary, captures = {}, []
for w, tile in TILES.items():
    for k in ("tile", "geometry"):
        raw = os.path.join(G, f"r02_{k}_{w}_raw.csv")
        if not os.path.exists(raw):
            continue
        rows = list(csv.reader(open(raw)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        summary[f"{w}_{k}"] = {h: f"{d[h][0]} {d[h][1]}".strip() for h in KEYS if h in d}
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            b = float(rd[0]) * UNIT.get(rd[1], 1.0) + float(wr[0]) * UNIT.get(wr[1], 1.0)
            captures.append({"workload": w, "tile": tile, "world": 1, "kernel": k, "dram_bytes": b,
                             "source": f"profiles/r02_{k}_{w}_raw.csv (tools/prof_run.py {w} {tile}, one launch)"})
        shutil.copy(raw, os.path.join(P, f"r02_{k}_{w}_raw.csv"))
        lines = os.path.join(G, f"r02_{k}_{w}_lines.txt")
        if os.path.exists(lines):
            shutil.copy(lines, os.path.join(P, f"r02_{k}_{w}_lines.txt"))
json.dump(summary, open(os.path.join(P, "r02_summary.json"), "w"), indent=1, sort_keys=True)
json.dump({"captures": captures}, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
for f in ["r02_launches_bench_c3.csv", "r02_compute_sanitizer.txt"] + [os.path.basename(x) for x in glob.glob(os.path.join(G, "r02_bench_*.json"))]:
    src = os.path.join(G, f)
    if os.path.exists(src) and os.path.getsize(src) > 0:
        shutil.copy(src, os.path.join(P, f))
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")} for k, v in summary.items()}, indent=1))
