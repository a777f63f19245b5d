"""Summarise an ncu report per CUDA source line: samples and instructions executed.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
acc = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> samples, inst, thread_inst, text
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].isdigit() and len(r) > 8:
        k = (cur_file, int(r[0]))
        try:
            acc[k][0] += int(r[4]); acc[k][1] += int(r[7]); acc[k][2] += int(r[8])
        except ValueError:
            pass
        acc[k][3] = r[1].strip()
tot_s = sum(v[0] for v in acc.values()) or 1
tot_i = sum(v[1] for v in acc.values()) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("--- by stall samples")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot_s:5.1f}% smp {100*v[1]/tot_i:5.1f}% inst  thr/inst {v[2]/max(v[1],1):4.1f}  {k[0]}:{k[1]:<4d} {v[3][:110]}")
