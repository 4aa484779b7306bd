#!/usr/bin/env python3
"""Per-source-line instruction / stall totals of one kernel from an .ncu-rep captured with --import-source on.

usage: ncu_lines.py report.ncu-rep kernel-regex [top_n]
Aggregates ncu's `--page source --print-source cuda,sass` CSV: the rows that carry a line number are per-line totals.
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    rec = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(rec["Instructions Executed"])
    except ValueError:
        continue
    smp = int(rec["# Samples"]) if rec["# Samples"].isdigit() else 0
    out.append((cur_file, int(r[0]), r[1].strip(), inst, smp, rec))
tot_i = sum(o[3] for o in out) or 1
tot_s = sum(o[4] for o in out) or 1
print(f"total warp-instructions {tot_i:,}   samples {tot_s:,}")
print("by instructions:")
for f, ln, src, inst, smp, rec in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"  {100*inst/tot_i:5.1f}% inst {100*smp/tot_s:5.1f}% smp  {f}:{ln:<4d} {src[:110]}")
print("by samples:")
for f, ln, src, inst, smp, rec in sorted(out, key=lambda o: -o[4])[:top]:
    st = {k: int(v) for k, v in rec.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {100*smp/tot_s:5.1f}% smp {100*inst/tot_i:5.1f}% inst  {f}:{ln:<4d} {src[:80]}  {top3}")
