#!/usr/bin/env python3
"""Per-phase instruction totals of k_hybrid_synth (or per-line top list of any kernel) from an .ncu-rep with source.
usage: ncu_regions.py report.ncu-rep kernel-regex units [file.cuh marker1 marker2 ...]
`units` = number of work units (granule-channels) the profiled launch processed; markers are source substrings whose lines
delimit the regions."""
import csv, subprocess, sys
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
fname = sys.argv[4] if len(sys.argv) > 4 else None
markers = sys.argv[5:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = cur = None
out = []
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    rec = dict(zip(hdr[4:], r[4:]))
    out.append((cur, int(r[0]), num(rec['Instructions Executed']), num(rec['Thread Instructions Executed']), r[1].strip(), num(rec['# Samples'])))
tot = sum(o[2] for o in out); tots = sum(o[5] for o in out) or 1
print(f"total warp-instructions {tot:,} = {tot/units:.1f} per unit; samples {tots:,}")
if fname:
    import os
    path = None
    for root, _, files in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))):
        if fname in files and 'simt_emu' not in root: path = os.path.join(root, fname)
    src = open(path).read().split('\n')
    bounds = []
    for m in markers:
        for i, l in enumerate(src):
            if m in l: bounds.append((m, i + 1)); break
    bounds.append(('<end>', len(src) + 1))
    for (m, a), (_, b) in zip(bounds, bounds[1:]):
        sel = [o for o in out if o[0] == fname and a <= o[1] < b]
        t = sum(o[2] for o in sel); s = sum(o[5] for o in sel); th = sum(o[3] for o in sel)
        print(f"  {m[:40]:40s} lines {a:4d}-{b:4d}: {t/units:8.1f}/unit {100*t/tot:5.1f}% inst {100*s/tots:5.1f}% smp  lanes {th/max(t,1):4.1f}")
    t = sum(o[2] for o in out if o[0] != fname)
    print(f"  other files: {t/units:.1f}/unit")
top = int(__import__('os').environ.get('TOP', '25'))
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"  {o[2]/units:7.1f}/unit lanes {o[3]/max(o[2],1):4.1f} smp {100*o[5]/tots:4.1f}%  {o[0]}:{o[1]} {o[4][:100]}")
