#!/usr/bin/env python3
"""Executed warp-instructions by SASS opcode of one kernel of an .ncu-rep (source page, SASS view).
usage: ncu_opcodes.py report.ncu-rep kernel-regex [units]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
tot = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r: continue
    if 'Source' in r and 'Instructions Executed' in r: hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    rec = dict(zip(hdr, r))
    try: n = int(rec['Instructions Executed'])
    except ValueError: continue
    src = rec['Source'].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    op = op.split('.')[0]
    tot[op] += n
    try: smp[op] += int(rec.get('# Samples', '0') or 0)
    except ValueError: pass
T = sum(tot.values()); S = sum(smp.values()) or 1
print(f"total {T:,} warp-instructions = {T/units:.1f} per unit")
for op, n in tot.most_common(28):
    print(f"  {op:10s} {n/units:9.1f}/unit {100*n/T:5.1f}%   samples {100*smp[op]/S:5.1f}%")
