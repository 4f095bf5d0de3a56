#!/usr/bin/env python
"""Hot-code footprint from an .ncu-rep: 128-byte instruction lines weighted by executions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
H = {h: i for i, h in enumerate(rows[1])}
data = rows[2:]
base = int(data[0][H['Address']], 16)
lines = {}
tot = 0
for r in data:
    a = (int(r[H['Address']], 16) - base) // 128
    e = int(r[H['Instructions Executed']])
    lines[a] = lines.get(a, 0) + e
    tot += e
vals = sorted(lines.values(), reverse=True)
print(f"static {len(data)} instrs = {len(data)*16/1024:.0f} KB in {len(lines)} lines; executed {tot:.3e}")
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99, 0.999]
mi = 0
for k, v in enumerate(vals):
    acc += v
    while mi < len(marks) and acc >= marks[mi] * tot:
        print(f"  {marks[mi]*100:5.1f}% of executed instrs within {k+1} lines = {(k+1)*128/1024:.1f} KB")
        mi += 1
