#!/usr/bin/env python
"""Per-source-line share of samples / executed instructions of an .ncu-rep, sorted by samples, plus per-file totals."""
import csv, subprocess, io, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; H = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": H = {h: i for i, h in enumerate(r)}; iE = r.index("Instructions Executed"); iS = r.index("# Samples"); continue
    if H and r and r[0].isdigit() and r[iE].isdigit(): out.append((int(r[iS]) if r[iS].isdigit() else 0, int(r[iE]), cur, int(r[0]), r[1].strip()[:80]))
ts = sum(o[0] for o in out); te = sum(o[1] for o in out)
print("samples", ts, "executed %.3e" % te)
by = {}
for s, e, f, ln, t in out:
    by.setdefault(f, [0, 0]); by[f][0] += s; by[f][1] += e
for f, (s, e) in by.items(): print(f"{f:28s} samples {100*s/ts:5.1f}%  exec {100*e/te:5.1f}%")
for s, e, f, ln, t in sorted(out, reverse=True)[:topn]:
    print(f"{100*s/ts:5.1f}% smp {100*e/te:5.1f}% ex  {f}:{ln} {t}")
