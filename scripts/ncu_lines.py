#!/usr/bin/env python
"""Per-source-line executed instructions / samples from an .ncu-rep (needs -lineinfo)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; H = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": H = {h: i for i, h in enumerate(r)}; iE = r.index("Instructions Executed"); iS = r.index("# Samples"); continue
    if H and r and r[0].isdigit():
        if r[iE].isdigit(): out.append((int(r[iE]), int(r[iS]) if r[iS].isdigit() else 0, cur, int(r[0]), r[1].strip()[:90]))
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print(f"total executed {tot:.3e} samples {ts}")
for e, s, f, ln, txt in sorted(out, reverse=True)[:topn]:
    print(f"{100*e/tot:5.1f}% ex {100*s/ts:5.1f}% smp  {f}:{ln:<4d} {txt}")
