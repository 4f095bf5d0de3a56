#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) : key metrics, stall reasons, opcode mix, hot spots."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max', 'smsp__inst_executed_op_branch.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in M: print(f"{k:62s} {M[k][0]:>18s} {M[k][1]}")
st = [(float(v[0].replace(',', '')), h) for h, v in M.items() if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and v[0]]
tot = sum(v for v, _ in st)
print("-- stall reasons (pc sampling, all samples)")
for v, h in sorted(st, reverse=True)[:9]: print(f"   {100*v/tot:5.1f}%  {h.replace('smsp__pcsamp_warps_issue_stalled_','')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
H = {h: i for i, h in enumerate(rows[1])}
data = rows[2:]
ti = sum(int(r[H['Instructions Executed']]) for r in data); ts = sum(int(r[H['# Samples']]) for r in data)
print(f"-- static SASS instrs {len(data)}  executed {ti:.3e}  samples {ts}")
c = Counter(); cs = Counter()
for r in data:
    t = r[H['Source']].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[H['Instructions Executed']]); cs[op] += int(r[H['# Samples']])
for op, n in c.most_common(16): print(f"   {op:8s} exec {100*n/ti:5.1f}%  samples {100*cs[op]/ts:5.1f}%")
if len(sys.argv) > 2:
    print("-- top SASS lines by samples")
    top = sorted(data, key=lambda r: -int(r[H['# Samples']]))[:int(sys.argv[2])]
    for r in top: print(f"   {int(r[H['# Samples']]):8d} {int(r[H['Instructions Executed']]):12d}  {r[H['Source']].strip()[:80]}")
