#!/usr/bin/env python3
"""Executed PTX instructions of the fast decoder's asm blocks per call and per output byte, on a text block of the bench
corpus -- from the PTX interpreter of the CPU tier (tests/ptx/interp.py), no GPU.  The SASS ptxas makes of a block is
about as long (it fuses sub+min, shr+add; it adds address arithmetic for the spilled predicates), so this is the
instruction budget of the bit ladders as written, without the C++ between them (~100 instructions per symbol)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lzma_b200 import corpus as K          # noqa: E402
from ptx import interp as I                # noqa: E402
import test_ptx_fast_decoder as T          # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128 << 10
blocks = I.extract(T.INV)
stats = {}
orig_run = I.run


def counting_run(block, values, sm, counters=None):
    c = {}
    out = orig_run(block, values, sm, c)
    name = next(k for k, v in blocks.items() if v is block)
    s = stats.setdefault(name, [0, 0])
    s[0] += 1
    s[1] += c["steps"]
    return out


I.run = counting_run
plain = K.text_block(1000, size)
f = T.Fast(blocks, K.compress_alone(plain), len(plain))
f.run()
n = len(f.out)
assert bytes(f.out) == plain[:n]
tot = sum(s[1] for s in stats.values())
print(f"{n} bytes decoded from text (1 MiB-corpus generator, seed 1000); {tot / n:.1f} PTX instructions per byte inside the asm blocks")
print(f"{'block':12s} {'calls':>8s} {'instr/call':>11s} {'instr/byte':>11s} {'share':>6s}")
for name, (calls, steps) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:12s} {calls:8d} {steps / calls:11.1f} {steps / n:11.2f} {100 * steps / tot:5.1f}%")
