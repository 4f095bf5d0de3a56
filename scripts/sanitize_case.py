"""Small mixed batch (valid, truncated, bit-flipped, bad headers, LZMA2 groups, HBM literal tables)
for compute-sanitizer memcheck: hostile input must never make the kernel touch memory out of bounds."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from lzma_b200 import batch as B  # noqa: E402
from lzma_b200 import corpus as K  # noqa: E402

rng = random.Random(7)
cs = cases.alone_cases(heavy=False)[:60] + cases.encoder_cases(heavy=False)[:40]
blk = K.mixed_block(3, 20_000)
s = K.compress_alone(blk)
for r in range(40):   # many corruptions of one stream, tight output capacity
    b = bytearray(s)
    for _ in range(rng.randrange(1, 4)):
        i = rng.randrange(13, len(b))
        b[i] ^= 1 << rng.randrange(8)
    cs.append((f"fuzz{r}", bytes(b), rng.choice([len(blk), len(blk) // 2, 3 * len(blk)])))
with B.Context([0]) as ctx:
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for name, st, dict_size, cap in cases.lzma2_cases():
        B.decode_lzma2_stream(ctx, st, dict_size)
print("sanitize case ran:", len(cs), "units; statuses", sorted({g.status for g in got}))
