"""Runs hostile inputs through the ASan/UBSan build of the lane-emulated kernel code
(spawned by test_emu_asan.py with libasan preloaded).  One unit per call, buffers of exactly
the declared size, so that any read or write outside a unit's ranges hits a redzone."""
import ctypes as C
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends  # noqa: E402
import cases  # noqa: E402
from lzma_b200 import batch as B  # noqa: E402
from lzma_b200 import corpus as K  # noqa: E402
from lzma_b200._lib import Result, Unit  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "liblzgpu_emu_asan.so"))
lib.emu_decode_batch.restype = C.c_int
lib.emu_decode_batch.argtypes = [C.POINTER(Unit), C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.POINTER(Result), C.c_int]
backends._EMU = lib
rng = random.Random(7)
cs = cases.alone_cases(heavy=False)[:70] + cases.encoder_cases(heavy=False)[:30]
blk = K.mixed_block(3, 12_000)
s = K.compress_alone(blk)
for r in range(60):
    b = bytearray(s)
    for _ in range(rng.randrange(1, 4)):
        i = rng.randrange(13, len(b))
        b[i] ^= 1 << rng.randrange(8)
    cs.append((f"fuzz{r}", bytes(b), rng.choice([len(blk), len(blk) // 2, 3 * len(blk)])))
for v in (0, 1, 17):
    ctx = backends.EmuContext(v)
    for name, st, cap in cs:
        units, in_buf, _, _ = B.build_alone_batch([st], [cap])
        in_exact = np.frombuffer(bytes(in_buf[:max(len(st), 1)]), dtype=np.uint8).copy()
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        units[0].out_cap = min(cap, out.nbytes)
        ctx.decode_batch(units, in_exact, out)
    for name, st, dict_size, cap in cases.lzma2_cases()[:12]:
        B.decode_lzma2_stream(ctx, st, dict_size)
print("ASAN_OK", len(cs))
