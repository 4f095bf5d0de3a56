"""Worker for test_multi_rank.py: one rank of a world_size-N job over gloo (CPU).

Mirrors what bench.py does under torchrun: every rank computes the same LPT sharding
(lzgpu_shard_units), decodes only its shard, and the ranks combine byte counts and
checksums with all_reduce -- there is no data-path collective.  Decoding here goes through
the lane-emulated kernel code (test infrastructure) because this tier has no GPU."""
import os
import sys
import zlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from backends import make_context  # noqa: E402
from lzma_b200 import _lib as L  # noqa: E402
from lzma_b200 import batch as B  # noqa: E402
from lzma_b200 import corpus as K  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 24
    plains = [K.text_block(100 + i, 20_000 + 3_000 * (i % 7)) if i % 5 else K.random_block(i, 9_000) for i in range(n)]
    streams = [K.compress_alone(p, preset=1) for p in plains]
    units, in_buf, out_size, _ = B.build_alone_batch(streams, [len(p) for p in plains])
    shard = B.shard_units(units, world)
    mine = [i for i in range(n) if shard[i] == rank]
    ctx = make_context("emu")
    out = np.zeros(out_size, dtype=np.uint8)
    res, _ = ctx.decode_batch([units[i] for i in mine], in_buf, out)
    ok = all(res[k].status == L.OK for k in range(len(mine)))
    crc = np.zeros(n, dtype=np.int64)
    nbytes = np.zeros(n, dtype=np.int64)
    for k, i in enumerate(mine):
        u = units[i]
        crc[i] = zlib.crc32(out[u.out_off:u.out_off + res[k].bytes_out].tobytes())
        nbytes[i] = res[k].bytes_out
    t_crc, t_n = torch.from_numpy(crc), torch.from_numpy(nbytes)
    t_ok = torch.tensor([1 if ok else 0])
    t_load = torch.tensor([float(sum(units[i].in_len for i in mine))])
    dist.all_reduce(t_crc)
    dist.all_reduce(t_n)
    dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
    t_max = t_load.clone()
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        want_crc = [zlib.crc32(p) for p in plains]
        assert t_ok.item() == 1
        assert t_n.tolist() == [len(p) for p in plains], "every unit decoded by exactly one rank"
        assert t_crc.tolist() == want_crc, "sharded result differs from the plaintext"
        total = sum(u.in_len for u in units)
        assert t_max.item() <= total / world + max(u.in_len for u in units), "shards are balanced by compressed size"
        print("MULTI_RANK_OK", world, len(mine))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
