"""`.xz` front-end: container parsing on the CPU tier (block table vs liblzma's own
view of the file, block payloads through the oracle), whole files through the GPU on the GPU tier."""
import lzma
import shutil
import struct
import subprocess

import pytest

from lzma_b200 import corpus as K
from lzma_b200 import xz as X
from oracle import oracle as O


def _xz_cli(data: bytes, *args: str) -> bytes:
    return subprocess.run(["xz", "-c", *args], input=data, stdout=subprocess.PIPE, check=True).stdout


def _files():
    text = K.text_block(11, 700_000)
    mixed = K.mixed_block(12, 300_000)
    out = [
        ("py-crc64", text, lzma.compress(text, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, preset=6)),
        ("py-crc32", mixed, lzma.compress(mixed, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC32, preset=1)),
        ("py-sha256", text[:50_000], lzma.compress(text[:50_000], format=lzma.FORMAT_XZ, check=lzma.CHECK_SHA256)),
        ("py-none", mixed[:10_000], lzma.compress(mixed[:10_000], format=lzma.FORMAT_XZ, check=lzma.CHECK_NONE)),
        ("empty", b"", lzma.compress(b"", format=lzma.FORMAT_XZ)),
        ("two-streams", text[:40_000] + mixed[:30_000],
         lzma.compress(text[:40_000], format=lzma.FORMAT_XZ) + b"\0" * 8 + lzma.compress(mixed[:30_000], format=lzma.FORMAT_XZ)),
    ]
    if shutil.which("xz"):
        out.append(("cli-blocks", text, _xz_cli(text, "-6", "--block-size=131072")))        # 6 blocks, CRC64
        out.append(("cli-threads", text + mixed, _xz_cli(text + mixed, "-T2", "--block-size=262144", "-C", "crc32")))
    return out


def test_scan_matches_liblzma_and_oracle():
    for name, plain, data in _files():
        if name != "two-streams":      # (python's lzma stops at stream padding; xz and this parser accept it)
            assert lzma.decompress(data) == plain, name
        streams = X.scan_xz(data)
        got = b""
        for st in streams:
            for b in st.blocks:
                payload = data[b.data_off:b.data_off + b.data_len]
                r = O.lzma2(payload, b.dict_size, b.uncompressed_size + 16)
                assert r.status == O.OK and len(r.data) == b.uncompressed_size, name
                assert X.check_ok(b.check_type, b.check, r.data), name
                got += r.data
        assert got == plain, name
    if shutil.which("xz"):
        nblk = sum(len(st.blocks) for st in X.scan_xz(dict((n, d) for n, _, d in _files())["cli-blocks"]))
        assert nblk == 6


def test_crc64_matches_known_value():
    assert X.crc64(b"123456789") == 0x995DC9BBDF1939FA     # CRC-64/XZ check value


def test_container_errors():
    data = lzma.compress(K.text_block(3, 20_000), format=lzma.FORMAT_XZ)
    for mutate in (lambda b: b[:-1], lambda b: b[:6] + bytes([b[6] ^ 1]) + b[7:], lambda b: b[:-12] + bytes([b[-12] ^ 1]) + b[-11:],
                   lambda b: b[:-20] + bytes([b[-20] ^ 0x40]) + b[-19:], lambda b: b[:13] + bytes([b[13] ^ 4]) + b[14:]):
        with pytest.raises(X.XZError):
            X.scan_xz(mutate(data))
    bcj = lzma.compress(b"x" * 1000, format=lzma.FORMAT_XZ, filters=[{"id": lzma.FILTER_X86}, {"id": lzma.FILTER_LZMA2, "preset": 1}])
    with pytest.raises(X.XZError, match="unsupported filter"):
        X.scan_xz(bcj)


@pytest.mark.gpu
def test_xz_files_on_gpu():
    from lzma_b200 import batch as B
    files = _files()
    with B.Context([0]) as ctx:
        outs = X.decode_xz_files(ctx, [d for _, _, d in files])      # every block of every file in one batch
        for (name, plain, _), got in zip(files, outs):
            assert got == plain, name
        # a flipped bit in a block's check field is caught by the host-side verification
        name, plain, data = files[0]
        st = X.scan_xz(data)[0]
        b = st.blocks[0]
        pos = b.data_off + b.data_len + (-(b.header_size + b.data_len) % 4) + len(b.check) - 1
        bad = data[:pos] + bytes([data[pos] ^ 1]) + data[pos + 1:]
        with pytest.raises(X.XZError, match="integrity"):
            X.decode_xz(ctx, bad)
        assert X.decode_xz(ctx, bad, verify=False) == plain
        # a flipped bit in the compressed payload is reported by the decoder or by the check
        pos = b.data_off + b.data_len // 2
        bad = data[:pos] + bytes([data[pos] ^ 0x10]) + data[pos + 1:]
        with pytest.raises(X.XZError):
            X.decode_xz(ctx, bad)


def test_crc_combine_host_arithmetic():
    """lzgpu_crc32_combine / lzgpu_crc64_combine (host arithmetic of the C ABI, no device needed):
    crc(A || B) from crc(A), crc(B), |B| -- against zlib.crc32 and the CRC-64/XZ routine, ragged lengths."""
    import random
    import zlib
    from lzma_b200 import _lib as L
    lib = L.lib()
    rng = random.Random(5)
    for la, lb in [(0, 0), (0, 5), (5, 0), (1, 1), (3, 4096), (70_001, 13), (65_536, 65_536), (1_000_003, 999_983)]:
        a, b = rng.randbytes(la), rng.randbytes(lb)
        assert lib.lzgpu_crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b), (la, lb)
        assert lib.lzgpu_crc64_combine(X.crc64(a), X.crc64(b), lb) == X.crc64(a + b), (la, lb)
    parts = [rng.randbytes(rng.randrange(0, 5000)) for _ in range(40)]   # fold of many pieces, as an .xz block of many units
    acc32 = acc64 = 0
    for p in parts:
        acc32 = lib.lzgpu_crc32_combine(acc32, zlib.crc32(p), len(p))
        acc64 = lib.lzgpu_crc64_combine(acc64, X.crc64(p), len(p))
    assert acc32 == zlib.crc32(b"".join(parts)) and acc64 == X.crc64(b"".join(parts))


@pytest.mark.gpu
def test_block_checks_on_the_device():
    """lzgpu_decode_batch_sums: CRC-32 / CRC-64 of each flagged unit computed by the GPU == zlib.crc32 / CRC-64/XZ of
    the decoded bytes (ragged sizes, odd offsets, a failed unit: checksum of what it decoded); lzgpu_plan_crc64 for
    device-resident plans; and a multi-block, multi-unit .xz file whose CRC-64 block checks are verified from the
    per-unit sums with no host pass (a unit that decodes to the right length but wrong bytes must be caught)."""
    import zlib
    import numpy as np
    from lzma_b200 import _lib as L
    from lzma_b200 import batch as B
    with B.Context([0]) as ctx:
        sizes = [0, 1, 7, 16, 255, 4097, 65_537, 300_001, 1 << 20]
        plains = [K.text_block(900 + i, n) if n else b"" for i, n in enumerate(sizes)]
        streams = [K.compress_alone(p) for p in plains]
        bad = bytearray(streams[-2]); bad[len(bad) // 2] ^= 0x20
        streams.append(bytes(bad)); plains.append(None)
        units, in_buf, out_size, _ = B.build_alone_batch(streams, [max(len(p), 1) if p is not None else 300_001 for p in plains])
        off = 0
        for k, u in enumerate(units):
            u.out_off = off + (k % 5)
            off = (u.out_off + u.out_cap + 15) // 16 * 16
            u.flags |= (L.UF_SUM_CRC32, L.UF_SUM_CRC64, 0)[k % 3]
        out = np.zeros(off + 16, dtype=np.uint8)
        res, _, sums = ctx.decode_batch_sums(units, in_buf, out)
        for k, (u, r) in enumerate(zip(units, res)):
            got = out[u.out_off:u.out_off + r.bytes_out].tobytes()
            if plains[k] is not None:
                assert r.status == L.OK and got == plains[k]
            want = (zlib.crc32(got), X.crc64(got), 0)[k % 3]
            assert int(sums[k]) == want, (k, len(got))
        # device-resident plan: CRC-64 of every unit
        import torch
        for u in units:
            u.flags = 0
        d_in = torch.from_numpy(in_buf).cuda()
        d_out = torch.zeros(off + 16, dtype=torch.uint8, device="cuda")
        plan = ctx.plan(units, in_buf.nbytes, off + 16)
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
        c64, c32 = plan.crc64(d_out.data_ptr()), plan.crc32(d_out.data_ptr())
        res, _ = plan.results()
        o = d_out.cpu().numpy()
        for k, u in enumerate(units):
            got = o[u.out_off:u.out_off + res[k].bytes_out].tobytes()
            assert int(c64[k]) == X.crc64(got) and int(c32[k]) == zlib.crc32(got), k
        plan.close()
        # .xz: 5 blocks; block 2's payload is a stream with 3 dictionary resets (3 units -> one block check folded from 3 sums)
        text = K.text_block(950, 1_200_000)
        if shutil.which("xz"):
            data = _xz_cli(text, "-6", "--block-size=262144", "-C", "crc64")
            assert X.decode_xz(ctx, data) == text
            data32 = _xz_cli(text, "-1", "--block-size=400000", "-C", "crc32")
            assert X.decode_xz_files(ctx, [data, data32]) == [text, text]
            # the check field of a block replaced by the CRC-64 of different bytes: caught from the device sums
            st = X.scan_xz(data)[0]
            b = st.blocks[1]
            cpos = b.data_off + b.data_len + (-(b.header_size + b.data_len) % 4)
            forged = data[:cpos] + struct.pack("<Q", X.crc64(b"other bytes")) + data[cpos + 8:]
            with pytest.raises(X.XZError, match="integrity"):
                X.decode_xz(ctx, forged)
