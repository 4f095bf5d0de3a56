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
