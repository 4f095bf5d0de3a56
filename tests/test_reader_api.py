"""The reference's own tests, restated against the host-side mirror of its API
(reader1_test.go:15-107, reader2_test.go:12-29).  CPU tier: the readers are driven over the
lane-emulated kernel code; GPU tier: over liblzgpu.so."""
import hashlib
import io
import os

import pytest

import cases
import lzma_b200 as lzma
from backends import make_context
from lzma_b200 import corpus as K
from lzma_b200 import errors as E
from lzma_b200.reader1 import DecodeDictSize, DecodeProp, DecodeUnpackSize, NewLZMADecompressorForSevenZip, NewReader1
from lzma_b200.reader2 import DecodeDictSize2, NewLZMA2DecompressorForSevenZip, NewReader2

randomFileMD5 = "b2d18c4275c394a729607ff9fe0caae7"   # reader1_test.go:107


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def ctx(request):
    c = make_context(request.param)
    yield c
    c.close()


TEST_READER1 = [  # reader1_test.go:26-67: (name, inputFile, constructor ok, io.Copy ok)
    ("correct_file_with_size", "a.lzma", True, True),
    ("correct_file_with_eos", "a_eos.lzma", True, True),
    ("correct_file_with_eos_and_size", "a_eos_and_size.lzma", True, True),
    ("correct_file_lp1_lc2_pb1", "a_lp1_lc2_pb1.lzma", True, True),
    ("bad_file", "bad_corrupted.lzma", True, False),
    ("bad_file_with_eos_and_incorrect_size", "bad_eos_incorrect_size.lzma", True, False),
    ("bad_file_with_incorrect_size", "bad_incorrect_size.lzma", True, False),
]


@pytest.mark.parametrize("name,inputFile,ok1,ok2", TEST_READER1)
def test_TestReader1(ctx, name, inputFile, ok1, ok2):
    reader, err = NewReader1(io.BytesIO(cases.asset(inputFile)), ctx)
    assert (err is None) == ok1
    sink = io.BytesIO()
    _, err = lzma.io_copy(sink, reader)
    assert (err is None) == ok2
    if ok2:
        assert hashlib.md5(sink.getvalue()).hexdigest() == "57a42eb7f425c13fa644f2618a097ab7"
    else:
        assert E.Is(err, E.ErrResultError)      # the class the reference returns


def test_TestReader1WithFileVerification(ctx):
    r, err = NewReader1(io.BytesIO(cases.asset("randomfile.dat.lzma")), ctx)
    assert err is None
    actualSummator = hashlib.md5()
    _, err = lzma.io_copy(actualSummator, r)
    assert err is None
    assert actualSummator.hexdigest() == randomFileMD5


def test_TestReader2WithFileVerification(ctx):
    r, err = NewReader2(io.BytesIO(cases.asset("randomfile.dat.lzma2")), 0, ctx)
    assert err is None
    actualSummator = hashlib.md5()
    lzma.io_copy(actualSummator, r)                # the reference ignores this error (reader2_test.go:23)
    assert actualSummator.hexdigest() == randomFileMD5


def test_constructor_errors(ctx):
    r, err = NewReader1(io.BytesIO(b""), ctx)
    assert err is E.EOF                                            # reader1.go:78-81
    r, err = NewReader1(io.BytesIO(bytes([225]) + bytes(20)), ctx)
    assert E.Is(err, E.ErrIncorrectProperties) and str(err).startswith("decode prop: ")
    r, err = NewReader1(io.BytesIO(bytes([0x5D, 0, 0])), ctx)
    assert E.Is(err, E.EOF) and str(err).startswith("decode dict size: ")
    r, err = NewReader1(io.BytesIO(bytes([0x5D]) + bytes(4) + bytes(3)), ctx)
    assert E.Is(err, E.EOF) and str(err).startswith("decode unpack size: ")
    good = cases.asset("a.lzma")
    r, err = NewReader1(io.BytesIO(good[:13] + b"\x01" + good[14:]), ctx)
    assert E.Is(err, E.ErrResultError) and str(err).startswith("rangeDec.Init: ")   # range_decoder.go:32-34
    r, err = NewReader1(io.BytesIO(good[:15]), ctx)
    assert E.Is(err, E.EOF) and str(err).startswith("rangeDec.Init: ")
    r, err = NewReader2(io.BytesIO(b""), 0, ctx)
    assert err is E.ErrUnexpectedEOF                               # reader2.go:103-110
    r, err = NewReader2(io.BytesIO(b"\xe0\x00"), 0, ctx)
    assert err is E.ErrUnexpectedEOF                               # reader2.go:121-128


def test_helpers():
    assert DecodeProp(0x5D) == (3, 2, 0, None)         # (lc, pb, lp, err)
    assert DecodeProp(225)[3] is E.ErrIncorrectProperties
    assert DecodeDictSize(bytes([0, 0, 0x80, 0])) == (8 << 20, None)
    assert DecodeDictSize(bytes(4)) == (4096, None)
    assert DecodeUnpackSize(b"\xff" * 8) == (1 << 64) - 1
    assert DecodeDictSize2(24) == 16 << 20


def test_truncated_stream_is_clean_eof(ctx):
    """Q1: the reference turns input exhaustion into a clean end of stream."""
    d = K.text_block(1, 50_000)
    s = K.compress_alone(d)
    r, err = NewReader1(io.BytesIO(s[:len(s) // 2]), ctx)
    assert err is None
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, r)
    assert err is None and 0 < n < len(d) and d.startswith(sink.getvalue())


def test_lzma2_reader_multi_unit_and_truncation(ctx):
    blocks = [K.text_block(i, 200_000) for i in range(3)] + [K.random_block(1, 70_000)]
    s = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    r, err = NewReader2(io.BytesIO(s), 1 << 20, ctx)
    assert err is None
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, r)
    assert err is None and sink.getvalue() == b"".join(blocks)
    r, err = NewReader2(io.BytesIO(s[:len(s) - 40_000]), 1 << 20, ctx)
    assert err is None
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, r)
    assert err is E.ErrUnexpectedEOF and b"".join(blocks).startswith(sink.getvalue())


class _CountingIO(io.BytesIO):
    bytes_read = 0

    def read(self, n=-1):
        c = super().read(n)
        self.bytes_read += len(c)
        return c


def test_lzma2_reader_waves(ctx):
    """Read in waves: bounded look-ahead, same bytes, same errors, early stop reads only a prefix."""
    blocks = [K.text_block(20 + i, 150_000) for i in range(6)]
    s = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    plain = b"".join(blocks)
    for wave in (1, 200_000, 10 << 20):
        src = _CountingIO(s)
        r, err = NewReader2(src, 1 << 20, ctx)
        assert err is None
        r.wave_bytes = wave
        sink = io.BytesIO()
        n, err = lzma.io_copy(sink, r)
        assert err is None and sink.getvalue() == plain, wave
    # early stop: only the first wave's input (and the read-ahead block) has been consumed
    big = K.lzma2_with_resets([K.random_block(40 + i, 700_000) for i in range(6)], dict_size=1 << 20)
    src = _CountingIO(big)
    r, err = NewReader2(src, 1 << 20, ctx)
    r.wave_bytes = 1
    r.decode_ahead = False                  # (with decode-ahead one more wave is read in the background)
    buf = bytearray(1000)
    assert r.Read(buf)[0] == 1000
    assert src.bytes_read < len(big) // 2
    # truncation inside a later wave
    r, err = NewReader2(io.BytesIO(s[:len(s) - 30_000]), 1 << 20, ctx)
    r.wave_bytes = 1
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, r)
    assert err is E.ErrUnexpectedEOF and plain.startswith(sink.getvalue()) and n > 4 * 150_000


class _Closer(io.BytesIO):
    closed_calls = 0

    def close(self):
        self.closed_calls += 1


def test_sevenzip_adapters(ctx):
    d = K.text_block(3, 80_000)
    s = K.compress_alone(d)
    src = _Closer(s[13:])
    rc, err = NewLZMADecompressorForSevenZip(s[:5], len(d), [src], ctx)
    assert err is None
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, rc)
    assert err is None and sink.getvalue() == d
    assert rc.Close() is None and src.closed_calls == 1
    assert rc.Close() is E.errAlreadyClosed and rc.Read(bytearray(4)) == (0, E.errAlreadyClosed)
    assert NewLZMADecompressorForSevenZip(s[:5], len(d), [], ctx)[1] is E.errNeedOneReader
    # bad body: errors are wrapped "lzma: error reading: %w" (readcloser.go:36-38)
    rc, err = NewLZMADecompressorForSevenZip(cases.asset("bad_corrupted.lzma")[:5], 327, [_Closer(cases.asset("bad_corrupted.lzma")[13:])], ctx)
    assert err is None
    n, err = lzma.io_copy(io.BytesIO(), rc)
    assert E.Is(err, E.ErrResultError) and str(err).startswith("lzma: error reading: ")
    l2 = K.compress_raw_lzma2(d, dict_size=1 << 20)
    rc, err = NewLZMA2DecompressorForSevenZip(bytes([18]), 0, [_Closer(l2)], ctx)   # 18 -> 1 MiB
    assert err is None
    sink = io.BytesIO()
    n, err = lzma.io_copy(sink, rc)
    assert err is None and sink.getvalue() == d
    assert NewLZMA2DecompressorForSevenZip(b"", 0, [_Closer(l2)], ctx)[1] is E.errInsufficientProperties


def test_decode_folders_one_call(ctx):
    """all folders of an archive -- LZMA (headerless, props 5 bytes) and LZMA2 (props 1 byte, several
    units each), a corrupt one, an untrusted size field, bad properties -- in ONE batch call; each folder's bytes
    and error class equal the oracle's for the same folder read on its own."""
    from lzma_b200.sevenzip import Folder, decode_folders
    from oracle import oracle as O
    folders, want = [], []
    for i in range(40):
        lc, lp, pb = [(3, 0, 2), (0, 2, 0), (4, 0, 4), (1, 1, 1)][i % 4]
        d = (K.text_block, K.mixed_block)[i % 2](4000 + i, 30_000 + 1777 * i)
        s = K.compress_alone(d, lc, lp, pb, 1 << 18, preset=1 + i % 6)
        folders.append(Folder(False, s[:5], len(d), s[13:]))
        want.append((d, None))
    for i in range(24):
        blocks = [K.text_block(4100 + 3 * i + j, 40_000 + 999 * j) if (i + j) % 3 else K.random_block(4100 + i + j, 70_000) for j in range(3)]
        s = K.lzma2_with_resets(blocks, dict_size=1 << 20, preset=1 + i % 4)
        folders.append(Folder(True, bytes([0x10]), 0, s))        # 0x10 = 1 MiB
        want.append((b"".join(blocks), None))
    bad = cases.asset("bad_corrupted.lzma")
    folders.append(Folder(False, bad[:5], DecodeUnpackSize(bad[5:13]), bad[13:]))
    want.append((None, E.ErrResultError))
    a = cases.asset("a.lzma")
    folders.append(Folder(False, a[:5], 1 << 50, a[13:]))        # a header that claims 2^50 bytes must not allocate them
    want.append((None, "any"))
    folders.append(Folder(True, b"", 0, b"\0"))
    want.append((b"", E.errInsufficientProperties))
    folders.append(Folder(False, bytes([225, 0, 0, 1, 0]), 10, b"\0" * 10))
    want.append((b"", E.ErrIncorrectProperties))
    trunc = K.lzma2_with_resets([K.text_block(4300, 90_000)], dict_size=1 << 20)
    folders.append(Folder(True, bytes([0x10]), 0, trunc[:len(trunc) // 2]))
    want.append((None, E.ErrUnexpectedEOF))
    assert len(folders) >= 64
    got = decode_folders(ctx, folders)
    for i, ((data, err), (wd, we)) in enumerate(zip(got, want)):
        if we is None:
            assert err is None and data == wd, i
        elif we == "any":
            r = O.lzma_raw(folders[i].packed, 3, 0, 2, 1 << 23, 1 << 50, 1 << 20)
            assert (err is None) == (r.status in (O.OK, O.OK_INPUT_EXHAUSTED)) and data == r.data, i
        else:
            assert E.Is(err, we), (i, err)
    # folder by folder through the sevenzip constructors: the same bytes
    for i in (0, 41, 45):
        f = folders[i]
        ctor = NewLZMA2DecompressorForSevenZip if f.lzma2 else NewLZMADecompressorForSevenZip
        rc, err = ctor(f.props, f.unpack_size, [io.BytesIO(f.packed)], ctx)
        assert err is None
        sink = io.BytesIO()
        _, err = lzma.io_copy(sink, rc)
        assert err is None and sink.getvalue() == got[i][0]


def test_lzma2_reader_decode_ahead(ctx):
    """N1: while a wave is served the next one is decoded on a second thread; small and large reads, waves of one
    unit, a failing unit in a later wave, and a caller that stops early (the ahead thread must not be left hanging)."""
    blocks = [K.text_block(4400 + i, 60_000 + 4321 * i) for i in range(9)]
    stream = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    plain = b"".join(blocks)
    for ahead in (True, False):
        for wave, bufsize in ((1, 7_000), (150_000, 1 << 20), (1 << 30, 32 * 1024)):
            r, err = NewReader2(io.BytesIO(stream), 1 << 20, ctx)
            assert err is None
            r.wave_bytes, r.decode_ahead = wave, ahead
            sink = io.BytesIO()
            _, err = lzma.io_copy(sink, r, bufsize)
            assert err is None and sink.getvalue() == plain, (ahead, wave)
    # corruption inside the 6th unit: the bytes before it are delivered, then ErrResultError
    units, _, _ = lzma.scan_lzma2(stream, 1 << 20)
    bad = bytearray(stream)
    bad[units[5].in_off + units[5].in_len // 2] ^= 0x40
    r, err = NewReader2(io.BytesIO(bytes(bad)), 1 << 20, ctx)
    r.wave_bytes = 1
    sink = io.BytesIO()
    _, err = lzma.io_copy(sink, r, 10_000)
    assert err is not None and sink.getvalue()[:units[5].out_off] == plain[:units[5].out_off]
    # stop after the first bytes
    r, err = NewReader2(io.BytesIO(stream), 1 << 20, ctx)
    r.wave_bytes = 1
    buf = bytearray(100)
    n, err = r.Read(buf)
    assert n == 100 and err is None and bytes(buf) == plain[:100]
    del r
