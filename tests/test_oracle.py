"""Pins the oracle: the reference's own golden values, then liblzma cross-checks."""
import hashlib
import itertools
import lzma

import pytest

from cases import asset
from lzma_b200 import corpus as K
from oracle import oracle as O

RANDOM_MD5 = "b2d18c4275c394a729607ff9fe0caae7"   # reader1_test.go:107
A_MD5 = "57a42eb7f425c13fa644f2618a097ab7"        # liblzma and the survey's model agree


@pytest.mark.parametrize("name", ["a.lzma", "a_eos.lzma", "a_eos_and_size.lzma", "a_lp1_lc2_pb1.lzma"])
def test_good_assets(name):  # reader1_test.go:26-49: constructor and io.Copy succeed
    r = O.lzma_alone(asset(name), 4096)
    assert r.status == O.OK and len(r.data) == 327
    assert hashlib.md5(r.data).hexdigest() == A_MD5
    assert r.data.startswith(b"LZMA decoder test example")
    assert r.data == lzma.decompress(asset(name), format=lzma.FORMAT_ALONE)


@pytest.mark.parametrize("name,site", [("bad_corrupted.lzma", 652), ("bad_eos_incorrect_size.lzma", 636),
                                       ("bad_incorrect_size.lzma", 46)])
def test_bad_assets(name, site):  # reader1_test.go:50-67: constructor succeeds, io.Copy errors
    r = O.lzma_alone(asset(name), 4096)
    assert r.status == O.RESULT_ERROR and r.err_site == site
    with pytest.raises(lzma.LZMAError):
        lzma.decompress(asset(name), format=lzma.FORMAT_ALONE)


def test_randomfile_lzma():  # TestReader1WithFileVerification, reader1_test.go:85-105
    r = O.lzma_alone(asset("randomfile.dat.lzma"), 2 << 20)
    assert r.status == O.OK and hashlib.md5(r.data).hexdigest() == RANDOM_MD5


def test_randomfile_lzma2():  # TestReader2WithFileVerification, reader2_test.go:12-29 (dictSize 0)
    r = O.lzma2(asset("randomfile.dat.lzma2"), 0, 2 << 20)
    assert r.status == O.OK and hashlib.md5(r.data).hexdigest() == RANDOM_MD5


def test_header_helpers():
    assert O.decode_prop(0x5D) == (3, 2, 0)          # (lc, pb, lp): reader1.go:210-221
    assert O.decode_prop(0x37) == (1, 1, 1)
    assert O.decode_prop(224) == (8, 4, 4)
    assert O.decode_prop(225) is None
    assert O.decode_dict_size(bytes([0, 0, 0x80, 0])) == 8 << 20
    assert O.decode_dict_size(bytes([1, 0, 0, 0])) == 4096       # clamp, reader1.go:199-201
    assert O.decode_unpack_size(bytes([0x47, 1, 0, 0, 0, 0, 0, 0])) == 327
    assert O.decode_unpack_size(b"\xff" * 8) == O.UNKNOWN_SIZE
    assert O.decode_dict_size2(24) == 16 << 20                   # reader2.go:296-298
    assert O.decode_dict_size2(0) == 4096 and O.decode_dict_size2(1) == 6144


def test_liblzma_cross_check_props():
    blk = K.mixed_block(3, 40_000) + K.text_block(5, 20_000)
    for lc, lp, pb in itertools.product(range(5), range(5), range(5)):
        if lc + lp > 4:
            continue
        s = K.compress_alone(blk, lc, lp, pb, 1 << 16, preset=3)
        r = O.lzma_alone(s, len(blk))
        assert r.status == O.OK and r.data == blk and r.bytes_in == len(s), (lc, lp, pb)


def test_liblzma_cross_check_lzma2_chunks():
    # LZMA-coded LZMA2 chunks, state resets and dictionary resets are NOT pinned by any
    # reference test (its only LZMA2 asset holds uncompressed chunks): pin them via liblzma.
    blocks = [K.text_block(i, 300_000) for i in range(3)] + [K.random_block(1, 100_000), K.mixed_block(2, 200_000)]
    s = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    r = O.lzma2(s, 1 << 20, 2 << 20)
    assert r.status == O.OK and r.data == b"".join(blocks) and r.bytes_in == len(s)
    f = [{"id": lzma.FILTER_LZMA2, "dict_size": 1 << 20}]
    assert lzma.decompress(s, format=lzma.FORMAT_RAW, filters=f) == r.data


def test_truncated_input_is_clean_eof():  # Q1: decompress.go:35-38 + reader1.go:246-249
    s = K.compress_alone(K.text_block(1, 50_000))
    r = O.lzma_alone(s[:len(s) // 2], 60_000)
    assert r.status == O.OK_INPUT_EXHAUSTED and 0 < len(r.data) < 50_000
    assert K.text_block(1, 50_000).startswith(r.data)
