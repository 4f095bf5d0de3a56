"""CPU tier: the C-ABI library loads, exports every symbol include/lzgpu.h declares, its host-side
helpers (header parse, LZMA2 scanner, scheduler) behave like the reference's, and decoding without
a GPU fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

import cases
from lzma_b200 import _lib as L
from lzma_b200 import batch as B
from lzma_b200 import corpus as K
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lzgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)   # declarations only, not prose
    declared = set(re.findall(r"\b(lzgpu_[a-z0-9_]+)\s*\(", hdr))
    bound = {s[0] for s in L.SYMBOLS}
    assert declared == bound, declared ^ bound
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lzgpu_abi_version() == 2


def test_struct_layout():
    assert C.sizeof(L.Unit) == 64 and C.sizeof(L.Result) == 32
    assert L.Unit.dict_size.offset == 40 and L.Unit.kind.offset == 44 and L.Unit.flags.offset == 52


def test_header_helpers_match_oracle():
    lib = L.lib()
    for d in range(256):
        lc, pb, lp = C.c_uint8(), C.c_uint8(), C.c_uint8()
        st = lib.lzgpu_decode_prop(d, C.byref(lc), C.byref(pb), C.byref(lp))
        want = O.decode_prop(d)
        if want is None:
            assert st == L.INCORRECT_PROPERTIES
        else:
            assert st == L.OK and (lc.value, pb.value, lp.value) == want
        assert lib.lzgpu_decode_dict_size2(d % 41) == O.decode_dict_size2(d % 41)
    for b in (bytes(4), b"\x01\x00\x00\x00", b"\x00\x00\x80\x00", b"\xff\xff\xff\xff"):
        assert lib.lzgpu_decode_dict_size(b) == O.decode_dict_size(b)
    assert lib.lzgpu_decode_unpack_size(b"\x47\x01" + bytes(6)) == 327


def test_parse_alone_header():
    st, u = B.parse_alone_header(cases.asset("a_lp1_lc2_pb1.lzma"))
    assert st == L.OK and (u.lc, u.lp, u.pb) == (1, 1, 1) and u.dict_size == 1 << 16 and u.unpack_size == 327
    st, u = B.parse_alone_header(cases.asset("a_eos.lzma"))
    assert st == L.OK and u.unpack_size == L.UNKNOWN_SIZE
    assert B.parse_alone_header(b"")[0] == L.UNEXPECTED_EOF            # reader1.go:78-81
    assert B.parse_alone_header(bytes([225]))[0] == L.INCORRECT_PROPERTIES
    assert B.parse_alone_header(bytes([0x5D, 0, 0]))[0] == L.UNEXPECTED_EOF


def test_scan_lzma2_units():
    blocks = [K.text_block(i, 300_000) for i in range(3)] + [K.random_block(1, 100_000)]
    s = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    units, total, sst = B.scan_lzma2(s, 1 << 20)
    assert sst == L.OK and total == sum(map(len, blocks))
    assert len(units) == 4
    assert [u.out_off for u in units] == [0, 300_000, 600_000, 900_000]
    assert units[0].flags & L.UF_LZMA2_FRESH and not units[1].flags & L.UF_LZMA2_FRESH
    assert units[-1].flags & L.UF_LZMA2_LAST and not units[0].flags & L.UF_LZMA2_LAST
    assert sum(u.in_len for u in units) == len(s)
    assert all(u.lit_bits == 3 for u in units[:3])
    # the asset: 22 uncompressed chunks, first one resets the dictionary -> one unit
    units, total, sst = B.scan_lzma2(cases.asset("randomfile.dat.lzma2"), 0)
    assert len(units) == 1 and total == 1 << 20 and units[0].dict_size == 8 << 20   # reader2.go:88-91
    # truncated stream
    units, total, sst = B.scan_lzma2(s[:len(s) // 2], 1 << 20)
    assert sst == L.UNEXPECTED_EOF and units[-1].flags & L.UF_LZMA2_LAST


def test_shard_units_balances_by_compressed_size():
    import random
    rng = random.Random(0)
    units = []
    for i in range(1000):
        u = L.Unit()
        u.in_len = rng.randrange(1000, 400_000)
        units.append(u)
    for n in (1, 2, 4, 8):
        sh = B.shard_units(units, n)
        loads = [sum(u.in_len for u, s in zip(units, sh) if s == r) for r in range(n)]
        assert set(sh) == set(range(n))
        assert max(loads) - min(loads) <= 400_000
        assert sh == B.shard_units(units, n)     # deterministic: every rank computes the same plan


def test_no_device_means_hard_error():
    lib = L.lib()
    if lib.lzgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(L.LzgpuError) as e:
        B.Context()
    assert e.value.code == L.E_NO_DEVICE and "no CPU" in str(e.value)
    assert not lib.lzgpu_alloc_pinned(4096)             # pinned buffers need a device too
    assert b"no CUDA device" in lib.lzgpu_last_error()
    lib.lzgpu_free_pinned(None)
