"""GPU tier: liblzgpu.so through the C ABI vs the oracle, on a real B200."""
import ctypes as C
import hashlib
import zlib

import numpy as np
import pytest

import cases
from check import same_outcome
from lzma_b200 import _lib as L
from lzma_b200 import batch as B
from lzma_b200 import corpus as K
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = B.Context([0])
    yield c
    c.close()


def test_reference_assets(ctx):
    """The reference's own test table (reader1_test.go:15-107, reader2_test.go:12-29)."""
    names = cases.ALONE_ASSETS
    got = B.decode_alone_streams(ctx, [cases.asset(n) for n in names])
    for n, g in zip(names, got):
        same_outcome(O.lzma_alone(cases.asset(n), 2 << 20), g.status, g.err_site, g.data, n)
    by = dict(zip(names, got))
    for n in ("a.lzma", "a_eos.lzma", "a_eos_and_size.lzma", "a_lp1_lc2_pb1.lzma"):
        assert hashlib.md5(by[n].data).hexdigest() == "57a42eb7f425c13fa644f2618a097ab7"
    assert hashlib.md5(by["randomfile.dat.lzma"].data).hexdigest() == "b2d18c4275c394a729607ff9fe0caae7"
    assert (by["bad_corrupted.lzma"].status, by["bad_corrupted.lzma"].err_site) == (L.RESULT_ERROR, 652)
    assert (by["bad_eos_incorrect_size.lzma"].status, by["bad_eos_incorrect_size.lzma"].err_site) == (L.RESULT_ERROR, 636)
    assert (by["bad_incorrect_size.lzma"].status, by["bad_incorrect_size.lzma"].err_site) == (L.RESULT_ERROR, 46)
    st, site, data = B.decode_lzma2_stream(ctx, cases.asset("randomfile.dat.lzma2"), 0)
    assert st == L.OK and hashlib.md5(data).hexdigest() == "b2d18c4275c394a729607ff9fe0caae7"


def test_alone_cases(ctx):
    """Mixed batch: every lc/lp/pb liblzma can write, EOS-only / EOS+size / size-only streams, dictionary
    wrap, runs, truncations, bit flips, bad headers -- one batch, one bad unit must not poison the rest."""
    cs = cases.alone_cases(heavy=True)
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)


def test_encoder_cases(ctx):
    """All 225 prop bytes the reference accepts (lc+lp up to 12: literal tables in HBM), size-only
    streams, odd dictionaries (Q3), a match at position 0 (Q4), corrupt streams with large tables."""
    cs = cases.encoder_cases(heavy=True)
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)


def test_lzma2_cases(ctx):
    for name, s, dict_size, cap in cases.lzma2_cases():
        st, site, data = B.decode_lzma2_stream(ctx, s, dict_size)
        same_outcome(O.lzma2(s, dict_size, cap + (1 << 20)), st, site, data, name, strict_site=False)


def test_corruption_fuzz(ctx):
    """1 200 hostile variants of valid streams in ONE batch (damage met while the fast decoder runs): status,
    error site and delivered bytes equal the oracle's for every one, and no unit writes outside its own output
    range (64-byte canaries between the ranges; compute-sanitizer is not available on this pool)."""
    cs = cases.fuzz_cases()
    units, in_buf, _, _ = B.build_alone_batch([c[1] for c in cs], [c[2] for c in cs])
    off = 64
    for u in units:
        u.out_off = off
        off = _ru16(off + u.out_cap) + 64
    out = np.full(off, 0xA5, dtype=np.uint8)
    res, _ = ctx.decode_batch(units, in_buf, out)
    kinds = {}
    prev_end = 0
    for (name, s, cap), u, r in zip(cs, units, res):
        want = O.lzma_alone(s, cap)
        same_outcome(want, r.status, r.err_site, out[u.out_off:u.out_off + r.bytes_out].tobytes(), name)
        assert r.bytes_out <= u.out_cap
        assert (out[prev_end:u.out_off] == 0xA5).all(), f"{name}: bytes before the unit's output range were written"
        prev_end = u.out_off + u.out_cap
        kinds[want.status_name] = kinds.get(want.status_name, 0) + 1
    assert (out[prev_end:] == 0xA5).all()
    assert len(kinds) >= 4, kinds       # OK, input exhausted, result error, overflow ... all occur


def test_unknown_size_retry(ctx):
    d = K.text_block(123, 3_000_000)     # ratio > the first capacity guess
    g = B.decode_alone_streams(ctx, [K.compress_alone(d, preset=1)])[0]
    assert g.status == L.OK and g.data == d


def test_device_resident_plan(ctx):
    """plan_create / plan_launch / plan_results with buffers already in HBM (what bench.py times)."""
    import torch
    plains = [K.text_block(i, 100_000 + 1000 * i) for i in range(40)]
    streams = [K.compress_alone(p) for p in plains]
    units, in_buf, out_size, _ = B.build_alone_batch(streams, [len(p) for p in plains])
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.zeros(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, out_size)
    assert plan.launch_count == 1
    for _ in range(2):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
    torch.cuda.synchronize()
    res, st = plan.results()
    out = d_out.cpu().numpy()
    for k, p in enumerate(plains):
        assert res[k].status == L.OK and res[k].bytes_out == len(p)
        assert out[units[k].out_off:units[k].out_off + len(p)].tobytes() == p
    assert st.kernel_ms > 0
    plan.close()


def test_device_crc32_of_units(ctx):
    """lzgpu_plan_crc32 (on-device verification) == zlib.crc32 of each unit's decoded bytes: ragged sizes around
    the slicing and alignment boundaries, an empty unit, a failed unit (CRC of the bytes decoded before the
    error), unaligned output offsets."""
    import torch
    sizes = [0, 1, 3, 4, 5, 15, 16, 17, 255, 256, 257, 1023, 4096, 4097, 65_537, 300_001, 1 << 20]
    plains = [K.text_block(700 + i, n) if n else b"" for i, n in enumerate(sizes)]
    streams = [K.compress_alone(p) for p in plains]
    bad = bytearray(streams[-2]); bad[len(bad) // 2] ^= 0x10
    streams.append(bytes(bad)); plains.append(None)
    units, in_buf, out_size, _ = B.build_alone_batch(streams, [max(len(p), 1) if p is not None else 300_001 for p in plains])
    off = 0
    for k, u in enumerate(units):          # odd output offsets: the CRC kernel must not assume alignment
        u.out_off = off + (k % 7)
        off = _ru16(u.out_off + u.out_cap)
    out_size = off + 16
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.zeros(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, out_size)
    plan.launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
    crc = plan.crc32(d_out.data_ptr())
    res, _ = plan.results()
    out = d_out.cpu().numpy()
    for k, p in enumerate(plains):
        got = out[units[k].out_off:units[k].out_off + res[k].bytes_out].tobytes()
        if p is not None:
            assert res[k].status == L.OK and got == p, k
        assert int(crc[k]) == zlib.crc32(got), (k, len(got))
    plan.close()
    # many units at once (more than the CRC kernel's grid, every CTA busy from its first cycle): 5 distinct
    # streams, each unit decoding into its own range
    distinct, n_units, size = 5, 5000, 48_000
    plains = [K.text_block(800 + i, size) for i in range(distinct)]
    t_units, in_buf, _, _ = B.build_alone_batch([K.compress_alone(p) for p in plains], [size] * distinct)
    units = (L.Unit * n_units)()
    for k in range(n_units):
        C.memmove(C.byref(units[k]), C.byref(t_units[k % distinct]), C.sizeof(L.Unit))
        units[k].out_off, units[k].out_cap = k * size, size
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.zeros(n_units * size + 16, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, n_units * size + 16)
    plan.launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
    for _ in range(3):
        crc = plan.crc32(d_out.data_ptr())
        want = np.array([zlib.crc32(plains[k % distinct]) for k in range(n_units)], dtype=np.uint32)
        assert np.array_equal(crc, want), np.nonzero(crc != want)[0][:8]
    plan.close()


def _ru16(x):
    return (x + 15) // 16 * 16


def test_full_size_batch_properties(ctx):
    """BASELINE-sized units (1 MiB) in a batch larger than the SM count: checksum per unit against the
    plaintext's, and the decoded size; every unit of the same stream must give the same bytes."""
    distinct = 12
    plains = [K.text_block(1000 + i, 1 << 20) for i in range(distinct)]
    streams = [K.compress_alone(p) for p in plains]
    n = 300
    pick = [i % distinct for i in range(n)]
    got = B.decode_alone_streams(ctx, [streams[i] for i in pick], [1 << 20] * n)
    crcs = [zlib.crc32(p) for p in plains]
    for i, g in zip(pick, got):
        assert g.status == L.OK and len(g.data) == 1 << 20 and zlib.crc32(g.data) == crcs[i]


@pytest.mark.parametrize("push", ["push", "streamed"])
@pytest.mark.parametrize("shift", [0, 3])
def test_pinned_host_buffers(ctx, shift, push, monkeypatch):
    """Pinned caller buffers take the zero-copy route (units read the compressed input straight from host
    memory and write their decoded bytes to the caller's buffer themselves, block by block -- or, with
    LZGPU_NO_PUSH_D2H=1, finished blocks are copied out by the host while the kernel runs and the unit tails by one
    kernel); pageable buffers and LZGPU_NO_ZEROCOPY_IN=1 / LZGPU_NO_TAIL_KERNEL=1 take the slab route.  Same bytes,
    same results, whatever the alignment."""
    import torch
    if push == "streamed":
        monkeypatch.setenv("LZGPU_NO_PUSH_D2H", "1")
    distinct = 6
    plains = [K.text_block(2000 + i, (1 << 20) - 37 * i) for i in range(distinct)]
    streams = [K.compress_alone(p) for p in plains]
    bad = bytearray(streams[0]); bad[len(bad) // 2] ^= 0x55
    n = 70
    pick = [i % distinct for i in range(n)]
    ss = [streams[i] for i in pick] + [bytes(bad), streams[1][:5000]]
    units, in_np, out_size, _ = B.build_alone_batch(ss, [1 << 20] * len(ss))
    pin_in = torch.empty(in_np.size + 64, dtype=torch.uint8).pin_memory()
    pin_out = torch.empty(out_size + 64, dtype=torch.uint8).pin_memory()
    a_in = pin_in.numpy()[shift:shift + in_np.size]
    a_in[:] = in_np
    a_out = pin_out.numpy()[shift:shift + out_size]
    a_out[:] = 0xEE
    res, st = ctx.decode_batch(units, a_in, a_out)
    monkeypatch.setenv("LZGPU_NO_ZEROCOPY_IN", "1")
    monkeypatch.setenv("LZGPU_NO_TAIL_KERNEL", "1")
    b_out = np.full(out_size, 0xEE, dtype=np.uint8)
    res2, st2 = ctx.decode_batch(units, in_np, b_out)
    assert st.h2d_ms < st2.h2d_ms or st2.h2d_ms == 0
    for k, (r, r2, u) in enumerate(zip(res, res2, units)):
        assert (r.status, r.err_site, r.bytes_out, r.bytes_in, r.final_code) == (r2.status, r2.err_site, r2.bytes_out, r2.bytes_in, r2.final_code), k
        assert a_out[u.out_off:u.out_off + r.bytes_out].tobytes() == b_out[u.out_off:u.out_off + r.bytes_out].tobytes(), k
        if k < n:
            assert r.status == L.OK and a_out[u.out_off:u.out_off + r.bytes_out].tobytes() == plains[pick[k]], k
    assert res[n].status != L.OK or a_out[units[n].out_off:units[n].out_off + res[n].bytes_out].tobytes() != plains[0]
    assert res[n + 1].status == L.OK_INPUT_EXHAUSTED


def test_push_mode_falls_back_to_streamed_copies(monkeypatch):
    """Push mode times its block pushes; a shard that found the link to the host too busy (here: any shard, threshold 0)
    makes its device take the copy-engine path from the next call on.  Both paths must leave the same bytes."""
    import torch
    monkeypatch.setenv("LZGPU_PUSH_SLOW_KC", "0")
    plains = [K.text_block(2300 + i, 1 << 20) for i in range(4)]
    streams = [K.compress_alone(p) for p in plains]
    n = 80                                             # 80 MiB: enough blocks to judge, enough bytes for the streamed path
    ss = [streams[i % 4] for i in range(n)]
    units, in_np, out_size, _ = B.build_alone_batch(ss, [1 << 20] * n)
    pin_in = torch.from_numpy(in_np.copy()).pin_memory()
    pin_out = torch.empty(out_size, dtype=torch.uint8).pin_memory()
    with B.Context([0]) as c:
        outs, d2h = [], []
        for _ in range(3):
            pin_out.fill_(0xEE)
            res, st = c.decode_batch(units, pin_in.numpy(), pin_out.numpy())
            assert all(r.status == L.OK and r.bytes_out == 1 << 20 for r in res)
            outs.append(pin_out.numpy().copy())
            d2h.append(st.d2h_ms)
    for k, u in enumerate(units):
        want = plains[k % 4]
        for o in outs:
            assert o[u.out_off:u.out_off + (1 << 20)].tobytes() == want, k
    # first call pushed (nothing follows the kernel), later ones streamed (the unit tails follow it): told apart by the
    # time between the kernel's end and the end of the copies -- a timing, so not worth a failure when it is unclear
    if not (d2h[0] < d2h[1] and d2h[0] < d2h[2]):
        pytest.skip(f"bytes are right, but the two paths could not be told apart by their timing: {d2h}")


def test_library_pinned_buffers(ctx):
    """lzgpu_alloc_pinned: what a cgo caller uses for its buffers; the batch call takes the zero-copy route on them."""
    lib = L.lib()
    plains = [K.text_block(2100 + i, 700_000) for i in range(3)]
    streams = [K.compress_alone(p) for p in plains] * 16
    units, in_np, out_size, _ = B.build_alone_batch(streams, [700_000] * len(streams))
    p_in, p_out = lib.lzgpu_alloc_pinned(in_np.size), lib.lzgpu_alloc_pinned(out_size)
    assert p_in and p_out
    try:
        a_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_uint8)), (in_np.size,))
        a_out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_uint8)), (out_size,))
        a_in[:] = in_np
        res, st = ctx.decode_batch(units, a_in, a_out)
        assert st.h2d_ms < 0.5          # nothing was copied ahead of the kernel
        for k, (r, u) in enumerate(zip(res, units)):
            assert r.status == L.OK and a_out[u.out_off:u.out_off + r.bytes_out].tobytes() == plains[k % 3], k
    finally:
        lib.lzgpu_free_pinned(p_in)
        lib.lzgpu_free_pinned(p_out)


def test_large_literal_tables_in_hbm(ctx):
    """lc+lp > 4 needs literal tables beyond shared memory (the reference accepts any prop < 225,
    reader1.go:210-221).  liblzma cannot write such streams, so build one by re-labelling: a stream
    coded with lc=4,lp=0 read as lc=4 via a RAW unit whose declared lp is larger is NOT equivalent,
    hence the check here is oracle == GPU on the same bytes, whatever they decode to."""
    d = K.text_block(5, 60_000)
    s = K.compress_alone(d, 4, 0, 2, 1 << 16)
    for prop in (4 + 9 * (4 + 5 * 2), 8 + 9 * (0 + 5 * 0), 8 + 9 * (4 + 5 * 4)):   # lc4 lp4, lc8 lp0, lc8 lp4
        t = bytes([prop]) + s[1:]
        want = O.lzma_alone(t, 200_000)
        g = B.decode_alone_streams(ctx, [t], [200_000])[0]
        same_outcome(want, g.status, g.err_site, g.data, f"prop{prop}")


@pytest.mark.parametrize("variant", [0, 1, 33])
def test_tuning_variants_agree(variant, monkeypatch):
    """Every decoder tuning variant (LZGPU_VARIANT, lzgpu_core.cuh V_*) is bit-exact."""
    monkeypatch.setenv("LZGPU_VARIANT", str(variant))
    with B.Context([0]) as c:
        cs = cases.alone_cases(heavy=False)
        got = B.decode_alone_streams(c, [x[1] for x in cs], [x[2] for x in cs])
        for (name, s, cap), g in zip(cs, got):
            same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, f"v{variant}:{name}")
        for name, s, dict_size, cap in cases.lzma2_cases()[:6]:
            st, site, data = B.decode_lzma2_stream(c, s, dict_size)
            same_outcome(O.lzma2(s, dict_size, cap + (1 << 20)), st, site, data, f"v{variant}:{name}", strict_site=False)


def test_lzma2_stream_many_units(ctx):
    """BASELINE config 3 at reduced size: one raw LZMA2 stream with a dictionary reset every 1 MiB;
    the host scanner cuts it into units that decode in parallel.  Compared with the plaintext."""
    n_blocks = 48
    blocks = [K.text_block(500 + i, 1 << 20) for i in range(4)]
    parts = [K.compress_raw_lzma2(b) for b in blocks]
    seq = [i % 4 for i in range(n_blocks)]
    stream = b"".join(parts[i][:-1] for i in seq[:-1]) + parts[seq[-1]]
    units, total, sst = B.scan_lzma2(stream, 8 << 20)
    assert len(units) == n_blocks and total == n_blocks << 20 and sst == L.OK
    st, site, data = B.decode_lzma2_stream(ctx, stream, 8 << 20)
    assert st == L.OK and len(data) == total
    for k, i in enumerate(seq):
        assert zlib.crc32(data[k << 20:(k + 1) << 20]) == zlib.crc32(blocks[i]), f"block {k}"


def test_mixed_kinds_one_batch(ctx):
    """BASELINE config 4: .lzma units, headerless LZMA1 units, LZMA2 groups (with uncompressed chunks),
    incompressible data and corrupt streams in ONE lzgpu_decode_batch call; a bad unit must not poison
    its neighbours."""
    plain = [K.text_block(1, 90_000), K.random_block(2, 50_000), K.mixed_block(3, 120_000)]
    alone = [K.compress_alone(p, lc, lp, pb) for p, (lc, lp, pb) in zip(plain, [(3, 0, 2), (0, 4, 0), (4, 0, 4)])]
    l2_data = [K.text_block(4, 300_000), K.random_block(5, 100_000), K.text_block(6, 150_000)]
    l2 = K.lzma2_with_resets(l2_data, dict_size=1 << 20)
    l2_units, l2_total, _ = B.scan_lzma2(l2, 1 << 20)
    bad = [cases.asset("bad_corrupted.lzma"), cases.asset("bad_incorrect_size.lzma"), alone[0][:5000]]
    blobs, units, off, out_off = [], [], 0, 0

    def add(u, blob, cap):
        nonlocal off, out_off
        u.in_off += off
        u.out_off += out_off
        units.append(u)

    layout = []
    for s in alone + bad:
        st, u = B.parse_alone_header(s)
        u.kind = L.KIND_LZMA1_ALONE
        u.in_off, u.in_len, u.out_off, u.out_cap = off, len(s), out_off, 200_000
        units.append(u)
        blobs.append(s)
        layout.append(("alone", s, 200_000))
        off += (len(s) + 15) & ~15
        blobs.append(b"\0" * (off - sum(map(len, blobs))))
        out_off += 200_016
    raw = L.Unit()   # sevenzip-style headerless unit: props from the caller
    s = alone[2]
    raw.kind, raw.lc, raw.lp, raw.pb, raw.dict_size, raw.unpack_size = L.KIND_LZMA1_RAW, 4, 0, 4, 8 << 20, len(plain[2])
    raw.in_off, raw.in_len, raw.out_off, raw.out_cap = off, len(s) - 13, out_off, len(plain[2])
    units.append(raw)
    blobs.append(s[13:])
    off += (len(s) - 13 + 15) & ~15
    blobs.append(b"\0" * (off - sum(map(len, blobs))))
    out_off += (len(plain[2]) + 15) & ~15
    l2_base_in, l2_base_out = off, out_off
    for u in l2_units:
        u.in_off += l2_base_in
        u.out_off += l2_base_out
        units.append(u)
    blobs.append(l2)
    in_buf = np.frombuffer(b"".join(blobs) + bytes(16), dtype=np.uint8)
    out_buf = np.zeros(out_off + l2_total + 16, dtype=np.uint8)
    res, st = ctx.decode_batch(units, in_buf, out_buf)
    k = 0
    for kind, s, cap in layout:
        want = O.lzma_alone(s, cap)
        r, u = res[k], units[k]
        same_outcome(want, r.status, r.err_site, out_buf[u.out_off:u.out_off + r.bytes_out].tobytes(), f"unit{k}")
        k += 1
    r, u = res[k], units[k]
    assert r.status == L.OK and out_buf[u.out_off:u.out_off + r.bytes_out].tobytes() == plain[2]
    k += 1
    got = b""
    for u in l2_units:
        assert res[k].status == L.OK
        got += out_buf[u.out_off:u.out_off + res[k].bytes_out].tobytes()
        k += 1
    assert got == b"".join(l2_data)
    assert st.launches >= 3      # three literal-table classes (lc+lp = 3, 4) and both kinds ran


def test_batch_over_all_visible_gpus():
    """lzgpu_decode_batch sharding over every GPU the box has (1 on the default test box)."""
    with B.Context() as c:
        plains = [K.text_block(900 + i, 60_000 + 5_000 * (i % 5)) for i in range(40)]
        got = B.decode_alone_streams(c, [K.compress_alone(p, preset=1) for p in plains])
        devs = set()
        for p, g in zip(plains, got):
            assert g.status == L.OK and g.data == p
            devs.add(g.device)
        assert len(devs) == c.n_devices or c.n_devices > len(plains)
        # the same over pinned caller buffers (every GPU reads its shard's input from host memory) and with
        # enough output per GPU (>= 32 MiB) for the streamed D2H + tail kernel
        import torch
        big = [K.text_block(950 + i, 1 << 20) for i in range(4)]
        streams = [K.compress_alone(b, preset=1) for b in big]
        n = 36 * c.n_devices
        units, in_np, out_size, _ = B.build_alone_batch([streams[i % 4] for i in range(n)], [1 << 20] * n)
        h_in = torch.empty(in_np.size, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(out_size, dtype=torch.uint8).pin_memory()
        h_in.numpy()[:] = in_np
        res, st = c.decode_batch(units, h_in.numpy(), h_out.numpy())
        out = h_out.numpy()
        assert st.devices == c.n_devices
        for k, (r, u) in enumerate(zip(res, units)):
            assert r.status == L.OK and out[u.out_off:u.out_off + r.bytes_out].tobytes() == big[k % 4], k


def test_go_binding_marshalling(ctx, tmp_path):
    """go/lzgpu.go cannot be compiled here (no Go toolchain); tests/cpp/go_marshal_test.c restates its
    Unit -> C.lzgpu_unit marshalling field by field and decodes an xz-made (lc=3) LZMA2 stream with dictionary
    resets on the GPU, both with the scanner's table sizes carried through and with them dropped (the round-1
    Go file dropped lit_bits and every such stream failed with site 9101; the library now derives them)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", "go_marshal_test")
    assert os.path.exists(exe), "build it: make -C lzma_b200/csrc reader"
    blocks = [K.text_block(3100, 300_000), K.random_block(3101, 80_000), K.mixed_block(3102, 250_000), K.text_block(3103, 1 << 20)]
    stream = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    (tmp_path / "s.lzma2").write_bytes(stream)
    (tmp_path / "p.bin").write_bytes(b"".join(blocks))
    r = subprocess.run([exe, str(tmp_path / "s.lzma2"), str(tmp_path / "p.bin"), str(1 << 20)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "A (ScanLZMA2" in r.stdout and "B (table sizes not carried)" in r.stdout


def test_units_without_table_sizes(ctx):
    """LZMA2 units whose lit_bits / pos_bits were never filled in (flags without UF_BITS_KNOWN): lzgpu_decode_batch
    walks the chunk headers itself.  lc/lp/pb change between the blocks, pb up to 4 (full posState tables)."""
    combos = [(3, 0, 2), (0, 2, 0), (4, 0, 4), (1, 3, 3), (2, 2, 1)]
    blocks = [K.text_block(3200 + i, 150_000 + 7 * i) for i in range(len(combos))]
    parts = [K.compress_raw_lzma2(b, lc, lp, pb, 1 << 20) for b, (lc, lp, pb) in zip(blocks, combos)]
    stream = b"".join(p[:-1] for p in parts[:-1]) + parts[-1]
    units, total, sst = B.scan_lzma2(stream, 1 << 20)
    assert len(units) == len(combos) and sst == L.OK
    for u, (lc, lp, pb) in zip(units, combos):
        assert u.flags & L.UF_BITS_KNOWN and u.lit_bits == lc + lp and u.pos_bits == pb
    in_buf = np.frombuffer(stream, dtype=np.uint8)
    want = b"".join(blocks)
    for strip in (False, True):
        us = (L.Unit * len(units))(*units)
        if strip:
            for u in us:
                u.lit_bits = u.pos_bits = 0
                u.flags &= ~L.UF_BITS_KNOWN
        out = np.zeros(total + 16, dtype=np.uint8)
        res, _ = ctx.decode_batch(us, in_buf, out)
        assert all(r.status == L.OK for r in res), [(r.status, r.err_site) for r in res]
        assert out[:total].tobytes() == want


def test_output_gaps_are_left_alone(ctx):
    """Bytes of the caller's output buffer that belong to no unit -- the gaps between small units' ranges and
    whatever lies behind bytes_out inside a range -- must come back untouched (ranges closer than
    64 KiB used to be merged into one D2H copy that overwrote the gaps with stale device memory; on several
    GPUs that also raced with the neighbour's copy).  Runs on every visible GPU."""
    with B.Context() as all_ctx:
        for rounds in range(2):       # the second round finds the device slab dirty from the first
            plains = [K.text_block(3300 + i, 20_000 + 997 * i) for i in range(24)]
            streams = [K.compress_alone(p) for p in plains]
            caps = [len(p) + (0 if i % 3 else 4_000) for i, p in enumerate(plains)]     # some units do not fill their range
            units, in_buf, _, _ = B.build_alone_batch(streams, caps)
            off = 100
            for i, u in enumerate(units):
                u.out_off = off
                off += u.out_cap + (1 + 61 * i) % 4096      # gaps of 1 .. 4 KiB
            out = np.full(off + 64, 0xC3 if rounds == 0 else 0x3C, dtype=np.uint8)
            fill = out[0]
            res, _ = all_ctx.decode_batch(units, in_buf, out)
            prev = 0
            for p, u, r in zip(plains, units, res):
                assert r.status == L.OK and out[u.out_off:u.out_off + r.bytes_out].tobytes() == p
                assert (out[prev:u.out_off] == fill).all(), "gap before a unit's range was overwritten"
                assert (out[u.out_off + r.bytes_out:u.out_off + u.out_cap] == fill).all(), "bytes behind bytes_out were overwritten"
                prev = u.out_off + u.out_cap
            assert (out[prev:] == fill).all()


def test_lzma2_uncompressed_chunks_every_alignment(ctx):
    """LZMA2 uncompressed chunks (reader2.go:252-294) go through the warp's 16-byte copy: every source /
    destination alignment, sizes around the vector and tail boundaries, and a stream that mixes them with
    LZMA chunks and dictionary resets.  Compared with the oracle and with the plaintext."""
    rng = np.random.default_rng(99)

    def raw_chunks(data: bytes, first_reset: bool, step: int) -> bytes:
        out, pos, first = bytearray(), 0, True
        while pos < len(data):
            n = min(step, len(data) - pos, 1 << 16)
            out += bytes([1 if (first and first_reset) else 2, (n - 1) >> 8, (n - 1) & 0xFF]) + data[pos:pos + n]
            pos += n
            first = False
        return bytes(out)

    streams, plains = [], []
    for size in [1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 100, 255, 256, 4095, 4096, 4097, 65_535, 65_536, 200_003]:
        for step in (1 << 16, 1000 + size % 13):
            d = bytes(rng.integers(0, 256, size, dtype=np.uint8))
            streams.append(raw_chunks(d, True, step) + b"\0")
            plains.append(d)
    # mixed: LZMA chunks, then incompressible data stored uncompressed by liblzma itself, dictionary resets between
    blocks = [K.text_block(3400, 70_001), K.random_block(3401, 150_000), K.text_block(3402, 33_333), K.random_block(3403, 65_537)]
    streams.append(K.lzma2_with_resets(blocks, dict_size=1 << 20))
    plains.append(b"".join(blocks))
    for shift in (0, 1, 5, 11):      # destination alignment: units laid out back to back at odd offsets
        all_units, blob, out_off = [], bytearray(b"\0" * shift), shift
        spans = []
        for s in streams:
            units, total, sst = B.scan_lzma2(s, 1 << 20)
            assert sst == L.OK
            for u in units:
                u.in_off += len(blob)
                u.out_off += out_off
            spans.append((out_off, total))
            all_units += units
            blob += s
            out_off += total + (shift | 1)
        in_buf = np.frombuffer(bytes(blob) + b"\0" * 16, dtype=np.uint8)
        out = np.full(out_off + 32, 0x5A, dtype=np.uint8)
        res, _ = ctx.decode_batch(all_units, in_buf, out)
        assert all(r.status == L.OK for r in res), [(k, r.status, r.err_site) for k, r in enumerate(res) if r.status != L.OK][:5]
        prev = 0
        for (o, n), p in zip(spans, plains):
            assert out[o:o + n].tobytes() == p, (shift, n)
            assert (out[prev:o] == 0x5A).all()
            prev = o + n
    want = O.lzma2(streams[-1], 1 << 20, len(plains[-1]) + (1 << 20))
    assert want.status == O.OK and want.data == plains[-1]


def test_long_and_overlapping_matches(ctx):
    """Matches longer than 32 bytes and matches that overlap themselves (distance < length) take the general
    copy code: runs of one byte, short periods (2..40), long repeats at far distances, at every residue."""
    rng = np.random.default_rng(5)
    parts = []
    for per in list(range(1, 41)) + [63, 64, 65, 100, 272, 273, 274]:
        pat = bytes(rng.integers(0, 256, per, dtype=np.uint8))
        parts.append(pat * (700 // per + 3))
        parts.append(bytes(rng.integers(0, 256, 7 + per % 5, dtype=np.uint8)))
    far = bytes(rng.integers(0, 256, 5000, dtype=np.uint8))
    parts += [far, K.text_block(3500, 3000), far[100:4000], b"\0" * 3000, far[:273], far[1000:1500]]
    d = b"".join(parts)
    cs = []
    for lc, lp, pb in [(3, 0, 2), (0, 0, 0), (4, 0, 4)]:
        s = K.compress_alone(d, lc, lp, pb, 1 << 16, preset=9)
        cs.append((f"overlap lc{lc}lp{lp}pb{pb}", s, len(d)))
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)
        assert g.data == d


def _hetero_streams(n_distinct):
    """Streams of very different lengths and compressibility (text, noise, mixed; 24 KiB ... 320 KiB): what makes units
    of one SM finish at different times."""
    plains = []
    for i in range(n_distinct):
        size = (24 + 37 * (i % 9)) << 10
        kind = (K.text_block, K.random_block, K.mixed_block)[i % 3]
        plains.append(kind(7000 + i, size))
    return plains, [K.compress_alone(p, size_mode=("eos", "eos+size")[i % 2]) for i, p in enumerate(plains)]


@pytest.mark.parametrize("n_units,rotate", [(2072, "1"), (2500, "3"), (5000, "16"), (700, "2")])
def test_sm_scheduler_time_slicing(n_units, rotate, monkeypatch):
    """The SM-resident scheduler (lzgpu_sm_kernel): up to 14 units per CTA, units exchanged between warps at every
    `rotate`-th refill of the input stage, units handed to idle warps when others finish, later waves started from
    the global counter.  Every unit's bytes must be those of the plaintext, whatever warp decoded which part, and
    equal to what the one-warp CTAs (LZGPU_SCHED=0) produce."""
    plains, streams = _hetero_streams(27)
    pick = [(k * 7 + k // 27) % 27 for k in range(n_units)]
    crcs = [zlib.crc32(p) for p in plains]
    outs = {}
    for sched in ("1", "0"):
        monkeypatch.setenv("LZGPU_SCHED", sched)
        monkeypatch.setenv("LZGPU_ROTATE", rotate)
        with B.Context([0]) as c:
            got = B.decode_alone_streams(c, [streams[i] for i in pick], [len(plains[i]) for i in pick])
        for k, (i, g) in enumerate(zip(pick, got)):
            assert g.status == L.OK and len(g.data) == len(plains[i]) and zlib.crc32(g.data) == crcs[i], (sched, k, i, g.status, g.err_site)
        outs[sched] = [(g.status, g.err_site, g.bytes_in) for g in got]
    assert outs["1"] == outs["0"]


def test_sm_scheduler_with_bad_units_and_lzma2(monkeypatch):
    """Corrupt LZMA1 units (they end early: their warps go idle and take over others' units) and LZMA2 groups (not
    time-sliced) next to sliced LZMA1 units in the same CTAs, in ONE call."""
    monkeypatch.setenv("LZGPU_ROTATE", "1")
    plains, streams = _hetero_streams(12)
    l2_blocks = [K.text_block(4242 + i, (40 + 30 * i) << 10) for i in range(4)]
    l2 = K.lzma2_with_resets(l2_blocks, dict_size=1 << 20)
    l2_plain = b"".join(l2_blocks)
    l2_units, l2_total, sst = B.scan_lzma2(l2, 1 << 20)
    assert sst == L.OK and l2_total == len(l2_plain) and len(l2_units) == 4
    n_alone, n_l2 = 1500, 120
    blobs, units, layout, off, out_off = [], [], [], 0, 0
    for k in range(n_alone):
        i = k % 12
        sb = bytearray(streams[i])
        if k % 5 == 0:
            sb[13 + 40 + (k % 97)] ^= 0x5A      # damage inside the range-coded part
        sb = bytes(sb)
        st, u = B.parse_alone_header(sb)
        u.kind = L.KIND_LZMA1_ALONE
        u.in_off, u.in_len, u.out_off, u.out_cap = off, len(sb), out_off, len(plains[i])
        units.append(u)
        layout.append((sb, len(plains[i])))
        blobs.append(sb + bytes(-len(sb) % 16))
        off += len(blobs[-1])
        out_off += (len(plains[i]) + 31) & ~15
        if k % (n_alone // n_l2) == 0:            # an LZMA2 stream's groups in between
            for t in l2_units:
                u2 = L.Unit()
                C.memmove(C.byref(u2), C.byref(t), C.sizeof(L.Unit))
                u2.in_off += off
                u2.out_off += out_off
                units.append(u2)
                layout.append(None)
            blobs.append(l2 + bytes(-len(l2) % 16))
            off += len(blobs[-1])
            out_off += (l2_total + 31) & ~15
    in_buf = np.frombuffer(b"".join(blobs) + bytes(16), dtype=np.uint8)
    out_buf = np.zeros(out_off + 16, dtype=np.uint8)
    with B.Context([0]) as c:
        res, st = c.decode_batch(units, in_buf, out_buf)
    k, n_bad = 0, 0
    while k < len(units):
        if layout[k] is not None:
            sb, cap = layout[k]
            r, u = res[k], units[k]
            want = O.lzma_alone(sb, cap)
            n_bad += r.status != L.OK
            same_outcome(want, r.status, r.err_site, out_buf[u.out_off:u.out_off + r.bytes_out].tobytes(), f"unit {k}")
            k += 1
        else:
            got = b""
            for j in range(4):
                assert res[k + j].status == L.OK, (k, j, res[k + j].status, res[k + j].err_site)
                got += out_buf[units[k + j].out_off:units[k + j].out_off + res[k + j].bytes_out].tobytes()
            assert got == l2_plain
            k += 4
    assert n_bad > 100


@pytest.mark.parametrize("rotate", ["1", "16"])
def test_sm_scheduler_slices_lzma2_groups(rotate, monkeypatch):
    """LZMA2 groups are time-sliced too (the chunk walk's state travels with the unit): one raw LZMA2 stream of 2 400
    groups of very different sizes -- text, incompressible data (uncompressed chunks) and mixed -- 16 per SM, exchanged
    between warps inside their LZMA chunks.  Compared with the plaintext, group by group."""
    monkeypatch.setenv("LZGPU_ROTATE", rotate)
    blocks = []
    for i in range(30):
        size = (20 + 41 * (i % 7)) << 10
        blocks.append((K.text_block, K.random_block, K.mixed_block)[i % 3](9100 + i, size))
    parts = [K.compress_raw_lzma2(b, dict_size=1 << 20) for b in blocks]
    n_groups = 2400
    seq = [(k * 11 + k // 30) % 30 for k in range(n_groups)]
    stream = b"".join(parts[i][:-1] for i in seq[:-1]) + parts[seq[-1]]
    with B.Context([0]) as c:
        units, total, sst = B.scan_lzma2(stream, 1 << 20)
        assert sst == L.OK and len(units) == n_groups and total == sum(len(blocks[i]) for i in seq)
        st, site, data = B.decode_lzma2_stream(c, stream, 1 << 20)
    assert st == L.OK and len(data) == total
    off = 0
    for k, i in enumerate(seq):
        n = len(blocks[i])
        assert data[off:off + n] == blocks[i], f"group {k} (block {i})"
        off += n
