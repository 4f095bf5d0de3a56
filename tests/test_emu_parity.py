"""CPU tier: the kernel's unit-decode code (lzgpu_unit.cuh / lzgpu_core.cuh) compiled for the host
with the warp's lanes emulated, against the oracle.  Checks decoder logic and the deferred-store
copy protocol without a GPU; the GPU tier (test_gpu_parity.py) repeats it through liblzgpu.so."""
import numpy as np
import pytest

import cases
from backends import make_context
from check import same_outcome
from lzma_b200 import _lib as L
from lzma_b200 import batch as B
from oracle import oracle as O


@pytest.fixture(scope="module", params=[0, 1, 33])
def ctx(request):
    """Every tuning variant of the decoder (lzgpu_core.cuh V_*) must give identical results."""
    return make_context("emu", request.param)


def test_alone_cases(ctx):
    cs = cases.alone_cases(heavy=False)
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)


def test_lzma2_cases(ctx):
    for name, s, dict_size, cap in cases.lzma2_cases():
        st, site, data = B.decode_lzma2_stream(ctx, s, dict_size)
        same_outcome(O.lzma2(s, dict_size, cap + (1 << 20)), st, site, data, name, strict_site=False)


def test_encoder_cases(ctx):
    """Props beyond liblzma's limits, odd dictionaries (Q3), match at position 0 (Q4)."""
    cs = cases.encoder_cases(heavy=False)
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)


def test_corruption_fuzz(ctx):
    """Hostile variants of valid streams (a third of the GPU tier's set): status, site and bytes == oracle."""
    if ctx.variant not in (1, 33):
        pytest.skip("fuzz on the two shipped decoders only (the others are tuning experiments)")
    cs = cases.fuzz_cases(400)
    got = B.decode_alone_streams(ctx, [c[1] for c in cs], [c[2] for c in cs])
    for (name, s, cap), g in zip(cs, got):
        same_outcome(O.lzma_alone(s, cap), g.status, g.err_site, g.data, name)


def test_gpu_tier_cases_on_the_emulation(ctx):
    """Three GPU-tier tests whose logic does not need a device, run on the emulation too: LZMA2 units whose
    table sizes were not filled in (compact and full posState tables), uncompressed chunks at every alignment,
    long and self-overlapping matches through the general copy code."""
    import test_gpu_parity as T
    T.test_units_without_table_sizes(ctx)
    T.test_long_and_overlapping_matches(ctx)
    if ctx.variant == 33:
        T.test_lzma2_uncompressed_chunks_every_alignment(ctx)
