// tests/emu/emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the *same* unit-decode code the CUDA kernel runs
// (lzma_b200/csrc/lzgpu_unit.cuh + lzgpu_core.cuh) for the host, with the 32 lanes
// of a warp emulated by loops and the deferred window stores reproduced exactly
// (a value is captured when the kernel would issue the load and written when the
// kernel would issue the store).  It lets the CPU-only test tier check the
// decoder's logic and the copy protocol against the oracle.  It is never linked
// into liblzgpu.so and is not a fallback: the product fails without a GPU.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../lzma_b200/csrc/lzgpu_prep.h"
#include "../../lzma_b200/csrc/lzgpu_unit.cuh"

using namespace lzgpu;

template <int kV>
static void run_one(const lzgpu_unit &u, const UnitIO &io, uint16_t *P, uint16_t *L, uint32_t bits, lzgpu_result &r) {
    if (u.kind == LZGPU_KIND_LZMA2_GROUP) run_unit_lzma2<kV>(u, io, P, L, bits, r);
    else run_unit_lzma1<kV>(u, io, P, L, r);
}

extern "C" int emu_decode_batch(const lzgpu_unit *units, int64_t n, const uint8_t *in_base, uint64_t in_size,
                                uint8_t *out_base, uint64_t out_size, lzgpu_result *results, int variant) {
    for (int64_t i = 0; i < n; i++) {
        lzgpu_unit u = units[i];
        if (u.in_off > in_size || u.in_len > in_size - u.in_off || u.out_off > out_size || u.out_cap > out_size - u.out_off)
            return LZGPU_E_INVALID;
        lzgpu_result r;
        bool alone = false;
        if (u.kind == LZGPU_KIND_LZMA2_GROUP && !(u.flags & LZGPU_UF_BITS_KNOWN)) derive_lzma2_bits(in_base, u);
        if (!prepare_unit(u, r, &alone)) { results[i] = r; continue; }
        const uint32_t bits = u.lit_bits;
        // same class choice as the device launch (lzgpu.cu, launch_decode): compact posState tables when pb <= 2
        const bool pb2 = u.pos_bits <= 2;
        const size_t fixed = pb2 ? Lay<2>::FIXED : Lay<4>::FIXED;
        std::vector<uint16_t> probs(fixed + ((size_t)0x300 << bits) + 64, 0xDEAD);
        UnitIO io;
        io.in = in_base + u.in_off;
        io.in_len = u.in_len;
        io.out = out_base + u.out_off;
        io.out_cap = u.out_cap;
        memset(&r, 0, sizeof r);
        r.status = LZGPU_NOT_RUN;
        uint16_t *P = probs.data(), *L = probs.data() + fixed;
        alignas(16) uint8_t stage[128];
        io.stage = stage;
        alignas(16) uint8_t inbuf[kF2Stage];
        io.inbuf = inbuf;
        io.progress = nullptr;
        io.hout = nullptr;
        io.push_stat = nullptr;
#define EMU_RUN(V) do { if (pb2) run_one<(V) | V_PB2>(u, io, P, L, bits, r); else run_one<(V)>(u, io, P, L, bits, r); } while (0)
        switch (variant) {
            case 0: EMU_RUN(0); break;
            case 33: EMU_RUN(33); break;
            default: EMU_RUN(1); break;
        }
#undef EMU_RUN
        if (alone) r.bytes_in += 13;
        r.device = -1;
        results[i] = r;
    }
    return LZGPU_E_OK;
}
