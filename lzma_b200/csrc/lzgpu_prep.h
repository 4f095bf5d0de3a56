// lzgpu_prep.h -- host-side normalisation of a caller's unit into what the kernel sees.
// Shared by lzgpu.cu (plan_create) and the lane-emulation test harness.
#pragma once
#include <string.h>

#include "../../include/lzgpu.h"

namespace lzgpu {

// Returns true when the unit must be run on the device; false when its outcome is
// already decided (preset.status set).  *was_alone: a 13-byte header was rebased away.
inline bool prepare_unit(lzgpu_unit &u, lzgpu_result &preset, bool *was_alone) {
    memset(&preset, 0, sizeof preset);
    preset.status = LZGPU_NOT_RUN;
    *was_alone = false;
    if (u.kind == LZGPU_KIND_LZMA1_ALONE) {
        // header already parsed (lzgpu_parse_alone_header); the device sees a RAW unit
        if (u.in_len < 13) { preset.status = LZGPU_UNEXPECTED_EOF; return false; }
        u.in_off += 13;
        u.in_len -= 13;
        u.kind = LZGPU_KIND_LZMA1_RAW;
        *was_alone = true;
    }
    if (u.kind == LZGPU_KIND_LZMA1_RAW) {
        if (u.lc > 8 || u.lp > 4 || u.pb > 4) { preset.status = LZGPU_INCORRECT_PROPERTIES; return false; }
        u.lit_bits = (uint8_t)(u.lc + u.lp);
        u.pos_bits = u.pb;
        if (u.dict_size < 4096u) u.dict_size = 4096u;    // lzmaDicMin clamp, reader1.go:199-201
    } else {
        if (u.dict_size < 4096u) u.dict_size = 8u << 20; // validateDictSize, reader2.go:88-91
        if (u.lit_bits > 12) u.lit_bits = 12;
        // table sizes not derived from the chunk headers (by lzgpu_scan_lzma2 or lzgpu_decode_batch): the
        // literal tables are the caller's word, the posState tables the full ones
        if (!(u.flags & LZGPU_UF_BITS_KNOWN) || u.pos_bits > 4) u.pos_bits = 4;
    }
    return true;
}

// Walk the chunk headers of an LZMA2 unit held in host memory (Reader2.startChunk's framing, reader2.go:100-214)
// and set lit_bits / pos_bits to the largest lc+lp / pb any LZMA chunk of it uses, starting from the properties
// in force before the unit (u.lc / u.lp / u.pb).  Truncated or malformed framing just ends the walk: the device
// reports it.
inline void derive_lzma2_bits(const uint8_t *in, lzgpu_unit &u) {
    const uint8_t *p = in + u.in_off, *const end = p + u.in_len;
    uint32_t props = ((uint32_t)u.pb * 5 + u.lp) * 9 + u.lc, lit = 0, pos = 0;
    while (p < end) {
        const uint32_t c = *p;
        if (c == 0 || (c >= 3 && c < 0x80)) break;
        const uint32_t hl = c < 0x80 ? 3 : (c < 0xC0 ? 5 : 6);
        if ((uint64_t)(end - p) < hl) break;
        uint64_t payload;
        if (c >= 0x80) {
            payload = ((((uint32_t)p[3] << 8) | p[4]) + 1u);
            if (c >= 0xC0) props = p[5];
            if (props < 225) {
                const uint32_t lb = props % 9 + (props / 9) % 5, pb = (props / 9) / 5;
                if (lb > lit) lit = lb;
                if (pb > pos) pos = pb;
            }
        } else {
            payload = ((((uint32_t)p[1] << 8) | p[2]) + 1u);
        }
        if ((uint64_t)(end - p) - hl < payload) break;
        p += hl + payload;
    }
    u.lit_bits = (uint8_t)lit;
    u.pos_bits = (uint8_t)pos;
    u.flags |= LZGPU_UF_BITS_KNOWN;
}

}  // namespace lzgpu
