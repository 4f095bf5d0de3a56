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
        if (u.dict_size < 4096u) u.dict_size = 4096u;    // lzmaDicMin clamp, reader1.go:199-201
    } else {
        if (u.dict_size < 4096u) u.dict_size = 8u << 20; // validateDictSize, reader2.go:88-91
        if (u.lit_bits > 12) u.lit_bits = 12;
    }
    return true;
}

}  // namespace lzgpu
