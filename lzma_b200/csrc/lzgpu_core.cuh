// lzgpu_core.cuh -- the in-stream LZMA symbol decoder that lane 0 of each warp runs,
// plus the per-lane phases of the warp-cooperative window copy.
//
// Follows (restated, not translated) the reference's live hot loop
// (*Reader1).decompress, decompress.go:8-1136, whose hand-inlined bit steps are the
// twins range_decoder.go:57-134, bit_tree_decoder.go:26-135, len_decoder.go:34-60,
// reader1.go:256-426; the 12-state machine state.go:153-187; window semantics
// window.go:31-95.  The ORDER of the size / EOS / distance checks is the
// reference's, because it decides which error a malformed stream reports.
//
// Differences in mechanism (results are identical):
//  * the output buffer in HBM *is* the dictionary window: there is no circular
//    buffer and no second ReadPending copy (window.go:101-133);
//  * match copies are executed by all 32 lanes and their stores are deferred
//    until the next copy, so the serial decoder never waits for a window load
//    unless the very next symbol is a literal that needs its context bytes;
//  * probability tables live in shared memory in our own layout (P_* below).
//
// The file compiles for the device (nvcc) and, with the LZ_HD macros degrading to
// plain inline functions, for the host lane-emulation harness under tests/emu/
// (test infrastructure; never linked into liblzgpu.so).
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/lzgpu.h"

#if defined(__CUDACC__)
#define LZ_HD __host__ __device__ __forceinline__
#else
#define LZ_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define LZ_WARP_SYNC() __syncwarp()
#define LZ_LD_IN8(p) __ldg(reinterpret_cast<const unsigned char *>(p))
#define LZ_LD_IN32(p) __ldg(reinterpret_cast<const unsigned int *>(p))
#define LZ_BSWAP32(x) __byte_perm((x), 0u, 0x0123u)
#define LZ_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define LZ_FUNNEL_L(lo, hi, sh) __funnelshift_l((lo), (hi), (sh))   /* ((hi:lo) << sh) >> 32, sh < 32 */
#define LZ_CLZ(x) ((uint32_t)__clz((int)(x)))
#define LZ_FUNNEL_R(lo, hi, sh) __funnelshift_r((lo), (hi), (sh))   /* ((hi:lo) >> (sh & 31)) low word */
#define LZ_SHR_CLAMP(x, sh) __funnelshift_rc((x), 0u, (sh))          /* x >> min(sh, 32) */
// predicated global load: no branch, so lane 0's instruction stream stays straight-line
// cp.async (LDGSTS): 4 aligned bytes global -> shared with no register and no scoreboard involved;
// completion is awaited explicitly (wait_group) where the bytes are needed
#define LZ_CP_ASYNC4(sdst, gsrc)                                                        \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory")
#define LZ_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define LZ_CP_WAIT() asm volatile("cp.async.wait_group 0;" ::: "memory")
// window byte -> 32-bit register, issued now, first touched when a literal needs it (anything
// the compiler inserts in between -- a mask, a move -- would stall on the load right here)
#define LZ_LD_WIN8(dst, p) asm volatile("ld.global.u8 %0, [%1];" : "=r"(dst) : "l"(p) : "memory")
#define LZ_LD_WIN32_IF(dst, p, cond)  /* window (global), written by this kernel: coherent load */ \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u32 %0, [%1];\n\t}" \
                 : "+r"(dst) : "l"(p), "r"((uint32_t)(cond)) : "memory")
#define LZ_LD_IN32_IF(dst, p, cond)                                                     \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}" \
                 : "+r"(dst) : "l"(p), "r"((uint32_t)(cond)))
#define LZ_LIKELY(x) __builtin_expect(!!(x), 1)
#define LZ_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define LZ_WARP_SYNC() ((void)0)
#define LZ_LD_IN8(p) (*(const uint8_t *)(p))
#define LZ_LD_IN32(p) (*(const uint32_t *)(p))
#define LZ_BSWAP32(x) __builtin_bswap32(x)
#define LZ_PREFETCH_L2(p) ((void)0)
#define LZ_FUNNEL_L(lo, hi, sh) ((uint32_t)(((((uint64_t)(hi)) << 32 | (uint64_t)(lo)) << (sh)) >> 32))
#define LZ_CLZ(x) ((uint32_t)__builtin_clz(x))
#define LZ_FUNNEL_R(lo, hi, sh) ((uint32_t)(((((uint64_t)(hi)) << 32 | (uint64_t)(lo)) >> ((sh) & 31))))
#define LZ_SHR_CLAMP(x, sh) ((sh) >= 32 ? 0u : ((uint32_t)(x) >> (sh)))
#define LZ_LD_IN32_IF(dst, p, cond) do { if (cond) (dst) = *(const uint32_t *)(p); } while (0)
#define LZ_LD_WIN32_IF(dst, p, cond) do { if (cond) (dst) = *(const uint32_t *)(p); } while (0)
#define LZ_LD_WIN8(dst, p) ((dst) = *(const uint8_t *)(p))
#define LZ_CP_ASYNC4(sdst, gsrc) memcpy((sdst), (gsrc), 4)   /* value captured at issue, like the copy */
#define LZ_CP_COMMIT() ((void)0)
#define LZ_CP_WAIT() ((void)0)
#define LZ_LIKELY(x) __builtin_expect(!!(x), 1)
#define LZ_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif

namespace lzgpu {

// ---- probability-table layout, in uint16 units (Appendix A of SURVEY.md lists the
// reference's tables: state.go:3-33).  isRep/G0/G1/G2 are interleaved per state.
// The posState-indexed tables hold 1 << kPB positions: kPB = 4 fits every pb the reference
// accepts (reader1.go:210-221); units whose pb is <= 2 (what xz / 7-zip write by default) run a
// kernel instantiated with kPB = 2, whose fixed tables take 2 400 bytes instead of 3 744 -- with
// lc+lp = 3 that is the difference between 13 and 14 units resident per SM.
template <int kPB>
struct Lay {
    static constexpr uint32_t PB = kPB, NPS = 1u << kPB;
    static constexpr uint32_t IS_MATCH = 0;                    // [12][NPS] (state<<kPB)+posState  decompress.go:23,26
    static constexpr uint32_t IS_REP0_LONG = 12 * NPS;         // [12][NPS]                        decompress.go:716
    static constexpr uint32_t REP4 = 24 * NPS;                 // [12][4] isRep, isRepG0, isRepG1, isRepG2 of one state
    // one length coder: choice, choice2 (+6 pad), low [NPS][8], mid [NPS][8], high [256]
    static constexpr uint32_t LEN_CHOICE = 0, LEN_CHOICE2 = 1, LEN_LOW = 8, LEN_MID = 8 + 8 * NPS,
                              LEN_HIGH = 8 + 16 * NPS, LEN_SIZE = 8 + 16 * NPS + 256;
    static constexpr uint32_t LEN0 = REP4 + 48;                // match length coder   decompress.go:218-429
    static constexpr uint32_t LEN1 = LEN0 + LEN_SIZE;          // rep length coder     decompress.go:870-1118
    static constexpr uint32_t POS_SLOT = LEN1 + LEN_SIZE;      // [4][64]              decompress.go:434-486
    static constexpr uint32_t POS_DEC = POS_SLOT + 256;        // [115] (+13 pad)      decompress.go:491-546
    static constexpr uint32_t ALIGN = POS_DEC + 128;           // [16]                 decompress.go:580-625
    static constexpr uint32_t FIXED = ALIGN + 16;              // 1 872 (kPB 4) / 1 200 (kPB 2) cells
    static constexpr uint32_t LIT = FIXED;                     // 0x300 << (lc+lp)     decompress.go:56-57
};
static_assert(Lay<4>::FIXED == 1872 && Lay<2>::FIXED == 1200, "table layout");

constexpr uint32_t kTop = 1u << 24;
constexpr uint32_t kProbInit = 1024;

// what lane 0 hands to the warp when it leaves decode_run()
enum : uint32_t { OP_COPY = 0, OP_COPY_Q4 = 1, OP_DONE = 2, OP_SWITCH = 3 };

// Lane 0's decoder registers.
struct Dec {
    uint32_t range, code;
    uint32_t inb_hi, inb_lo;        // 64-bit input lookahead, next byte in the top 8 bits of inb_hi
    uint32_t inbits;                // bits in the lookahead (real bytes first, then phantom zeros)
    uint32_t phantom;               // zero BITS appended after the real input ran out
    const uint8_t *ip, *in_end;     // next byte to load / end of the real input
    uint32_t nextw;                 // fast mode: the aligned word at ip-4, loaded ahead, not yet appended
    uint32_t rep0, rep1, rep2, rep3, state;
    uint32_t wpos, dict_size;       // window.pos (wrapped, Q3) and window.size
    uint32_t full;                  // window.isFull
    uint8_t *outp, *out_end;        // write cursor; out_end = start + min(size, cap)
    const uint8_t *fast_in_end;     // fast decoder may start a symbol while ip <= fast_in_end ...
    uint8_t *fast_out_end;          // ... and outp <= fast_out_end
    uint32_t end_is_size;           // out_end is the declared unpack size (else the caller's cap)
    uint32_t size_defined;          // state.unpackSizeDefined
    uint32_t lc, lp_mask, pos_mask;
    uint32_t prev_byte, mbyte;      // literal context: byte at -1 and at -(rep0+1)
    uint32_t ctx_a, ctx_b;          // the same two bytes when a window copy has just fetched them:
    uint32_t ctx_pending;           // 1 = values loaded into ctx_a/b, 2 = offsets into `stage` (cp.async)
    const uint8_t *stage;           // V_STAGE: the warp's shared staging buffer for window-copy sources
    int32_t status, site;
    // V_CHAIN fast decoder (lzgpu_fast2.cuh): compressed input staged in shared memory, read a byte ahead
    uint32_t nb, ips, lims;         // next input byte (already loaded); its shared address; last address a symbol may start at
    uint32_t sP, sL, sIn, sStage;   // shared-window addresses of the fixed tables, the literal tables, the input stage, the copy stage
    const uint8_t *g0;              // global address of the byte staged at sIn
    // streamed D2H (host-buffer entry point): decoded bytes that are final, published in 64 KiB blocks
    uint32_t *prog;                 // host-mapped counter of this unit, or null
    const uint8_t *out0;            // start of the unit's output
    uint32_t pub;                   // blocks published (or pushed) so far
    uint8_t *hout;                  // push mode: where this unit's output goes in the caller's pinned buffer (device view), or null
    uint32_t *push_stat;            // push mode: [0] += kilo-cycles spent pushing whole blocks, [1] += blocks (both x 32: every lane adds)
};

// Input is consumed through a 64-bit lookahead register so that the per-bit
// normalisation is branch-free; LZ_FILL() tops it up to >= 4 bytes and is placed so
// that at most 5 adaptive bit steps run between two fills.  Five steps cannot consume
// more than 4 bytes: a step shrinks the range by at most 31/2048 (probabilities stay
// within [31, 2017]), i.e. log2(range) drops by < 6.05, each normalisation adds 8 and
// log2(range) lives in [24, 32) between steps, so k steps normalise at most
// floor((8 + 6.05 k) / 8) times: 4 for k = 5.  When the real input runs out the
// lookahead is padded with zero bytes that are counted in d.phantom: the reference
// stops (io.EOF from ReadByte, decompress.go:35-38) exactly when the decoder would
// consume the first of them, i.e. when d.phantom > d.inbits, which is tested before
// anything a symbol decides becomes visible.
LZ_HD void rc_put(Dec &d, uint32_t w, uint32_t nbits) {   // append the top nbits of w (inbits <= 32)
    // lookahead = hi:lo, valid bits at the top; new bits go right below them
    const uint32_t k = d.inbits;
    if (k == 32) { d.inb_lo |= w; }
    else if (k == 0) { d.inb_hi |= w; }
    else { d.inb_hi |= w >> k; d.inb_lo |= w << (32 - k); }
    d.inbits = k + nbits;
}
LZ_HD void rc_fill(Dec &d) {
    const uint64_t rem = (uint64_t)(d.in_end - d.ip);
    if (LZ_LIKELY(rem >= 4 && ((uintptr_t)d.ip & 3u) == 0)) {
        const uint32_t w = LZ_BSWAP32(LZ_LD_IN32(d.ip));
        if (((uintptr_t)d.ip & 127u) == 0) LZ_PREFETCH_L2(d.ip + 256);
        d.ip += 4;
        rc_put(d, w, 32);
        return;
    }
    // unaligned start or the last few bytes: byte by byte up to a word boundary
    while (d.ip < d.in_end && (d.inbits < 32 || (((uintptr_t)d.ip & 3u) != 0 && d.inbits < 64))) {
        const uint32_t b = LZ_LD_IN8(d.ip);
        if (d.inbits < 32) d.inb_hi |= b << (24 - d.inbits);
        else d.inb_lo |= b << (56 - d.inbits);
        d.ip++;
        d.inbits += 8;
    }
    if (d.inbits < 32) {   // real input exhausted: pad with phantom zeros
        d.phantom += 32 - d.inbits;
        d.inbits = 32;
    }
}
// decode_run<kV, kFast>.
//  kFast: the straight-line decoder used while >= kFastInMargin input bytes and
//         >= kFastOutMargin output bytes remain (one symbol can never need more): no input
//         exhaustion, no size/capacity checks, and the lookahead is topped up WITHOUT a branch from a
//         word that was loaded one top-up earlier.  Taken branches cost a lone warp an
//         instruction-fetch bubble each (ncu: stall_no_instructions), hence the effort.
//         The careful decoder (kFast = false) handles the head and tail of a unit.
//  kV tuning variants, chosen per launch (LZGPU_VARIANT):
//   V_FAST      allow the fast decoder at all (off: the careful decoder runs everywhere -- a test
//               mode that drives every stream through the tail/error-checking code)
//   V_PREFETCH  bit trees fetch BOTH children of the current node (one aligned 32-bit LDS) before
//               the bit is known: shared-memory latency leaves the serial chain at two more
//               instructions per bit (helps a lone warp, not a contended one)
//   V_STAGE     window-copy sources are staged in shared memory by cp.async instead of being held
//               in registers until the deferred store: no load result is outstanding when the
//               decoder resumes, so nothing in it can be made to wait on the window
//   V_CHAIN     the fast decoder is the one in lzgpu_fast2.cuh (latency-optimised dependency chain: both
//               children of a tree node are loaded before the bit is known, every conditional update is
//               one predicated instruction, compressed input is staged in shared memory and read one
//               byte ahead so that there are no lookahead top-ups).  Literal tables use the layout
//               described there (also by the careful decoder of the same instantiation).
//   V_PB2       the posState-indexed tables hold 4 positions instead of 16 (Lay<2>): chosen by the host for
//               units whose pb is <= 2; not a tuning knob
enum : int { V_FAST = 1, V_PREFETCH = 4, V_STAGE = 16, V_CHAIN = 32, V_PB2 = 64 };
#define LZ_LAY(kV) Lay<((kV) & V_PB2) ? 2 : 4>
#ifndef LZGPU_F2_STAGE
#define LZGPU_F2_STAGE 512
#endif
constexpr uint32_t kF2Stage = LZGPU_F2_STAGE;   // bytes of compressed input staged per refill (V_CHAIN): a multiple of 512
constexpr uint32_t kF2Margin = 41;        // a symbol consumes <= 21 bytes; the byte-ahead read adds 1
constexpr uint32_t kF2MinInput = 128;     // do not (re)enter the V_CHAIN fast decoder with less input left
constexpr uint32_t kFastInMargin = 64;    // >= 48 bit steps of one symbol + one word loaded ahead + slack
constexpr uint32_t kFastOutMargin = 274;  // longest match is 273

#define LZ_FILL32() do { if (LZ_UNLIKELY(d.inbits < 32)) rc_fill(d); } while (0)
#if defined(__CUDA_ARCH__)
// fast-mode top-up, all predicated: append the word in hand below the valid bits, count it, fetch
// the next word, advance.  lo is 0 whenever inbits < 32, so lo = (w:0) >> inbits (low word).
#define LZ_FILL_FAST()                                                              \
    asm("{\n\t.reg .pred q;\n\t.reg .b32 w, t;\n\t"                               \
        "setp.lt.u32 q, %2, 32;\n\t"                                                \
        "prmt.b32 w, %3, 0, 0x0123;\n\t"                                            \
        "shr.u32 t, w, %2;\n\t"            /* 0 when inbits >= 32 */               \
        "or.b32 %0, %0, t;\n\t"                                                     \
        "@q shf.r.wrap.b32 %1, 0, w, %2;\n\t"                                       \
        "@q add.u32 %2, %2, 32;\n\t"                                                \
        "@q ld.global.nc.u32 %3, [%4];\n\t"                                         \
        "@q add.u64 %4, %4, 4;\n\t}"                                                \
        : "+r"(d.inb_hi), "+r"(d.inb_lo), "+r"(d.inbits), "+r"(d.nextw), "+l"(d.ip))
#else
#define LZ_FILL_FAST()                                                              \
    do {                                                                            \
        const bool f_ = d.inbits < 32;      /* then inb_lo == 0 */                 \
        const uint32_t w_ = LZ_BSWAP32(d.nextw);                                    \
        d.inb_hi |= LZ_SHR_CLAMP(w_, d.inbits);   /* adds nothing when inbits >= 32 */ \
        d.inb_lo = f_ ? LZ_FUNNEL_R(0u, w_, d.inbits) : d.inb_lo;                   \
        d.inbits += f_ ? 32u : 0u;                                                  \
        LZ_LD_IN32_IF(d.nextw, d.ip, f_);                                           \
        d.ip += f_ ? 4 : 0;                                                         \
    } while (0)
#endif
#define LZ_FILL() do { if (kFast) LZ_FILL_FAST(); else LZ_FILL32(); } while (0)
#define LZ_SHIFT8()                                                                 \
    do {                                                                            \
        d.range <<= 8;                                                              \
        d.code = (d.code << 8) | (d.inb_hi >> 24);                                  \
        d.inb_hi = (d.inb_hi << 8) | (d.inb_lo >> 24);                              \
        d.inb_lo <<= 8;                                                             \
        d.inbits -= 8;                                                              \
    } while (0)
#define LZ_EXHAUSTED() (!kFast && d.phantom > d.inbits)

// real input bytes consumed so far, given the start of the input
LZ_HD uint64_t rc_consumed(const Dec &d, const uint8_t *start) {
    const uint32_t unread = d.inbits > d.phantom ? (d.inbits - d.phantom) >> 3 : 0;
    return (uint64_t)(d.ip - start) - unread;
}

// May the fast decoder start a symbol here?  (ip is where the careful decoder loads next;
// the fast one keeps one more word loaded ahead, hence the +4.)
template <int kV>
LZ_HD bool fast_possible(const Dec &d) {
    if (kV & V_CHAIN) {
#if defined(__CUDA_ARCH__)
        // all buffered bytes are real (phantom == 0); the next unconsumed byte is at ip - inbits/8
        return d.phantom == 0 && (uint64_t)(d.in_end - d.ip) + (d.inbits >> 3) >= kF2MinInput && d.outp <= d.fast_out_end;
#else
        return false;   // host lane emulation: the careful decoder runs everything (same table layout)
#endif
    }
    return ((uintptr_t)d.ip & 3u) == 0 && d.phantom == 0 && d.ip + 4 <= d.fast_in_end && d.outp <= d.fast_out_end;
}
LZ_HD void set_fast_limits(Dec &d) {
    d.fast_in_end = d.in_end - kFastInMargin;
    d.fast_out_end = d.out_end - kFastOutMargin;
    if ((uint64_t)(d.in_end - d.ip) < kFastInMargin + 8) d.fast_in_end = d.ip - 8;   // never
    if ((uint64_t)(d.out_end - d.outp) < kFastOutMargin) d.fast_out_end = d.outp - 1;  // never
}

// Range-coder preamble: rangeDecoder.Init, range_decoder.go:27-46.
// 0 ok, 1 first byte != 0, -1 fewer than 5 bytes.
LZ_HD int rc_init(Dec &d) {
    d.range = 0xFFFFFFFFu;
    d.code = 0;
    d.inb_hi = d.inb_lo = 0;
    d.inbits = 0;
    d.phantom = 0;
    if ((uint64_t)(d.in_end - d.ip) < 1) return -1;
    if (LZ_LD_IN8(d.ip) != 0) return 1;
    if ((uint64_t)(d.in_end - d.ip) < 5) { d.ip = d.in_end; return -1; }
    uint32_t c = 0;
    for (int i = 1; i < 5; i++) c = (c << 8) | LZ_LD_IN8(d.ip + i);
    d.code = c;
    d.ip += 5;
    return 0;
}

// Normalise AFTER the bit, as the reference does (range_decoder.go:64-75).  Select form:
// no branch, so no convergence barrier and no fetch bubble in lane 0's instruction stream.
#define LZ_NORM()                                                                   \
    do {                                                                            \
        const uint32_t sh_ = d.range < kTop ? 8u : 0u;                              \
        d.range <<= sh_;                                                            \
        d.code = LZ_FUNNEL_L(d.inb_hi, d.code, sh_);                                \
        d.inb_hi = LZ_FUNNEL_L(d.inb_lo, d.inb_hi, sh_);                            \
        d.inb_lo <<= sh_;                                                           \
        d.inbits -= sh_;                                                            \
    } while (0)

// One adaptive bit (DecodeBit, range_decoder.go:57-98) followed by the normalisation, select form.
// Probability update: p + ((2048 - p) >> 5) for a 0, p - (p >> 5) for a 1; the latter equals
// p + ((31 - p) >> 5) with an arithmetic shift, so both are p + ((k - p) >> 5).
#if defined(__CUDA_ARCH__) && !defined(LZ_NO_PTX_BIT)
// Device: the step spelled out in PTX so that the conditional updates are single predicated
// instructions (the C form below compiles to select + operate pairs: ~3 more per bit).
#define LZ_BIT(PP, BIT) LZ_BIT_V(PP, *(PP), BIT)
#define LZ_BIT_V(PP, PVAL, BIT)                                                     \
    do {                                                                            \
        uint16_t *pp_ = (PP);                                                       \
        const uint32_t p_ = (PVAL);                                                 \
        uint32_t pn_, b01_;                                                         \
        asm("{\n\t.reg .pred one, nz;\n\t.reg .b32 bd, t, k;\n\t"                   \
            "shr.u32 t, %0, 11;\n\t"                                                \
            "mul.lo.u32 bd, t, %7;\n\t"                                             \
            "setp.ge.u32 one, %1, bd;\n\t"                                          \
            "sub.u32 t, %0, bd;\n\t"                                                \
            "selp.b32 %0, t, bd, one;\n\t"                                          \
            "@one sub.u32 %1, %1, bd;\n\t"                                          \
            "selp.b32 k, 31, 2048, one;\n\t"                                        \
            "sub.s32 k, k, %7;\n\t"                                                 \
            "shr.s32 k, k, 5;\n\t"                                                  \
            "add.s32 %6, %7, k;\n\t"                                                \
            "selp.u32 %5, 1, 0, one;\n\t"                                           \
            "setp.lt.u32 nz, %0, 0x1000000;\n\t"                                    \
            "@nz shl.b32 %0, %0, 8;\n\t"                                            \
            "@nz shf.l.wrap.b32 %1, %2, %1, 8;\n\t"                                 \
            "@nz shf.l.wrap.b32 %2, %3, %2, 8;\n\t"                                 \
            "@nz shl.b32 %3, %3, 8;\n\t"                                            \
            "@nz add.u32 %4, %4, -8;\n\t}"                                          \
            : "+r"(d.range), "+r"(d.code), "+r"(d.inb_hi), "+r"(d.inb_lo), "+r"(d.inbits), \
              "=r"(b01_), "=r"(pn_)                                                 \
            : "r"(p_));                                                             \
        *pp_ = (uint16_t)pn_;                                                       \
        (BIT) = b01_;                                                               \
    } while (0)
// matched-literal flavour: additionally OFFS ^= (bit ? 0 : OLD) as one predicated xor
#define LZ_BIT_MLIT(PP, BIT, OFFS, OLD)                                             \
    do {                                                                            \
        uint16_t *pp_ = (PP);                                                       \
        const uint32_t p_ = *pp_;                                                   \
        uint32_t pn_, b01_;                                                         \
        asm("{\n\t.reg .pred one, nz;\n\t.reg .b32 bd, t, k;\n\t"                   \
            "shr.u32 t, %0, 11;\n\t"                                                \
            "mul.lo.u32 bd, t, %8;\n\t"                                             \
            "setp.ge.u32 one, %1, bd;\n\t"                                          \
            "sub.u32 t, %0, bd;\n\t"                                                \
            "selp.b32 %0, t, bd, one;\n\t"                                          \
            "@one sub.u32 %1, %1, bd;\n\t"                                          \
            "@!one xor.b32 %7, %7, %9;\n\t"                                         \
            "selp.b32 k, 31, 2048, one;\n\t"                                        \
            "sub.s32 k, k, %8;\n\t"                                                 \
            "shr.s32 k, k, 5;\n\t"                                                  \
            "add.s32 %6, %8, k;\n\t"                                                \
            "selp.u32 %5, 1, 0, one;\n\t"                                           \
            "setp.lt.u32 nz, %0, 0x1000000;\n\t"                                    \
            "@nz shl.b32 %0, %0, 8;\n\t"                                            \
            "@nz shf.l.wrap.b32 %1, %2, %1, 8;\n\t"                                 \
            "@nz shf.l.wrap.b32 %2, %3, %2, 8;\n\t"                                 \
            "@nz shl.b32 %3, %3, 8;\n\t"                                            \
            "@nz add.u32 %4, %4, -8;\n\t}"                                          \
            : "+r"(d.range), "+r"(d.code), "+r"(d.inb_hi), "+r"(d.inb_lo), "+r"(d.inbits), \
              "=r"(b01_), "=r"(pn_), "+r"(OFFS)                                     \
            : "r"(p_), "r"(OLD));                                                   \
        *pp_ = (uint16_t)pn_;                                                       \
        (BIT) = b01_;                                                               \
    } while (0)
#else
#define LZ_BIT(PP, BIT) LZ_BIT_V(PP, *(PP), BIT)
#define LZ_BIT_V(PP, PVAL, BIT)                                                     \
    do {                                                                            \
        uint16_t *pp_ = (PP);                                                       \
        const uint32_t p_ = (PVAL);                                                 \
        const uint32_t bound_ = (d.range >> 11) * p_;                               \
        const bool one_ = d.code >= bound_;                                         \
        d.range = one_ ? d.range - bound_ : bound_;                                 \
        d.code = one_ ? d.code - bound_ : d.code;                                   \
        *pp_ = (uint16_t)(p_ + (uint32_t)((int32_t)((one_ ? 31u : 2048u) - p_) >> 5)); \
        (BIT) = one_ ? 1u : 0u;                                                     \
        LZ_NORM();                                                                  \
    } while (0)
#define LZ_BIT_MLIT(PP, BIT, OFFS, OLD)                                             \
    do {                                                                            \
        LZ_BIT(PP, BIT);                                                            \
        (OFFS) ^= (BIT) ? 0u : (OLD);                                               \
    } while (0)
#endif

// Children of node m are entries 2m and 2m+1: one aligned 32-bit load (all P_* bases and
// sub-table strides are even, the shared array is 16-byte aligned).
#define LZ_PAIR(TP, M) (*reinterpret_cast<const uint32_t *>((TP) + 2u * (M)))
#define LZ_PICK(PAIR, BIT) (((PAIR) >> ((BIT) << 4)) & 0xFFFFu)

// MSB-first bit tree (BitTreeDecode, bit_tree_decoder.go:26-76), NBITS constant, unrolled.
// FILL_MASK: bit i set = top the lookahead up before tree bit i.
#define LZ_TREE(PROBS, NBITS, OUT, FILL_MASK)                                       \
    do {                                                                            \
        uint16_t *tp_ = (PROBS);                                                    \
        uint32_t m_ = 1, b_;                                                        \
        if (kV & V_PREFETCH) {                                                      \
            uint32_t pv_ = tp_[1];                                                  \
            _Pragma("unroll") for (int i_ = 0; i_ < (NBITS); i_++) {                \
                if (((FILL_MASK) >> i_) & 1) LZ_FILL();                             \
                uint32_t pair_ = 0;                                                 \
                if (i_ + 1 < (NBITS)) pair_ = LZ_PAIR(tp_, m_);                     \
                LZ_BIT_V(tp_ + m_, pv_, b_);                                        \
                m_ = (m_ << 1) | b_;                                                \
                pv_ = LZ_PICK(pair_, b_);                                           \
            }                                                                       \
        } else {                                                                    \
            _Pragma("unroll") for (int i_ = 0; i_ < (NBITS); i_++) {                \
                if (((FILL_MASK) >> i_) & 1) LZ_FILL();                             \
                LZ_BIT(tp_ + m_, b_);                                               \
                m_ = (m_ << 1) | b_;                                                \
            }                                                                       \
        }                                                                           \
        (OUT) = m_ - (1u << (NBITS));                                               \
    } while (0)

// LSB-first bit tree (BitTreeReverseDecode, bit_tree_decoder.go:82-135), runtime NBITS <= MAXBITS
#define LZ_TREE_REV(PROBS, NBITS, OUT, MAXBITS)                                     \
    do {                                                                            \
        uint16_t *tp_ = (PROBS);                                                    \
        uint32_t m_ = 1, b_, s_ = 0;                                                \
        const uint32_t nb_ = (NBITS);                                               \
        if (kV & V_PREFETCH) {                                                      \
            uint32_t pv_ = tp_[1];                                                  \
            _Pragma("unroll") for (uint32_t i_ = 0; i_ < (MAXBITS); i_++) {         \
                if (i_ < nb_) {                                                     \
                    if (i_ == 4) LZ_FILL();                                         \
                    uint32_t pair_ = 0;                                             \
                    if (i_ + 1 < nb_) pair_ = LZ_PAIR(tp_, m_);                     \
                    LZ_BIT_V(tp_ + m_, pv_, b_);                                    \
                    m_ = (m_ << 1) | b_;                                            \
                    s_ |= b_ << i_;                                                 \
                    pv_ = LZ_PICK(pair_, b_);                                       \
                }                                                                   \
            }                                                                       \
        } else {                                                                    \
            _Pragma("unroll") for (uint32_t i_ = 0; i_ < (MAXBITS); i_++) {         \
                if (i_ < nb_) {                                                     \
                    if (i_ == 4) LZ_FILL();                                         \
                    LZ_BIT(tp_ + m_, b_);                                           \
                    m_ = (m_ << 1) | b_;                                            \
                    s_ |= b_ << i_;                                                 \
                }                                                                   \
            }                                                                       \
        }                                                                           \
        (OUT) = s_;                                                                 \
    } while (0)

// lenDecoder.Decode (len_decoder.go:34-60); WHICH = 0 low, 1 mid, 2 high.
// Entered right after a fill: choice, choice2 and a 3-bit tree are 5 steps; the 8-bit high tree
// tops up again before its bits 3 and 7 (0x88).
#define LZ_LEN(LP, POS_STATE, LEN, WHICH)                                           \
    do {                                                                            \
        uint16_t *lp_ = (LP);                                                       \
        uint32_t lb_, lv_;                                                          \
        LZ_BIT(lp_ + Y::LEN_CHOICE, lb_);                                              \
        if (lb_ == 0) {                                                             \
            LZ_TREE(lp_ + Y::LEN_LOW + ((POS_STATE) << 3), 3, lv_, 0);                 \
            (LEN) = lv_; (WHICH) = 0;                                               \
        } else {                                                                    \
            LZ_BIT(lp_ + Y::LEN_CHOICE2, lb_);                                         \
            if (LZ_LIKELY(lb_ == 0)) {                                              \
                LZ_TREE(lp_ + Y::LEN_MID + ((POS_STATE) << 3), 3, lv_, 0);             \
                (LEN) = 8 + lv_; (WHICH) = 1;                                       \
            } else {                                                                \
                LZ_TREE(lp_ + Y::LEN_HIGH, 8, lv_, 0x88);                              \
                (LEN) = 16 + lv_; (WHICH) = 2;                                      \
            }                                                                       \
        }                                                                           \
    } while (0)

// An error return.  The reference would have returned io.EOF from the failing ReadByte
// before reaching any later check, so input exhaustion is tested first.
#define LZ_FAIL(ST, SITE)                                                 \
    do {                                                                  \
        if (LZ_EXHAUSTED()) goto input_eof;                               \
        d.status = (ST); d.site = (SITE); return OP_DONE;                 \
    } while (0)

// Runs lane 0's serial decoder until a symbol needs the warp (a match / rep /
// short rep: OP_COPY or OP_COPY_Q4 with len and dist set, window position already
// advanced) or the unit part ends (OP_DONE with d.status / d.site set).
// P: fixed tables (shared memory), L: literal tables (shared or global).
template <int kV, bool kFast>
LZ_HD uint32_t decode_run(Dec &d, uint16_t *P, uint16_t *L, uint32_t &out_len, uint32_t &out_dist) {
    using Y = LZ_LAY(kV);
    for (;;) {
        // fast decoder: leave when the margins are gone (at_end is then impossible);
        // careful decoder: hand over as soon as the fast one may run
        if (kFast) {
            if (LZ_UNLIKELY(d.ip > d.fast_in_end || d.outp > d.fast_out_end)) return OP_SWITCH;
        } else if ((kV & V_FAST) && fast_possible<kV>(d)) {
            return OP_SWITCH;
        }
        const bool at_end = !kFast && (d.outp == d.out_end);
        // decompress.go:14-20
        if (at_end && d.end_is_size && d.code == 0) LZ_FAIL(LZGPU_OK, 0);

        LZ_FILL();
        const uint32_t pos_state = d.wpos & d.pos_mask;          // :22
        const uint32_t state2 = (d.state << Y::PB) + pos_state;      // :23
        uint32_t bit;
        LZ_BIT(P + Y::IS_MATCH + state2, bit);                    // :25-42

        if (bit == 0) {  // literal, :44-175
            if (LZ_UNLIKELY(at_end)) {
                if (d.end_is_size) LZ_FAIL(LZGPU_RESULT_ERROR, 46);
                LZ_FAIL(LZGPU_OUTPUT_OVERFLOW, 0);
            }
            // context bytes: straight from the last window copy's loads if there was one (first use
            // of those registers, so this is where a literal-after-match waits for the window)
            uint32_t prevb = d.prev_byte, matchb = d.mbyte;
            if ((kV & V_STAGE) && d.ctx_pending == 2) {
                // staged by cp.async: every lane waits for its own chunks, then the warp syncs so
                // that each lane may read what the others fetched (all lanes are here together)
                LZ_CP_WAIT();
                LZ_WARP_SYNC();
                prevb = d.stage[d.ctx_a];
                matchb = d.stage[d.ctx_b];
            } else if (d.ctx_pending) {
                prevb = d.ctx_a;
                matchb = d.ctx_b;
            }
            d.ctx_pending = 0;
            uint16_t *pr = L + 0x300u * (((d.wpos & d.lp_mask) << d.lc) + (prevb >> (8 - d.lc)));  // :56-57
            // Plain and matched literals in one straight-line tree walk: `offs` is 0x100 while
            // the decoded prefix still equals the match byte's (matched mode, state >= 7,
            // :59-114) and drops to 0 at the first mismatch, after which the index is the
            // plain one (:127-166).  Index = offs + match_bit + sym = ((1 + matchBit) << 8) + sym.
            uint32_t sym = 1;
            if (kV & V_CHAIN) {
                // V_CHAIN table layout: plain node m at [m]; matched node at [0x100 + 2m + matchBit]
                // (the reference's is [((1 + matchBit) << 8) + m]; a private permutation of the same cells)
                uint32_t matched = d.state >= 7 ? 1u : 0u;
#pragma unroll 1
                for (int i = 0; i < 8; i++) {
                    if (i == 4) LZ_FILL();
                    const uint32_t mbit = (matchb >> (7 - i)) & 1u;
                    LZ_BIT(pr + (matched ? 0x100u + 2u * sym + mbit : sym), bit);
                    sym = (sym << 1) | bit;
                    matched &= (bit == mbit) ? 1u : 0u;
                }
            } else {
            uint32_t offs = d.state >= 7 ? 0x100u : 0u;
            uint32_t mb = matchb;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i == 4) LZ_FILL();
                mb += mb;
                const uint32_t old = offs;
                offs &= mb;                                      // match bit, if still in matched mode
                LZ_BIT_MLIT(pr + offs + old + sym, bit, offs, old);   // offs stays set only while bit == match bit
                sym = (sym << 1) | bit;
            }
            }
            if (LZ_UNLIKELY(LZ_EXHAUSTED())) goto input_eof;
            sym &= 0xFF;
            *d.outp++ = (uint8_t)sym;                             // PutByte, :168
            d.prev_byte = sym;
            d.wpos++;
            if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
            d.state = d.state < 4 ? 0 : (d.state < 10 ? d.state - 3 : d.state - 6);  // stateUpdateLiteral
            continue;
        }

        uint32_t len, which = 0, trunc_site;
        uint16_t *rep4 = P + Y::REP4 + (d.state << 2);
        LZ_BIT(rep4 + 0, bit);                                    // isRep, :195-213
        if (bit == 0) {  // simple match, :215-668
            d.rep3 = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0;    // :216
            LZ_FILL();
            LZ_LEN(P + Y::LEN0, pos_state, len, which);            // :218-429
            d.state = d.state < 7 ? 7 : 10;                       // stateUpdateMatch, :431
            const uint32_t len_state = len > 3 ? 3 : len;         // :434-437
            uint32_t slot;
            LZ_TREE(P + Y::POS_SLOT + (len_state << 6), 6, slot, 0x21);  // :441-486, fills before bits 0, 5
            if (LZ_UNLIKELY(slot < 4)) {
                d.rep0 = slot;                                    // :488-489
            } else {
                const uint32_t nd = (slot >> 1) - 1;
                uint32_t dist = (2 | (slot & 1)) << nd, v;
                if (LZ_UNLIKELY(slot < 14)) {                     // :494-546
                    // own sub-table layout: slot s starts at dist - 4; the reference's is dist - slot (:496)
                    LZ_TREE_REV(P + Y::POS_DEC + dist - 4, nd, v, 5);
                    dist += v;
                } else {                                          // :548-628
                    uint32_t res = 0;
                    {
                        // DecodeDirectBits (:549-576) normalises when the halved range drops below
                        // 2^24: first after g = 8 - clz(range) halvings, then after every 8th.
                        // At most 4 normalisations for 26 bits: one top-up covers them.
                        LZ_FILL();
                        uint32_t n = nd - 4;                       // 1..26
                        uint32_t g = 8 - LZ_CLZ(d.range);          // 1..8 halvings to the next normalisation
#pragma unroll 1
                        for (;;) {
                            const uint32_t k = n < g ? n : g;
                            // k halvings without normalisation: step j compares against range >> j
                            // (floor of floor = floor), so only `code` carries from step to step
                            const uint32_t r0 = d.range;
                            uint32_t acc = 0;
#pragma unroll
                            for (uint32_t j = 1; j <= 8; j++) {
#if defined(__CUDA_ARCH__) && !defined(LZ_NO_PTX_BIT)
                                // 5 instructions per step: the compare carries the "step is live" test
                                asm("{\n\t.reg .pred live, one;\n\t.reg .b32 rj;\n\t"
                                    "setp.le.u32 live, %3, %4;\n\t"
                                    "shr.u32 rj, %2, %3;\n\t"
                                    "setp.ge.and.u32 one, %0, rj, live;\n\t"
                                    "@one sub.u32 %0, %0, rj;\n\t"
                                    "@one or.b32 %1, %1, %5;\n\t}"
                                    : "+r"(d.code), "+r"(acc)
                                    : "r"(r0), "r"(j), "r"(k), "r"(1u << (8 - j)));
#else
                                const uint32_t rj = r0 >> j;
                                const bool one = (j <= k) && d.code >= rj;
                                d.code -= one ? rj : 0u;
                                acc |= one ? (1u << (8 - j)) : 0u;
#endif
                            }
                            d.range = r0 >> k;
                            res = (res << k) | (acc >> (8 - k));
                            n -= k;
                            if (k == g) LZ_SHIFT8();
                            g = 8;
                            if (n == 0) break;
                        }
                    }
                    dist += res << 4;
                    LZ_FILL();
                    LZ_TREE_REV(P + Y::ALIGN, 4, v, 4);            // :580-625
                    dist += v;
                }
                d.rep0 = dist;
            }
            if (LZ_UNLIKELY(d.rep0 == 0xFFFFFFFFu)) {             // EOS marker, :633-645
                if (d.code == 0) {
                    // sizeDefined && bytesLeft > 0 (a cap below the declared size implies bytes left)
                    if (d.size_defined && !(d.end_is_size && at_end)) LZ_FAIL(LZGPU_RESULT_ERROR, 636);
                    LZ_FAIL(LZGPU_OK, 0);
                }
                LZ_FAIL(LZGPU_RESULT_ERROR, 643);
            }
            if (LZ_UNLIKELY(at_end)) {                            // :647-649
                if (d.end_is_size) LZ_FAIL(LZGPU_RESULT_ERROR, 648);
                LZ_FAIL(LZGPU_OUTPUT_OVERFLOW, 0);
            }
            if (LZ_UNLIKELY(d.rep0 >= d.dict_size || !(d.full || d.rep0 <= d.wpos)))  // :651-653 (Q4 as written)
                LZ_FAIL(LZGPU_RESULT_ERROR, 652);
            len += 2;                                             // :656
            trunc_site = 662;
        } else {  // rep match, :685-1118
            if (LZ_UNLIKELY(at_end)) {                            // :686-688
                if (d.end_is_size) LZ_FAIL(LZGPU_RESULT_ERROR, 687);
                LZ_FAIL(LZGPU_OUTPUT_OVERFLOW, 0);
            }
            if (LZ_UNLIKELY(d.wpos == 0 && !d.full)) LZ_FAIL(LZGPU_RESULT_ERROR, 691);  // IsEmpty, :690-692
            bool short_rep = false;
            LZ_BIT(rep4 + 1, bit);                                // isRepG0, :694-772
            if (bit == 0) {
                LZ_BIT(P + Y::IS_REP0_LONG + state2, bit);         // :715-755
                short_rep = (bit == 0);
            } else {
                uint32_t dist;
                LZ_BIT(rep4 + 2, bit);                            // isRepG1, :777-813
                if (bit == 0) {
                    dist = d.rep1; d.rep1 = d.rep0; d.rep0 = dist;
                } else {
                    LZ_BIT(rep4 + 3, bit);                        // isRepG2, :816-861 (5th step since the fill)
                    if (bit == 0) { dist = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0; d.rep0 = dist; }
                    else { dist = d.rep3; d.rep3 = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0; d.rep0 = dist; }
                }
            }
            if (short_rep) {                                      // :735-739
                d.state = d.state < 7 ? 9 : 11;                   // stateUpdateShortRep
                len = 1;
                trunc_site = 0;
            } else {
                LZ_FILL();
                LZ_LEN(P + Y::LEN1, pos_state, len, which);        // :870-1101
                d.state = d.state < 7 ? 8 : 11;                   // stateUpdateRep
                len += 2;
                trunc_site = which == 0 ? 941 : (which == 1 ? 1035 : 1111);
            }
        }
        if (LZ_UNLIKELY(LZ_EXHAUSTED())) goto input_eof;

        // copy: decompress.go:656-668 / :934-947 / :1028-1041 / :1104-1117
        const uint64_t avail = kFast ? ~(uint64_t)0 : (uint64_t)(d.out_end - d.outp);
        if (kFast) {
        } else if (d.end_is_size) {
            if (LZ_UNLIKELY((uint32_t)avail < len)) LZ_FAIL(LZGPU_RESULT_ERROR, trunc_site);  // uint32(bytesLeft) < length (Q10)
        } else if (LZ_UNLIKELY(avail < len)) {
            LZ_FAIL(LZGPU_OUTPUT_OVERFLOW, 0);
        }
        const uint32_t dist = d.rep0 + 1;
        uint32_t op = OP_COPY;
        if (LZ_UNLIKELY(!d.full && dist > d.wpos)) {
            // Fewer than `dist` bytes since the dictionary start.  dist == wpos+1 is the
            // reference's off-by-one (Q4): the byte before the start reads as 0 in a fresh
            // window.  Anything further can only come from a rep after an LZMA2 dictionary
            // reset without state reset (Q5): bounds-checked here, a documented deviation.
            if (dist != d.wpos + 1) LZ_FAIL(LZGPU_RESULT_ERROR, LZGPU_SITE_REP_BEFORE_DICT);
            op = OP_COPY_Q4;
        }
        d.wpos += len;
        if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
        out_len = len;
        out_dist = dist;
        return op;
    }
input_eof:
    d.status = LZGPU_OK_INPUT_EXHAUSTED;
    d.site = 0;
    return OP_DONE;
}

// ---- per-lane phases of the warp-cooperative window copy (window.CopyMatch,
// window.go:55-87: byte-serial semantics, overlap replicates with period dist).
// Byte i of a match comes from dst[src_index(i) - dist]; src_index(i) < dist always,
// so every source byte predates the match: loads never depend on this match's stores.
LZ_HD uint32_t src_index(uint32_t i, uint32_t dist) { return i < dist ? i : i % dist; }
// the same when the caller already knows (warp-uniformly) that the match does not overlap itself
LZ_HD uint32_t src_index_far(uint32_t i) { return i; }

}  // namespace lzgpu
