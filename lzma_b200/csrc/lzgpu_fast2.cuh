// lzgpu_fast2.cuh -- the V_CHAIN fast decoder: the same symbol loop as decode_run<kV, true>
// (lzgpu_core.cuh; reference decompress.go:8-1136), rebuilt around the LATENCY of the serial
// dependency chain, because that -- not bandwidth, not occupancy -- bounds a unit.
//
// What the ncu source view of the previous fast decoder showed (profiles/r01_final_*): one adaptive
// bit took ~55-60 cycles for a lone warp although its range arithmetic needs ~24, because the chain
//     bit -> tree index -> address (4 ALU ops) -> LDS (~30 cycles) -> multiply -> compare -> bit
// sits on the critical path.  Here
//   * bit trees keep the address of the CHILDREN PAIR of the current node (y = base + 4m) and load both
//     children before the bit is known; the decided bit then costs one select (value) and one
//     predicated add (next pair address) -- the loads have a whole step to arrive;
//   * matched literals (decompress.go:59-114) walk a path that is known from the match byte alone, so
//     the probability of the next matched node and of the plain node a mismatch would lead to are
//     both loaded ahead, from addresses computed off the critical path; a mismatch branches into the
//     plain ladder at the same depth (one uniform branch per literal);
//   * the probability store needs no address arithmetic (two predicated stores off the pair address);
//   * compressed input is staged in shared memory (512 B per refill) and consumed through a
//     byte-ahead register: a normalisation is five predicated instructions and there are no
//     lookahead top-ups at all (the previous decoder spent ~1.8 instructions per bit on them);
//   * single bits (isMatch, isRep, ...) get their probability loaded before the preceding step.
//
// Literal-table layout of this variant (private; every cell still starts at 1024, state.go:79-121):
// per context 0x300 cells; plain tree node m (1..255) at cell m; matched node for (prefix m, match
// bit b) at cell 0x100 + 2m + b.  The careful decoder of the same instantiation uses the same cells.
//
// Device only: the host lane emulation (tests/emu) runs the careful decoder for this variant.
#pragma once
#include "lzgpu_core.cuh"

namespace lzgpu {

#if defined(__CUDA_ARCH__)

// ---- PTX building blocks.  Operand convention of every block:
//   %0 range  %1 code  %2 nb (next input byte)  %3 ips (shared address of that byte)   -- all "+r"
//   %4 the block's result, %5.. its inputs.
#define F2_REGS                                                                         \
    ".reg .pred one, nz, q0, q1, mbp, ne;\n\t"                                          \
    ".reg .b32 t, t2, tn, u3, bd, k, pn, p, pz, lo, hi, ya, yb, yc, nS;\n\t"

// range >> 11 of the NEXT step is selected from the two shifts of the un-normalised range (>> 3 if the
// normalisation will shift it left by 8, else >> 11) as soon as that range exists, instead of being taken
// from the normalised range: the serial chain  bound -> bit -> range -> normalise? -> shift -> >> 11 -> bound
// loses one link for one more instruction (ptxas folds the select into a predicated shift).  tn carries it
// from step to step inside one asm block; F2_T0 starts a block.  Measured: lone warp -2.2 %, bench shape -0.8 %.
#ifndef F2_TN
#define F2_TN 1
#endif
#if F2_TN
#define F2_T0 "shr.u32 tn, %0, 11;\n\t"
#define F2_TLOAD ""
#define F2_TAHEAD "shr.u32 u3, %0, 3;\n\tshr.u32 tn, %0, 11;\n\t"
#define F2_TSEL "@nz mov.b32 tn, u3;\n\t"
#else
#define F2_T0 ""
#define F2_TLOAD "shr.u32 tn, %0, 11;\n\t"
#define F2_TAHEAD ""
#define F2_TSEL ""
#endif

// code -= bound for a decoded 1, as  code = min(code, code - bound)  in unsigned arithmetic (code < bound: the difference
// wraps above 2^31 > code): ptxas fuses the pair into ONE add-min instruction (VIADDMNMX) that does not wait for the
// bit's predicate, where the predicated subtraction pays the 13-cycle guard latency.
#ifndef F2_CMIN
#define F2_CMIN 1
#endif
#if F2_CMIN
#define F2_CSUB(Q) "sub.u32 t2, %1, bd;\n\tmin.u32 %1, %1, t2;\n\t"
#else
#define F2_CSUB(Q) "@" Q " sub.u32 %1, %1, bd;\n\t"
#endif

// DecodeBit (range_decoder.go:57-98) on probability register P, predicate Q = the bit, followed by the
// normalisation: consume the byte in hand, fetch the one after it.  Ordered along the critical path.
// (The input address is advanced BEFORE the load: an add placed after it would have to wait until the
// load has read its address register -- ~10 cycles in every step, measured.)
#define F2_CORE(P, Q)                                                                   \
    F2_TLOAD                                                                            \
    "mul.lo.u32 bd, tn, " P ";\n\t"                                                     \
    "sub.s32 k, 31, " P ";\n\t"                                                         \
    "setp.ge.u32 " Q ", %1, bd;\n\t"                                                    \
    "sub.u32 t, %0, bd;\n\t"                                                            \
    "selp.b32 %0, t, bd, " Q ";\n\t"                                                    \
    "setp.lt.u32 nz, %0, 0x1000000;\n\t"                                                \
    F2_TAHEAD                                                                           \
    F2_CSUB(Q)                                                                          \
    F2_TSEL                                                                             \
    "@nz shl.b32 %0, %0, 8;\n\t"                                                        \
    "@nz add.u32 %3, %3, 1;\n\t"                                                        \
    "@nz mad.lo.u32 %1, %1, 256, %2;\n\t"                                               \
    "@nz ld.shared.u8 %2, [%3];\n\t"
// pn = P + ((Q ? 31 : 2048) - P) >> 5   (p - (p >> 5) for a 1, p + ((2048 - p) >> 5) for a 0)
#define F2_UPD(P, Q)                                                                    \
    "@!" Q " sub.s32 k, 2048, " P ";\n\t"                                               \
    "shr.s32 k, k, 5;\n\t"                                                              \
    "add.s32 pn, " P ", k;\n\t"
#define F2_NORM ""   /* part of F2_CORE */

#define F2_LD(Y) "ld.shared.u16 lo, [" Y "];\n\tld.shared.u16 hi, [" Y "+2];\n\t"
#define F2_NOLD(Y) ""
// One level of a heap-ordered bit tree.  On entry: P = probability of the current node, lo / hi =
// its children's, loaded from [YC] / [YC+2] (YC = base + 4m).  The node itself lives at YP + (QP ? 2 : 0).
// On exit YN = pair address of the chosen child's children (loaded if LOADS), PN = the chosen child's.
// Instruction ORDER follows the critical path (range -> bound -> bit -> range -> normalised range): ptxas
// breaks scheduling ties by source order.
#define F2_STEP_BODY(YC, YN, QC, NS, LOADS, P, PN, STORE)                               \
    F2_TLOAD                                                                            \
    "mul.lo.u32 bd, tn, " P ";\n\t"                                                     \
    "mad.lo.u32 " YN ", " YC ", 2, " NS ";\n\t"                                         \
    "sub.s32 k, 31, " P ";\n\t"                                                         \
    "setp.ge.u32 " QC ", %1, bd;\n\t"                                                   \
    "sub.u32 t, %0, bd;\n\t"                                                            \
    "selp.b32 %0, t, bd, " QC ";\n\t"                                                   \
    "setp.lt.u32 nz, %0, 0x1000000;\n\t"                                                \
    F2_TAHEAD                                                                           \
    "selp.b32 " PN ", hi, lo, " QC ";\n\t"                                              \
    F2_TSEL                                                                             \
    F2_CSUB(QC)                                                                         \
    "@" QC " add.u32 " YN ", " YN ", 4;\n\t"                                            \
    "@nz shl.b32 %0, %0, 8;\n\t"                                                        \
    "@nz add.u32 %3, %3, 1;\n\t"                                                        \
    "@nz mad.lo.u32 %1, %1, 256, %2;\n\t"                                               \
    LOADS(YN)                                                                           \
    "@nz ld.shared.u8 %2, [%3];\n\t"                                                    \
    "@!" QC " sub.s32 k, 2048, " P ";\n\t"                                              \
    "shr.s32 k, k, 5;\n\t"                                                              \
    "add.s32 pn, " P ", k;\n\t"                                                         \
    STORE
#define F2_STEP(YP, QP, YC, YN, QC, NS, LOADS, P, PN)                                   \
    F2_STEP_BODY(YC, YN, QC, NS, LOADS, P, PN,                                          \
                 "@" QP " st.shared.u16 [" YP "+2], pn;\n\t"                            \
                 "@!" QP " st.shared.u16 [" YP "], pn;\n\t")
// the root level: node at [BASE+2]
#define F2_STEP0(BASE, YC, YN, QC, NS, LOADS, P, PN)                                    \
    F2_STEP_BODY(YC, YN, QC, NS, LOADS, P, PN, "st.shared.u16 [" BASE "+2], pn;\n\t")
// levels by depth; the pair-address registers rotate yb -> yc -> ya, the predicates and the
// probability registers (p, pz) alternate
#define F2_L0(BASE, LOADS) F2_STEP0(BASE, "yb", "yc", "q0", "nS", LOADS, "p", "pz")
#define F2_L1(LOADS) F2_STEP("yb", "q0", "yc", "ya", "q1", "nS", LOADS, "pz", "p")
#define F2_L2(LOADS) F2_STEP("yc", "q1", "ya", "yb", "q0", "nS", LOADS, "p", "pz")
#define F2_L3(LOADS) F2_STEP("ya", "q0", "yb", "yc", "q1", "nS", LOADS, "pz", "p")
#define F2_L4(LOADS) F2_STEP("yb", "q1", "yc", "ya", "q0", "nS", LOADS, "p", "pz")
#define F2_L5(LOADS) F2_STEP("yc", "q0", "ya", "yb", "q1", "nS", LOADS, "pz", "p")
#define F2_L6(LOADS) F2_STEP("ya", "q1", "yb", "yc", "q0", "nS", LOADS, "p", "pz")
#define F2_L7(LOADS) F2_STEP("yb", "q0", "yc", "ya", "q1", "nS", LOADS, "pz", "p")
// pair address after n levels: n=3 -> yb, 4 -> yc, 6 -> yb, 8 -> ya; node index m = (y - base) >> 2
#define F2_ROOT(BASE)                                                                   \
    "neg.s32 nS, " BASE ";\n\t"                                                         \
    "ld.shared.u16 p, [" BASE "+2];\n\t"                                                \
    "ld.shared.u16 lo, [" BASE "+4];\n\t"                                               \
    "ld.shared.u16 hi, [" BASE "+6];\n\t"                                               \
    "add.u32 yb, " BASE ", 4;\n\t"

#define F2_IO(d) "+r"((d).range), "+r"((d).code), "+r"((d).nb), "+r"((d).ips)

__device__ __forceinline__ uint32_t f2_lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t f2_lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// one adaptive bit whose probability PV was loaded from shared address A earlier
#define F2_BIT(d, PV, A, BIT)                                                           \
    asm volatile("{\n\t" F2_REGS                                                        \
                 F2_T0 F2_CORE("%6", "one") F2_UPD("%6", "one")                         \
                 "st.shared.u16 [%5], pn;\n\t"                                          \
                 "selp.u32 %4, 1, 0, one;\n\t"                                          \
                 F2_NORM "}"                                                            \
                 : F2_IO(d), "=&r"(BIT) : "r"(A), "r"(PV) : "memory")

// 6-level tree at byte address BASE (posSlot, decompress.go:441-486): result = node index 64..127
#define F2_TREE6(d, OUT, BASE)                                                          \
    asm volatile("{\n\t" F2_REGS F2_ROOT("%5") F2_T0                                    \
                 F2_L0("%5", F2_LD) F2_L1(F2_LD) F2_L2(F2_LD) F2_L3(F2_LD) F2_L4(F2_LD) F2_L5(F2_NOLD) \
                 "add.u32 t, yb, nS;\n\tshr.u32 %4, t, 2;\n\t}"                          \
                 : F2_IO(d), "=&r"(OUT) : "r"(BASE) : "memory")
// 4-level tree (align, decompress.go:580-625, LSB first): result = node index 16..31, bits MSB-first.
// The root and its children (P0, PLO, PHI) were loaded by the caller before the direct bits.
#define F2_TREE4(d, OUT, BASE, P0, PLO, PHI)                                            \
    asm volatile("{\n\t" F2_REGS                                                        \
                 "neg.s32 nS, %5;\n\tmov.b32 p, %6;\n\tmov.b32 lo, %7;\n\tmov.b32 hi, %8;\n\tadd.u32 yb, %5, 4;\n\t" F2_T0 \
                 F2_L0("%5", F2_LD) F2_L1(F2_LD) F2_L2(F2_LD) F2_L3(F2_NOLD)            \
                 "add.u32 t, yc, nS;\n\tshr.u32 %4, t, 2;\n\t}"                          \
                 : F2_IO(d), "=&r"(OUT) : "r"(BASE), "r"(P0), "r"(PLO), "r"(PHI) : "memory")

// lenDecoder.Decode (len_decoder.go:34-60; live copies decompress.go:218-429, 870-1118).
// SLEN (%5) = byte address of the coder, SLOW (%6) = byte address of its low tree for this posState
// (mid tree = SLOW + MID, high tree = SLEN + HI: immediates that depend on the table layout, Lay<kPB>).  Result 0..271.  The choice cells and the low
// tree's root and children are loaded by F2_LEN_LOADS, which the caller places as early as it can.
#define F2_LEN_LOADS                                                                    \
    "ld.shared.u16 p, [%5];\n\t"                                                        \
    "ld.shared.u16 pc2, [%5+2];\n\t"                                                    \
    "ld.shared.u16 pr0, [%6+2];\n\t"                                                    \
    "ld.shared.u16 lo, [%6+4];\n\t"                                                     \
    "ld.shared.u16 hi, [%6+6];\n\t"
#define F2_LEN_BODY(L, MID, HI)                                                                  \
    F2_CORE("p", "one") F2_UPD("p", "one")                                              \
    "st.shared.u16 [%5], pn;\n\t"                                                       \
    "@one bra.uni " L "CH2;\n\t"                                                        \
    "mov.b32 bs, %6;\n\t"                                                               \
    "mov.b32 p, pr0;\n\t"                                                               \
    "mov.u32 %4, 0xfffffff8;\n\t"                                                       \
    "bra.uni " L "T3;\n\t"                                                              \
    L "CH2:\n\t"                                                                        \
    "add.u32 bs, %6, " MID ";\n\t"                                                          \
    "ld.shared.u16 pr0, [bs+2];\n\t"                                                    \
    "ld.shared.u16 lo, [bs+4];\n\t"                                                     \
    "ld.shared.u16 hi, [bs+6];\n\t"                                                     \
    F2_CORE("pc2", "one") F2_UPD("pc2", "one")                                          \
    "st.shared.u16 [%5+2], pn;\n\t"                                                     \
    "@one bra.uni " L "HI;\n\t"                                                         \
    "mov.b32 p, pr0;\n\t"                                                               \
    "mov.u32 %4, 0;\n\t"                                                                \
    L "T3:\n\t"                                                                         \
    "neg.s32 nS, bs;\n\t"                                                               \
    "add.u32 yb, bs, 4;\n\t"                                                            \
    F2_L0("bs", F2_LD) F2_L1(F2_LD) F2_L2(F2_NOLD)                                      \
    "add.u32 t, yb, nS;\n\t"                                                            \
    "shr.u32 t, t, 2;\n\t"                                                              \
    "add.u32 %4, %4, t;\n\t"                                                            \
    "bra.uni " L "END;\n\t"                                                             \
    L "HI:\n\t"                                                                         \
    "add.u32 bs, %5, " HI ";\n\t"                                                          \
    F2_ROOT("bs")                                                                       \
    F2_L0("bs", F2_LD) F2_L1(F2_LD) F2_L2(F2_LD) F2_L3(F2_LD) F2_L4(F2_LD) F2_L5(F2_LD) F2_L6(F2_LD) F2_L7(F2_NOLD) \
    "add.u32 t, ya, nS;\n\t"                                                            \
    "shr.u32 t, t, 2;\n\t"                                                              \
    "add.u32 %4, t, 0xffffff10;\n\t"                                                    \
    L "END:\n\t"
#define F2_LEN(d, OUT, SLEN, SLOW)                                                      \
    asm volatile("{\n\t" F2_REGS ".reg .b32 bs, pc2, pr0;\n\t"                          \
                 F2_LEN_LOADS F2_T0 F2_LEN_BODY("F2_LEN_", "%7", "%8") "}"                    \
                 : F2_IO(d), "=&r"(OUT) : "r"(SLEN), "r"(SLOW), "n"(2u * (Y::LEN_MID - Y::LEN_LOW)), "n"(2u * Y::LEN_HIGH) : "memory")
// isRep (decompress.go:195-213; cell at AREP, value PREP loaded earlier) and, when it decodes 0, the
// match length: the length coder's cells are in flight while isRep decodes.  OUT = 0xFFFFFFFF for a rep.
#define F2_ISREP_LEN(d, OUT, SLEN, SLOW, AREP, PREP)                                    \
    asm volatile("{\n\t" F2_REGS ".reg .b32 bs, pc2, pr0;\n\t"                          \
                 F2_LEN_LOADS F2_T0                                                     \
                 F2_CORE("%8", "one") F2_UPD("%8", "one")                               \
                 "st.shared.u16 [%7], pn;\n\t"                                          \
                 "mov.u32 %4, 0xffffffff;\n\t"                                          \
                 "@one bra.uni F2_RL_END;\n\t"                                          \
                 F2_LEN_BODY("F2_RL_", "%9", "%10") "}"                                 \
                 : F2_IO(d), "=&r"(OUT) : "r"(SLEN), "r"(SLOW), "r"(AREP), "r"(PREP),                  \
                   "n"(2u * (Y::LEN_MID - Y::LEN_LOW)), "n"(2u * Y::LEN_HIGH) : "memory")

// ---- literal (decompress.go:49-169).  %5 = byte address S of the context's 0x300 cells,
// %6 = 0x100 | match byte, %7 != 0: matched mode (state >= 7).  Result = the byte.
// Matched level i: x = (0x100 | mb) >> (7 - i) = 2 * prefix + match bit; u = S + 2x; its cell is [u + 512];
// the plain node a mismatch leads to is x ^ 1 = cell [u ^ 2] (S is 4-byte aligned).
#define F2_MLEVEL(X, U, PC, XN, UN, PNX, SHN, MIS)                                      \
    "shr.u32 " XN ", %6, " SHN ";\n\t"                                                  \
    "mad.lo.u32 " UN ", " XN ", 2, %5;\n\t"                                             \
    "ld.shared.u16 " PNX ", [" UN "+512];\n\t"                                          \
    "xor.b32 v, " U ", 2;\n\t"                                                          \
    "ld.shared.u16 pmis, [v];\n\t"                                                      \
    "and.b32 t, " X ", 1;\n\t"                                                          \
    "setp.ne.u32 mbp, t, 0;\n\t"                                                        \
    F2_CORE(PC, "one") F2_UPD(PC, "one")                                                \
    "st.shared.u16 [" U "+512], pn;\n\t"                                                \
    "xor.pred ne, one, mbp;\n\t"                                                        \
    F2_NORM                                                                             \
    "@ne bra.uni " MIS ";\n\t"
// mismatch at level i: enter plain level i+1 at node v (YP / QP / YC are that level's names)
#define F2_MIS(LABEL, YP, QP, YC, TARGET, P)                                            \
    LABEL ":\n\t"                                                                       \
    "mov.b32 " P ", pmis;\n\t"                                                            \
    "mov.b32 " YP ", v;\n\t"                                                            \
    "setp.lt.u32 " QP ", %5, 0;\n\t"                                                    \
    "mad.lo.u32 " YC ", v, 2, nS;\n\t"                                                  \
    F2_LD(YC)                                                                           \
    "bra.uni " TARGET ";\n\t"

#define F2_LIT(d, OUT, S, MB, MATCHED)                                                  \
    asm volatile("{\n\t" F2_REGS ".reg .b32 xa, xb, ua, ub, pa, pb, v, pmis;\n\t"       \
                 "neg.s32 nS, %5;\n\t" F2_T0                                            \
                 "setp.ne.u32 one, %7, 0;\n\t"                                          \
                 "@one bra.uni F2_LIT_M;\n\t"                                           \
                 "ld.shared.u16 p, [%5+2];\n\t"                                         \
                 "ld.shared.u16 lo, [%5+4];\n\t"                                        \
                 "ld.shared.u16 hi, [%5+6];\n\t"                                        \
                 "add.u32 yb, %5, 4;\n\t"                                               \
                 /* plain literal: one branch-free block, so that ptxas overlaps the levels */ \
                 F2_L0("%5", F2_LD) F2_L1(F2_LD) F2_L2(F2_LD) F2_L3(F2_LD) F2_L4(F2_LD) F2_L5(F2_LD) F2_L6(F2_LD) F2_L7(F2_NOLD) \
                 "add.u32 t, ya, nS;\n\t"                                               \
                 "shr.u32 t, t, 2;\n\t"                                                 \
                 "and.b32 %4, t, 255;\n\t"                                              \
                 "bra.uni F2_LIT_END;\n\t"                                              \
                 /* the same ladder with an entry per level, for a matched literal after its mismatch */ \
                 "F2_PL1:\n\t" F2_L1(F2_LD)                                             \
                 "F2_PL2:\n\t" F2_L2(F2_LD)                                             \
                 "F2_PL3:\n\t" F2_L3(F2_LD)                                             \
                 "F2_PL4:\n\t" F2_L4(F2_LD)                                             \
                 "F2_PL5:\n\t" F2_L5(F2_LD)                                             \
                 "F2_PL6:\n\t" F2_L6(F2_LD)                                             \
                 "F2_PL7:\n\t" F2_L7(F2_NOLD)                                           \
                 "add.u32 t, ya, nS;\n\t"                                               \
                 "shr.u32 t, t, 2;\n\t"                                                 \
                 "and.b32 %4, t, 255;\n\t"                                              \
                 "bra.uni F2_LIT_END;\n\t"                                              \
                 "F2_LIT_M:\n\t"                                                        \
                 "shr.u32 xa, %6, 7;\n\t"                                               \
                 "mad.lo.u32 ua, xa, 2, %5;\n\t"                                        \
                 "ld.shared.u16 pa, [ua+512];\n\t"                                      \
                 F2_MLEVEL("xa", "ua", "pa", "xb", "ub", "pb", "6", "F2_MIS0")          \
                 F2_MLEVEL("xb", "ub", "pb", "xa", "ua", "pa", "5", "F2_MIS1")          \
                 F2_MLEVEL("xa", "ua", "pa", "xb", "ub", "pb", "4", "F2_MIS2")          \
                 F2_MLEVEL("xb", "ub", "pb", "xa", "ua", "pa", "3", "F2_MIS3")          \
                 F2_MLEVEL("xa", "ua", "pa", "xb", "ub", "pb", "2", "F2_MIS4")          \
                 F2_MLEVEL("xb", "ub", "pb", "xa", "ua", "pa", "1", "F2_MIS5")          \
                 F2_MLEVEL("xa", "ua", "pa", "xb", "ub", "pb", "0", "F2_MIS6")          \
                 /* level 7: x = 0x100 | mb in xb; the byte is its prefix with the decoded bit */ \
                 F2_CORE("pb", "one") F2_UPD("pb", "one")                               \
                 "st.shared.u16 [ub+512], pn;\n\t"                                      \
                 F2_NORM                                                                \
                 "and.b32 t, xb, 254;\n\t"                                              \
                 "selp.u32 k, 1, 0, one;\n\t"                                           \
                 "or.b32 %4, t, k;\n\t"                                                 \
                 "bra.uni F2_LIT_END;\n\t"                                              \
                 F2_MIS("F2_MIS0", "yb", "q0", "yc", "F2_PL1", "pz")                        \
                 F2_MIS("F2_MIS1", "yc", "q1", "ya", "F2_PL2", "p")                        \
                 F2_MIS("F2_MIS2", "ya", "q0", "yb", "F2_PL3", "pz")                        \
                 F2_MIS("F2_MIS3", "yb", "q1", "yc", "F2_PL4", "p")                        \
                 F2_MIS("F2_MIS4", "yc", "q0", "ya", "F2_PL5", "pz")                        \
                 F2_MIS("F2_MIS5", "ya", "q1", "yb", "F2_PL6", "p")                        \
                 F2_MIS("F2_MIS6", "yb", "q0", "yc", "F2_PL7", "pz")                        \
                 "F2_LIT_END:\n\t}"                                                     \
                 : F2_IO(d), "=&r"(OUT) : "r"(S), "r"(MB), "r"(MATCHED) : "memory")

// plain literal whose root cell and its children (P0, PLO, PHI at S+2 / S+4 / S+6) were loaded ahead: after a
// literal the state is < 7, so a literal that follows is a plain one whose context is the byte just decoded --
// its cells are fetched while isMatch is still being decoded
#define F2_LIT_PRE(d, OUT, S, P0, PLO, PHI)                                             \
    asm volatile("{\n\t" F2_REGS                                                        \
                 "neg.s32 nS, %5;\n\tmov.b32 p, %6;\n\tmov.b32 lo, %7;\n\tmov.b32 hi, %8;\n\tadd.u32 yb, %5, 4;\n\t" F2_T0 \
                 F2_L0("%5", F2_LD) F2_L1(F2_LD) F2_L2(F2_LD) F2_L3(F2_LD) F2_L4(F2_LD) F2_L5(F2_LD) F2_L6(F2_LD) F2_L7(F2_NOLD) \
                 "add.u32 t, ya, nS;\n\tshr.u32 t, t, 2;\n\tand.b32 %4, t, 255;\n\t}"  \
                 : F2_IO(d), "=&r"(OUT) : "r"(S), "r"(P0), "r"(PLO), "r"(PHI) : "memory")

// consume one input byte outside an adaptive step (direct bits, decompress.go:549-576)
#define F2_SHIFT8(d)                                                                    \
    asm volatile("add.u32 %3, %3, 1;\n\t"                                               \
                 "shl.b32 %0, %0, 8;\n\t"                                               \
                 "mad.lo.u32 %1, %1, 256, %2;\n\t"                                      \
                 "ld.shared.u8 %2, [%3];"                                               \
                 : F2_IO(d) : : "memory")

// eight equiprobable bits between two normalisations: thresholds R >> 1 .. R >> 8, result MSB first.
// A step is  code = min(code, code - (R >> j))  in unsigned arithmetic: when code < R >> j the difference wraps to more
// than 2^31 > code, so the minimum keeps code (R >> j < 2^31).  The serial chain per bit is SUB -> MIN (10 cycles); the
// predicated form (setp -> @p sub) pays the 13-cycle guard latency in every step (18 cycles, profiles/r02_ubench.txt).
// The bits are read off afterwards (a step that changed code decided 1), outside the chain.
// prefetch of the match source before the align bits are decoded
#ifndef F2_PF_SRC
#define F2_PF_SRC 1
#endif
#ifndef F2_DMIN
#define F2_DMIN 1
#endif
#if F2_DMIN
#define F2_DSTEP(J, M, CI, CO)                                                          \
    "shr.u32 rj, %2, " J ";\n\t"                                                        \
    "sub.u32 tj, " CI ", rj;\n\t"                                                       \
    "min.u32 " CO ", " CI ", tj;\n\t"                                                   \
    "setp.ne.u32 one, " CO ", " CI ";\n\t"                                              \
    "@one or.b32 %1, %1, " M ";\n\t"
#define F2_DIRECT8(CODE, ACC, R)                                                        \
    asm("{\n\t.reg .pred one;\n\t.reg .b32 rj, tj, c<8>;\n\tmov.u32 %1, 0;\n\t"         \
        F2_DSTEP("1", "128", "%0", "c1") F2_DSTEP("2", "64", "c1", "c2") F2_DSTEP("3", "32", "c2", "c3") \
        F2_DSTEP("4", "16", "c3", "c4") F2_DSTEP("5", "8", "c4", "c5") F2_DSTEP("6", "4", "c5", "c6") \
        F2_DSTEP("7", "2", "c6", "c7") F2_DSTEP("8", "1", "c7", "%0") "}"                \
        : "+r"(CODE), "=&r"(ACC) : "r"(R))
// the first K (< 8) of those steps only: a dead step subtracts 0
#define F2_DSTEP_IF(J, M, CI, CO)                                                       \
    "setp.le.u32 live, " J ", %3;\n\t"                                                  \
    "shr.u32 rj, %2, " J ";\n\t"                                                        \
    "selp.b32 rj, rj, 0, live;\n\t"                                                     \
    "sub.u32 tj, " CI ", rj;\n\t"                                                       \
    "min.u32 " CO ", " CI ", tj;\n\t"                                                   \
    "setp.ne.u32 one, " CO ", " CI ";\n\t"                                              \
    "@one or.b32 %1, %1, " M ";\n\t"
#define F2_DIRECT_PART(CODE, ACC, R, K)                                                 \
    asm("{\n\t.reg .pred one, live;\n\t.reg .b32 rj, tj, c<8>;\n\tmov.u32 %1, 0;\n\t"   \
        F2_DSTEP("1", "128", "%0", "c1") F2_DSTEP_IF("2", "64", "c1", "c2") F2_DSTEP_IF("3", "32", "c2", "c3") \
        F2_DSTEP_IF("4", "16", "c3", "c4") F2_DSTEP_IF("5", "8", "c4", "c5") F2_DSTEP_IF("6", "4", "c5", "c6") \
        F2_DSTEP_IF("7", "2", "c6", "%0") "}"                                           \
        : "+r"(CODE), "=&r"(ACC) : "r"(R), "r"(K))
#else
#define F2_DSTEP(J, M)                                                                  \
    "shr.u32 rj, %2, " J ";\n\t"                                                        \
    "setp.ge.u32 one, %0, rj;\n\t"                                                      \
    "@one sub.u32 %0, %0, rj;\n\t"                                                      \
    "@one or.b32 %1, %1, " M ";\n\t"
#define F2_DIRECT8(CODE, ACC, R)                                                        \
    asm("{\n\t.reg .pred one;\n\t.reg .b32 rj;\n\tmov.u32 %1, 0;\n\t"                   \
        F2_DSTEP("1", "128") F2_DSTEP("2", "64") F2_DSTEP("3", "32") F2_DSTEP("4", "16") \
        F2_DSTEP("5", "8") F2_DSTEP("6", "4") F2_DSTEP("7", "2") F2_DSTEP("8", "1") "}"  \
        : "+r"(CODE), "=&r"(ACC) : "r"(R))
// the first K (< 8) of those steps only
#define F2_DSTEP_IF(J, M)                                                               \
    "setp.le.u32 live, " J ", %3;\n\t"                                                  \
    "shr.u32 rj, %2, " J ";\n\t"                                                        \
    "setp.ge.and.u32 one, %0, rj, live;\n\t"                                            \
    "@one sub.u32 %0, %0, rj;\n\t"                                                      \
    "@one or.b32 %1, %1, " M ";\n\t"
#define F2_DIRECT_PART(CODE, ACC, R, K)                                                 \
    asm("{\n\t.reg .pred one, live;\n\t.reg .b32 rj;\n\tmov.u32 %1, 0;\n\t"             \
        F2_DSTEP("1", "128") F2_DSTEP_IF("2", "64") F2_DSTEP_IF("3", "32") F2_DSTEP_IF("4", "16") \
        F2_DSTEP_IF("5", "8") F2_DSTEP_IF("6", "4") F2_DSTEP_IF("7", "2") "}"             \
        : "+r"(CODE), "=&r"(ACC) : "r"(R), "r"(K))
#endif

#define F2_FAIL(ST, SITE) do { d.status = (ST); d.site = (SITE); return OP_DONE; } while (0)

// predicated byte store / load of the window copy: no branch, no convergence barrier
#define F2_ST8_IF(PTR, VAL, LANE, N)                                                    \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q st.global.u8 [%0], %1;\n\t}" \
                 : : "l"(PTR), "r"(VAL), "r"(LANE), "r"(N) : "memory")
#define F2_LD8_IF(VAL, PTR, LANE, N)                                                    \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q ld.global.u8 %0, [%1];\n\t}" \
                 : "+r"(VAL) : "l"(PTR), "r"(LANE), "r"(N) : "memory")

// Same contract as decode_run<kV, true>, except that the common window copy (<= 32 bytes, not
// overlapping itself) is done right here: returns when a symbol needs the general copy code (OP_COPY /
// OP_COPY_Q4 with len and dist set), the unit part ends (OP_DONE) or the staged input / the output
// margin is used up (OP_SWITCH: the caller refills the stage or hands over to the careful decoder).
// (__forceinline__: compiled as a function of its own -- one copy for the LZMA1 and the LZMA2 path -- the decoder state
// lives in local memory across the call and the bench shape takes 253 ms instead of 110)
template <int kV>
__device__ __forceinline__ uint32_t decode_fast2(Dec &d, WarpCopy &wc, uint32_t &out_len, uint32_t &out_dist) {
    // Every lane holds the same decoder state, but the compiler cannot know: whatever derives from a
    // memory load counts as divergent, and each `if` below would get a convergence barrier (BSSY /
    // BSYNC, ~15 cycles a branch, measured).  Passing the state through a shuffle from lane 0 once
    // per entry makes it provably uniform; inside the loop it is only touched by asm blocks whose
    // inputs are uniform, so every branch of the loop compiles to a plain uniform branch.
#define F2_U32(x) (x) = __shfl_sync(0xffffffffu, (x), 0)
#define F2_U64P(T, x)                                                                   \
    do {                                                                                \
        const uint64_t v_ = (uint64_t)(uintptr_t)(x);                                   \
        const uint32_t lo_ = __shfl_sync(0xffffffffu, (uint32_t)v_, 0), hi_ = __shfl_sync(0xffffffffu, (uint32_t)(v_ >> 32), 0); \
        (x) = (T)(uintptr_t)(((uint64_t)hi_ << 32) | lo_);                              \
    } while (0)
    F2_U32(d.range); F2_U32(d.code); F2_U32(d.nb); F2_U32(d.ips); F2_U32(d.lims);
    F2_U32(d.sP); F2_U32(d.sL); F2_U32(d.sStage);
    F2_U32(d.rep0); F2_U32(d.rep1); F2_U32(d.rep2); F2_U32(d.rep3); F2_U32(d.state);
    F2_U32(d.wpos); F2_U32(d.dict_size); F2_U32(d.full);
    F2_U32(d.size_defined); F2_U32(d.lc); F2_U32(d.lp_mask); F2_U32(d.pos_mask);
    F2_U32(d.prev_byte); F2_U32(d.mbyte); F2_U32(d.ctx_a); F2_U32(d.ctx_b); F2_U32(d.ctx_pending);
    F2_U64P(uint8_t *, d.outp); F2_U64P(uint8_t *, d.fast_out_end);
    F2_U32(wc.pend_len); F2_U32(wc.pend_staged); F2_U32(wc.pend_off); F2_U32(wc.stage_sel);
    F2_U64P(uint8_t *, wc.pend_dst);
#undef F2_U32
#undef F2_U64P
    using Y = LZ_LAY(kV);
    const uint32_t sP = d.sP;
    const uint32_t lane = LZ_LANE();
    // the next symbol's isMatch / isRep probabilities are loaded as soon as the current symbol has fixed
    // the state and the position they depend on (nothing in between touches those cells)
    uint32_t pos_state = d.wpos & d.pos_mask;                    // decompress.go:22
    uint32_t a_im = sP + 2u * Y::IS_MATCH + 2u * ((d.state << Y::PB) + pos_state);   // :23
    uint32_t a_rep = sP + 2u * Y::REP4 + 8u * d.state;
    uint32_t p_im = f2_lds16(a_im);
    uint32_t p_rep = f2_lds16(a_rep);
#define F2_NEXT_CTX()                                                                   \
    do {                                                                                \
        pos_state = d.wpos & d.pos_mask;                                                \
        a_im = sP + 2u * Y::IS_MATCH + 2u * ((d.state << Y::PB) + pos_state);                \
        a_rep = sP + 2u * Y::REP4 + 8u * d.state;                                        \
        p_im = f2_lds16(a_im);                                                          \
        p_rep = f2_lds16(a_rep);                                                        \
    } while (0)
    uint32_t pl_valid = 0, pl_S = 0, pl_p = 0, pl_lo = 0, pl_hi = 0;   // root cells of a literal after a literal
    for (;;) {
        if (LZ_UNLIKELY(d.ips > d.lims || d.outp > d.fast_out_end)) return OP_SWITCH;

        uint32_t bit;
        F2_BIT(d, p_im, a_im, bit);                              // :25-42

        if (bit == 0) {  // literal, :44-175
            uint32_t sym;
            if (pl_valid) {
                // the previous symbol was a literal too: state < 7 (plain literal), nothing pending in the window,
                // and this literal's root cells were loaded before isMatch was decoded
                d.state = (d.state > 3u ? d.state : 3u) - 3u;    // stateUpdateLiteral for state < 7
                d.wpos++;
                if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
                F2_NEXT_CTX();
                F2_LIT_PRE(d, sym, pl_S, pl_p, pl_lo, pl_hi);
            } else {
            uint32_t prevb = d.prev_byte, matchb = d.mbyte;
            if (d.ctx_pending) {                                 // a window copy came right before: its source
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // words are (being) staged in shared memory
                __syncwarp();                                    //  other lanes' copies become visible
                prevb = f2_lds8(d.ctx_a);
                matchb = f2_lds8(d.ctx_b);
            }
            d.ctx_pending = 0;
            const uint32_t S = d.sL + 0x600u * (((d.wpos & d.lp_mask) << d.lc) + (prevb >> (8 - d.lc)));  // :56-57
            const uint32_t matched = d.state >= 7 ? 1u : 0u;
            {                                                    // stateUpdateLiteral (state.go:153-163), branch-free
                uint32_t ns = (d.state > 3u ? d.state : 3u) - 3u;
                if (d.state >= 10u) ns -= 3u;
                d.state = ns;
            }
            d.wpos++;
            if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
            F2_NEXT_CTX();
            F2_LIT(d, sym, S, 0x100u | matchb, matched);
            }
            *d.outp++ = (uint8_t)sym;                             // PutByte, :168
            d.prev_byte = sym;
            // root cells of the literal that may follow (context: this byte, the position after it)
            pl_S = d.sL + 0x600u * (((d.wpos & d.lp_mask) << d.lc) + (sym >> (8 - d.lc)));
            pl_p = f2_lds16(pl_S + 2);
            pl_lo = f2_lds16(pl_S + 4);
            pl_hi = f2_lds16(pl_S + 6);
            pl_valid = 1;
            continue;
        }
        pl_valid = 0;

        uint32_t len;
        const uint32_t state2 = (a_im - sP - 2u * Y::IS_MATCH) >> 1;   // of THIS symbol (isRep0Long, :716)
        const uint32_t a_rep_cur = a_rep, pos_state_cur = pos_state;
        // isRep (:195-213) and, for a simple match, its length (:218-429) in one block
        F2_ISREP_LEN(d, len, sP + 2u * Y::LEN0, sP + 2u * (Y::LEN0 + Y::LEN_LOW) + 16u * pos_state_cur, a_rep, p_rep);
        if (LZ_LIKELY(len != 0xFFFFFFFFu)) {  // simple match, :215-668
            d.rep3 = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0;    // :216
            d.state = d.state < 7 ? 7 : 10;                       // stateUpdateMatch, :431
            const uint32_t len_state = len > 3 ? 3 : len;         // :434-437
            len += 2;                                             // :656
            // window position after this match: fixes the next symbol's contexts (the checks below read
            // the position before it, kept in wpos0)
            const uint32_t wpos0 = d.wpos;
            d.wpos += len;
            if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full |= 2; }   // bit 1: became full by THIS match
            F2_NEXT_CTX();
            uint32_t slot;
            F2_TREE6(d, slot, sP + 2u * Y::POS_SLOT + (len_state << 7));   // :441-486
            slot -= 64;
            if (LZ_UNLIKELY(slot < 4)) {
                d.rep0 = slot;                                    // :488-489
            } else {
                const uint32_t nd = (slot >> 1) - 1;
                uint32_t dist = (2 | (slot & 1)) << nd, v;
                if (LZ_UNLIKELY(slot < 14)) {                     // :494-546, LSB first
                    const uint32_t tb = sP + 2u * (Y::POS_DEC + dist - 4);   // own sub-table layout (lzgpu_core.cuh)
                    uint32_t m = 1;
                    v = 0;
#pragma unroll 1
                    for (uint32_t i = 0; i < nd; i++) {
                        const uint32_t a = tb + 2u * m;
                        const uint32_t pv = f2_lds16(a);
                        F2_BIT(d, pv, a, bit);
                        m = (m << 1) | bit;
                        v |= bit << i;
                    }
                    dist += v;
                } else {                                          // :548-628
                    // the align tree's root and its children: in flight during the direct bits
                    const uint32_t al0 = f2_lds16(sP + 2u * Y::ALIGN + 2), al2 = f2_lds16(sP + 2u * Y::ALIGN + 4),
                                   al3 = f2_lds16(sP + 2u * Y::ALIGN + 6);
                    // DecodeDirectBits (:549-576) normalises when the halved range drops below 2^24: first
                    // after g = 8 - clz(range) halvings, then after every 8th.  Between two normalisations
                    // step j compares against range >> j, so only `code` is carried (4 instructions a bit).
                    // The first run is made a full one: range < 2^(24+g), so range << (8-g) fits 32 bits and
                    // its first 8-g thresholds (>= range > code) decide 0 and leave code alone.
                    uint32_t res = 0, acc;
                    uint32_t n = nd - 4;                           // 1..26
                    const uint32_t g = 8 - LZ_CLZ(d.range);        // 1..8
                    if (LZ_LIKELY(n >= g)) {
                        F2_DIRECT8(d.code, acc, d.range << (8 - g));
                        res = acc;
                        d.range >>= g;
                        n -= g;
                        F2_SHIFT8(d);
#pragma unroll 1
                        while (n >= 8) {
                            F2_DIRECT8(d.code, acc, d.range);
                            res = (res << 8) | acc;
                            d.range >>= 8;
                            n -= 8;
                            F2_SHIFT8(d);
                        }
                    }
                    if (n) {                                       // last run: n < 8 halvings, no normalisation after it
                        F2_DIRECT_PART(d.code, acc, d.range, n);
                        res = (res << n) | (acc >> (8 - n));
                        d.range >>= n;
                    }
                    dist += res << 4;
#if F2_PF_SRC
                    // the match source is now known to within the 4 align bits: ask for its line(s) before they are
                    // decoded -- a literal that follows the match waits for the source bytes (prevByte, matchByte),
                    // and the align tree is 150 cycles of head start on a 260 (L2) ... 800 (DRAM) cycle fetch
                    {
                        const uint8_t *pf = d.outp - dist - 16u;
                        asm volatile("prefetch.global.L1 [%0];\n\tprefetch.global.L1 [%0+48];" : : "l"(pf));
                    }
#endif
                    uint32_t m;
                    F2_TREE4(d, m, sP + 2u * Y::ALIGN, al0, al2, al3);   // :580-625
                    dist += __brev(m) >> 28;
                }
                d.rep0 = dist;
            }
            // One test for everything unusual about the distance: ordinary ones are below the number of
            // bytes the window held BEFORE this match.  The EOS marker (0xFFFFFFFF) is never ordinary.
            const uint32_t full0 = d.full & 1u;
            if (LZ_UNLIKELY(d.rep0 >= (full0 ? d.dict_size : wpos0))) {
                if (d.rep0 == 0xFFFFFFFFu) {                      // EOS marker, :633-645 (never at the declared end here)
                    if (d.code == 0) {
                        if (d.size_defined) F2_FAIL(LZGPU_RESULT_ERROR, 636);
                        F2_FAIL(LZGPU_OK, 0);
                    }
                    F2_FAIL(LZGPU_RESULT_ERROR, 643);
                }
                if (d.rep0 >= d.dict_size || !(full0 || d.rep0 <= wpos0))   // :651-653 (Q4 as written)
                    F2_FAIL(LZGPU_RESULT_ERROR, 652);
                // rep0 == wpos0 on a window that is not full: the reference's off-by-one (Q4)
                d.full = (d.full | (d.full >> 1)) & 1u;
                out_len = len;
                out_dist = d.rep0 + 1;
                return OP_COPY_Q4;
            }
            d.full = (d.full | (d.full >> 1)) & 1u;
        } else {  // rep match, :685-1118 (rare here: the generic single-bit steps)
            if (LZ_UNLIKELY(d.wpos == 0 && !d.full)) F2_FAIL(LZGPU_RESULT_ERROR, 691);  // IsEmpty, :690-692
            bool short_rep = false;
            uint32_t a = a_rep_cur + 2, pv = f2_lds16(a);
            F2_BIT(d, pv, a, bit);                                // isRepG0, :694-772
            if (bit == 0) {
                a = sP + 2u * Y::IS_REP0_LONG + 2u * state2;
                pv = f2_lds16(a);
                F2_BIT(d, pv, a, bit);                            // :715-755
                short_rep = (bit == 0);
            } else {
                uint32_t dist;
                a = a_rep_cur + 4;
                pv = f2_lds16(a);
                F2_BIT(d, pv, a, bit);                            // isRepG1, :777-813
                if (bit == 0) {
                    dist = d.rep1; d.rep1 = d.rep0; d.rep0 = dist;
                } else {
                    a = a_rep_cur + 6;
                    pv = f2_lds16(a);
                    F2_BIT(d, pv, a, bit);                        // isRepG2, :816-861
                    if (bit == 0) { dist = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0; d.rep0 = dist; }
                    else { dist = d.rep3; d.rep3 = d.rep2; d.rep2 = d.rep1; d.rep1 = d.rep0; d.rep0 = dist; }
                }
            }
            if (short_rep) {                                      // :735-739
                d.state = d.state < 7 ? 9 : 11;                   // stateUpdateShortRep
                len = 1;
            } else {
                F2_LEN(d, len, sP + 2u * Y::LEN1, sP + 2u * (Y::LEN1 + Y::LEN_LOW) + 16u * pos_state_cur);   // :870-1101
                d.state = d.state < 7 ? 8 : 11;                   // stateUpdateRep
                len += 2;
            }
            const uint32_t dist = d.rep0 + 1;
            if (LZ_UNLIKELY(!d.full && dist > d.wpos)) {
                if (dist != d.wpos + 1) F2_FAIL(LZGPU_RESULT_ERROR, LZGPU_SITE_REP_BEFORE_DICT);   // Q5
                d.wpos += len;
                if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
                out_len = len;
                out_dist = dist;
                return OP_COPY_Q4;                                // Q4
            }
            d.wpos += len;
            if (LZ_UNLIKELY(d.wpos >= d.dict_size)) { d.wpos -= d.dict_size; d.full = 1; }
            F2_NEXT_CTX();
        }

        // copy: decompress.go:656-668 / :934-947 / :1028-1041 / :1104-1117 (sizes cannot be exceeded here)
        const uint32_t dist = d.rep0 + 1;
        if (LZ_UNLIKELY(len > 32 || dist <= len)) {   // long or self-overlapping: the general copy code
            out_len = len;
            out_dist = dist;
            return OP_COPY;
        }
        // window.CopyMatch (window.go:55-87) for the common case.  The source bytes (plus the byte after
        // them, which a matched literal would need) are fetched by cp.async as the aligned 4-byte words
        // that cover them, into one half of the copy stage; nothing waits for them here.  The PREVIOUS
        // match's bytes are moved from the other half to the window first -- they were held back for the
        // same reason.  (Loads into registers instead would keep a scoreboard busy across the loop, and
        // ptxas then makes every following symbol wait for them at its first branch: measured, 86 cycles
        // per symbol.)
        uint8_t *dst = d.outp;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        {   // (a pending copy is always a staged one here: the general copy code does not defer for this decoder)
            uint32_t v = 0;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q ld.shared.u8 %0, [%1];\n\t}"
                         : "+r"(v) : "r"(d.sStage + wc.pend_off + lane), "r"(lane), "r"(wc.pend_len) : "memory");
            F2_ST8_IF(wc.pend_dst + lane, v, lane, wc.pend_len);
        }
        __syncwarp();                                             // those stores precede the fetches below
        const uint8_t *src = dst - dist;
        const uint32_t off = (uint32_t)(uintptr_t)src & 3u;
        const uint32_t nw = (off + len + 4) >> 2;                 // words covering src[0 .. len] (len + 1 bytes): <= 9
        const uint32_t sbuf = d.sStage + wc.stage_sel;
        asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t"
                     "@q cp.async.ca.shared.global [%0], [%1], 4;\n\t"
                     "cp.async.commit_group;\n\t}"
                     : : "r"(sbuf + 4u * lane), "l"(src - off + 4u * lane), "r"(lane), "r"(nw) : "memory");
        d.ctx_a = sbuf + off + len - 1;                           // context of a literal that may follow: last byte of
        d.ctx_b = sbuf + off + len;                               // the match, byte at -(rep0+1) after it
        d.ctx_pending = 2;
        wc.pend_len = len;
        wc.pend_dst = dst;
        wc.pend_staged = 1;
        wc.pend_off = wc.stage_sel + off;
        wc.pend_dist = 0xFFFFFFFFu;                               // not self-overlapping (general commit code)
        wc.stage_sel ^= 64u;
        d.outp = dst + len;
    }
#undef F2_NEXT_CTX
}

#else   // host passes: never called (fast_possible<V_CHAIN> is false off the device)
template <int kV>
LZ_HD uint32_t decode_fast2(Dec &, WarpCopy &, uint32_t &, uint32_t &) { return OP_DONE; }
#endif

}  // namespace lzgpu
