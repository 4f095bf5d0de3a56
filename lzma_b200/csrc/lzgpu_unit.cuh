// lzgpu_unit.cuh -- one warp decodes one unit.
//
// Written once in "warp-uniform" style and compiled two ways:
//   * device (nvcc): the 32 lanes of a warp execute it together: every lane runs the serial
//     range decoder redundantly on identical data (same issue cost under SIMT, no divergence,
//     nothing to broadcast; same-address shared / global accesses are broadcasts), and each lane
//     moves its own byte of a window copy;
//   * host lane emulation (tests/emu/, test infrastructure only): LZ_FOR_LANES
//     loops over 32 virtual lanes and per-lane variables become arrays, so the
//     protocol (deferred stores, what lane 0 may read when) can be checked
//     bit-exactly on a machine with no GPU.
//
// Reference behaviour reproduced here: Reader1.initialize + Read loop
// (reader1.go:149-159, 223-254) for LZMA1 units; Reader2.startChunk / Read /
// uncompressedRead (reader2.go:100-294) for LZMA2 groups.
#pragma once
#include "lzgpu_core.cuh"

#if defined(__CUDA_ARCH__)
#define LZ_LANE() (threadIdx.x & 31u)
#define LZ_FOR_LANES(l) for (uint32_t l = LZ_LANE(), once_ = 1; once_; once_ = 0)
#define LZ_IF_LANE0_ONLY   /* every lane stores the same value to the same address: no branch needed */
#define LZ_SYNC() __syncwarp()
#define LZ_LANEVAR(T, name) T name
#define LZ_LV(name, l) name
#define LZ_DEV __device__ __forceinline__
// Per-lane work is written WITHOUT divergent branches and WITHOUT generic-space accesses at lane-dependent
// addresses: either one, anywhere in the kernel, makes ptxas treat every later load as possibly divergent and
// bracket every branch of the (uniform) decoder with convergence barriers (BSSY / BSYNC / BREAK: 3.5 % of the
// instructions plus their latency, measured).  Loops over bytes have a uniform trip count; the lane's test
// only predicates a load / store that names its state space.
#define LZ_STG8_IF(ptr, val, cond)   /* window (global) */                              \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}" \
                 : : "l"(ptr), "r"((uint32_t)(val)), "r"((uint32_t)(cond)) : "memory")
#define LZ_LDG8_IF(dst, ptr, cond)                                                      \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" \
                 : "+r"(dst) : "l"(ptr), "r"((uint32_t)(cond)) : "memory")
#define LZ_LDIN8_IF(dst, ptr, cond)  /* compressed input (global, read-only) */         \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.nc.u8 %0, [%1];\n\t}" \
                 : "+r"(dst) : "l"(ptr), "r"((uint32_t)(cond)) : "memory")
#define LZ_STAGE8(wc, off, dst)      /* copy stage (shared) */                          \
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(dst) : "r"((wc).s_stage + (off)) : "memory")
#define LZ_STG128_IF(ptr, a, b, c, d, cond)   /* window (global), 16-byte aligned */    \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.global.v4.u32 [%0], {%1, %2, %3, %4};\n\t}" \
                 : : "l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"((uint32_t)(cond)) : "memory")
#else
#define LZ_STG8_IF(ptr, val, cond) do { if (cond) *(uint8_t *)(ptr) = (uint8_t)(val); } while (0)
#define LZ_LDG8_IF(dst, ptr, cond) do { if (cond) (dst) = *(const uint8_t *)(ptr); } while (0)
#define LZ_LDIN8_IF(dst, ptr, cond) do { if (cond) (dst) = *(const uint8_t *)(ptr); } while (0)
#define LZ_STAGE8(wc, off, dst) ((dst) = (wc).stage[off])
#define LZ_FOR_LANES(l) for (uint32_t l = 0; l < 32; l++)
#define LZ_IF_LANE0_ONLY
#define LZ_SYNC() ((void)0)
#define LZ_LANEVAR(T, name) T name[32]
#define LZ_LV(name, l) name[l]
#define LZ_DEV inline
#endif

namespace lzgpu {

// Window-copy bookkeeping (same in every lane) + the per-lane deferred byte.
struct WarpCopy {
    uint32_t pend_len;     // bytes fetched but not yet stored (0..32)
    uint8_t *pend_dst;
    uint32_t pend_staged;  // the bytes wait in `stage` (cp.async) rather than in pend_val
    uint32_t pend_off;     // staged: offset of the source's first byte in `stage`
    uint32_t pend_dist;    // staged: match distance (period of an overlapping copy)
    uint8_t *stage;        // shared staging buffer of the warp: 2 x 64 bytes (V_STAGE uses the first, V_CHAIN alternates)
    uint32_t stage_sel;    // V_CHAIN: which half the NEXT copy stages into (0 / 64)
    uint32_t s_stage;      // shared-window address of `stage` (device)
    uint8_t *out_limit;    // end of the unit's output range (staging over-reads <= 3 bytes)
    LZ_LANEVAR(uint32_t, pend_val);   // the byte, in a 32-bit register
};

}  // namespace lzgpu

#include "lzgpu_fast2.cuh"   // needs WarpCopy and the lane macros

namespace lzgpu {

// Store the previous match's bytes.  They reach memory before anything can read them: this runs
// ahead of every window fetch.
LZ_DEV void wc_commit(WarpCopy &wc) {
    if (wc.pend_len) {
        if (wc.pend_staged) {
            LZ_CP_WAIT();
            LZ_SYNC();   // each lane reads bytes other lanes' cp.async fetched
            LZ_FOR_LANES(l) {
                const uint32_t live = l < wc.pend_len;
                uint32_t v;
                LZ_STAGE8(wc, wc.pend_off + (live ? src_index(l, wc.pend_dist) : 0u), v);
                LZ_STG8_IF(wc.pend_dst + l, v, live);
            }
        } else {
            LZ_FOR_LANES(l) {
                LZ_STG8_IF(wc.pend_dst + l, LZ_LV(wc.pend_val, l), l < wc.pend_len);
            }
        }
        wc.pend_len = 0;
    }
    LZ_SYNC();
}

// Warp-cooperative copy of n bytes of compressed input (global, read-only) into the window (global), any
// alignment on either side: an LZMA2 uncompressed chunk (uncompressedRead, reader2.go:252-294 -> window.ReadFrom,
// window.go:142-155).  Bytes up to the first 16-byte boundary of dst and the last < 20 go one byte per lane; the
// body is one 16-byte store per lane, its source taken as the five aligned 32-bit words that cover it and
// funnel-shifted into place (consecutive chunks of a stream sit 3 header bytes apart, so source and destination
// are rarely congruent).  Four vectors per lane are in flight per trip.  No word is loaded that does not lie
// entirely inside [src rounded down to 4, src + n).
template <bool kWindow>   // kWindow: the source is window bytes this kernel wrote (coherent loads), not read-only input
LZ_DEV void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n) {
#if defined(__CUDA_ARCH__)
    const uint32_t lane = LZ_LANE();
    uint32_t head = (uint32_t)((16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    {
        uint32_t v = 0;
        if (kWindow) LZ_LDG8_IF(v, src + (lane < head ? lane : 0u), lane < head);
        else LZ_LDIN8_IF(v, src + (lane < head ? lane : 0u), lane < head);
        LZ_STG8_IF(dst + lane, v, lane < head);
    }
    dst += head; src += head; n -= head;
    const uint32_t a = (uint32_t)(uintptr_t)src & 3u, sh = 8u * a;
    const uint8_t *s_al = src - a;
    const uint32_t avail = n + a;                                       // bytes from s_al to the end of the source
    const uint32_t nv = a ? (avail >= 4u ? (avail - 4u) >> 4 : 0u) : avail >> 4;   // vectors whose words lie inside
    for (uint32_t j0 = 0; j0 < nv; j0 += 128) {                         // uniform trip count
        uint32_t w[4][5];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t j = j0 + 32u * u + lane, live = j < nv;
            const uint8_t *q = s_al + 16u * (live ? j : 0u);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                w[u][k] = 0;
                if (kWindow) LZ_LD_WIN32_IF(w[u][k], q + 4 * k, live && (k < 4 || a != 0u));
                else LZ_LD_IN32_IF(w[u][k], q + 4 * k, live && (k < 4 || a != 0u));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t j = j0 + 32u * u + lane, live = j < nv;
            LZ_STG128_IF(dst + 16u * (live ? j : 0u), LZ_FUNNEL_R(w[u][0], w[u][1], sh), LZ_FUNNEL_R(w[u][1], w[u][2], sh),
                         LZ_FUNNEL_R(w[u][2], w[u][3], sh), LZ_FUNNEL_R(w[u][3], w[u][4], sh), live);
        }
    }
    const uint32_t done = 16u * nv, rest = n - done;                    // < 20 (a != 0) or < 16
    {
        uint32_t v = 0;
        if (kWindow) LZ_LDG8_IF(v, src + done + (lane < rest ? lane : 0u), lane < rest);
        else LZ_LDIN8_IF(v, src + done + (lane < rest ? lane : 0u), lane < rest);
        LZ_STG8_IF(dst + done + lane, v, lane < rest);
    }
#else
    memcpy(dst, src, n);
#endif
}
LZ_DEV void warp_copy_in(uint8_t *dst, const uint8_t *src, uint32_t n) { warp_copy<false>(dst, src, n); }
// Push mode (the caller's output buffer is pinned and mapped): the unit itself writes its decoded bytes to host memory,
// over PCIe, as they become final -- [from, to) of its output range; no D2H copy follows the kernel.
LZ_DEV void push_out(const Dec &d, uint64_t from, uint64_t to) {
#if defined(__CUDA_ARCH__)
    __threadfence_block();   // the bytes were stored by all lanes (and, after an exchange, by other warps of the CTA)
    __syncwarp();
    // whole 64 KiB blocks between congruent 16-byte-aligned addresses (the usual case: units laid out on 16-byte
    // boundaries in both buffers): 16-byte loads, eight per lane in flight -- a block is 16 round trips to L2, not 32
    // trips of five 4-byte loads per vector
    const uint32_t aligned = __shfl_sync(0xffffffffu, (uint32_t)((((uintptr_t)(d.hout + from) | (uintptr_t)(d.out0 + from) | (uintptr_t)(to - from)) & 15u) == 0), 0);
    if (aligned) {
        // timed: the host switches to copy-engine transfers when the link is so busy (eight GPUs of a node writing
        // to host memory at once) that these stores hold the decoder up
        const long long t0 = clock64();
        const uint32_t lane = LZ_LANE();
        const uint8_t *s = d.out0 + from;
        uint8_t *t = d.hout + from;
        const uint64_t nv = (to - from) >> 4;
        for (uint64_t j0 = 0; j0 < nv; j0 += 256) {                     // uniform trip count
            uint32_t v[8][4];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint64_t j = j0 + 32u * u + lane;
                const uint32_t live = j < nv;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0;
                asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                             : "+r"(v[u][0]), "+r"(v[u][1]), "+r"(v[u][2]), "+r"(v[u][3]) : "l"(s + 16u * (live ? j : 0u)), "r"(live) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint64_t j = j0 + 32u * u + lane;
                const uint32_t live = j < nv;
                LZ_STG128_IF(t + 16u * (live ? j : 0u), v[u][0], v[u][1], v[u][2], v[u][3], live);
            }
        }
        if (d.push_stat) {
            const uint32_t kc = (uint32_t)((unsigned long long)(clock64() - t0) >> 10), nb = (uint32_t)((to - from) >> 16);
            asm volatile("red.global.add.u32 [%0], %1;\n\tred.global.add.u32 [%0+4], %2;" : : "l"(d.push_stat), "r"(kc), "r"(nb) : "memory");
        }
        return;
    }
    while (from < to) {
        const uint64_t n = to - from < (1u << 20) ? to - from : (uint64_t)(1u << 20);
        warp_copy<true>(d.hout + from, d.out0 + from, (uint32_t)n);
        from += n;
    }
#endif
}

LZ_DEV void probs_fill(uint16_t *p, uint32_t n) {  // initProbs, prob.go:3-7
#if defined(__CUDA_ARCH__)
    // uniform trip count, predicated store that names its state space (tables live in shared memory, or in
    // the HBM workspace when lc+lp > 4)
    if (__isShared(p)) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(p) + 2u * LZ_LANE();
        for (uint32_t b = 0; b < n; b += 32)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %1, %2;\n\t@q st.shared.u16 [%0], %3;\n\t}"
                         : : "r"(sa + 2u * b), "r"(b + LZ_LANE()), "r"(n), "r"(kProbInit) : "memory");
    } else {
        for (uint32_t b = 0; b < n; b += 32)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %1, %2;\n\t@q st.global.u16 [%0], %3;\n\t}"
                         : : "l"(p + b + LZ_LANE()), "r"(b + LZ_LANE()), "r"(n), "r"(kProbInit) : "memory");
    }
#else
    for (uint32_t i = 0; i < n; i++) p[i] = (uint16_t)kProbInit;
#endif
}

// state.Reset (state.go:79-121): every table back to 1024, rep0..3 = 0, state = 0.
template <int kV>
LZ_DEV void coder_reset(Dec &d, uint16_t *P, uint16_t *L, uint32_t lit_bits) {
    probs_fill(P, LZ_LAY(kV)::FIXED);
    probs_fill(L, 0x300u << lit_bits);
    d.rep0 = d.rep1 = d.rep2 = d.rep3 = 0;
    d.state = 0;
    LZ_SYNC();
}

// V_CHAIN addresses its tables and the input stage through 32-bit shared-window addresses
template <int kV>
LZ_DEV void set_shared_addrs(Dec &d, uint16_t *P, uint8_t *inbuf, uint8_t *stage) {
#if defined(__CUDA_ARCH__)
    d.sP = (uint32_t)__cvta_generic_to_shared(P);
    d.sL = d.sP + 2u * LZ_LAY(kV)::LIT;
    d.sIn = (uint32_t)__cvta_generic_to_shared(inbuf);
    d.sStage = (uint32_t)__cvta_generic_to_shared(stage);
#else
    d.sP = d.sL = d.sIn = d.sStage = 0;
#endif
    d.nb = d.ips = d.lims = 0;
    d.g0 = nullptr;
}

LZ_DEV void set_props(Dec &d, uint32_t lc, uint32_t lp, uint32_t pb) {
    d.lc = lc;
    d.lp_mask = (1u << lp) - 1;
    d.pos_mask = (1u << pb) - 1;
}

// V_CHAIN: (re)fill the shared input stage from the next unconsumed byte `gpos` and point the byte-ahead
// register at it.  False when too little input is left for the fast decoder to be worth entering.
template <int kV>
LZ_DEV bool f2_enter(Dec &d, const uint8_t *gpos, uint8_t *inbuf) {
#if defined(__CUDA_ARCH__)
    if ((uint64_t)(d.in_end - gpos) < kF2MinInput) return false;
    const uint8_t *g0 = (const uint8_t *)((uintptr_t)gpos & ~(uintptr_t)15);
    const uint64_t span = (uint64_t)(d.in_end - g0);
    const uint32_t avail = span >= kF2Stage ? kF2Stage : ((uint32_t)span & ~15u);   // whole 16-byte chunks of this unit only
    const uint32_t l = LZ_LANE();
    {   // one 16-byte cp.async (LDGSTS) per lane, global -> shared with no register in between; predicated, not
        // branched (avail >= 112: lanes beyond it name chunk 0 and copy nothing).  It is awaited right away -- a
        // second stage to fill ahead would cost the 14th unit per SM, and a refill is 0.1 % of a unit's time.
#pragma unroll
        for (uint32_t c = 0; c < kF2Stage / 512u; c++) {
            const uint32_t i = l + 32u * c;
            const bool live = 16u * i + 16u <= avail;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t"
                         "@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                         : : "r"(d.sIn + 16u * i), "l"(g0 + 16u * (live ? i : 0u)), "r"((uint32_t)live) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        const uint8_t *pf = g0 + kF2Stage + 128u * (l & (kF2Stage / 128u - 1u));
        if (pf >= d.in_end) pf = g0;                 // (same prefetch from every lane group: harmless)
        LZ_PREFETCH_L2(pf);
    }
    __syncwarp();
    if (d.ctx_pending == 1) {   // context bytes held as values (careful decoder's convention): fold them in
        d.prev_byte = d.ctx_a;
        d.mbyte = d.ctx_b;
        d.ctx_pending = 0;
    }
    const uint32_t off = (uint32_t)(gpos - g0);
    d.g0 = g0;
    d.ips = d.sIn + off;
    d.lims = d.sIn + avail - kF2Margin;
    d.nb = inbuf[off];
    return true;
#else
    return false;
#endif
}
// Streamed D2H: tell the host how much of this unit's output is final (stored and never touched again),
// in 64 KiB blocks.  Called from the stage-refill path only (every few hundred input bytes).
template <class Yield>
LZ_DEV void publish_progress(Dec &d, const WarpCopy &wc, Yield &yield) {
#if defined(__CUDA_ARCH__)
    if (d.hout) {
        const uint8_t *fin = wc.pend_len ? wc.pend_dst : d.outp;   // a pending window copy is not stored yet
        const uint32_t blocks = (uint32_t)((uint64_t)(fin - d.out0) >> 16);
        if (blocks != d.pub) {
            yield.push(d, (uint64_t)d.pub << 16, (uint64_t)blocks << 16);   // this warp, or the CTA's pusher warp
            d.pub = blocks;
        }
    } else if (d.prog) {
        const uint8_t *fin = wc.pend_len ? wc.pend_dst : d.outp;   // a pending window copy is not stored yet
        const uint32_t blocks = (uint32_t)((uint64_t)(fin - d.out0) >> 16);
        if (blocks != d.pub) {
            __threadfence_system();      // every lane: its window stores are visible before the counter is
            __syncwarp();
            *(volatile uint32_t *)d.prog = blocks;   // every lane, same value
            d.pub = blocks;
        }
    }
#endif
}
// ... and back: the careful decoder's lookahead starts empty at the next unconsumed byte
LZ_DEV void f2_leave(Dec &d) {
    if (d.ctx_pending == 2) {   // context bytes of the last window copy still in the copy stage: fetch them
        LZ_CP_WAIT();
        LZ_SYNC();
        d.ctx_a = d.stage[d.ctx_a - d.sStage];
        d.ctx_b = d.stage[d.ctx_b - d.sStage];
        d.ctx_pending = 1;
    }
    d.ip = d.g0 + (d.ips - d.sIn);
    d.inb_hi = d.inb_lo = 0;
    d.inbits = 0;
    d.phantom = 0;
}

// Decode symbols until the range-coded part ends (Reader1.Read driving
// decompress(), reader1.go:223-254).  On return d.status / d.site are set and no
// store is pending.
// Time slicing (the SM-resident scheduler of lzgpu.cu): a unit may be taken off its warp where the fast decoder refills
// its input stage -- the one point of the symbol loop where everything that is live sits in Dec / WarpCopy and in the
// unit's own shared memory.  run_lzma then returns RUN_YIELD; called again with resume = true (by any warp of the CTA, on
// the same Dec / WarpCopy contents) it carries on with that refill.
enum : int { RUN_DONE = 0, RUN_YIELD = 1 };
struct NoYield {
    LZ_DEV bool want(const Dec &) { return false; }
    LZ_DEV void push(const Dec &d, uint64_t from, uint64_t to) { push_out(d, from, to); }
};

template <int kV, class Yield = NoYield>
LZ_DEV int run_lzma(Dec &d, WarpCopy &wc, uint16_t *P, uint16_t *L, const uint8_t *dict_base, uint8_t *inbuf, Yield &yield,
                    bool resume = false) {
    bool fast = resume;   // which decoder runs (lzgpu_core.cuh, kFast)
    set_fast_limits(d);
    d.stage = wc.stage;
    for (;;) {
        uint32_t op, len = 0, dist = 0;
        if (kV & V_CHAIN) {
            for (;;) {
                if (fast) {
                    if (!resume) {
                        op = decode_fast2<kV>(d, wc, len, dist);
                        if (op != OP_SWITCH) break;
                        // stage used up (refill) or the unit's tail reached (careful decoder from here on)
                        publish_progress(d, wc, yield);
                        if (yield.want(d)) {
                            LZ_CP_WAIT();   // window-copy sources this warp's cp.async is still fetching into the copy stage
                            LZ_SYNC();
                            return RUN_YIELD;
                        }
                    }
                    resume = false;
                    if (d.outp <= d.fast_out_end && f2_enter<kV>(d, d.g0 + (d.ips - d.sIn), inbuf)) continue;
                    f2_leave(d);
                    fast = false;
                } else {
                    op = decode_run<kV, false>(d, P, L, len, dist);
                    if (op != OP_SWITCH) break;
                    fast = f2_enter<kV>(d, d.ip - (d.inbits >> 3), inbuf);   // true: fast_possible() checked the input left
                }
            }
            if (fast && op == OP_DONE) { f2_leave(d); fast = false; }
        } else
        for (;;) {
            if (fast) {
                op = decode_run<kV, true>(d, P, L, len, dist);
                if (op != OP_SWITCH) break;
                d.ip -= 4;                       // forget the word loaded ahead
                fast = false;
            } else {
                op = decode_run<kV, false>(d, P, L, len, dist);
                if (op != OP_SWITCH) break;
                d.nextw = LZ_LD_IN32(d.ip);      // fast decoder keeps one aligned word in hand
                d.ip += 4;
                fast = true;
            }
        }
        if (op == OP_DONE) break;
        uint8_t *dst = d.outp;

        wc_commit(wc);

        if (op == OP_COPY) {
            // Byte i of the match comes from src[src_index(i)], always a byte that predates the match
            // (period-dist replication when the match overlaps itself, window.go:73-86).
            const uint8_t *src = dst - dist;
            const bool far = dist > len;   // the common case: src_index(i) == i
            if ((kV & V_STAGE) && len <= 32 && dst + len + 4 <= wc.out_limit) {
                // cp.async the 4-byte-aligned words covering src[0 .. need) into shared memory; the
                // stores (wc_commit) and a following literal's context bytes read them from there.
                const uint32_t need = far ? len + 1 : dist;      // + the byte a matched literal needs
                const uint32_t off = (uint32_t)((uintptr_t)src & 3u);
                const uint8_t *a0 = src - off;
                const uint32_t nch = (off + need + 3) >> 2;      // <= 9
                LZ_FOR_LANES(l) {
                    if (l < nch) LZ_CP_ASYNC4(wc.stage + 4 * l, a0 + 4 * l);   // (V_STAGE experiments only)
                }
                LZ_CP_COMMIT();
                wc.pend_len = len;
                wc.pend_dst = dst;
                wc.pend_staged = 1;
                wc.pend_off = off;
                wc.pend_dist = dist;
                d.ctx_a = off + (far ? len - 1 : src_index(len - 1, dist));
                d.ctx_b = off + (far ? len : src_index(len, dist));
                d.ctx_pending = 2;
            } else if (kV & V_CHAIN) {
                // This decoder does its common copies itself (lzgpu_fast2.cuh); what arrives here is rare
                // (longer than 32 bytes, self-overlapping, or at the head / tail of a unit): copied at once,
                // the context of a following literal read back from the window.
                // All loads, then all stores: every source byte predates the match (src_index < dist), so the up
                // to nine bytes a lane moves (len <= 273) are in flight together -- one round trip to L2, not nine.
                // An overlapping match repeats its first `dist` bytes: lane l's k-th byte is at (l + 32 k) mod dist,
                // stepped without a division.
                const uint32_t step = far ? 32u : 32u % dist;
                LZ_FOR_LANES(l) {
                    uint32_t si = far ? l : l % dist;
                    uint32_t v[9];
#pragma unroll
                    for (uint32_t k = 0; k < 9; k++) {
                        const uint32_t live = l + 32u * k < len;
                        v[k] = 0;
                        LZ_LDG8_IF(v[k], src + (live ? si : 0u), live);
                        si += step;
                        if (!far && si >= dist) si -= dist;
                    }
#pragma unroll
                    for (uint32_t k = 0; k < 9; k++) LZ_STG8_IF(dst + l + 32u * k, v[k], l + 32u * k < len);
                }
                LZ_SYNC();
                d.prev_byte = dst[len - 1];
                d.mbyte = dst[(int64_t)len - (int64_t)dist];
                d.ctx_pending = 0;
            } else {
                if (len <= 32) {
                    // deferred in registers: load now, store at the next commit
                    LZ_FOR_LANES(l) {
                        const uint32_t live = l < len;
                        LZ_LDG8_IF(LZ_LV(wc.pend_val, l), src + (live ? (far ? l : src_index(l, dist)) : 0u), live);
                    }
                    wc.pend_len = len;
                    wc.pend_dst = dst;
                    wc.pend_staged = 0;
                } else {
                    for (uint32_t b = 0; b < len; b += 32) {
                        LZ_FOR_LANES(l) {
                            const uint32_t i = b + l, live = i < len;
                            uint32_t v = 0;
                            LZ_LDG8_IF(v, src + (live ? (far ? i : src_index(i, dist)) : 0u), live);
                            LZ_STG8_IF(dst + i, v, live);
                        }
                    }
                }
                // context for a literal that may follow: the last byte of the match and the byte at
                // -(rep0+1) after it.  One load site, results untouched until a literal asks for them.
                LZ_LD_WIN8(d.ctx_a, src + (far ? len - 1 : src_index(len - 1, dist)));
                LZ_LD_WIN8(d.ctx_b, src + (far ? len : src_index(len, dist)));
                d.ctx_pending = 1;
            }
            d.outp = dst + len;
        } else {  // OP_COPY_Q4: dist == bytes since dictionary start + 1; byte "-1" reads as 0
            for (uint32_t i = 0; i < len; i++) {   // every lane writes the same bytes
                const uint8_t *src = dst + i - dist;
                dst[i] = src < dict_base ? (uint8_t)0 : *src;
            }
            d.prev_byte = dst[len - 1];
            const uint8_t *m = dst + len - dist;
            d.mbyte = m < dict_base ? (uint8_t)0 : *m;
            d.ctx_pending = 0;
            d.outp = dst + len;
            LZ_SYNC();
        }
    }
    if (kV & V_CHAIN) { if (fast) f2_leave(d); }
    else if (fast) d.ip -= 4;   // the word loaded ahead was never consumed
    wc_commit(wc);
    return RUN_DONE;
}
template <int kV>
LZ_DEV void run_lzma(Dec &d, WarpCopy &wc, uint16_t *P, uint16_t *L, const uint8_t *dict_base, uint8_t *inbuf = nullptr) {
    NoYield ny;
    run_lzma<kV, NoYield>(d, wc, P, L, dict_base, inbuf, ny, false);
}

template <int kV>
LZ_DEV void reload_context(Dec &d, const uint8_t *dict_base) {
    const uint64_t hist = (uint64_t)(d.outp - dict_base);
    d.prev_byte = hist > 0 ? d.outp[-1] : 0;
    d.mbyte = ((uint64_t)d.rep0 + 1 <= hist) ? d.outp[-(int64_t)((uint64_t)d.rep0 + 1)] : 0;
    d.ctx_pending = 0;
}

struct UnitIO {
    const uint8_t *in;       // unit's compressed bytes
    uint64_t in_len;
    uint8_t *out;            // unit's output
    uint64_t out_cap;
    uint8_t *stage;          // 64 bytes of shared memory, 4-byte aligned (window-copy staging)
    uint8_t *inbuf;          // kF2Stage bytes of shared memory, 16-byte aligned (V_CHAIN input stage)
    uint32_t *progress;      // host-mapped progress counter of this unit (streamed D2H), or null
    uint8_t *hout;           // push mode: the unit's output range in the caller's pinned buffer (device view), or null
    uint32_t *push_stat;     // push mode: two counters of the launch (Dec::push_stat), or null
};

// LZMA1 unit (kind RAW; ALONE units are converted by the host): set-up, symbol loop, verdict.  The three are separate so
// that the SM-resident scheduler can run the loop in time slices.
template <int kV>
LZ_DEV bool lzma1_start(const lzgpu_unit &u, const UnitIO &io, uint16_t *P, uint16_t *L, Dec &d, WarpCopy &wc) {
    wc.pend_len = 0;
    wc.pend_dst = io.out;
    wc.pend_staged = 0;
    wc.pend_off = wc.pend_dist = 0;
    wc.stage = io.stage;
    wc.stage_sel = 0;
    wc.out_limit = io.out + io.out_cap;
    set_props(d, u.lc, u.lp, u.pb);
    set_shared_addrs<kV>(d, P, io.inbuf, io.stage);
    wc.s_stage = d.sStage;
    d.prog = io.progress;
    d.hout = io.hout;
    d.push_stat = io.push_stat;
    d.out0 = io.out;
    d.pub = 0;
    d.dict_size = u.dict_size;
    d.wpos = 0;
    d.full = 0;
    d.prev_byte = 0;
    d.mbyte = 0;
    d.ctx_a = d.ctx_b = 0;
    d.ctx_pending = 0;
    d.outp = io.out;
    d.size_defined = u.unpack_size != LZGPU_UNKNOWN_SIZE;   // state.go:135-151
    d.end_is_size = d.size_defined && u.unpack_size <= io.out_cap;
    d.out_end = io.out + (d.end_is_size ? u.unpack_size : io.out_cap);
    d.ip = io.in;
    d.in_end = io.in + io.in_len;
    d.status = LZGPU_OK;
    d.site = 0;
    d.nextw = 0;
    coder_reset<kV>(d, P, L, (uint32_t)u.lc + u.lp);

    const int32_t r = rc_init(d);
    if (r < 0) { d.status = LZGPU_UNEXPECTED_EOF; return false; }        // "rangeDec.Init: %w" of io.EOF
    if (r > 0) { d.status = LZGPU_RESULT_ERROR; d.site = LZGPU_SITE_RC_INIT; return false; }
    return true;
}
LZ_DEV void lzma1_finish(const Dec &d, const uint8_t *in, const uint8_t *out, lzgpu_result &res) {
    if (d.hout) push_out(d, (uint64_t)d.pub << 16, (uint64_t)(d.outp - out));
    LZ_IF_LANE0_ONLY {
        res.status = d.status;
        res.err_site = d.site;
        res.bytes_out = (uint64_t)(d.outp - out);
        res.bytes_in = rc_consumed(d, in);
        res.final_code = d.code;
    }
}
template <int kV>
LZ_DEV void run_unit_lzma1(const lzgpu_unit &u, const UnitIO &io, uint16_t *P, uint16_t *L,
                           lzgpu_result &res) {
    Dec d;
    WarpCopy wc;
    if (lzma1_start<kV>(u, io, P, L, d, wc)) run_lzma<kV>(d, wc, P, L, io.out, io.inbuf);
    lzma1_finish(d, io.in, io.out, res);
}

// LZMA2 group: walk the chunks (Reader2.startChunk + Read, reader2.go:100-250).
// Everything the walk keeps between two chunks -- and inside an LZMA chunk, while run_lzma decodes it -- lives in
// Lz2Walk, so that the walk can be left where run_lzma yields and taken up again by another warp (time slicing).
struct Lz2Walk {
    const uint8_t *in, *ip, *in_end, *payload, *dict_base;
    uint8_t *out, *out_limit;
    uint64_t in_len, in_rem, consumed;
    uint32_t have_coder, props, csz, short_payload, flags, lit_bits_cap;
    int32_t status, site;
};

template <int kV>
LZ_DEV void lzma2_start(const lzgpu_unit &u, const UnitIO &io, uint16_t *P, uint32_t lit_bits_cap, Dec &d, WarpCopy &wc, Lz2Walk &w) {
    wc.pend_len = 0;
    wc.pend_dst = io.out;
    wc.pend_staged = 0;
    wc.pend_off = wc.pend_dist = 0;
    wc.stage = io.stage;
    wc.stage_sel = 0;
    wc.out_limit = io.out + io.out_cap;
    set_props(d, u.lc, u.lp, u.pb);
    set_shared_addrs<kV>(d, P, io.inbuf, io.stage);
    wc.s_stage = d.sStage;
    d.prog = io.progress;
    d.hout = io.hout;
    d.push_stat = io.push_stat;
    d.out0 = io.out;
    d.pub = 0;
    d.dict_size = u.dict_size;
    d.wpos = 0;
    d.full = 0;
    d.prev_byte = 0;
    d.mbyte = 0;
    d.ctx_a = d.ctx_b = 0;
    d.ctx_pending = 0;
    d.rep0 = d.rep1 = d.rep2 = d.rep3 = 0;
    d.state = 0;
    d.range = 0xFFFFFFFFu;
    d.code = 0;
    d.inb_hi = d.inb_lo = 0;
    d.inbits = 0;
    d.phantom = 0;
    d.nextw = 0;
    d.outp = io.out;
    d.out_end = io.out;
    d.ip = io.in;
    d.in_end = io.in;
    d.end_is_size = 1;
    d.size_defined = 1;
    d.status = LZGPU_OK;
    d.site = 0;

    w.in = io.in;
    w.in_len = io.in_len;
    w.ip = io.in;                              // chunk cursor (uniform)
    w.in_end = io.in + io.in_len;
    w.payload = io.in;
    w.out = io.out;
    w.out_limit = io.out + io.out_cap;
    w.dict_base = io.out;                      // window.Reset() moves it (window.go:135-140)
    w.have_coder = 0;                          // r.lzmaReader != nil, for this unit
    w.props = ((uint32_t)u.pb * 5 + u.lp) * 9 + u.lc;  // r.header[5]: persists between chunks (Q8)
    w.csz = 0;
    w.short_payload = 0;
    w.in_rem = 0;
    w.flags = u.flags;
    w.lit_bits_cap = lit_bits_cap;
    w.status = LZGPU_OK;
    w.site = 0;
    w.consumed = 0;
}

// Walks until the group ends (RUN_DONE: w.status / w.site / w.consumed hold the verdict) or run_lzma yields inside an
// LZMA chunk (RUN_YIELD; call again with resume = true and the same d / wc / w).
template <int kV, class Yield>
LZ_DEV int lzma2_walk(Dec &d, WarpCopy &wc, Lz2Walk &w, uint16_t *P, uint16_t *L, uint8_t *inbuf, Yield &yield, bool resume) {
    for (;;) {
        if (!resume) {
        // ---- startChunk: lane 0 reads the header, everybody gets the verdict ----
        const uint8_t *ip = w.ip;
        uint32_t ctrl = 0, hdr = 0;  // hdr: [0]=end [1]=eof ; usz, csz below
        uint32_t usz = 0, csz = 0, newprops = 0xFFFFFFFFu;
        {
            const uint64_t rem = (uint64_t)(w.in_end - ip);
            if (rem == 0) {
                hdr = 2;  // ran off the unit: fine between units, UnexpectedEOF at the stream's end
            } else {
                ctrl = LZ_LD_IN8(ip);
                if (ctrl == 0 || (ctrl >= 3 && ctrl < 0x80)) {
                    hdr = 1;  // end of stream; 0x03..0x7F too (Q6, reader2.go:185-198)
                } else {
                    const uint32_t hl = ctrl < 0x80 ? 3 : (ctrl < 0xC0 ? 5 : 6);   // chunkLength, :201-214
                    if (rem < hl) {
                        hdr = 3;  // truncated header -> io.ErrUnexpectedEOF (:121-128)
                    } else {
                        usz = ((uint32_t)LZ_LD_IN8(ip + 1) << 8) | LZ_LD_IN8(ip + 2);          // :130
                        if (ctrl >= 0x80) {
                            usz |= (ctrl & 0x1Fu) << 16;                                        // :141
                            csz = (((uint32_t)LZ_LD_IN8(ip + 3) << 8) | LZ_LD_IN8(ip + 4)) + 1; // :143-144 (no uint16 wrap: Q7)
                            if (ctrl >= 0xC0) newprops = LZ_LD_IN8(ip + 5);
                        }
                        usz += 1;
                    }
                }
            }
        }
        if (hdr != 0) {
            if (hdr == 1) { w.consumed = (uint64_t)(ip - w.in) + 1; w.status = LZGPU_OK; }
            else {
                w.consumed = hdr == 3 ? w.in_len : (uint64_t)(ip - w.in);
                w.status = (hdr == 3 || (w.flags & LZGPU_UF_LZMA2_LAST)) ? LZGPU_UNEXPECTED_EOF : LZGPU_OK;
            }
            break;
        }
        const uint32_t hl = ctrl < 0x80 ? 3 : (ctrl < 0xC0 ? 5 : 6);
        const uint8_t *payload = ip + hl;
        if (newprops != 0xFFFFFFFFu) w.props = newprops;

        if (ctrl == 1 || ctrl >= 0xE0) {  // dictionary reset (:132-134)
            w.dict_base = d.outp;
            d.wpos = 0;
            d.full = 0;
        }

        if (ctrl < 0x80) {  // ---- uncompressed chunk: uncompressedRead (:252-294) ----
            uint64_t n = (uint64_t)(w.in_end - payload);
            const bool short_payload = n < usz;
            if (!short_payload) n = usz;
            if ((uint64_t)(w.out_limit - d.outp) < n) { w.status = LZGPU_OUTPUT_OVERFLOW; w.consumed = (uint64_t)(payload - w.in); break; }
            uint8_t *dst = d.outp;
            warp_copy_in(dst, payload, (uint32_t)n);     // n <= 65 536
            LZ_SYNC();
            d.outp = dst + n;
            uint64_t wp = (uint64_t)d.wpos + n;      // window.ReadFrom, window.go:142-155
            while (wp >= d.dict_size) { wp -= d.dict_size; d.full = 1; }
            d.wpos = (uint32_t)wp;
            w.ip = payload + n;
            if (short_payload) {  // the next header read hits EOF (:103-110)
                w.status = LZGPU_UNEXPECTED_EOF; w.consumed = w.in_len; break;
            }
            continue;
        }

        // ---- LZMA chunk ----
        if (!w.have_coder) {
            // First LZMA chunk of the unit.  At the start of a stream the reference builds
            // a new coder from header[5] whatever the control byte (reader2.go:146-153).
            // Elsewhere the scanner only cuts units where the chunk resets the state.
            if (!(w.flags & LZGPU_UF_LZMA2_FRESH) && ctrl < 0xA0) {
                w.status = LZGPU_RESULT_ERROR; w.site = LZGPU_SITE_LZMA2_NO_STATE; w.consumed = (uint64_t)(ip - w.in); break;
            }
        }
        if (!w.have_coder || ctrl >= 0xA0) {
            if (!w.have_coder || ctrl >= 0xC0) {      // DecodeProp(header[5]) + Renew (:158-165)
                if (w.props >= 225) { w.status = LZGPU_INCORRECT_PROPERTIES; w.consumed = (uint64_t)(ip - w.in); break; }
                const uint32_t lc = w.props % 9, r = w.props / 9, pb = r / 5, lp = r % 5;
                if (lc + lp > w.lit_bits_cap || pb > LZ_LAY(kV)::PB) {   // tables of this launch too small: the unit lied
                    w.status = LZGPU_RESULT_ERROR; w.site = LZGPU_SITE_LZMA2_PROPS; w.consumed = (uint64_t)(ip - w.in); break;
                }
                set_props(d, lc, lp, pb);
            }
            uint32_t lb = d.lc, m = d.lp_mask;
            while (m) { lb++; m >>= 1; }
            coder_reset<kV>(d, P, L, lb);                // s.Reset() (:156-157) / newState
            w.have_coder = 1;
        }
        const uint64_t in_rem = (uint64_t)(w.in_end - payload);
        const bool short_payload = in_rem < csz;
        d.ip = payload;
        d.in_end = payload + (short_payload ? in_rem : (uint64_t)csz);   // limitByteReader(in, cs)
        const uint64_t cap_left = (uint64_t)(w.out_limit - d.outp);
        d.size_defined = 1;                          // Reopen -> SetUnpackSize(us) (reader1.go:166-176)
        d.end_is_size = usz <= cap_left;
        d.out_end = d.outp + (d.end_is_size ? (uint64_t)usz : cap_left);
        d.status = LZGPU_OK;
        d.site = 0;
        int32_t r = 0;
        r = rc_init(d);
        if (r != 0) {
            w.consumed = (uint64_t)(payload - w.in);
            if (r < 0) { w.status = LZGPU_UNEXPECTED_EOF; w.consumed = w.in_len; }
            else { w.status = LZGPU_RESULT_ERROR; w.site = LZGPU_SITE_RC_INIT; }
            break;
        }
        reload_context<kV>(d, w.dict_base);
        w.payload = payload;
        w.csz = csz;
        w.short_payload = short_payload ? 1u : 0u;
        w.in_rem = in_rem;
        }
        if (run_lzma<kV, Yield>(d, wc, P, L, w.dict_base, inbuf, yield, resume) == RUN_YIELD) return RUN_YIELD;
        resume = false;

        // what the chunk did, as seen by every lane
        const uint8_t *payload = w.payload;
        const uint32_t csz = w.csz;
        const bool short_payload = w.short_payload != 0;
        const uint64_t in_rem = w.in_rem;
        const uint32_t st = (uint32_t)d.status, st_site = (uint32_t)d.site;
        const bool complete = d.outp == d.out_end && d.end_is_size;
        const bool exact = rc_consumed(d, payload) == (uint64_t)(d.in_end - payload);

        w.consumed = (uint64_t)(payload - w.in) + (short_payload ? in_rem : (uint64_t)csz);
        if (st == LZGPU_OK || st == LZGPU_OK_INPUT_EXHAUSTED) {
            if (!complete) {
                // The coder wanted more input than the chunk holds.  At the end of a truncated
                // stream the reference reports io.ErrUnexpectedEOF at the next header read;
                // otherwise it would carry on with a short chunk (Q1/Q8): documented deviation.
                if (short_payload) { w.status = LZGPU_UNEXPECTED_EOF; w.consumed = w.in_len; }
                else { w.status = LZGPU_RESULT_ERROR; w.site = LZGPU_SITE_LZMA2_CHUNK_SIZE; }
                break;
            }
            if (st == LZGPU_OK && !exact) {
                // Decoded size reached with compressed bytes left over: the reference parses the
                // next header from the middle of the payload (Q8).  Documented deviation.
                w.status = LZGPU_RESULT_ERROR; w.site = LZGPU_SITE_LZMA2_CHUNK_SIZE;
                break;
            }
            if (short_payload) { w.status = LZGPU_UNEXPECTED_EOF; w.consumed = w.in_len; break; }
            w.ip = payload + csz;
            continue;
        }
        w.status = (int32_t)st;
        w.site = (int32_t)st_site;
        break;
    }
    return RUN_DONE;
}

LZ_DEV void lzma2_finish(const Dec &d, const Lz2Walk &w, lzgpu_result &res) {
    if (d.hout) push_out(d, (uint64_t)d.pub << 16, (uint64_t)(d.outp - w.out));
    LZ_IF_LANE0_ONLY {
        res.status = w.status;
        res.err_site = w.site;
        res.bytes_out = (uint64_t)(d.outp - w.out);
        res.bytes_in = w.consumed;
        res.final_code = d.code;
    }
}

template <int kV>
LZ_DEV void run_unit_lzma2(const lzgpu_unit &u, const UnitIO &io, uint16_t *P, uint16_t *L,
                           uint32_t lit_bits_cap, lzgpu_result &res) {
    Dec d;
    WarpCopy wc;
    Lz2Walk w;
    NoYield ny;
    lzma2_start<kV>(u, io, P, lit_bits_cap, d, wc, w);
    lzma2_walk<kV, NoYield>(d, wc, w, P, L, io.inbuf, ny, false);
    lzma2_finish(d, w, res);
}

}  // namespace lzgpu
