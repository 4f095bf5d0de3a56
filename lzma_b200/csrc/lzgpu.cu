// lzgpu.cu -- sm_100a decode kernel + the host side of the C ABI in include/lzgpu.h.
//
// Host side = what the reference's readers do around the hot loop, re-stated for a
// batch: header parsing (reader1.go:77-147,178-221), LZMA2 chunk framing
// (reader2.go:100-214), plus what the reference never needed: unit discovery, a
// size-balanced scheduler over GPUs, and the H2D / D2H plumbing.
// There is no CPU decode path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "lzgpu_prep.h"
#include "lzgpu_unit.cuh"

using namespace lzgpu;

#ifndef LZGPU_DEFAULT_VARIANT
#define LZGPU_DEFAULT_VARIANT 33
#endif

// ------------------------------------------------------------------ kernel
struct KArgs {
    const lzgpu_unit *units;     // device copy of the plan's units (ALONE already rebased to RAW)
    const int32_t *order;        // launch slot -> unit index (longest compressed first)
    const uint8_t *in_base;
    uint8_t *out_base;
    lzgpu_result *results;
    uint16_t *lit_ws;            // literal tables in HBM for units with lc+lp > 4
    uint64_t lit_ws_stride;      // uint16 elements per slot
    uint32_t lit_bits_cap;       // literal-table capacity of this launch, as lc+lp
    uint32_t slot0;              // first slot of this launch in `order`
    uint32_t stage_off;          // uint16 index of the 64-byte staging buffer inside the shared array
    uint32_t *progress;          // per unit: decoded bytes that are final in HBM, in 64 KiB blocks (host-mapped; may be null)
    uint8_t *hout_base;          // push mode: device view of the caller's pinned output buffer (else null) ...
    const uint64_t *hout_off;    // ... and where in it each unit's output goes (host-mapped, indexed like `units`)
    uint32_t *push_stat;         // ... and two counters: time spent pushing whole blocks, blocks pushed (see Dec)
};

// One warp per CTA, one unit per warp.  Fixed tables (3.7 KB) always in shared
// memory; literal tables in shared memory when lc+lp <= 4 (<= 24 KB), else in HBM.
#ifndef LZGPU_MIN_CTAS
#define LZGPU_MIN_CTAS 14   // 14 units of lc+lp = 3, pb <= 2 fit one SM's shared memory: the registers must allow as many
#endif
template <bool kLitGlobal, int kV>
__global__ void __launch_bounds__(32, LZGPU_MIN_CTAS) lzgpu_decode_kernel(const KArgs a) {
    extern __shared__ __align__(16) uint16_t smem_probs[];
    const uint32_t slot = a.slot0 + blockIdx.x;
    const int32_t ui = a.order[slot];
    const lzgpu_unit u = a.units[ui];
    uint16_t *P = smem_probs;
    uint16_t *L = kLitGlobal ? a.lit_ws + (size_t)blockIdx.x * a.lit_ws_stride : smem_probs + LZ_LAY(kV)::LIT;
    UnitIO io;
    io.in = a.in_base + u.in_off;
    io.in_len = u.in_len;
    io.out = a.out_base + u.out_off;
    io.out_cap = u.out_cap;
    io.stage = reinterpret_cast<uint8_t *>(smem_probs + a.stage_off);
    io.inbuf = io.stage + 128;
    io.progress = a.progress ? a.progress + ui : nullptr;
    io.hout = a.hout_base ? a.hout_base + a.hout_off[ui] : nullptr;
    io.push_stat = a.hout_base ? a.push_stat : nullptr;
    lzgpu_result &res = a.results[ui];
    if (u.kind == LZGPU_KIND_LZMA2_GROUP) run_unit_lzma2<kV>(u, io, P, L, a.lit_bits_cap, res);
    else run_unit_lzma1<kV>(u, io, P, L, res);
}

// ------------------------------------------------------------------ SM-resident scheduler
// One CTA per SM holds up to 14 units (their tables in its shared memory) and as many warps; which warp decodes which
// unit changes while the kernel runs.  Why: a warp is tied to one of the SM's four sub-partitions (warp id mod 4), and a
// sub-partition that holds 4 warps gives each of them 3/4 of the issue rate the warps of a 3-warp sub-partition get.
// With one-warp CTAs and 2 048 units per GPU (BASELINE config 5 sharded over 8 GPUs: ONE wave, 13.8 units per SM) the
// kernel lasts as long as the units that happened to land on the crowded sub-partitions, and when units finish at
// different times the warps that are left stay where they are, however unevenly.  Here
//   * a unit can leave its warp wherever the fast decoder refills its input stage (run_lzma, RUN_YIELD): its state goes
//     to a save area next to its tables and any warp of the CTA can take it from there;
//   * every `rotate_every` refills a warp offers its unit for exchange and takes the one another warp offers, so over
//     their lifetime all units of the SM see the same mix of sub-partitions and advance at the SM's average rate;
//   * when units finish, warps on a sub-partition that now has more active warps than its share hand their unit to an
//     idle warp of a sub-partition with fewer (targets: the live units spread evenly over the four);
//   * a warp whose unit is done takes the next unit of the launch from a global counter (longest compressed first,
//     as before); the first wave is assigned statically (unit w * grid + cta), one unit of every size band per CTA.
// Both kinds of unit are time-sliced: an LZMA2 group's chunk walk keeps its state in Lz2Walk, saved with the rest.
struct SmArgs {
    uint32_t count;          // units of this launch (slots [slot0, slot0 + count) of `order`)
    uint32_t n_slots;        // units resident per CTA = warps per CTA
    uint32_t slot_bytes;     // shared memory of one slot: tables, stages, save area
    uint32_t save_off;       // offset of the save area inside a slot
    uint32_t rotate_every;   // refills between two looks of a warp for a unit to exchange its own with (0: never)
    uint32_t rotate_mode;    // 0: exchange with a unit that has less work left and sits on a less crowded sub-partition; 1: with anybody
    uint32_t *next;          // units started beyond the first wave (device counter, zeroed before the launch)
};
struct SmCtl {
    uint32_t lock;
    uint32_t live;           // units resident in this CTA and not finished
    uint32_t dry;            // the launch has no more units to start
    uint32_t q_head, q_tail; // ring of slots waiting for a warp
    uint32_t sp_active[4];   // warps decoding, per sub-partition
    uint32_t sp_over[4];     // that sub-partition has more active warps than its share while another has room
    uint32_t uneven;         // the active warps are not spread evenly: exchanging units pays
    uint32_t ring[16];
    uint32_t rem[16];        // per slot: compressed bytes its unit has left (0xffffffff: a unit that is not sliced)
    uint32_t where[16];      // per slot: sub-partition of the warp decoding it (4: waiting in the ring / empty)
    uint32_t origin[16];     // per slot waiting in the ring: sub-partition of the warp that put it there
};
struct SmSaved {
    Dec d;
    WarpCopy wc;
    Lz2Walk w;               // LZMA2 groups only
    const uint8_t *in;
    uint8_t *out;
    int32_t ui;
    uint32_t lzma2;
};
constexpr uint32_t kSmCtlBytes = 320;   // >= sizeof(SmCtl), multiple of 16
constexpr uint32_t kSmSaveBytes = (sizeof(SmSaved) + 15u) & ~15u;
constexpr uint32_t kSmMaxSlots = 14;
static_assert(sizeof(SmCtl) <= kSmCtlBytes, "SmCtl");

__device__ __forceinline__ uint32_t sm_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// Atomics are executed by ALL lanes (a predicated atom becomes a branch around it in SASS, and one divergent branch
// anywhere makes ptxas bracket every branch of the decoder with convergence barriers): the lock is taken by whichever
// lane's exchange reads 0, and the unit counter advances by 32 per fetch; a warp-wide minimum gives every lane the
// same answer.
__device__ __forceinline__ uint32_t sm_fetch_global(uint32_t *p) {   // returns 0, 1, 2, ... over the calls of all warps
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(r) : "l"(p) : "memory");
    return __reduce_min_sync(0xffffffffu, r) >> 5;
}
__device__ __forceinline__ void sm_lock(SmCtl *c) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(&c->lock);
#ifdef LZGPU_SM_WATCHDOG
    uint32_t spins = 0;
#endif
    for (;;) {
        uint32_t r;
        asm volatile("atom.shared.exch.b32 %0, [%1], 1;" : "=r"(r) : "r"(a) : "memory");
        if (__reduce_min_sync(0xffffffffu, r) == 0u) break;
        __nanosleep(40);
#ifdef LZGPU_SM_WATCHDOG
        if (++spins > 2000000u) asm volatile("trap;");
#endif
    }
    __threadfence_block();
}
__device__ __forceinline__ void sm_unlock(SmCtl *c) {
    __threadfence_block();
    __syncwarp();
    *(volatile uint32_t *)&c->lock = 0u;   // every lane, same value
}
// How evenly the active warps are spread over the four sub-partitions.  lo_spare: the fewest active warps on a
// sub-partition that still has an idle warp (0xffffffff: none has).
struct SmLoad {
    uint32_t act[4], hi, lo, lo_spare;
};
__device__ __forceinline__ SmLoad sm_load(volatile SmCtl *c, uint32_t n_slots) {
    SmLoad l;
    l.hi = 0;
    l.lo = l.lo_spare = 0xffffffffu;
#pragma unroll
    for (uint32_t sp = 0; sp < 4; sp++) {
        const uint32_t cap = (n_slots + 3u - sp) >> 2;          // warps of the CTA on this sub-partition
        const uint32_t v = sm_uniform(c->sp_active[sp]);
        l.act[sp] = v;
        if (cap) {
            l.hi = v > l.hi ? v : l.hi;
            l.lo = v < l.lo ? v : l.lo;
            if (v < cap) l.lo_spare = v < l.lo_spare ? v : l.lo_spare;
        }
    }
    return l;
}
// under the lock, after sp_active changed: a sub-partition is `over` when moving one of its units to an idle warp
// elsewhere would leave both better off (two or more apart)
__device__ __forceinline__ void sm_rebalance(volatile SmCtl *c, uint32_t n_slots) {
    const SmLoad l = sm_load(c, n_slots);
#pragma unroll
    for (uint32_t sp = 0; sp < 4; sp++)
        c->sp_over[sp] = (l.act[sp] == l.hi && l.lo_spare != 0xffffffffu && l.hi >= l.lo_spare + 2u) ? 1u : 0u;
    c->uneven = (l.hi != l.lo && l.hi >= 2u) ? 1u : 0u;   // some warps share a sub-partition while others have more room
}
// Would a warp of sub-partition `sp`, decoding a unit with `rem` compressed bytes left, do well to exchange it for the
// unit at the head of the ring?  Yes when that puts the unit with MORE left on the LESS crowded sub-partition (the
// warp that offered the head unit is idle: its sub-partition counts one more when it takes a unit again).  Units on a
// crowded sub-partition fall behind, so between equal units the same rule is a rotation that keeps them level; between
// unequal ones it lets the long ones run where fewer warps share the issue slots, and the SM drains later.
// mode 1: always (plain rotation).
__device__ __forceinline__ uint32_t sm_accept(volatile SmCtl *c, uint32_t sp, uint32_t rem, uint32_t mode) {
    const uint32_t qh = c->q_head;
    if (qh == c->q_tail) return 0u;
    if (mode == 1u) return 1u;
    const uint32_t h = c->ring[qh & 15u], ca = c->sp_active[c->origin[h] & 3u] + 1u, cb = c->sp_active[sp], ra = c->rem[h];
    return ((ca > cb && ra > rem) || (ca < cb && ra < rem)) ? 1u : 0u;
}
// Asked at every refill of the fast decoder's input stage: should this unit leave its warp?
//   1  a unit waits in the ring and exchanging this one for it is a gain (sm_accept);
//   2  this sub-partition holds more active warps than its share and another has an idle warp: leave it to that one;
//   3  (every `every` refills, when no unit of the launch is left to start) some unit on a less crowded sub-partition
//      has less left than this one: put this one into the ring and wait for one of those warps to come by.
struct SmYield {
    volatile SmCtl *c;
    uint32_t sp, every, mode, n_slots, slot, refills, intent, rem;
    const uint8_t *unit_end;   // end of the unit's compressed bytes (an LZMA2 group: of its last chunk)
    // push mode: the decoding warp writes its finished blocks to the caller's buffer itself.  (A spare warp per CTA doing
    // it for the others was tried: 113.3 -> 114.2 ms end to end -- its copies and the queue cost more than the waits.)
    __device__ __forceinline__ void push(const Dec &d, uint64_t from, uint64_t to) { push_out(d, from, to); }
    __device__ __forceinline__ bool want(const Dec &d) {
        refills++;
        rem = (uint32_t)(unit_end - (d.g0 + (d.ips - d.sIn)));
        c->rem[slot] = rem;
        uint32_t it = 0;
        if (sm_accept(c, sp, rem, mode)) it = 1;
        else if (c->sp_over[sp]) it = 2;
        else if (every && refills >= every && c->dry && c->uneven) {
            refills = 0;
            if (mode == 1u) it = 3;
            else {
                const uint32_t mine = c->sp_active[sp];
                for (uint32_t u = 0; u < n_slots; u++) {
                    const uint32_t wu = c->where[u], ru = c->rem[u];
                    if (wu < 4u && c->sp_active[wu] < mine && ru != 0xffffffffu && ru + 1024u < rem) it = 3;
                }
            }
        }
        intent = sm_uniform(it);
        return intent != 0u;
    }
};

template <int kV>
__global__ void __launch_bounds__(32 * kSmMaxSlots, 1) lzgpu_sm_kernel(const KArgs a, const SmArgs s) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    SmCtl *ctl = reinterpret_cast<SmCtl *>(sm_raw);
    volatile SmCtl *vc = ctl;
    // (the warp index through a shuffle: what derives from threadIdx counts as divergent, and every branch of the
    // decoder below would get a convergence barrier)
    const uint32_t w = sm_uniform(threadIdx.x >> 5), sp = w & 3u, cta = blockIdx.x, grid = gridDim.x;
    // first wave: one unit of band w for the warps that have one (every thread computes the same counts)
    // (bands of `grid` units in launch order, longest first; odd bands are dealt in reverse, so that the CTAs' sums agree)
    uint32_t k0 = 0;
    for (uint32_t i = 0; i < s.n_slots; i++) k0 += (i * grid + ((i & 1u) ? grid - 1u - cta : cta) < s.count) ? 1u : 0u;
    {
        vc->lock = 0;
        vc->live = k0;
        vc->dry = (uint64_t)grid * s.n_slots >= s.count ? 1u : 0u;
        vc->q_head = vc->q_tail = 0;
        for (uint32_t q = 0; q < 4; q++) {
            uint32_t n = 0;
            for (uint32_t i = q; i < k0; i += 4) n++;
            vc->sp_active[q] = n;
            vc->sp_over[q] = 0;
        }
        const uint32_t lo = k0 >> 2, hi = (k0 + 3u) >> 2;
        vc->uneven = (lo != hi && hi >= 2u) ? 1u : 0u;
        for (uint32_t i = 0; i < 16; i++) {
            vc->rem[i] = 0xffffffffu;
            vc->where[i] = i < k0 ? (i & 3u) : 4u;
            vc->origin[i] = 0;
        }
    }
    __syncthreads();

    SmYield yield{vc, sp, s.rotate_every, s.rotate_mode, s.n_slots, 0, 0, 0, 0, nullptr};
#ifdef LZGPU_SM_WATCHDOG
    uint32_t idle_polls = 0;
#endif
    uint32_t idx = w * grid + ((w & 1u) ? grid - 1u - cta : cta), have = (w < s.n_slots && idx < s.count) ? 1u : 0u, fresh = 1, slot = w, last_slot = 0xffffffffu, patience = 0;
    for (;;) {
        if (!have) {
            // idle: take a waiting unit if no sub-partition with an idle warp has fewer active warps than this one
            // (not the unit this warp has just offered for exchange, unless nobody wanted it); leave when the CTA
            // has nothing left.  The look is taken without the lock: idle warps must not keep it from the others.
            uint32_t take = 0;
            const uint32_t live = sm_uniform(vc->live);
            if (sm_uniform(vc->q_head) != sm_uniform(vc->q_tail)) {
                SmLoad l = sm_load(vc, s.n_slots);
                if (l.act[sp] == l.lo_spare) {
                    sm_lock(ctl);
                    const uint32_t qh = sm_uniform(vc->q_head), qt = sm_uniform(vc->q_tail);
                    l = sm_load(vc, s.n_slots);
                    if (qh != qt && l.act[sp] == l.lo_spare) {
                        const uint32_t head = sm_uniform(vc->ring[qh & 15u]);
                        if (head != last_slot || patience == 0) {
                            take = 1;
                            slot = head;
                            vc->q_head = qh + 1;
                            vc->sp_active[sp] = l.act[sp] + 1;
                            vc->where[head] = sp;
                            sm_rebalance(vc, s.n_slots);
                        }
                    }
                    sm_unlock(ctl);
                }
            }
            if (!take) {
                if (live == 0) return;
                if (patience) patience--;
                __nanosleep(2000);
#ifdef LZGPU_SM_WATCHDOG
                if (++idle_polls > (uint32_t)LZGPU_SM_WATCHDOG * 500000u) asm volatile("trap;");
#endif
                continue;
            }
            have = 1;
            fresh = 0;
            last_slot = 0xffffffffu;
        }
        // ---- this warp decodes the unit in `slot`
        for (;;) {
            uint8_t *sb = sm_raw + kSmCtlBytes + (size_t)slot * s.slot_bytes;
            uint16_t *P = reinterpret_cast<uint16_t *>(sb), *L = P + LZ_LAY(kV)::LIT;
            SmSaved *sv = reinterpret_cast<SmSaved *>(sb + s.save_off);
            UnitIO io;
            io.stage = reinterpret_cast<uint8_t *>(P + a.stage_off);
            io.inbuf = io.stage + 128;
            Dec d;
            WarpCopy wc;
            Lz2Walk w;
            const uint8_t *u_in;
            uint8_t *u_out;
            int32_t ui;
            uint32_t lzma2;
            bool resume, run = true;
            if (fresh) {
                ui = a.order[a.slot0 + idx];
                const lzgpu_unit u = a.units[ui];
                io.in = a.in_base + u.in_off;
                io.in_len = u.in_len;
                io.out = a.out_base + u.out_off;
                io.out_cap = u.out_cap;
                io.progress = a.progress ? a.progress + ui : nullptr;
                io.hout = a.hout_base ? a.hout_base + a.hout_off[ui] : nullptr;
                io.push_stat = a.hout_base ? a.push_stat : nullptr;
                u_in = io.in;
                u_out = io.out;
                resume = false;
                lzma2 = u.kind == LZGPU_KIND_LZMA2_GROUP ? 1u : 0u;
                vc->where[slot] = sp;
                vc->rem[slot] = u.in_len > 0xfffffffeull ? 0xfffffffeu : (uint32_t)u.in_len;
                if (lzma2) {
                    lzma2_start<kV>(u, io, P, a.lit_bits_cap, d, wc, w);
                } else {
                    run = lzma1_start<kV>(u, io, P, L, d, wc);
                    if (!run) lzma1_finish(d, u_in, u_out, a.results[ui]);
                }
            } else {
                d = sv->d;
                wc = sv->wc;
                u_in = sv->in;
                u_out = sv->out;
                ui = sv->ui;
                lzma2 = sv->lzma2;
                if (lzma2) w = sv->w;
                resume = true;
            }
            if (run) {
                yield.refills = 0;
                int r;
                uint32_t action = 0;   // 1: carry on with the unit in `slot` (another one after an exchange), 2: idle
                for (;;) {
                    yield.slot = slot;
                    if (lzma2) {
                        yield.unit_end = w.in_end;
                        r = lzma2_walk<kV, SmYield>(d, wc, w, P, L, io.inbuf, yield, resume);
                    } else {
                        yield.unit_end = d.in_end;
                        r = run_lzma<kV, SmYield>(d, wc, P, L, u_out, io.inbuf, yield, resume);
                    }
                    if (r != RUN_YIELD) break;
                    resume = true;
                    sm_lock(ctl);
                    const uint32_t qh = sm_uniform(vc->q_head), qt = sm_uniform(vc->q_tail), over = sm_uniform(vc->sp_over[sp]);
                    uint32_t give = 0;                     // 1: exchange, 2: leave to an idle warp, 3: offer
                    if (sm_uniform(sm_accept(vc, sp, yield.rem, s.rotate_mode))) give = 1;
                    else if (over) give = 2;
                    else if (yield.intent == 3 && sm_uniform(vc->dry) && sm_uniform(vc->uneven)) give = 3;
                    if (give) {
                        sv->d = d;
                        sv->wc = wc;
                        if (lzma2) sv->w = w;
                        sv->in = u_in;
                        sv->out = u_out;
                        sv->ui = ui;
                        sv->lzma2 = lzma2;
                        vc->ring[qt & 15u] = slot;
                        vc->q_tail = qt + 1;
                        vc->where[slot] = 4u;
                        vc->origin[slot] = sp;
                        if (give == 1) {
                            const uint32_t head = sm_uniform(vc->ring[qh & 15u]);
                            vc->q_head = qh + 1;
                            vc->where[head] = sp;
                            slot = head;
                            action = 1;
                        } else {
                            vc->sp_active[sp] = sm_uniform(vc->sp_active[sp]) - 1;
                            sm_rebalance(vc, s.n_slots);
                            last_slot = slot;
                            patience = give == 2 ? 0u : 300u;
                            action = 2;
                        }
                    }
                    sm_unlock(ctl);
                    if (action) break;
                }
                if (action == 1) { fresh = 0; continue; }
                if (action == 2) { have = 0; break; }
                if (lzma2) lzma2_finish(d, w, a.results[ui]);
                else lzma1_finish(d, u_in, u_out, a.results[ui]);
            }
            // ---- the unit is done: start the launch's next one in this slot, or retire
            uint32_t nidx = 0xffffffffu;
            if (!sm_uniform(vc->dry)) {
                nidx = grid * s.n_slots + sm_fetch_global(s.next);
                if (nidx >= s.count) nidx = 0xffffffffu;
            }
            if (nidx != 0xffffffffu) {
                idx = nidx;
                fresh = 1;
                continue;
            }
            sm_lock(ctl);
            vc->dry = 1;
            vc->live = sm_uniform(vc->live) - 1;
            vc->sp_active[sp] = sm_uniform(vc->sp_active[sp]) - 1;
            vc->where[slot] = 4u;
            vc->rem[slot] = 0xffffffffu;
            sm_rebalance(vc, s.n_slots);
            sm_unlock(ctl);
            have = 0;
            last_slot = 0xffffffffu;
            patience = 0;
            break;
        }
    }
}

// Tuning variant of the decoder (lzgpu_core.cuh, V_*): LZGPU_VARIANT in the environment
// overrides the default; read once per plan.
static int decoder_variant() {
    const char *e = getenv("LZGPU_VARIANT");
    if (e && *e >= '0' && *e <= '9') return atoi(e);
    return LZGPU_DEFAULT_VARIANT;
}

// pb2: the launch's units all have pb <= 2 -> the instantiation with compact posState tables (V_PB2, Lay<2>).
template <bool kLitGlobal>
static void launch_decode(int variant, bool pb2, unsigned grid, size_t smem, cudaStream_t st, const KArgs &a) {
    if (kLitGlobal && (variant & V_CHAIN)) variant = 1;   // V_CHAIN needs its literal tables in shared memory
    if (kLitGlobal) pb2 = false;                          // (rare class: one layout is enough)
#define LZ_LAUNCH(V)                                                                                      \
    do {                                                                                                  \
        if (pb2) lzgpu_decode_kernel<kLitGlobal, kLitGlobal ? (V) : ((V) | V_PB2)><<<grid, 32, smem, st>>>(a);   \
        else lzgpu_decode_kernel<kLitGlobal, (V)><<<grid, 32, smem, st>>>(a);                             \
    } while (0)
    switch (variant) {
        case 33: LZ_LAUNCH(kLitGlobal ? 1 : 33); break;
        case 0: LZ_LAUNCH(0); break;
        default: LZ_LAUNCH(1); break;
    }
#undef LZ_LAUNCH
}

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(LZGPU_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    } while (0)

extern "C" int lzgpu_abi_version(void) { return LZGPU_ABI_VERSION; }

extern "C" int lzgpu_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_last_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" const char *lzgpu_last_error(void) { return g_last_error.c_str(); }

extern "C" const char *lzgpu_status_name(int s) {
    switch (s) {
        case LZGPU_OK: return "OK";
        case LZGPU_OK_INPUT_EXHAUSTED: return "OK_INPUT_EXHAUSTED";
        case LZGPU_RESULT_ERROR: return "RESULT_ERROR";
        case LZGPU_INCORRECT_PROPERTIES: return "INCORRECT_PROPERTIES";
        case LZGPU_UNEXPECTED_EOF: return "UNEXPECTED_EOF";
        case LZGPU_OUTPUT_OVERFLOW: return "OUTPUT_OVERFLOW";
        case LZGPU_DICT_OUT_OF_RANGE: return "DICT_OUT_OF_RANGE";
        case LZGPU_UNEXPECTED_LZMA2_CODE: return "UNEXPECTED_LZMA2_CODE";
        case LZGPU_NOT_RUN: return "NOT_RUN";
    }
    return "?";
}

// ------------------------------------------------------------------ header helpers
extern "C" int lzgpu_decode_prop(uint8_t d, uint8_t *lc, uint8_t *pb, uint8_t *lp) {
    if (d >= 9 * 5 * 5) return LZGPU_INCORRECT_PROPERTIES;   // reader1.go:211-213
    *lc = d % 9;
    d /= 9;
    *pb = d / 5;
    *lp = d % 5;
    return LZGPU_OK;
}
extern "C" uint32_t lzgpu_decode_dict_size(const uint8_t p[4]) {
    uint32_t d = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
    return d < 4096u ? 4096u : d;                            // lzmaDicMin, reader1.go:199-201
}
extern "C" uint64_t lzgpu_decode_unpack_size(const uint8_t h[8]) {
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= (uint64_t)h[i] << (8 * i);
    return n;
}
extern "C" uint32_t lzgpu_decode_dict_size2(uint8_t b) { return (uint32_t)(2 | (b & 1)) << (b / 2 + 11); }

extern "C" int lzgpu_parse_alone_header(const uint8_t *in, uint64_t in_len, lzgpu_unit *u) {
    if (in_len < 1) return LZGPU_UNEXPECTED_EOF;             // bare io.EOF, reader1.go:78-81
    uint8_t lc, pb, lp;
    if (lzgpu_decode_prop(in[0], &lc, &pb, &lp) != LZGPU_OK) return LZGPU_INCORRECT_PROPERTIES;
    if (in_len < 13) return LZGPU_UNEXPECTED_EOF;            // "decode dict size/unpack size: EOF"
    u->kind = LZGPU_KIND_LZMA1_ALONE;
    u->lc = lc; u->lp = lp; u->pb = pb;
    u->lit_bits = (uint8_t)(lc + lp);
    u->dict_size = lzgpu_decode_dict_size(in + 1);
    u->unpack_size = lzgpu_decode_unpack_size(in + 5);
    return LZGPU_OK;
}

// ------------------------------------------------------------------ LZMA2 scanner
extern "C" int64_t lzgpu_scan_lzma2(const uint8_t *in, uint64_t in_len, uint32_t dict_size,
                                    lzgpu_unit *units, int64_t max_units,
                                    uint64_t *total_out, int32_t *stream_status) {
    if (dict_size < 4096u) dict_size = 8u << 20;             // validateDictSize, reader2.go:88-91
    uint64_t pos = 0, out = 0;
    uint32_t props = 0;          // r.header[5], initially 0 (Q8)
    bool seen_lzma = false;      // r.lzmaReader != nil
    int64_t n = 0;
    int32_t sst = LZGPU_OK;

    lzgpu_unit cur;
    auto open_unit = [&](uint64_t at) {
        memset(&cur, 0, sizeof cur);
        cur.kind = LZGPU_KIND_LZMA2_GROUP;
        cur.in_off = at;
        cur.out_off = out;
        cur.dict_size = dict_size;
        const uint32_t p = props < 225 ? props : 0;
        cur.lc = (uint8_t)(p % 9);
        cur.lp = (uint8_t)((p / 9) % 5);
        cur.pb = (uint8_t)((p / 9) / 5);
        cur.lit_bits = 0;
        cur.pos_bits = 0;
        cur.flags = (seen_lzma ? 0u : LZGPU_UF_LZMA2_FRESH) | LZGPU_UF_BITS_KNOWN;
    };
    auto close_unit = [&](uint64_t end, bool last) {
        cur.in_len = end - cur.in_off;
        cur.out_cap = out - cur.out_off;
        cur.unpack_size = cur.out_cap;
        if (last) cur.flags |= LZGPU_UF_LZMA2_LAST;
        if (n < max_units && units) units[n] = cur;
        n++;
    };
    open_unit(0);

    for (;;) {
        if (pos >= in_len) { sst = LZGPU_UNEXPECTED_EOF; break; }          // reader2.go:103-110
        const uint32_t c = in[pos];
        if (c == 0 || (c >= 3 && c < 0x80)) { pos += 1; break; }           // end of stream (and Q6)
        const uint32_t hl = c < 0x80 ? 3 : (c < 0xC0 ? 5 : 6);
        if (in_len - pos < hl) { pos = in_len; sst = LZGPU_UNEXPECTED_EOF; break; }
        uint32_t usz = ((uint32_t)in[pos + 1] << 8 | in[pos + 2]);
        uint32_t pl;
        if (c >= 0x80) {
            usz |= (c & 0x1Fu) << 16;
            pl = (((uint32_t)in[pos + 3] << 8) | in[pos + 4]) + 1;
        }
        usz += 1;
        if (c < 0x80) pl = usz;

        // A unit may begin at a dictionary reset provided the first LZMA chunk from here
        // on resets the coder state too (0xE0.. does; after 0x01 look ahead).
        if ((c == 1 || c >= 0xE0) && pos != cur.in_off) {
            bool cut = true;
            if (c == 1) {
                uint64_t q = pos;
                for (;;) {
                    if (q >= in_len) break;
                    const uint32_t cc = in[q];
                    if (cc == 0 || (cc >= 3 && cc < 0x80)) break;
                    if (cc >= 0x80) { cut = cc >= 0xA0; break; }
                    if (q != pos && cc == 1) break;
                    if (in_len - q < 3) break;
                    q += 3 + (((uint64_t)in[q + 1] << 8 | in[q + 2]) + 1);
                }
            }
            if (cut) {
                close_unit(pos, false);
                open_unit(pos);
            }
        }
        if (c >= 0xC0) props = in[pos + 5];
        if (c >= 0x80) {
            if (props < 225) {
                const uint32_t lb = props % 9 + (props / 9) % 5, pbits = (props / 9) / 5;
                if (lb > cur.lit_bits) cur.lit_bits = (uint8_t)lb;
                if (pbits > cur.pos_bits) cur.pos_bits = (uint8_t)pbits;
            }
            seen_lzma = true;
        }
        out += usz;
        if (in_len - pos - hl < pl) { pos = in_len; sst = LZGPU_UNEXPECTED_EOF; break; }
        pos += hl + pl;
    }
    // the last unit owns everything up to the end of the buffer, so that the device walk
    // sees exactly what the reference would read
    close_unit(in_len, true);
    if (total_out) *total_out = out;
    if (stream_status) *stream_status = sst;
    return n;
}

// ------------------------------------------------------------------ scheduler
extern "C" int lzgpu_shard_units(const lzgpu_unit *units, int64_t n, int n_shards, int32_t *shard_of_unit) {
    if (n_shards < 1 || n < 0 || (n > 0 && (!units || !shard_of_unit))) return fail(LZGPU_E_INVALID, "shard_units: bad arguments");
    std::vector<int64_t> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return units[a].in_len > units[b].in_len; });
    std::vector<uint64_t> load((size_t)n_shards, 0);
    for (int64_t k = 0; k < n; k++) {
        int best = 0;
        for (int s = 1; s < n_shards; s++) if (load[s] < load[best]) best = s;
        shard_of_unit[idx[k]] = best;
        load[best] += units[idx[k]].in_len + 64;   // +64: every unit costs a table reset even when tiny
    }
    return LZGPU_E_OK;
}

// ------------------------------------------------------------------ context / plan
struct DevState {
    int device = -1;
    cudaStream_t stream = nullptr;
    // grow-only staging for the host-buffer entry point
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint64_t in_cap = 0, out_cap = 0;
    // launches of different table classes of one batch run side by side (fork / join around the caller's stream)
    static constexpr int kAux = 4;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[kAux] = {nullptr, nullptr, nullptr, nullptr};
    // streamed D2H: second stream, host-mapped progress counters (grow-only)
    cudaStream_t copy_stream = nullptr;
    uint32_t *h_progress = nullptr, *d_progress = nullptr;
    uint64_t *h_tails = nullptr, *d_tails = nullptr;   // host-mapped (source offset, destination offset, length) per unit
    uint64_t progress_cap = 0;
    // descriptors / results of the batch in flight (grow-only): the host-buffer entry point allocates nothing per
    // call (cudaMalloc / cudaFree serialise against the other GPUs' work in a multi-GPU process: 12 ms seen)
    uint8_t *d_desc = nullptr;
    uint64_t desc_cap = 0;
    // per-unit checksums of lzgpu_plan_crc32 / lzgpu_plan_crc64 (grow-only, for the same reason)
    uint8_t *d_sum = nullptr;
    uint64_t sum_cap = 0;
    // literal tables of units with lc+lp > 4 (grow-only, for the same reason: a cudaFree per call was seen to take
    // 0.1 - 0.8 s now and then)
    uint8_t *d_litws = nullptr;
    uint64_t litws_cap = 0;
    // push mode found the link to the host too busy (see run_shard): copy engines until the next probe
    bool push_slow = false;
    uint32_t push_skipped = 0;
};

struct lzgpu_ctx {
    std::vector<DevState> devs;
    std::mutex mu;
};

struct Launch {
    uint32_t lit_bits;   // class
    bool lit_global;
    bool pb2;            // compact posState tables (every unit of the launch has pb <= 2)
    uint32_t slot0, count;
    size_t smem;
    uint32_t sm_slots = 0, sm_grid = 0, sm_slot_bytes = 0;   // SM-resident scheduler geometry (0: one-warp CTAs)
};

struct lzgpu_plan {
    lzgpu_ctx *ctx = nullptr;
    int dev_index = 0;
    int64_t n = 0;
    uint64_t in_size = 0, out_size = 0;
    std::vector<lzgpu_unit> units;        // device view (ALONE rebased)
    std::vector<lzgpu_result> preset;     // results decided on the host (status != NOT_RUN)
    std::vector<uint8_t> was_alone;       // unit arrived as LZMA1_ALONE (13-byte header rebased away)
    std::vector<int32_t> order;
    std::vector<Launch> launches;
    lzgpu_unit *d_units = nullptr;
    int32_t *d_order = nullptr;
    lzgpu_result *d_results = nullptr;
    uint16_t *d_lit_ws = nullptr;
    uint32_t *d_progress = nullptr;       // optional, set by the host-buffer entry point
    uint8_t *d_hout_base = nullptr;       // push mode (host-buffer entry point, pinned output): see KArgs
    const uint64_t *d_hout_off = nullptr;
    uint32_t *d_next = nullptr;           // one unit counter per launch (SM-resident scheduler)
    uint64_t lit_ws_stride = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t last_stream = nullptr;
    bool launched = false;
    bool borrowed = false;                // d_units / d_order / d_results live in the device's descriptor arena
    int variant = 0;
    int max_ctas_per_sm = 0;              // 0: as many as fit (14 at lc3 lp0 pb2)
    uint32_t rotate_every = 16;           // SM-resident scheduler: refills between two looks for an exchange (LZGPU_ROTATE)
    uint32_t rotate_mode = 0;             // 0: exchange towards "more work left on the less crowded sub-partition"; 1: blind (LZGPU_ROTATE_MODE)
};

extern "C" int lzgpu_ctx_create(const int *devices, int n_devices, lzgpu_ctx **out) {
    if (!out) return fail(LZGPU_E_INVALID, "ctx_create: null out");
    const int avail = lzgpu_device_count();
    if (avail <= 0) return fail(LZGPU_E_NO_DEVICE, "no CUDA device visible: this library has no CPU decode path (" + g_last_error + ")");
    std::vector<int> ids;
    if (n_devices <= 0 || !devices) for (int i = 0; i < avail; i++) ids.push_back(i);
    else for (int i = 0; i < n_devices; i++) ids.push_back(devices[i]);
    lzgpu_ctx *c = new lzgpu_ctx();
    for (int id : ids) {
        if (id < 0 || id >= avail) { lzgpu_ctx_destroy(c); return fail(LZGPU_E_INVALID, "ctx_create: device ordinal out of range"); }
        DevState d;
        d.device = id;
        cudaError_t e = cudaSetDevice(id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { lzgpu_ctx_destroy(c); return fail(LZGPU_E_CUDA, std::string("ctx_create: ") + cudaGetErrorString(e)); }
        c->devs.push_back(d);
    }
    *out = c;
    return LZGPU_E_OK;
}

extern "C" void lzgpu_ctx_destroy(lzgpu_ctx *c) {
    if (!c) return;
    for (auto &d : c->devs) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        for (int k = 0; k < DevState::kAux; k++) {
            if (d.aux[k]) cudaStreamDestroy(d.aux[k]);
            if (d.join_ev[k]) cudaEventDestroy(d.join_ev[k]);
        }
        if (d.fork_ev) cudaEventDestroy(d.fork_ev);
        if (d.h_progress) cudaFreeHost(d.h_progress);
        if (d.h_tails) cudaFreeHost(d.h_tails);
        if (d.d_desc) cudaFree(d.d_desc);
        if (d.d_sum) cudaFree(d.d_sum);
        if (d.d_litws) cudaFree(d.d_litws);
        if (d.d_in) cudaFree(d.d_in);
        if (d.d_out) cudaFree(d.d_out);
    }
    delete c;
}

extern "C" int lzgpu_ctx_device_count(const lzgpu_ctx *c) { return c ? (int)c->devs.size() : 0; }

extern "C" void lzgpu_plan_destroy(lzgpu_plan *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->devs[p->dev_index].device);
    if (!p->borrowed) {
        if (p->d_next) cudaFree(p->d_next);
        if (p->d_units) cudaFree(p->d_units);
        if (p->d_order) cudaFree(p->d_order);
        if (p->d_results) cudaFree(p->d_results);
    }
    if (p->d_lit_ws && !p->borrowed) cudaFree(p->d_lit_ws);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}

// shared memory of one unit: probability tables, 2 x 64 bytes of window-copy staging, the V_CHAIN input stage
static size_t probs_elems(uint32_t lit_bits, bool lit_global, bool pb2) {
    return (size_t)(pb2 && !lit_global ? Lay<2>::FIXED : Lay<4>::FIXED) + (lit_global ? 0 : ((size_t)0x300 << lit_bits));
}
static size_t smem_bytes(uint32_t lit_bits, bool lit_global, bool pb2) {
    return sizeof(uint16_t) * probs_elems(lit_bits, lit_global, pb2) + 128 + kF2Stage;
}

static int ensure(uint8_t *&ptr, uint64_t &cap, uint64_t need);
static int plan_create_impl(lzgpu_ctx *ctx, int dev_index, const lzgpu_unit *units, int64_t n,
                            uint64_t in_size, uint64_t out_size, lzgpu_plan **out, bool use_arena);
extern "C" int lzgpu_plan_create(lzgpu_ctx *ctx, int dev_index, const lzgpu_unit *units, int64_t n,
                                 uint64_t in_size, uint64_t out_size, lzgpu_plan **out) {
    return plan_create_impl(ctx, dev_index, units, n, in_size, out_size, out, false);
}
static int plan_create_impl(lzgpu_ctx *ctx, int dev_index, const lzgpu_unit *units, int64_t n,
                            uint64_t in_size, uint64_t out_size, lzgpu_plan **out, bool use_arena) {
    if (!ctx || !out || n < 0 || (n > 0 && !units)) return fail(LZGPU_E_INVALID, "plan_create: bad arguments");
    if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(LZGPU_E_INVALID, "plan_create: dev_index out of range");
    if (n > INT32_MAX) return fail(LZGPU_E_INVALID, "plan_create: too many units");
    CUDA_TRY(cudaSetDevice(ctx->devs[dev_index].device));
    lzgpu_plan *p = new lzgpu_plan();
    p->ctx = ctx;
    p->dev_index = dev_index;
    p->variant = decoder_variant();
    if (const char *e = getenv("LZGPU_MAX_CTAS_PER_SM")) p->max_ctas_per_sm = atoi(e);
    if (const char *e = getenv("LZGPU_ROTATE")) p->rotate_every = (uint32_t)atoi(e);
    if (const char *e = getenv("LZGPU_ROTATE_MODE")) p->rotate_mode = (uint32_t)atoi(e);
    p->n = n;
    p->in_size = in_size;
    p->out_size = out_size;
    p->units.assign(units, units + n);
    p->preset.resize((size_t)n);
    p->was_alone.assign((size_t)n, 0);
    std::vector<int32_t> runnable;
    for (int64_t i = 0; i < n; i++) {
        lzgpu_unit &u = p->units[(size_t)i];
        lzgpu_result &r = p->preset[(size_t)i];
        if (u.in_off > in_size || u.in_len > in_size - u.in_off || u.out_off > out_size || u.out_cap > out_size - u.out_off) {
            lzgpu_plan_destroy(p);
            return fail(LZGPU_E_INVALID, "plan_create: unit " + std::to_string(i) + " lies outside the buffers");
        }
        if (u.kind != LZGPU_KIND_LZMA1_ALONE && u.kind != LZGPU_KIND_LZMA1_RAW && u.kind != LZGPU_KIND_LZMA2_GROUP) {
            lzgpu_plan_destroy(p);
            return fail(LZGPU_E_INVALID, "plan_create: unknown unit kind");
        }
        bool alone = false;
        const bool run = prepare_unit(u, r, &alone);
        r.device = ctx->devs[dev_index].device;
        p->was_alone[(size_t)i] = alone;
        if (!run) continue;
        runnable.push_back((int32_t)i);
    }
    // launch order: by table class (literal-table size; posState tables compact or full), then longest
    // compressed input first (LPT within the GPU)
    auto cls = [](const lzgpu_unit &u) -> uint32_t { return u.lit_bits <= 4 ? 2u * u.lit_bits + (u.pos_bits > 2 ? 1u : 0u) : 100u; };
    std::stable_sort(runnable.begin(), runnable.end(), [&](int32_t a, int32_t b) {
        const lzgpu_unit &x = p->units[(size_t)a], &y = p->units[(size_t)b];
        const uint32_t cx = cls(x), cy = cls(y);
        if (cx != cy) return cx < cy;
        return x.in_len > y.in_len;
    });
    p->order = runnable;
    uint32_t s = 0;
    uint64_t ws_slots = 0;
    uint32_t ws_bits = 0;
    while (s < runnable.size()) {
        const lzgpu_unit &u0 = p->units[(size_t)runnable[s]];
        const bool g = u0.lit_bits > 4;
        uint32_t e = s;
        uint32_t maxbits = 0;
        while (e < runnable.size()) {
            const lzgpu_unit &ue = p->units[(size_t)runnable[e]];
            if (g ? ue.lit_bits <= 4 : cls(ue) != cls(u0)) break;
            maxbits = std::max<uint32_t>(maxbits, ue.lit_bits);
            e++;
        }
        if (g) {
            // HBM literal tables: bound the workspace by launching in waves
            const uint64_t per = (uint64_t)0x300 << maxbits;
            const uint64_t max_slots = std::max<uint64_t>(1, ((uint64_t)2 << 30) / (per * 2));
            ws_bits = std::max(ws_bits, maxbits);
            for (uint32_t w = s; w < e; w += (uint32_t)max_slots) {
                const uint32_t cnt = (uint32_t)std::min<uint64_t>(max_slots, e - w);
                p->launches.push_back({maxbits, true, false, w, cnt, smem_bytes(maxbits, true, false)});
                ws_slots = std::max<uint64_t>(ws_slots, cnt);
            }
        } else {
            const bool pb2 = u0.pos_bits <= 2;
            p->launches.push_back({maxbits, false, pb2, s, e - s, smem_bytes(maxbits, false, pb2)});
        }
        s = e;
    }
    // SM-resident scheduler (lzgpu_sm_kernel) for the launches of the default decoder whose tables live in shared memory:
    // one CTA per SM, `sm_slots` units resident per CTA.  LZGPU_SCHED=0 keeps the one-warp CTAs.
    {
        const bool sched_on = !(getenv("LZGPU_SCHED") && atoi(getenv("LZGPU_SCHED")) == 0);
        int nsm = 0;
        if (sched_on && p->variant == 33 && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->devs[dev_index].device) == cudaSuccess && nsm > 0) {
            bool any = false;
            for (size_t li = 0; li < p->launches.size() && li < 62; li++) {   // (counters 62, 63: push-mode statistics)
                Launch &L = p->launches[li];
                if (L.lit_global || L.count == 0 || L.count > (1u << 26)) continue;   // (the unit counter advances by 32 per fetch)
                const uint32_t slot_bytes = (uint32_t)((L.smem + 15u) & ~(size_t)15u) + kSmSaveBytes;
                uint32_t smax = std::min<uint32_t>(kSmMaxSlots, (232448u - kSmCtlBytes) / slot_bytes);
                if (p->max_ctas_per_sm > 0) smax = std::min<uint32_t>(smax, (uint32_t)p->max_ctas_per_sm);
                if (smax == 0) continue;
                L.sm_grid = std::min<uint32_t>((uint32_t)nsm, L.count);
                L.sm_slots = std::min<uint32_t>(smax, (L.count + L.sm_grid - 1) / L.sm_grid);
                L.sm_slot_bytes = slot_bytes;
                any = true;
            }
            if (any) {
                cudaFuncSetAttribute(lzgpu_sm_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
                cudaFuncSetAttribute(lzgpu_sm_kernel<33 | V_PB2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
            }
        }
    }
    auto bail = [&](cudaError_t e, const char *what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(e);
        lzgpu_plan_destroy(p);
        return fail(e == cudaErrorMemoryAllocation ? LZGPU_E_NOMEM : LZGPU_E_CUDA, m);
    };
    cudaError_t e;
    if (n > 0 && use_arena) {
        DevState &ds = ctx->devs[dev_index];
        const uint64_t need = (sizeof(lzgpu_unit) + sizeof(lzgpu_result) + sizeof(int32_t)) * (uint64_t)n + 256 + 512;
        if (ds.desc_cap < need) {
            if (ds.d_desc) cudaFree(ds.d_desc);
            ds.d_desc = nullptr; ds.desc_cap = 0;
            const uint64_t want = need + (need >> 2);
            if ((e = cudaMalloc(&ds.d_desc, want)) != cudaSuccess) return bail(e, "cudaMalloc descriptor arena");
            ds.desc_cap = want;
        }
        p->borrowed = true;
        p->d_units = reinterpret_cast<lzgpu_unit *>(ds.d_desc);
        p->d_results = reinterpret_cast<lzgpu_result *>(ds.d_desc + sizeof(lzgpu_unit) * (size_t)n);
        p->d_order = reinterpret_cast<int32_t *>(ds.d_desc + (sizeof(lzgpu_unit) + sizeof(lzgpu_result)) * (size_t)n);
        p->d_next = reinterpret_cast<uint32_t *>(ds.d_desc + (sizeof(lzgpu_unit) + sizeof(lzgpu_result) + sizeof(int32_t)) * (size_t)n);   // 64 counters
    }
    if (n > 0) {
        if (!use_arena) {
        if ((e = cudaMalloc(&p->d_units, sizeof(lzgpu_unit) * (size_t)n)) != cudaSuccess) return bail(e, "cudaMalloc units");
        if ((e = cudaMalloc(&p->d_results, sizeof(lzgpu_result) * (size_t)n)) != cudaSuccess) return bail(e, "cudaMalloc results");
        if ((e = cudaMalloc(&p->d_order, sizeof(int32_t) * std::max<size_t>(1, runnable.size()))) != cudaSuccess) return bail(e, "cudaMalloc order");
        if ((e = cudaMalloc(&p->d_next, sizeof(uint32_t) * 64)) != cudaSuccess) return bail(e, "cudaMalloc counters");
        }
        if ((e = cudaMemcpy(p->d_units, p->units.data(), sizeof(lzgpu_unit) * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload units");
        if (!runnable.empty() && (e = cudaMemcpy(p->d_order, runnable.data(), sizeof(int32_t) * runnable.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload order");
        if ((e = cudaMemcpy(p->d_results, p->preset.data(), sizeof(lzgpu_result) * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload results");
    }
    if (ws_slots) {
        p->lit_ws_stride = (uint64_t)0x300 << ws_bits;
        if (use_arena) {
            DevState &ds = ctx->devs[dev_index];
            if (ensure(ds.d_litws, ds.litws_cap, ws_slots * p->lit_ws_stride * 2)) return bail(cudaErrorMemoryAllocation, "cudaMalloc literal workspace");
            p->d_lit_ws = reinterpret_cast<uint16_t *>(ds.d_litws);
        } else if ((e = cudaMalloc(&p->d_lit_ws, ws_slots * p->lit_ws_stride * 2)) != cudaSuccess) return bail(e, "cudaMalloc literal workspace");
    }
    if ((e = cudaEventCreate(&p->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&p->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    *out = p;
    return LZGPU_E_OK;
}

extern "C" int lzgpu_plan_launch_count(const lzgpu_plan *p) { return p ? (int)p->launches.size() : 0; }

extern "C" int lzgpu_plan_launch(lzgpu_plan *p, const uint8_t *d_in, uint8_t *d_out, void *stream) {
    if (!p) return fail(LZGPU_E_INVALID, "plan_launch: null plan");
    DevState &ds = p->ctx->devs[p->dev_index];
    CUDA_TRY(cudaSetDevice(ds.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ds.stream;
    CUDA_TRY(cudaEventRecord(p->ev0, st));
    if (p->d_hout_base && p->d_next) CUDA_TRY(cudaMemsetAsync(p->d_next + 62, 0, 2 * sizeof(uint32_t), st));   // push-mode statistics (before the fork below)
    // A mixed batch has one launch per table class (literal-table size x posState-table size), and a launch lasts as
    // long as its longest unit: classes are independent, so they run side by side on auxiliary streams, forked from
    // and joined back into the caller's stream.  (Launches that share the HBM literal workspace stay on one stream.)
    const bool fork = p->launches.size() > 1;
    if (fork) {
        if (!ds.fork_ev) {
            CUDA_TRY(cudaEventCreateWithFlags(&ds.fork_ev, cudaEventDisableTiming));
            for (int k = 0; k < DevState::kAux; k++) {
                CUDA_TRY(cudaStreamCreateWithFlags(&ds.aux[k], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&ds.join_ev[k], cudaEventDisableTiming));
            }
        }
        CUDA_TRY(cudaEventRecord(ds.fork_ev, st));
        for (int k = 0; k < DevState::kAux; k++) CUDA_TRY(cudaStreamWaitEvent(ds.aux[k], ds.fork_ev, 0));
    }
    int rr = 1;
    for (const Launch &L : p->launches) {
        cudaStream_t ls = st;
        if (fork) { ls = L.lit_global ? ds.aux[0] : ds.aux[rr]; if (!L.lit_global) rr = rr % (DevState::kAux - 1) + 1; }
        KArgs a;
        a.units = p->d_units;
        a.order = p->d_order;
        a.in_base = d_in;
        a.out_base = d_out;
        a.results = p->d_results;
        a.lit_ws = p->d_lit_ws;
        a.lit_ws_stride = p->lit_ws_stride;
        a.lit_bits_cap = L.lit_bits;
        a.slot0 = L.slot0;
        a.stage_off = (uint32_t)probs_elems(L.lit_bits, L.lit_global, L.pb2);
        a.progress = p->d_progress;
        a.hout_base = p->d_hout_base;
        a.hout_off = p->d_hout_off;
        a.push_stat = p->d_next ? p->d_next + 62 : nullptr;   // the last two of the plan's 64 counters
        size_t smem = L.smem;
        if (p->max_ctas_per_sm >= 5) {
            // occupancy cap (experiments, and the single-wave heuristic of plan_create): asking for more shared memory
            // than a unit needs is how a launch of one-warp CTAs limits how many of them share an SM
            const size_t want = ((233472u / (unsigned)p->max_ctas_per_sm) - 1024u) & ~(size_t)15;
            if (want > smem && want <= 48u * 1024u) smem = want;
        }
        const size_t li = (size_t)(&L - p->launches.data());
        if (L.sm_slots && p->d_next) {
            SmArgs sa;
            sa.count = L.count;
            sa.n_slots = L.sm_slots;
            sa.slot_bytes = L.sm_slot_bytes;
            sa.save_off = L.sm_slot_bytes - kSmSaveBytes;
            sa.rotate_every = p->rotate_every;
            sa.rotate_mode = p->rotate_mode;
            sa.next = p->d_next + li;
            CUDA_TRY(cudaMemsetAsync(sa.next, 0, sizeof(uint32_t), ls));
            const size_t sm_bytes = kSmCtlBytes + (size_t)L.sm_slots * L.sm_slot_bytes;
            const unsigned threads = 32u * L.sm_slots;
            if (L.pb2) lzgpu_sm_kernel<33 | V_PB2><<<L.sm_grid, threads, sm_bytes, ls>>>(a, sa);
            else lzgpu_sm_kernel<33><<<L.sm_grid, threads, sm_bytes, ls>>>(a, sa);
        } else if (L.lit_global) launch_decode<true>(p->variant, false, L.count, smem, ls, a);
        else launch_decode<false>(p->variant, L.pb2, L.count, smem, ls, a);
        CUDA_TRY(cudaGetLastError());
    }
    if (fork)
        for (int k = 0; k < DevState::kAux; k++) {
            CUDA_TRY(cudaEventRecord(ds.join_ev[k], ds.aux[k]));
            CUDA_TRY(cudaStreamWaitEvent(st, ds.join_ev[k], 0));
        }
    CUDA_TRY(cudaEventRecord(p->ev1, st));
    p->last_stream = st;
    p->launched = true;
    return LZGPU_E_OK;
}

// grow-only device buffer: 0 ok, -1 out of memory
static int ensure(uint8_t *&ptr, uint64_t &cap, uint64_t need) {
    if (need <= cap) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    const uint64_t want = need + (need >> 3) + 4096;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    cap = want;
    return 0;
}

// ------------------------------------------------------------------ pinned buffers for callers without CUDA bindings
extern "C" void *lzgpu_alloc_pinned(uint64_t size) {
    if (lzgpu_device_count() <= 0) { fail(LZGPU_E_NO_DEVICE, "alloc_pinned: no CUDA device visible (" + g_last_error + ")"); return nullptr; }
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, size ? size : 1, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(e == cudaErrorMemoryAllocation ? LZGPU_E_NOMEM : LZGPU_E_CUDA, std::string("alloc_pinned: ") + cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
extern "C" void lzgpu_free_pinned(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
}

// ------------------------------------------------------------------ on-device verification
// CRC-32 (IEEE 802.3, reflected 0xEDB88320: zlib's crc32, the .xz CHECK_CRC32) of every unit's decoded bytes,
// so that a batch too large to copy back (BASELINE config 5: 64 GiB) is verified where it lies.
// One CTA per unit.  The unit's bytes are cut into 256 slices of equal length L, RIGHT-aligned (the leading
// slices may be short or empty); each thread runs slicing-by-4 over its slice, then the 256 partial CRCs are
// folded in a tree with  crc(A || B) = crc(A) * x^(8|B|) mod P  xor  crc(B)  (every right operand is full,
// so step k multiplies by the one power x^(8 L 2^k)).
namespace {
// W = uint32_t: CRC-32 (IEEE 802.3, reflected 0xEDB88320: zlib's crc32, the .xz CHECK_CRC32);
// W = uint64_t: CRC-64/XZ (ECMA-182, reflected 0xC96C5795D7870F42: the .xz CHECK_CRC64, xz's default).
template <typename W> struct CrcP;
template <> struct CrcP<uint32_t> { static constexpr uint32_t poly = 0xEDB88320u; };
template <> struct CrcP<uint64_t> { static constexpr uint64_t poly = 0xC96C5795D7870F42ull; };

template <typename W>
__host__ __device__ inline W crc_mulmod(W a, W b) {   // a * b mod P, reflected bit order (bit 0 = highest power)
    W m = (W)1 << (8 * sizeof(W) - 1), p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ CrcP<W>::poly : b >> 1;
    }
    return p;
}
template <typename W>
__host__ __device__ inline W crc_x_pow_8n(uint64_t n) {             // x^(8n) mod P
    const W x1 = (W)1 << (8 * sizeof(W) - 2);
    W sq = crc_mulmod<W>(x1, x1);                   // x^2
    sq = crc_mulmod<W>(sq, sq);                     // x^4
    sq = crc_mulmod<W>(sq, sq);                     // x^8
    W p = (W)1 << (8 * sizeof(W) - 1);              // x^0
    while (n) {
        if (n & 1) p = crc_mulmod<W>(sq, p);
        sq = crc_mulmod<W>(sq, sq);
        n >>= 1;
    }
    return p;
}
// crc(A || B) from crc(A), crc(B) and |B|: the pre- and post-inversions cancel except for the terms below
template <typename W>
inline W crc_combine(W a, W b, uint64_t len_b) {
    return len_b ? (W)(crc_mulmod<W>(crc_x_pow_8n<W>(len_b), a) ^ b) : a;
}

// One CTA per unit.  The unit's bytes are cut into 256 slices of equal length L, RIGHT-aligned (the leading slices may
// be short or empty); each thread runs slicing-by-4 over its slice, then the 256 partial CRCs are folded in a tree with
// crc(A || B) = crc(A) * x^(8|B|) mod P  xor  crc(B)  (every right operand is full, so step k multiplies by the one
// power x^(8 L 2^k)).  `want`: only units whose flags have this bit are summed (0: all); the others' slots are left alone.
template <typename W>
__global__ void __launch_bounds__(256) lzgpu_crc_kernel(const lzgpu_unit *units, const lzgpu_result *results,
                                                        const uint8_t *out_base, W *crc, int64_t n, uint32_t want) {
    __shared__ W T[4][256];
    __shared__ W part[256];
    __shared__ W powk[8];
    const uint32_t t = threadIdx.x;
    {   // slicing tables: T[0] the byte table, T[j][b] = T[j-1][b] advanced by one zero byte
        W c = t;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CrcP<W>::poly : c >> 1;
        T[0][t] = c;
        __syncthreads();
        for (int j = 1; j < 4; j++) { c = (c >> 8) ^ T[0][c & 0xFF]; T[j][t] = c; }
        __syncthreads();   // every warp reads every entry
    }
    for (int64_t ui = blockIdx.x; ui < n; ui += gridDim.x) {
        if (want && !(units[ui].flags & want)) continue;   // uniform over the CTA
        const uint64_t total = results[ui].status == LZGPU_NOT_RUN ? 0 : results[ui].bytes_out;
        const uint8_t *base = out_base + units[ui].out_off;
        const uint64_t L = ((total + 255) / 256 + 3) & ~(uint64_t)3;
        if (t == 0) {
            W pw = crc_x_pow_8n<W>(L);
            for (int k = 0; k < 8; k++) { powk[k] = pw; pw = crc_mulmod<W>(pw, pw); }
        }
        // slice t = [total - (256 - t) L, total - (255 - t) L) clipped at 0
        const uint64_t back_hi = (uint64_t)(256 - t) * L, back_lo = (uint64_t)(255 - t) * L;
        const uint64_t lo = back_hi >= total ? 0 : total - back_hi, hi = back_lo >= total ? 0 : total - back_lo;
        W c = 0;
        if (hi > lo) {
            const uint8_t *p = base + lo, *e = base + hi;
            c = ~(W)0;
            while (p < e && ((uintptr_t)p & 15)) c = (c >> 8) ^ T[0][(c ^ *p++) & 0xFF];
            for (; p + 16 <= e; p += 16) {
                const uint4 w = *reinterpret_cast<const uint4 *>(p);
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    c ^= ws[q];
                    const uint32_t l = (uint32_t)c;
                    W nx = T[3][l & 0xFF] ^ T[2][(l >> 8) & 0xFF] ^ T[1][(l >> 16) & 0xFF] ^ T[0][l >> 24];
                    if (sizeof(W) == 8) nx ^= (W)((uint64_t)c >> 32);
                    c = nx;
                }
            }
            while (p < e) c = (c >> 8) ^ T[0][(c ^ *p++) & 0xFF];
            c = ~c;
        }
        part[t] = c;
        __syncthreads();
        for (int k = 0; k < 8; k++) {
            const uint32_t step = 1u << k;
            if ((t & (2 * step - 1)) == 0) part[t] = crc_mulmod<W>(powk[k], part[t]) ^ part[t + step];
            __syncthreads();
        }
        if (t == 0) crc[ui] = part[0];
        __syncthreads();
    }
}

// checksums of a launched plan's units into host memory; stream = the plan's
template <typename W>
int plan_crc(lzgpu_plan *p, const uint8_t *d_out, W *crc, const char *what) {
    if (!p || !p->launched || (p->n > 0 && (!d_out || !crc))) return fail(LZGPU_E_INVALID, std::string(what) + ": plan was not launched / null argument");
    if (p->n == 0) return LZGPU_E_OK;
    DevState &ds = p->ctx->devs[p->dev_index];
    CUDA_TRY(cudaSetDevice(ds.device));
    if (ensure(ds.d_sum, ds.sum_cap, sizeof(W) * (uint64_t)p->n)) return fail(LZGPU_E_NOMEM, std::string(what) + ": cudaMalloc of the checksum array failed");
    W *d_crc = reinterpret_cast<W *>(ds.d_sum);
    const unsigned grid = (unsigned)std::min<int64_t>(p->n, 148 * 16);
    lzgpu_crc_kernel<W><<<grid, 256, 0, p->last_stream>>>(p->d_units, p->d_results, d_out, d_crc, p->n, 0u);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(crc, d_crc, sizeof(W) * (size_t)p->n, cudaMemcpyDeviceToHost, p->last_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->last_stream);
    if (e != cudaSuccess) return fail(LZGPU_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return LZGPU_E_OK;
}
}  // namespace

extern "C" int lzgpu_plan_crc32(lzgpu_plan *p, const uint8_t *d_out, uint32_t *crc) { return plan_crc<uint32_t>(p, d_out, crc, "plan_crc32"); }
extern "C" int lzgpu_plan_crc64(lzgpu_plan *p, const uint8_t *d_out, uint64_t *crc) { return plan_crc<uint64_t>(p, d_out, crc, "plan_crc64"); }
extern "C" uint32_t lzgpu_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) { return crc_combine<uint32_t>(crc_a, crc_b, len_b); }
extern "C" uint64_t lzgpu_crc64_combine(uint64_t crc_a, uint64_t crc_b, uint64_t len_b) { return crc_combine<uint64_t>(crc_a, crc_b, len_b); }

extern "C" int lzgpu_plan_results(lzgpu_plan *p, lzgpu_result *results, lzgpu_stats *stats) {
    if (!p || !p->launched) return fail(LZGPU_E_INVALID, "plan_results: plan was not launched");
    DevState &ds = p->ctx->devs[p->dev_index];
    CUDA_TRY(cudaSetDevice(ds.device));
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    std::vector<lzgpu_result> tmp((size_t)p->n);
    if (p->n) CUDA_TRY(cudaMemcpy(tmp.data(), p->d_results, sizeof(lzgpu_result) * (size_t)p->n, cudaMemcpyDeviceToHost));
    uint64_t bi = 0, bo = 0;
    for (int64_t i = 0; i < p->n; i++) {
        lzgpu_result r = tmp[(size_t)i];
        r.device = ds.device;
        if (p->preset[(size_t)i].status != LZGPU_NOT_RUN) r = p->preset[(size_t)i];
        else if (p->was_alone[(size_t)i]) r.bytes_in += 13;   // the .lzma header consumed on the host
        bi += r.bytes_in;
        bo += r.bytes_out;
        if (results) results[i] = r;
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->kernel_ms = ms;
        stats->bytes_in = bi;
        stats->bytes_out = bo;
        stats->launches = (int32_t)p->launches.size();
        stats->devices = 1;
    }
    return LZGPU_E_OK;
}

// ------------------------------------------------------------------ host-buffer batch
namespace {

struct Shard {
    std::vector<int64_t> idx;          // unit indices (caller's numbering)
    std::vector<lzgpu_unit> units;     // rebased into the shard's device slabs
    struct Run { uint64_t host_off, dev_off, len; };
    std::vector<Run> in_runs, out_runs;
    uint64_t in_bytes = 0, out_bytes = 0;
    int rc = 0;
    std::string err;
    double kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
    int launches = 0, launches_extra = 0;
};

// Coalesce [off, off+len) ranges (sorted by off) into few large copies; returns the
// device offset of each range.
// `gap`: ranges closer than this are merged into one copy.  Output ranges are merged only when they touch
// (gap 0): a D2H copy of a merged run writes every byte between its ends, and bytes of the caller's buffer
// that belong to no unit of this shard (padding, or another GPU's units) must stay untouched.
void layout_ranges(const std::vector<std::pair<uint64_t, uint64_t>> &ranges, std::vector<uint64_t> &dev_off,
                   std::vector<Shard::Run> &runs, uint64_t &total, uint64_t gap) {
    const size_t n = ranges.size();
    std::vector<size_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return ranges[a].first < ranges[b].first; });
    dev_off.assign(n, 0);
    total = 0;
    for (size_t k = 0; k < n; k++) {
        const auto &r = ranges[ord[k]];
        if (!runs.empty()) {
            Shard::Run &last = runs.back();
            const uint64_t last_end = last.host_off + last.len;
            if (r.first <= last_end + gap) {
                const uint64_t new_end = std::max(last_end, r.first + r.second);
                dev_off[ord[k]] = last.dev_off + (r.first - last.host_off);
                last.len = new_end - last.host_off;
                total = last.dev_off + last.len;
                continue;
            }
        }
        // new run; keep host and device offsets congruent mod 16 so 4-byte alignment is preserved
        uint64_t d = (total + 255) & ~(uint64_t)255;
        d += r.first & 15;
        runs.push_back({r.first, d, r.second});
        dev_off[ord[k]] = d;
        total = d + r.second;
    }
}

// Pinned (device-accessible) host memory under unified addressing: the device pointer of [p, p + len), else null.
const uint8_t *device_view_of_host(const uint8_t *p, uint64_t len) {
    if (!p || !len) return nullptr;
    cudaPointerAttributes a, b;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || cudaPointerGetAttributes(&b, p + len - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer || !b.devicePointer) return nullptr;
    if ((const uint8_t *)b.devicePointer - (const uint8_t *)a.devicePointer != (ptrdiff_t)(len - 1)) return nullptr;
    return (const uint8_t *)a.devicePointer;
}

__global__ void lzgpu_tail_copy_kernel(const uint64_t *desc, const uint8_t *src_base, uint8_t *dst_base) {
    // desc[3k..3k+2] = source offset, destination offset, length of unit k's tail; grid = (units, parts)
    const uint64_t so = desc[3 * blockIdx.x], dof = desc[3 * blockIdx.x + 1], len = desc[3 * blockIdx.x + 2];
    const uint8_t *s = src_base + so;
    uint8_t *d = dst_base + dof;
    const uint64_t tid = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.y * blockDim.x;
    if ((((uintptr_t)s ^ (uintptr_t)d) & 15) == 0) {
        const uint64_t head = min(len, (uint64_t)((16 - ((uintptr_t)s & 15)) & 15));
        const uint64_t nv = (len - head) >> 4;
        for (uint64_t i = tid; i < head; i += nth) d[i] = s[i];
        const uint4 *sv = reinterpret_cast<const uint4 *>(s + head);
        uint4 *dv = reinterpret_cast<uint4 *>(d + head);
        for (uint64_t i = tid; i < nv; i += nth) dv[i] = sv[i];
        for (uint64_t i = head + (nv << 4) + tid; i < len; i += nth) d[i] = s[i];
    } else {
        for (uint64_t i = tid; i < len; i += nth) d[i] = s[i];
    }
}

void run_shard(lzgpu_ctx *ctx, int dev_index, Shard &sh, const uint8_t *in_base, uint64_t in_size, uint8_t *out_base,
               uint64_t out_size, lzgpu_result *results, uint64_t *sums) {
    DevState &ds = ctx->devs[dev_index];
    // LZGPU_TRACE=1: host-side timeline of this shard on stderr (ms since the shard started)
    static const bool trace = getenv("LZGPU_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (trace) fprintf(stderr, "[lzgpu dev %d] %8.3f ms  %s\n", ds.device,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), what);
    };
    auto cuda_fail = [&](cudaError_t e, const char *what) {
        sh.rc = e == cudaErrorMemoryAllocation ? LZGPU_E_NOMEM : LZGPU_E_CUDA;
        sh.err = std::string(what) + ": " + cudaGetErrorString(e);
    };
    cudaError_t e = cudaSetDevice(ds.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t n = sh.idx.size();
    if (n == 0) return;
    // device layout
    std::vector<std::pair<uint64_t, uint64_t>> ir(n), orr(n);
    for (size_t k = 0; k < n; k++) {
        ir[k] = {sh.units[k].in_off, sh.units[k].in_len};
        orr[k] = {sh.units[k].out_off, sh.units[k].out_cap};
    }
    std::vector<uint64_t> ioff, ooff, host_out_off(n);
    for (size_t k = 0; k < n; k++) host_out_off[k] = sh.units[k].out_off;
    // Compressed input in pinned host memory is read by the units straight over PCIe (each byte is read
    // once, 512 bytes at a time, at ~1/400 of the link's rate): no H2D copy ahead of the kernel.  Pageable
    // input is packed into a device slab first.  (LZGPU_NO_ZEROCOPY_IN=1 forces the slab.)
    const uint8_t *zc_in = getenv("LZGPU_NO_ZEROCOPY_IN") ? nullptr : device_view_of_host(in_base, in_size);
    uint8_t *zc_out = getenv("LZGPU_NO_TAIL_KERNEL") ? nullptr : const_cast<uint8_t *>(device_view_of_host(out_base, out_size));
    if (zc_in) sh.in_bytes = in_size;
    else layout_ranges(ir, ioff, sh.in_runs, sh.in_bytes, 64 << 10);
    layout_ranges(orr, ooff, sh.out_runs, sh.out_bytes, 0);
    for (size_t k = 0; k < n; k++) { if (!zc_in) sh.units[k].in_off = ioff[k]; sh.units[k].out_off = ooff[k]; }
    if ((!zc_in && ensure(ds.d_in, ds.in_cap, sh.in_bytes + 16)) || ensure(ds.d_out, ds.out_cap, sh.out_bytes + 16)) {
        sh.rc = LZGPU_E_NOMEM;
        sh.err = "cudaMalloc of the shard's input/output slabs failed";
        return;
    }
    lzgpu_plan *plan = nullptr;
    int rc = plan_create_impl(ctx, dev_index, sh.units.data(), (int64_t)n, zc_in ? in_size : sh.in_bytes + 16, sh.out_bytes + 16, &plan, true);
    if (rc != LZGPU_E_OK) { sh.rc = rc; sh.err = g_last_error; return; }
    mark("plan created");
    // Streamed D2H for large shards: the kernel publishes, per unit, how many 64 KiB blocks of its output
    // are final (host-mapped counters); this thread polls them while the kernel runs and sends finished
    // blocks to the caller's buffer on a second stream, so that only each unit's tail is left to copy
    // when the kernel ends.  (LZGPU_NO_STREAM_D2H=1 restores kernel-then-copy.)
    const uint64_t kBlock = 64 << 10;
    // Push mode: when the caller's output buffer is pinned and mapped, every unit writes its decoded bytes there itself,
    // over PCIe, 64 KiB block by block as they become final and its tail when it ends (push_out, lzgpu_unit.cuh): no
    // D2H copy, no polling thread, nothing left to do when the kernel ends.  (LZGPU_NO_PUSH_D2H=1: the streamed copies.)
    // Push mode holds the decoding warps up when the host cannot take the bytes as fast as they come (measured: eight
    // GPUs of one node, 78 GB/s of output against the ~69 GB/s the host ingests: 135 ms per step where the copy engines
    // need 125).  The units time their block pushes; a shard that needed more than kPushSlowKc kilo-cycles per 64 KiB
    // block (one GPU: ~130) makes this device use the streamed copies, and push mode is probed again every 64th call.
    uint32_t kPushSlowKc = 600;
    if (const char *e = getenv("LZGPU_PUSH_SLOW_KC")) kPushSlowKc = (uint32_t)atoi(e);   // (tests: 0 = every shard counts as slow)
    bool push = zc_out != nullptr && !getenv("LZGPU_NO_PUSH_D2H");
    if (push && ds.push_slow && !getenv("LZGPU_PUSH_D2H")) {
        if (++ds.push_skipped < 64) push = false;
        else ds.push_skipped = 0;
    }
    bool stream_out = !push && sh.out_bytes >= ((uint64_t)32 << 20) && !getenv("LZGPU_NO_STREAM_D2H");
    if (stream_out || push) {
        if (stream_out && !ds.copy_stream && cudaStreamCreateWithFlags(&ds.copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); stream_out = false; }
        if ((stream_out || push) && ds.progress_cap < n) {
            if (ds.h_progress) cudaFreeHost(ds.h_progress);
            if (ds.h_tails) cudaFreeHost(ds.h_tails);
            ds.h_progress = nullptr; ds.d_progress = nullptr; ds.h_tails = nullptr; ds.d_tails = nullptr; ds.progress_cap = 0;
            const size_t want = n + (n >> 2) + 64;
            if (cudaHostAlloc(&ds.h_progress, want * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer(&ds.d_progress, ds.h_progress, 0) != cudaSuccess ||
                cudaHostAlloc(&ds.h_tails, want * 3 * sizeof(uint64_t), cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer(&ds.d_tails, ds.h_tails, 0) != cudaSuccess) {
                cudaGetLastError();
                if (ds.h_progress) cudaFreeHost(ds.h_progress);
                if (ds.h_tails) cudaFreeHost(ds.h_tails);
                ds.h_progress = nullptr; ds.d_progress = nullptr; ds.h_tails = nullptr; ds.d_tails = nullptr;
                stream_out = false;
                push = false;
            } else {
                ds.progress_cap = want;
            }
        }
        if (stream_out) {
            memset(ds.h_progress, 0, n * sizeof(uint32_t));
            plan->d_progress = ds.d_progress;
        }
        if (push) {
            for (size_t k = 0; k < n; k++) ds.h_tails[k] = host_out_off[k];
            plan->d_hout_base = zc_out;
            plan->d_hout_off = ds.d_tails;
        }
    }
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    cudaEventRecord(e0, ds.stream);
    for (const auto &r : sh.in_runs) {
        e = cudaMemcpyAsync(ds.d_in + r.dev_off, in_base + r.host_off, r.len, cudaMemcpyHostToDevice, ds.stream);
        if (e != cudaSuccess) { cuda_fail(e, "H2D"); break; }
    }
    cudaEventRecord(e1, ds.stream);
    if (sh.rc == 0) {
        rc = lzgpu_plan_launch(plan, zc_in ? zc_in : ds.d_in, ds.d_out, ds.stream);
        if (rc != LZGPU_E_OK) { sh.rc = rc; sh.err = g_last_error; }
    }
    cudaEventRecord(e2, ds.stream);
    mark("kernel enqueued");
    // Only the bytes a unit decoded go back to the caller (not its whole capacity: what lies behind bytes_out
    // in the device slab is left over from earlier batches), so the final copies wait for the results.
    std::vector<uint32_t> copied(n, 0);      // streamed: 64 KiB blocks of unit k already sent
    auto send = [&](size_t k, uint64_t from, uint64_t to) {   // bytes [from, to) of unit k's output
        if (to <= from || sh.rc != 0) return;
        cudaError_t ce = cudaMemcpyAsync(out_base + host_out_off[k] + from, ds.d_out + sh.units[k].out_off + from, to - from,
                                         cudaMemcpyDeviceToHost, ds.copy_stream);
        if (ce != cudaSuccess) cuda_fail(ce, "D2H (streamed)");
    };
    if (sh.rc == 0 && stream_out) {
        volatile const uint32_t *prog = ds.h_progress;
        int idle = 0;
        for (;;) {
            const cudaError_t q = cudaEventQuery(e2);
            if (q != cudaErrorNotReady) { if (q != cudaSuccess) cuda_fail(q, "decode kernel"); break; }
            bool progressed = false;
            for (size_t k = 0; k < n && sh.rc == 0; k++) {
                const uint32_t have = prog[k];
                if (have > copied[k]) {
                    const uint64_t cap = sh.units[k].out_cap;
                    send(k, std::min<uint64_t>(copied[k] * kBlock, cap), std::min<uint64_t>(have * kBlock, cap));
                    copied[k] = have;
                    progressed = true;
                }
            }
            if (sh.rc != 0) break;
            // a 1 MiB unit finishes a block every ~7 ms: nothing is lost by sleeping between scans, and the
            // host core is free for the caller (or for the other GPUs' shard threads)
            if (progressed) idle = 0;
            else if (++idle > 2) std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
    }
    std::vector<lzgpu_result> tmp(n);
    lzgpu_stats st;
    memset(&st, 0, sizeof st);
    if (sh.rc == 0) {
        rc = lzgpu_plan_results(plan, tmp.data(), &st);   // waits for the kernel
        if (rc != LZGPU_E_OK) { sh.rc = rc; sh.err = g_last_error; }
    }
    if (sh.rc == 0 && push && plan->d_next) {
        uint32_t ps[2] = {0, 0};
        if (cudaMemcpy(ps, plan->d_next + 62, sizeof ps, cudaMemcpyDeviceToHost) == cudaSuccess && ps[1] >= 32u * 64u) {
            const uint32_t kc_per_block = ps[0] / ps[1];
            ds.push_slow = kc_per_block > kPushSlowKc;
            if (trace) fprintf(stderr, "[lzgpu dev %d] push mode: %u kilo-cycles per 64 KiB block over %u blocks%s\n", ds.device,
                               kc_per_block, ps[1] / 32u, ds.push_slow ? " -> streamed copies from the next call on" : "");
        } else cudaGetLastError();
    }
    if (sh.rc == 0 && stream_out) {
        if (zc_out) {
            // the tails, by ONE kernel that writes the caller's (pinned) buffer over PCIe: a thousand small
            // cudaMemcpyAsync calls cost more host time than the bytes take to move
            for (size_t k = 0; k < n; k++) {
                const uint64_t done = std::min<uint64_t>(tmp[k].bytes_out, sh.units[k].out_cap), from = std::min<uint64_t>(copied[k] * kBlock, done);
                ds.h_tails[3 * k] = sh.units[k].out_off + from;
                ds.h_tails[3 * k + 1] = host_out_off[k] + from;
                ds.h_tails[3 * k + 2] = done - from;
            }
            lzgpu_tail_copy_kernel<<<dim3((unsigned)n, 2), 256, 0, ds.copy_stream>>>(ds.d_tails, ds.d_out, zc_out);
            const cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) cuda_fail(ce, "tail copy kernel");
        } else {
            for (size_t k = 0; k < n && sh.rc == 0; k++) {
                const uint64_t done = std::min<uint64_t>(tmp[k].bytes_out, sh.units[k].out_cap);
                send(k, std::min<uint64_t>(copied[k] * kBlock, done), done);
            }
        }
        mark("kernel finished, tails enqueued");
        cudaEventRecord(e3, ds.copy_stream);
    } else {
        if (sh.rc == 0 && !push) {
            // one copy per run of units that are adjacent in the caller's buffer and filled to their capacity
            std::vector<size_t> ord(n);
            std::iota(ord.begin(), ord.end(), 0);
            std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return host_out_off[a] < host_out_off[b]; });
            uint64_t h0 = 0, d0 = 0, len = 0;
            auto flush = [&]() {
                if (len && sh.rc == 0) {
                    e = cudaMemcpyAsync(out_base + h0, ds.d_out + d0, len, cudaMemcpyDeviceToHost, ds.stream);
                    if (e != cudaSuccess) cuda_fail(e, "D2H");
                }
                len = 0;
            };
            for (size_t q = 0; q < n; q++) {
                const size_t k = ord[q];
                const uint64_t done = std::min<uint64_t>(tmp[k].bytes_out, sh.units[k].out_cap);
                if (len && host_out_off[k] == h0 + len && sh.units[k].out_off == d0 + len) len += done;
                else { flush(); h0 = host_out_off[k]; d0 = sh.units[k].out_off; len = done; }
                if (done < sh.units[k].out_cap) flush();
            }
            flush();
        }
        cudaEventRecord(e3, ds.stream);
    }
    if (sh.rc == 0 && sums) {
        // checksums of the decoded bytes where they lie (LZGPU_UF_SUM_*): the units' CRCs are computed by the GPU
        // while the tails / the output travel back on the copy stream -- no host pass over the payload
        uint32_t any = 0;
        for (size_t k = 0; k < n; k++) any |= sh.units[k].flags;
        if (any & (LZGPU_UF_SUM_CRC32 | LZGPU_UF_SUM_CRC64)) {
            const uint64_t need = (sizeof(uint64_t) + sizeof(uint32_t)) * (uint64_t)n;
            if (ensure(ds.d_sum, ds.sum_cap, need)) { sh.rc = LZGPU_E_NOMEM; sh.err = "cudaMalloc of the checksum array failed"; }
            else {
                uint64_t *d64 = reinterpret_cast<uint64_t *>(ds.d_sum);
                uint32_t *d32 = reinterpret_cast<uint32_t *>(ds.d_sum + sizeof(uint64_t) * n);
                const unsigned grid = (unsigned)std::min<size_t>(n, 148 * 16);
                if (any & LZGPU_UF_SUM_CRC32) lzgpu_crc_kernel<uint32_t><<<grid, 256, 0, ds.stream>>>(plan->d_units, plan->d_results, ds.d_out, d32, (int64_t)n, LZGPU_UF_SUM_CRC32);
                if (any & LZGPU_UF_SUM_CRC64) lzgpu_crc_kernel<uint64_t><<<grid, 256, 0, ds.stream>>>(plan->d_units, plan->d_results, ds.d_out, d64, (int64_t)n, LZGPU_UF_SUM_CRC64);
                std::vector<uint8_t> h(need);
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), ds.d_sum, need, cudaMemcpyDeviceToHost, ds.stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(ds.stream);
                if (e != cudaSuccess) cuda_fail(e, "checksum kernels");
                else {
                    const uint64_t *h64 = reinterpret_cast<const uint64_t *>(h.data());
                    const uint32_t *h32 = reinterpret_cast<const uint32_t *>(h.data() + sizeof(uint64_t) * n);
                    for (size_t k = 0; k < n; k++)
                        sums[sh.idx[k]] = (sh.units[k].flags & LZGPU_UF_SUM_CRC64) ? h64[k] : (sh.units[k].flags & LZGPU_UF_SUM_CRC32) ? h32[k] : 0;
                    sh.launches_extra = ((any & LZGPU_UF_SUM_CRC32) ? 1 : 0) + ((any & LZGPU_UF_SUM_CRC64) ? 1 : 0);
                }
            }
        }
    }
    if (stream_out) {
        e = cudaStreamSynchronize(ds.copy_stream);
        if (e != cudaSuccess && sh.rc == 0) cuda_fail(e, "D2H (streamed) sync");
    }
    e = cudaStreamSynchronize(ds.stream);
    if (e != cudaSuccess && sh.rc == 0) cuda_fail(e, "decode kernel / stream sync");
    mark("streams idle");
    if (sh.rc == 0) {
        for (size_t k = 0; k < n; k++) results[sh.idx[k]] = tmp[k];
        sh.kernel_ms = st.kernel_ms;
        sh.launches = st.launches + sh.launches_extra;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e2, e3);
        sh.h2d_ms = a;
        sh.d2h_ms = b;
    }
    mark("results read");
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    lzgpu_plan_destroy(plan);
    mark("plan destroyed");
}

}  // namespace

static int decode_batch_impl(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n, const uint8_t *in_base, uint64_t in_size,
                             uint8_t *out_base, uint64_t out_size, lzgpu_result *results, lzgpu_stats *stats, uint64_t *sums);
extern "C" int lzgpu_decode_batch(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n,
                                  const uint8_t *in_base, uint64_t in_size,
                                  uint8_t *out_base, uint64_t out_size,
                                  lzgpu_result *results, lzgpu_stats *stats) {
    return decode_batch_impl(ctx, units, n, in_base, in_size, out_base, out_size, results, stats, nullptr);
}
extern "C" int lzgpu_decode_batch_sums(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n,
                                       const uint8_t *in_base, uint64_t in_size,
                                       uint8_t *out_base, uint64_t out_size,
                                       lzgpu_result *results, lzgpu_stats *stats, uint64_t *sums) {
    if (n > 0 && !sums) return fail(LZGPU_E_INVALID, "decode_batch_sums: null sums");
    for (int64_t i = 0; i < n; i++) sums[i] = 0;
    return decode_batch_impl(ctx, units, n, in_base, in_size, out_base, out_size, results, stats, sums);
}
static int decode_batch_impl(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n, const uint8_t *in_base, uint64_t in_size,
                             uint8_t *out_base, uint64_t out_size, lzgpu_result *results, lzgpu_stats *stats, uint64_t *sums) {
    if (!ctx || n < 0 || (n > 0 && (!units || !results))) return fail(LZGPU_E_INVALID, "decode_batch: bad arguments");
    if (ctx->devs.empty()) return fail(LZGPU_E_NO_DEVICE, "decode_batch: context has no device");
    std::lock_guard<std::mutex> lock(ctx->mu);
    const auto t0 = std::chrono::steady_clock::now();
    for (int64_t i = 0; i < n; i++) {
        const lzgpu_unit &u = units[i];
        if (u.in_off > in_size || u.in_len > in_size - u.in_off || u.out_off > out_size || u.out_cap > out_size - u.out_off)
            return fail(LZGPU_E_INVALID, "decode_batch: unit " + std::to_string(i) + " lies outside the buffers");
        memset(&results[i], 0, sizeof results[i]);
        results[i].status = LZGPU_NOT_RUN;
    }
    const int nd = (int)ctx->devs.size();
    std::vector<int32_t> shard_of((size_t)n);
    int rc = lzgpu_shard_units(units, n, nd, shard_of.data());
    if (rc != LZGPU_E_OK) return rc;
    std::vector<Shard> shards((size_t)nd);
    for (int64_t i = 0; i < n; i++) {
        lzgpu_unit u = units[i];
        if (u.kind == LZGPU_KIND_LZMA1_ALONE) {
            // NewReader1 reads the 13-byte header eagerly (reader1.go:77-101)
            const int hs = lzgpu_parse_alone_header(in_base + u.in_off, u.in_len, &u);
            if (hs != LZGPU_OK) { results[i].status = hs; results[i].device = -1; continue; }
        } else if (u.kind == LZGPU_KIND_LZMA2_GROUP && !(u.flags & LZGPU_UF_BITS_KNOWN)) {
            // table sizes from the chunk headers themselves: a binding that does not carry lit_bits / pos_bits
            // through (they are not part of any reference interface) cannot make the unit fail
            derive_lzma2_bits(in_base, u);
        }
        Shard &s = shards[(size_t)shard_of[(size_t)i]];
        s.idx.push_back(i);
        s.units.push_back(u);
    }
    if (nd == 1) {
        run_shard(ctx, 0, shards[0], in_base, in_size, out_base, out_size, results, sums);
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; d++)
            th.emplace_back([&, d]() { run_shard(ctx, d, shards[(size_t)d], in_base, in_size, out_base, out_size, results, sums); });
        for (auto &t : th) t.join();
    }
    lzgpu_stats st;
    memset(&st, 0, sizeof st);
    for (int d = 0; d < nd; d++) {
        if (shards[(size_t)d].rc != 0) return fail(shards[(size_t)d].rc, "device " + std::to_string(ctx->devs[(size_t)d].device) + ": " + shards[(size_t)d].err);
        st.kernel_ms = std::max(st.kernel_ms, shards[(size_t)d].kernel_ms);
        st.h2d_ms = std::max(st.h2d_ms, shards[(size_t)d].h2d_ms);
        st.d2h_ms = std::max(st.d2h_ms, shards[(size_t)d].d2h_ms);
        st.launches += shards[(size_t)d].launches;
    }
    for (int64_t i = 0; i < n; i++) {
        st.bytes_in += results[i].bytes_in;
        st.bytes_out += results[i].bytes_out;
    }
    st.devices = nd;
    st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = st;
    return LZGPU_E_OK;
}
