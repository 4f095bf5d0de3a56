/*
 * tests/enc/lzma_test_encoder.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference package is decoder-only and liblzma refuses lc+lp > 4, odd dictionary
 * sizes, "known size without EOS marker" and a match at stream position 0.  This small
 * greedy LZMA1 encoder writes such streams so that the oracle and the CUDA path can be
 * compared on them (every prop byte < 225 is legal for the reference: reader1.go:210-221).
 * Format written: the classic LZMA bit stream (same model the decoder restates from
 * decompress.go): isMatch / isRep / isRepG0-2 / isRep0Long, length coder (choice, choice2,
 * low/mid/high), posSlot + reverse-tree / direct bits + align, 12-state machine.
 * Parsing: rep0..3 first, then a hash-chain match finder, greedy.  Not a compressor to be
 * proud of -- a generator of valid, varied streams.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint16_t prob;
#define PROB_INIT 1024
#define TOP (1u << 24)

typedef struct {
    uint64_t low;
    uint32_t range;
    uint8_t cache;
    uint64_t cache_size;
    uint8_t *out;
    size_t pos, cap;
    int overflow;
} renc;

static void put(renc *r, uint8_t b) { if (r->pos < r->cap) r->out[r->pos] = b; else r->overflow = 1; r->pos++; }
static void rc_init(renc *r, uint8_t *out, size_t cap) { memset(r, 0, sizeof *r); r->range = 0xFFFFFFFFu; r->cache_size = 1; r->out = out; r->cap = cap; }
static void shift_low(renc *r) {
    if ((uint32_t)r->low < 0xFF000000u || (int)(r->low >> 32) != 0) {
        uint8_t t = r->cache;
        do { put(r, (uint8_t)(t + (uint8_t)(r->low >> 32))); t = 0xFF; } while (--r->cache_size != 0);
        r->cache = (uint8_t)((uint32_t)r->low >> 24);
    }
    r->cache_size++;
    r->low = (uint32_t)r->low << 8;
}
static void rc_bit(renc *r, prob *p, int bit) {
    uint32_t bound = (r->range >> 11) * *p;
    if (!bit) { r->range = bound; *p += (2048 - *p) >> 5; }
    else { r->low += bound; r->range -= bound; *p -= *p >> 5; }
    while (r->range < TOP) { r->range <<= 8; shift_low(r); }
}
static void rc_direct(renc *r, uint32_t v, int nbits) {
    for (int i = nbits - 1; i >= 0; i--) {
        r->range >>= 1;
        if ((v >> i) & 1) r->low += r->range;
        while (r->range < TOP) { r->range <<= 8; shift_low(r); }
    }
}
static void rc_flush(renc *r) { for (int i = 0; i < 5; i++) shift_low(r); }
static void tree(renc *r, prob *p, int nbits, uint32_t v) {
    uint32_t m = 1;
    for (int i = nbits - 1; i >= 0; i--) { int b = (v >> i) & 1; rc_bit(r, &p[m], b); m = (m << 1) | b; }
}
static void tree_rev(renc *r, prob *p, int nbits, uint32_t v) {
    uint32_t m = 1;
    for (int i = 0; i < nbits; i++) { int b = v & 1; v >>= 1; rc_bit(r, &p[m], b); m = (m << 1) | b; }
}

typedef struct { prob choice, choice2, low[16][8], mid[16][8], high[256]; } lenc;
static void fillp(prob *p, size_t n) { for (size_t i = 0; i < n; i++) p[i] = PROB_INIT; }
static void len_init(lenc *l) { l->choice = l->choice2 = PROB_INIT; fillp(&l->low[0][0], 128); fillp(&l->mid[0][0], 128); fillp(l->high, 256); }
static void len_enc(renc *r, lenc *l, uint32_t len /* 0..271 */, uint32_t ps) {
    if (len < 8) { rc_bit(r, &l->choice, 0); tree(r, l->low[ps], 3, len); }
    else { rc_bit(r, &l->choice, 1);
        if (len < 16) { rc_bit(r, &l->choice2, 0); tree(r, l->mid[ps], 3, len - 8); }
        else { rc_bit(r, &l->choice2, 1); tree(r, l->high, 8, len - 16); } }
}

typedef struct {
    prob *lit;
    prob is_match[192], is_rep[12], g0[12], g1[12], g2[12], rep0long[192];
    prob pos_slot[4][64], pos_dec[115], align[16];
    lenc len, rep_len;
    uint32_t rep[4], state;
    int lc, lp, pb;
} model;

static uint32_t slot_of(uint32_t d) { /* d = distance - 1 */
    if (d < 4) return d;
    uint32_t n = 31 - (uint32_t)__builtin_clz(d);
    return (n << 1) + ((d >> (n - 1)) & 1);
}
static void enc_dist(renc *r, model *m, uint32_t d, uint32_t len /* 0-based */) {
    uint32_t ls = len > 3 ? 3 : len, slot = slot_of(d);
    tree(r, m->pos_slot[ls], 6, slot);
    if (slot >= 4) {
        uint32_t nd = (slot >> 1) - 1, base = (2 | (slot & 1)) << nd, rest = d - base;
        if (slot < 14) tree_rev(r, &m->pos_dec[base - slot], (int)nd, rest);
        else { rc_direct(r, rest >> 4, (int)nd - 4); tree_rev(r, m->align, 4, rest & 15); }
    }
}
static void enc_literal(renc *r, model *m, const uint8_t *data, uint64_t pos, uint32_t wpos) {
    uint32_t prev = pos ? data[pos - 1] : 0, sym = data[pos];
    prob *p = m->lit + 0x300u * (((wpos & ((1u << m->lp) - 1)) << m->lc) + (prev >> (8 - m->lc)));
    uint32_t s = 1;
    if (m->state >= 7) {
        uint32_t mb = data[pos - m->rep[0] - 1], offs = 0x100, c = sym;
        for (int i = 0; i < 8; i++) {
            mb <<= 1; c <<= 1;
            uint32_t mbit = mb & offs, bit = (c >> 8) & 1;
            rc_bit(r, &p[offs + mbit + s], (int)bit);
            s = (s << 1) | bit;
            offs &= bit ? mbit : ~mbit;
        }
    } else {
        for (int i = 7; i >= 0; i--) { int b = (sym >> i) & 1; rc_bit(r, &p[s], b); s = (s << 1) | (uint32_t)b; }
    }
    m->state = m->state < 4 ? 0 : (m->state < 10 ? m->state - 3 : m->state - 6);
}

/*
 * flags: 1 = write EOS marker, 2 = header carries the size (else all-ones),
 *        4 = start with a match of distance 1 at position 0 when the data begins with zeros (Q4),
 *        8 = raw (no 13-byte header)
 * Returns bytes written, 0 on overflow/alloc failure.
 */
size_t lzma_test_encode(const uint8_t *data, uint64_t n, int lc, int lp, int pb, uint32_t dict_size,
                        int flags, uint8_t *out, size_t out_cap) {
    model *m = (model *)calloc(1, sizeof(model));
    renc r;
    size_t lit_n = (size_t)0x300 << (lc + lp), hdr = (flags & 8) ? 0 : 13;
    if (!m || out_cap < hdr + 16) { free(m); return 0; }
    m->lit = (prob *)malloc(lit_n * sizeof(prob));
    const uint32_t HB = 16;
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) << HB), *chain = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    if (!m->lit || !head || !chain) { free(m->lit); free(head); free(chain); free(m); return 0; }
    memset(head, 0xFF, sizeof(int32_t) << HB);
    fillp(m->lit, lit_n); fillp(m->is_match, 192); fillp(m->is_rep, 12); fillp(m->g0, 12); fillp(m->g1, 12); fillp(m->g2, 12);
    fillp(m->rep0long, 192); fillp(&m->pos_slot[0][0], 256); fillp(m->pos_dec, 115); fillp(m->align, 16);
    len_init(&m->len); len_init(&m->rep_len);
    m->lc = lc; m->lp = lp; m->pb = pb;
    if (!(flags & 8)) {
        out[0] = (uint8_t)((pb * 5 + lp) * 9 + lc);
        for (int i = 0; i < 4; i++) out[1 + i] = (uint8_t)(dict_size >> (8 * i));
        for (int i = 0; i < 8; i++) out[5 + i] = (flags & 2) ? (uint8_t)(n >> (8 * i)) : 0xFF;
    }
    rc_init(&r, out + hdr, out_cap - hdr);
    if (dict_size < 4096) dict_size = 4096;
    const uint32_t pmask = (1u << pb) - 1;
    uint64_t pos = 0;
    uint32_t wpos = 0;   /* the reference's wrapped window position (Q3) */
#define ADV(k) do { wpos += (k); while (wpos >= dict_size) wpos -= dict_size; } while (0)
#define INS(p_) do { if ((p_) + 3 <= n) { uint32_t h_ = ((data[p_] | data[(p_) + 1] << 8 | data[(p_) + 2] << 16) * 2654435761u) >> (32 - HB); \
                     chain[p_] = head[h_]; head[h_] = (int32_t)(p_); } } while (0)
    if ((flags & 4) && n >= 4 && data[0] == 0 && data[1] == 0 && data[2] == 0 && data[3] == 0) {
        uint32_t len = 4;
        while (len < 273 && len < n && data[len] == 0) len++;
        uint32_t ps = 0;
        rc_bit(&r, &m->is_match[0], 1); rc_bit(&r, &m->is_rep[0], 0);
        len_enc(&r, &m->len, len - 2, ps); enc_dist(&r, m, 0, len - 2);
        m->rep[3] = m->rep[2]; m->rep[2] = m->rep[1]; m->rep[1] = m->rep[0]; m->rep[0] = 0; m->state = 7;
        for (uint32_t i = 0; i < len; i++) INS(pos + i);
        pos += len; ADV(len);
    }
    while (pos < n) {
        uint32_t ps = wpos & pmask, st2 = (m->state << 4) + ps;
        uint64_t avail = n - pos, maxlen = avail < 273 ? avail : 273;
        uint64_t maxd = pos < dict_size ? pos : dict_size; /* distances 1..maxd are valid */
        /* rep candidates */
        uint32_t best_rep = 0, best_rep_i = 0;
        for (uint32_t i = 0; i < 4; i++) {
            uint64_t d = (uint64_t)m->rep[i] + 1;
            if (d > maxd) continue;
            uint32_t l = 0;
            while (l < maxlen && data[pos + l] == data[pos + l - d]) l++;
            if (l > best_rep) { best_rep = l; best_rep_i = i; }
        }
        /* hash-chain candidates */
        uint32_t best = 0; uint64_t best_d = 0;
        if (avail >= 3) {
            uint32_t h = ((data[pos] | data[pos + 1] << 8 | data[pos + 2] << 16) * 2654435761u) >> (32 - HB);
            int32_t c = head[h]; int tries = 24;
            while (c >= 0 && tries-- > 0) {
                uint64_t d = pos - (uint64_t)c;
                if (d > maxd) break;
                uint32_t l = 0;
                while (l < maxlen && data[pos + l] == data[(uint64_t)c + l]) l++;
                if (l > best) { best = l; best_d = d; }
                c = chain[c];
            }
        }
        uint32_t used;
        if (best_rep >= 2 && best_rep + 1 >= best) {          /* rep match */
            rc_bit(&r, &m->is_match[st2], 1); rc_bit(&r, &m->is_rep[m->state], 1);
            if (best_rep_i == 0) { rc_bit(&r, &m->g0[m->state], 0); rc_bit(&r, &m->rep0long[st2], 1); }
            else {
                rc_bit(&r, &m->g0[m->state], 1);
                if (best_rep_i == 1) rc_bit(&r, &m->g1[m->state], 0);
                else { rc_bit(&r, &m->g1[m->state], 1); rc_bit(&r, &m->g2[m->state], best_rep_i == 2 ? 0 : 1); }
                uint32_t d = m->rep[best_rep_i];
                for (uint32_t i = best_rep_i; i > 0; i--) m->rep[i] = m->rep[i - 1];
                m->rep[0] = d;
            }
            len_enc(&r, &m->rep_len, best_rep - 2, ps);
            m->state = m->state < 7 ? 8 : 11;
            used = best_rep;
        } else if (best >= 3 || (best == 2 && best_d < 128)) { /* simple match */
            rc_bit(&r, &m->is_match[st2], 1); rc_bit(&r, &m->is_rep[m->state], 0);
            len_enc(&r, &m->len, best - 2, ps);
            enc_dist(&r, m, (uint32_t)(best_d - 1), best - 2);
            m->rep[3] = m->rep[2]; m->rep[2] = m->rep[1]; m->rep[1] = m->rep[0]; m->rep[0] = (uint32_t)(best_d - 1);
            m->state = m->state < 7 ? 7 : 10;
            used = best;
        } else if ((uint64_t)m->rep[0] + 1 <= maxd && data[pos] == data[pos - m->rep[0] - 1] && (pos & 3) == 1) { /* short rep now and then */
            rc_bit(&r, &m->is_match[st2], 1); rc_bit(&r, &m->is_rep[m->state], 1);
            rc_bit(&r, &m->g0[m->state], 0); rc_bit(&r, &m->rep0long[st2], 0);
            m->state = m->state < 7 ? 9 : 11;
            used = 1;
        } else {
            rc_bit(&r, &m->is_match[st2], 0);
            enc_literal(&r, m, data, pos, wpos);
            used = 1;
        }
        for (uint32_t i = 0; i < used; i++) INS(pos + i);
        pos += used; ADV(used);
    }
    if (flags & 1) { /* EOS marker: match with distance 0xFFFFFFFF + 1 */
        uint32_t ps = wpos & pmask, st2 = (m->state << 4) + ps;
        rc_bit(&r, &m->is_match[st2], 1); rc_bit(&r, &m->is_rep[m->state], 0);
        len_enc(&r, &m->len, 0, ps);
        enc_dist(&r, m, 0xFFFFFFFFu, 0);
    }
    rc_flush(&r);
    size_t total = r.overflow ? 0 : hdr + r.pos;
    free(m->lit); free(head); free(chain); free(m);
    return total;
}
