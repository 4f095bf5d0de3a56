/*
 * oracle/lzma_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the decode path of the
 * reference package kulaginds/lzma (pure Go).  It exists to CHECK the CUDA
 * path; nothing under lzma_b200/ may include, link or call it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it.
 *
 * Parity pins (see tests/test_oracle.py): the reference's own golden values
 * -- MD5 b2d18c42... of randomfile.dat.lzma / .lzma2 (reader1_test.go:107,
 * reader2_test.go:12-29), success of the four good a*.lzma assets and failure
 * of the three bad_* assets (reader1_test.go:26-67) -- plus liblzma
 * cross-checks on generated streams.  The Go toolchain is absent in this image,
 * so the reference itself cannot be run here (no oracle/_ref).
 *
 * Each function cites the reference file:line it follows.  The control flow
 * and, above all, the ORDER of the size / EOS / distance checks follow
 * decompress.go, because that order decides the error site.  The reference's
 * hand-inlined bit steps are restated once (rc_bit) instead of ~40 times.
 *
 * Reference quirks kept on purpose:
 *   Q1  input exhausted inside the symbol loop == clean end of stream
 *       (decompress.go:35-38, reader1.go:246-249)      -> ORC_OK_INPUT_EXHAUSTED
 *   Q3  posState / literal lp context use the WRAPPED window position
 *       (decompress.go:22,56; window.go:31-42)
 *   Q4  CheckDistance accepts distance == pos+1 on a non-full window
 *       (window.go:89-91) and reads buf[size-1]
 *   Q5  rep matches are guarded only by IsEmpty (decompress.go:690-692)
 *   Q6  LZMA2 control bytes 0x03..0x7F end the stream (reader2.go:185-198)
 *   Q7  LZMA2 compressed size is uint16 and wraps (reader2.go:21,143-144)
 *   Q8  LZMA2 framing is not validated; unread chunk payload is not skipped
 *   Q10 uint32(bytesLeft) < length truncation (decompress.go:657)
 * Not kept: Q2 (caller buffer larger than the dictionary corrupts output); the
 * oracle is the reference driven with small reads, which is the LZMA spec.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lzma_oracle.h"

/* types.go:12-35 */
#define K_BITS 11
#define K_MOVE 5
#define K_TOP (1u << 24)
#define PROB_INIT 1024
#define DIC_MIN 4096u

typedef uint16_t prob;

/* ---- input: io.ByteReader with an optional limit (bytereader.go:7-28) ---- */
typedef struct {
    const uint8_t *p, *end; /* underlying stream */
    int64_t limit;          /* <0: unlimited, else limitedByteReader.N */
} bytein;

static int in_byte(bytein *in) {
    if (in->limit == 0) return -1;       /* bytereader.go:20-22 */
    if (in->p >= in->end) return -1;     /* underlying io.EOF */
    if (in->limit > 0) in->limit--;
    return *in->p++;
}

/* ---- window.go:8-159 ---- */
typedef struct {
    uint8_t *buf;
    uint32_t pos, size;
    int is_full;
    /* linear sink standing in for ReadPending -> caller buffer */
    uint8_t *out;
    uint64_t out_pos, out_cap;
    int overflow;
} window;

static void win_emit(window *w, uint8_t b) {
    if (w->out_pos < w->out_cap) w->out[w->out_pos] = b; else w->overflow = 1;
    w->out_pos++;
}
/* window.go:31-42 */
static void win_put(window *w, uint8_t b) {
    w->buf[w->pos] = b;
    win_emit(w, b);
    if (++w->pos >= w->size) { w->pos -= w->size; w->is_full = 1; }
}
/* window.go:44-53 */
static uint8_t win_get(const window *w, uint32_t dist) {
    uint32_t i = w->pos - dist;
    if (dist > w->pos) i = w->size - dist + w->pos;
    return w->buf[i];
}
/* window.go:55-87: byte loop, both cursors wrap, overlap replicates */
static void win_copy(window *w, uint32_t dist, uint32_t len) {
    uint32_t from = dist <= w->pos ? w->pos - dist : w->size - dist + w->pos;
    uint32_t to = w->pos;
    w->pos += len;
    if (w->pos >= w->size) { w->pos -= w->size; w->is_full = 1; }
    for (; len > 0; len--) {
        uint8_t b = w->buf[from];
        w->buf[to] = b;
        win_emit(w, b);
        if (++from == w->size) from = 0;
        if (++to == w->size) to = 0;
    }
}
/* window.go:89-95 */
static int win_check_distance(const window *w, uint32_t d) { return w->is_full || d <= w->pos; }
static int win_is_empty(const window *w) { return w->pos == 0 && !w->is_full; }
/* window.go:135-140: Reset does NOT clear the bytes */
static void win_reset(window *w) { w->pos = 0; w->is_full = 0; }

/* ---- state.go:3-45 ---- */
typedef struct {
    prob *lit; uint32_t lit_cap;
    prob pos_slot[4][64];
    prob pos_dec[115];
    prob align[16];
    prob len_choice[2][2];        /* [0]=match len, [1]=rep len : choice, choice2 */
    prob len_low[2][16][8], len_mid[2][16][8], len_high[2][256];
    prob is_match[192], is_rep[12], is_rep_g0[12], is_rep_g1[12], is_rep_g2[12], is_rep0_long[192];
    uint8_t lc, pb, lp;
    int size_defined;
    uint64_t unpack_size, bytes_left;
    uint32_t rep0, rep1, rep2, rep3, state;
} lzstate;

static void fill(prob *p, size_t n) { for (size_t i = 0; i < n; i++) p[i] = PROB_INIT; }

/* state.go:79-121 */
static void st_reset(lzstate *s) {
    fill(s->lit, (size_t)0x300 << (s->lc + s->lp));
    fill(&s->pos_slot[0][0], 4 * 64); fill(s->align, 16); fill(s->pos_dec, 115);
    fill(s->is_match, 192); fill(s->is_rep, 12); fill(s->is_rep_g0, 12);
    fill(s->is_rep_g1, 12); fill(s->is_rep_g2, 12); fill(s->is_rep0_long, 192);
    for (int k = 0; k < 2; k++) {
        s->len_choice[k][0] = s->len_choice[k][1] = PROB_INIT;
        fill(&s->len_low[k][0][0], 128); fill(&s->len_mid[k][0][0], 128); fill(s->len_high[k], 256);
    }
    s->rep0 = s->rep1 = s->rep2 = s->rep3 = 0;
    s->state = 0;
}
/* state.go:47-77 (newState / Renew) */
static int st_renew(lzstate *s, uint8_t lc, uint8_t pb, uint8_t lp) {
    uint32_t need = (uint32_t)0x300 << (lc + lp);
    s->lc = lc; s->pb = pb; s->lp = lp;
    if (need > s->lit_cap) {
        free(s->lit);
        s->lit = (prob *)malloc((size_t)need * sizeof(prob));
        if (!s->lit) return -1;
        s->lit_cap = need;
    }
    st_reset(s);
    return 0;
}
/* state.go:123-151: "defined" unless all eight bytes are 0xFF */
static void st_set_unpack_size(lzstate *s, uint64_t n) {
    s->bytes_left = s->unpack_size = n;
    s->size_defined = n != UINT64_MAX;
}
/* state.go:153-187 */
static uint32_t upd_lit(uint32_t s) { return s < 4 ? 0 : (s < 10 ? s - 3 : s - 6); }
static uint32_t upd_match(uint32_t s) { return s < 7 ? 7 : 10; }
static uint32_t upd_rep(uint32_t s) { return s < 7 ? 8 : 11; }
static uint32_t upd_shortrep(uint32_t s) { return s < 7 ? 9 : 11; }

/* ---- range_decoder.go ---- */
typedef struct { uint32_t range, code; bytein *in; int eof; } rdec;

/* range_decoder.go:27-46.  0 ok, 1 first byte != 0, -1 EOF */
static int rc_init(rdec *rc, bytein *in) {
    rc->in = in; rc->range = 0xFFFFFFFFu; rc->code = 0; rc->eof = 0;
    int b = in_byte(in);
    if (b < 0) return -1;
    if (b != 0) return 1;
    for (int i = 0; i < 4; i++) {
        b = in_byte(in);
        if (b < 0) return -1;
        rc->code = (rc->code << 8) | (uint32_t)b;
    }
    return 0;
}
/* "Normalize" AFTER the bit, as the reference does (range_decoder.go:64-75). */
static inline int rc_norm(rdec *rc) {
    if (rc->range < K_TOP) {
        int b = in_byte(rc->in);
        if (b < 0) { rc->eof = 1; return -1; }
        rc->range <<= 8;
        rc->code = (rc->code << 8) | (uint32_t)b;
    }
    return 0;
}
/* range_decoder.go:57-98.  Returns 0/1, or -1 when the normalisation hit EOF
 * (the probability has already been updated, as in the reference). */
static inline int rc_bit(rdec *rc, prob *p) {
    uint32_t v = *p, bound = (rc->range >> K_BITS) * v;
    int bit;
    if (rc->code < bound) { *p = (prob)(v + (((1u << K_BITS) - v) >> K_MOVE)); rc->range = bound; bit = 0; }
    else { *p = (prob)(v - (v >> K_MOVE)); rc->code -= bound; rc->range -= bound; bit = 1; }
    if (rc_norm(rc) < 0) return -1;
    return bit;
}
/* range_decoder.go:100-134 / decompress.go:549-576 */
static int rc_direct(rdec *rc, int n, uint32_t *res) {
    uint32_t r = 0;
    for (; n > 0; n--) {
        rc->range >>= 1;
        rc->code -= rc->range;
        uint32_t t = 0u - (rc->code >> 31);
        rc->code += rc->range & t;
        r = (r << 1) + (t + 1);
        if (rc_norm(rc) < 0) return -1;
    }
    *res = r;
    return 0;
}
/* bit_tree_decoder.go:26-76 */
static int bt_fwd(rdec *rc, prob *probs, int nbits, uint32_t *out) {
    uint32_t m = 1;
    for (int i = 0; i < nbits; i++) { int b = rc_bit(rc, &probs[m]); if (b < 0) return -1; m = (m << 1) | (uint32_t)b; }
    *out = m - (1u << nbits);
    return 0;
}
/* bit_tree_decoder.go:82-135 */
static int bt_rev(rdec *rc, prob *probs, int nbits, uint32_t *out) {
    uint32_t m = 1, sym = 0;
    for (int i = 0; i < nbits; i++) { int b = rc_bit(rc, &probs[m]); if (b < 0) return -1; m = (m << 1) | (uint32_t)b; sym |= (uint32_t)b << i; }
    *out = sym;
    return 0;
}
/* len_decoder.go:34-60; live copies decompress.go:218-429 and :870-1118 */
static int len_decode(rdec *rc, lzstate *s, int k, uint32_t pos_state, uint32_t *len, int *which) {
    uint32_t v; int b;
    if ((b = rc_bit(rc, &s->len_choice[k][0])) < 0) return -1;
    if (b == 0) { if (bt_fwd(rc, s->len_low[k][pos_state], 3, &v) < 0) return -1; *len = v; *which = 0; return 0; }
    if ((b = rc_bit(rc, &s->len_choice[k][1])) < 0) return -1;
    if (b == 0) { if (bt_fwd(rc, s->len_mid[k][pos_state], 3, &v) < 0) return -1; *len = 8 + v; *which = 1; return 0; }
    if (bt_fwd(rc, s->len_high[k], 8, &v) < 0) return -1;
    *len = 16 + v; *which = 2;
    return 0;
}

/* Outcome of one run of the symbol loop. */
enum { RUN_EOF = 0, RUN_INPUT_EOF = 1, RUN_ERROR = 2 };

/*
 * decompress.go:8-1136 with needBytesCount = infinity: run until the loop
 * breaks with io.EOF, an input read fails, or ErrResultError is returned.
 * *site receives the decompress.go line of the failing return.
 */
static int decompress(lzstate *s, window *w, rdec *rc, int *site) {
    const uint32_t pos_mask = (1u << s->pb) - 1, lp_mask = (1u << s->lp) - 1;
    *site = 0;
    for (;;) {
        if (w->overflow) return RUN_ERROR; /* oracle-only: caller's buffer too small */
        if (s->size_defined && s->bytes_left == 0 && rc->code == 0) return RUN_EOF;   /* :14-20 */

        uint32_t pos_state = w->pos & pos_mask;                                       /* :22 */
        uint32_t state2 = (s->state << 4) + pos_state;                                /* :23 */
        int b = rc_bit(rc, &s->is_match[state2]);                                     /* :25-42 */
        if (b < 0) return RUN_INPUT_EOF;

        if (b == 0) { /* literal, :44-175 */
            if (s->size_defined && s->bytes_left == 0) { *site = 46; return RUN_ERROR; }
            uint32_t prev = win_is_empty(w) ? 0 : win_get(w, 1);                      /* :50-53 */
            uint32_t sym = 1;
            uint32_t lit_state = ((w->pos & lp_mask) << s->lc) + (prev >> (8 - s->lc)); /* :56 */
            prob *pr = &s->lit[0x300u * lit_state];
            if (s->state >= 7) {                                                      /* :59-114 */
                uint32_t mb = win_get(w, s->rep0 + 1);
                while (sym < 0x100) {
                    uint32_t mbit = (mb >> 7) & 1;
                    mb = (mb << 1) & 0xFF;
                    b = rc_bit(rc, &pr[((1 + mbit) << 8) + sym]);
                    if (b < 0) return RUN_INPUT_EOF;
                    sym = (sym << 1) | (uint32_t)b;
                    if (mbit != (uint32_t)b) break;
                }
            }
            while (sym < 0x100) {                                                     /* :127-166 */
                b = rc_bit(rc, &pr[sym]);
                if (b < 0) return RUN_INPUT_EOF;
                sym = (sym << 1) | (uint32_t)b;
            }
            win_put(w, (uint8_t)(sym - 0x100));                                       /* :168 */
            s->state = upd_lit(s->state);
            s->bytes_left--;
            continue;
        }

        uint32_t length; int which;
        b = rc_bit(rc, &s->is_rep[s->state]);                                         /* :195-213 */
        if (b < 0) return RUN_INPUT_EOF;
        if (b == 0) { /* simple match, :215-668 */
            s->rep3 = s->rep2; s->rep2 = s->rep1; s->rep1 = s->rep0;                   /* :216 */
            if (len_decode(rc, s, 0, pos_state, &length, &which) < 0) return RUN_INPUT_EOF;
            s->state = upd_match(s->state);                                           /* :431 */
            uint32_t len_state = length > 3 ? 3 : length, slot;                       /* :434-437 */
            if (bt_fwd(rc, s->pos_slot[len_state], 6, &slot) < 0) return RUN_INPUT_EOF;
            if (slot < 4) s->rep0 = slot;                                             /* :488-489 */
            else {
                uint32_t nd = (slot >> 1) - 1, dist = (2 | (slot & 1)) << nd, v;
                if (slot < 14) {                                                      /* :494-546 */
                    if (bt_rev(rc, &s->pos_dec[dist - slot], (int)nd, &v) < 0) return RUN_INPUT_EOF;
                    dist += v;
                } else {                                                              /* :548-628 */
                    if (rc_direct(rc, (int)nd - 4, &v) < 0) return RUN_INPUT_EOF;
                    dist += v << 4;
                    if (bt_rev(rc, s->align, 4, &v) < 0) return RUN_INPUT_EOF;
                    dist += v;
                }
                s->rep0 = dist;
            }
            if (s->rep0 == 0xFFFFFFFFu) {                                             /* :633-645 */
                if (rc->code == 0) {
                    if (s->size_defined && s->bytes_left > 0) { *site = 636; return RUN_ERROR; }
                    return RUN_EOF;
                }
                *site = 643; return RUN_ERROR;
            }
            if (s->size_defined && s->bytes_left == 0) { *site = 648; return RUN_ERROR; }
            if (s->rep0 >= w->size || !win_check_distance(w, s->rep0)) { *site = 652; return RUN_ERROR; }
            length += 2;                                                              /* :656 */
            if (s->size_defined && (uint32_t)s->bytes_left < length) {                /* :657-662 (Q10) */
                length = (uint32_t)s->bytes_left;
                win_copy(w, s->rep0 + 1, length);
                s->bytes_left -= length;
                *site = 662; return RUN_ERROR;
            }
            win_copy(w, s->rep0 + 1, length);
            s->bytes_left -= length;
            continue;
        }

        /* rep match, :685-1118 */
        if (s->size_defined && s->bytes_left == 0) { *site = 687; return RUN_ERROR; }
        if (win_is_empty(w)) { *site = 691; return RUN_ERROR; }
        b = rc_bit(rc, &s->is_rep_g0[s->state]);                                      /* :694-712 */
        if (b < 0) return RUN_INPUT_EOF;
        if (b == 0) {
            b = rc_bit(rc, &s->is_rep0_long[state2]);                                 /* :715-755 */
            if (b < 0) return RUN_INPUT_EOF;
            if (b == 0) { /* short rep, :735-739 */
                s->state = upd_shortrep(s->state);
                win_put(w, win_get(w, s->rep0 + 1));
                s->bytes_left--;
                continue;
            }
        } else {
            uint32_t dist;
            b = rc_bit(rc, &s->is_rep_g1[s->state]);                                  /* :777-813 */
            /* The reference rotates rep0..3 before a failing normalisation returns;
             * the stream has ended either way, so that order is unobservable. */
            if (b < 0) return RUN_INPUT_EOF;
            if (b == 0) { dist = s->rep1; s->rep1 = s->rep0; s->rep0 = dist; }
            else {
                b = rc_bit(rc, &s->is_rep_g2[s->state]);                              /* :816-861 */
                if (b < 0) return RUN_INPUT_EOF;
                if (b == 0) { dist = s->rep2; s->rep2 = s->rep1; s->rep1 = s->rep0; s->rep0 = dist; }
                else { dist = s->rep3; s->rep3 = s->rep2; s->rep2 = s->rep1; s->rep1 = s->rep0; s->rep0 = dist; }
            }
        }
        if (len_decode(rc, s, 1, pos_state, &length, &which) < 0) return RUN_INPUT_EOF; /* :870-1101 */
        s->state = upd_rep(s->state);
        length += 2;
        if (s->size_defined && (uint32_t)s->bytes_left < length) {                    /* :936,1030,1106 */
            static const int sites[3] = {941, 1035, 1111};
            length = (uint32_t)s->bytes_left;
            win_copy(w, s->rep0 + 1, length);
            s->bytes_left -= length;
            *site = sites[which]; return RUN_ERROR;
        }
        win_copy(w, s->rep0 + 1, length);
        s->bytes_left -= length;
    }
}

/* reader1.go:210-221.  Return order of the reference is (lc, pb, lp). */
int orc_decode_prop(uint8_t d, uint8_t *lc, uint8_t *pb, uint8_t *lp) {
    if (d >= 9 * 5 * 5) return ORC_INCORRECT_PROPERTIES;
    *lc = d % 9; d /= 9; *pb = d / 5; *lp = d % 5;
    return ORC_OK;
}
/* reader1.go:193-208 */
uint32_t orc_decode_dict_size(const uint8_t p[4]) {
    uint32_t d = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
    return d < DIC_MIN ? DIC_MIN : d;
}
/* reader1.go:178-191 */
uint64_t orc_decode_unpack_size(const uint8_t p[8]) {
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= (uint64_t)p[i] << (8 * i);
    return n;
}
/* reader2.go:296-298 */
uint32_t orc_decode_dict_size2(uint8_t b) { return (uint32_t)(2 | (b & 1)) << (b / 2 + 11); }

static void result_fill(orc_result *r, int status, int site, const window *w, const bytein *in,
                        const uint8_t *in0, const rdec *rc) {
    r->status = status; r->err_site = site;
    r->bytes_out = w ? w->out_pos : 0;
    r->bytes_in = in ? (uint64_t)(in->p - in0) : 0;
    r->final_code = rc ? rc->code : 0;
}

/* Reader1.initialize (reader1.go:149-159) + Read loop (reader1.go:223-254). */
static int run_reader1(bytein *in, const uint8_t *in0, uint8_t lc, uint8_t pb, uint8_t lp,
                       uint32_t dict_size, uint64_t unpack_size,
                       uint8_t *out, uint64_t out_cap, orc_result *res) {
    lzstate *s = (lzstate *)calloc(1, sizeof(lzstate));
    window w; memset(&w, 0, sizeof w);
    rdec rc; memset(&rc, 0, sizeof rc);
    int rv = ORC_OK, site = 0;
    w.buf = (uint8_t *)calloc(dict_size, 1); /* newWindow: make([]byte, dictSize), window.go:18-29 */
    w.size = dict_size; w.out = out; w.out_cap = out_cap;
    if (!s || !w.buf || st_renew(s, lc, pb, lp) < 0) { free(s); free(w.buf); return -1; }
    st_set_unpack_size(s, unpack_size);
    int e = rc_init(&rc, in);
    if (e < 0) rv = ORC_UNEXPECTED_EOF;          /* "rangeDec.Init: %w" of io.EOF, reader1.go:153-156 */
    else if (e > 0) { rv = ORC_RESULT_ERROR; site = 2033; } /* range_decoder.go:32-34 */
    else {
        int r = decompress(s, &w, &rc, &site);
        if (w.overflow) rv = ORC_OUTPUT_OVERFLOW;
        else rv = r == RUN_EOF ? ORC_OK : r == RUN_INPUT_EOF ? ORC_OK_INPUT_EXHAUSTED : ORC_RESULT_ERROR;
    }
    result_fill(res, rv, site, &w, in, in0, &rc);
    free(s->lit); free(s); free(w.buf);
    return 0;
}

/* NewReader1 (reader1.go:18-24, 77-101) followed by io.Copy. */
int orc_lzma_alone(const uint8_t *in_p, uint64_t in_len, uint8_t *out, uint64_t out_cap, orc_result *res) {
    bytein in = {in_p, in_p + in_len, -1};
    uint8_t lc, pb, lp;
    memset(res, 0, sizeof *res);
    if (in_len < 1) { res->status = ORC_UNEXPECTED_EOF; return 0; }     /* bare io.EOF, reader1.go:78-81 */
    if (orc_decode_prop(in_p[0], &lc, &pb, &lp) != ORC_OK) { res->status = ORC_INCORRECT_PROPERTIES; return 0; }
    if (in_len < 13) { res->status = ORC_UNEXPECTED_EOF; return 0; }    /* reader1.go:88-98 */
    uint32_t dict = orc_decode_dict_size(in_p + 1);
    uint64_t usz = orc_decode_unpack_size(in_p + 5);
    in.p += 13;
    return run_reader1(&in, in_p, lc, pb, lp, dict, usz, out, out_cap, res);
}

/* NewLZMADecompressorForSevenZip (reader1.go:32-61): no 13-byte header in the stream. */
int orc_lzma_raw(const uint8_t *in_p, uint64_t in_len, uint8_t lc, uint8_t lp, uint8_t pb,
                 uint32_t dict_size, uint64_t unpack_size,
                 uint8_t *out, uint64_t out_cap, orc_result *res) {
    bytein in = {in_p, in_p + in_len, -1};
    memset(res, 0, sizeof *res);
    if (dict_size < DIC_MIN) dict_size = DIC_MIN;
    return run_reader1(&in, in_p, lc, pb, lp, dict_size, unpack_size, out, out_cap, res);
}

/*
 * NewReader2 + io.Copy (reader2.go:26-41, 77-294).  One window and one
 * Reader1 state shared by all chunks, exactly as the reference does.
 */
int orc_lzma2(const uint8_t *in_p, uint64_t in_len, uint32_t dict_size,
              uint8_t *out, uint64_t out_cap, orc_result *res) {
    bytein in = {in_p, in_p + in_len, -1};
    window w; memset(&w, 0, sizeof w);
    rdec rc; memset(&rc, 0, sizeof rc);
    lzstate *s = NULL;                 /* r.lzmaReader == nil until the first LZMA chunk */
    uint8_t header5 = 0;               /* r.header[5]: persists between chunks, initially 0 (Q8) */
    int rv = ORC_OK, site = 0;
    memset(res, 0, sizeof *res);
    if (dict_size < DIC_MIN) dict_size = 8u << 20;      /* reader2.go:88-91 */
    w.buf = (uint8_t *)calloc(dict_size, 1);
    if (!w.buf) return -1;
    w.size = dict_size; w.out = out; w.out_cap = out_cap;

    for (;;) { /* startChunk, reader2.go:100-173 */
        int c = in_byte(&in);
        if (c < 0) { rv = ORC_UNEXPECTED_EOF; break; }                      /* :103-110 */
        int lz = c >> 5, type; /* decodeChunkType, :175-199 */
        if (c == 0) type = 0; else if (c == 1) type = 1; else if (c == 2) type = 2;
        else if (lz >= 4) type = lz - 1; /* 3..6 = NoReset, ResetState, NewProp, NewPropResetDict */
        else type = 0;                   /* 0x03..0x7F: treated as end of stream (Q6) */
        if (type == 0) { rv = ORC_OK; break; }
        int hl = type <= 2 ? 3 : (type <= 4 ? 5 : 6);                       /* chunkLength, :201-214 */
        uint8_t h[6]; h[0] = (uint8_t)c;
        if ((uint64_t)(in.end - in.p) < (uint64_t)(hl - 1)) { in.p = in.end; rv = ORC_UNEXPECTED_EOF; break; } /* :121-128 */
        for (int i = 1; i < hl; i++) h[i] = *in.p++;
        if (hl == 6) header5 = h[5];
        uint32_t usz = ((uint32_t)h[1] << 8) | h[2];                        /* :130 */
        if (type == 1 || type == 6) win_reset(&w);                          /* :132-134 */
        if (type <= 2) { /* uncompressed: uncompressedRead, :252-294 + window.ReadFrom :142-155 */
            usz++;
            uint64_t avail = (uint64_t)(in.end - in.p), n = usz < avail ? usz : avail;
            for (uint64_t i = 0; i < n; i++) win_put(&w, *in.p++);
            if (w.overflow) { rv = ORC_OUTPUT_OVERFLOW; break; }
            continue; /* a short payload surfaces as UNEXPECTED_EOF at the next header read */
        }
        usz |= (uint32_t)(c & 0x1F) << 16; usz++;                           /* :141-142 */
        uint16_t csz = (uint16_t)((((uint16_t)h[3] << 8) | h[4]) + 1);      /* :143-144, uint16 wrap (Q7) */
        in.limit = csz;                                                     /* limitByteReader(in, cs) */
        int e;
        if (!s) { /* first LZMA chunk ever: NewReader1ForReader2, :146-153 */
            uint8_t lc, pb, lp;
            if (orc_decode_prop(header5, &lc, &pb, &lp) != ORC_OK) { rv = ORC_INCORRECT_PROPERTIES; break; }
            s = (lzstate *)calloc(1, sizeof(lzstate));
            if (!s || st_renew(s, lc, pb, lp) < 0) { free(w.buf); free(s); return -1; }
        } else if (type == 4) st_reset(s);                                  /* :156-157 */
        else if (type >= 5) {                                               /* :158-165 */
            uint8_t lc, pb, lp;
            if (orc_decode_prop(header5, &lc, &pb, &lp) != ORC_OK) { rv = ORC_INCORRECT_PROPERTIES; break; }
            if (st_renew(s, lc, pb, lp) < 0) { free(w.buf); free(s->lit); free(s); return -1; }
        }
        st_set_unpack_size(s, usz);                                         /* Reopen, reader1.go:166-176 */
        e = rc_init(&rc, &in);
        if (e < 0) { rv = ORC_UNEXPECTED_EOF; break; }
        if (e > 0) { rv = ORC_RESULT_ERROR; site = 2033; break; }
        int r = decompress(s, &w, &rc, &site);
        in.limit = -1; /* the next header is read straight from the underlying stream (Q8) */
        if (w.overflow) { rv = ORC_OUTPUT_OVERFLOW; break; }
        if (r == RUN_ERROR) { rv = ORC_RESULT_ERROR; break; }
        /* RUN_EOF and RUN_INPUT_EOF both end the chunk; Reader2.Read calls startChunk (:234-241) */
    }
    result_fill(res, rv, site, &w, &in, in_p, &rc);
    if (s) { free(s->lit); free(s); }
    free(w.buf);
    return 0;
}

/* ---- threaded batch driver (bench baseline only) ---- */
#include <pthread.h>
typedef struct {
    const uint8_t *in_base; const uint64_t *in_off, *in_len;
    uint8_t *out_base; const uint64_t *out_off, *out_cap;
    orc_result *res; int n; int next; int bad; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int i = j->next < j->n ? j->next++ : -1;
        pthread_mutex_unlock(&j->mu);
        if (i < 0) break;
        orc_lzma_alone(j->in_base + j->in_off[i], j->in_len[i], j->out_base + j->out_off[i], j->out_cap[i], &j->res[i]);
        if (j->res[i].status != ORC_OK) { pthread_mutex_lock(&j->mu); j->bad++; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

int orc_lzma_alone_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                         orc_result *res, int n, int threads) {
    batch_job j = {in_base, in_off, in_len, out_base, out_off, out_cap, res, n, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int i = 0; i < threads; i++) if (pthread_create(&t[started], NULL, batch_worker, &j) == 0) started++;
    if (started == 0) batch_worker(&j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    free(t);
    return j.bad;
}
