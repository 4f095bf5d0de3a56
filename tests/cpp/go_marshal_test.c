/* go_marshal_test.c -- pins the marshalling of go/lzgpu.go (which cannot be compiled here: no Go toolchain).
 *
 * The Go binding builds []C.lzgpu_unit with make() (all bytes zero) and then assigns, per unit, exactly the
 * fields listed in marshal() of go/lzgpu.go.  This program does the same in C, field for field, for the two
 * ways a Go caller gets LZMA2 units:
 *   A  ScanLZMA2 -> []Unit -> DecodeBatch            (lit_bits / pos_bits / flags carried through)
 *   B  Units built by hand / by an older binding      (lit_bits = pos_bits = 0, LZGPU_UF_BITS_KNOWN clear:
 *      the round-1 Go file dropped lit_bits this way and every xz-made stream failed with site 9101)
 * and decodes an LZMA2 stream (argv[1]; expected plaintext argv[2]) on the GPU through lzgpu_decode_batch with
 * pageable buffers laid out like DecodeBatch lays them (16-byte aligned slots, +16 bytes of slack).
 * Exit code 0 = both ways give the plaintext.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lzgpu.h"

/* mirror of `type Unit struct` in go/lzgpu.go */
typedef struct {
    uint8_t Kind;
    const uint8_t *In;
    uint64_t InLen;
    uint64_t OutCap;
    uint8_t Lc, Lp, Pb;
    uint32_t DictSize;
    uint64_t UnpackSize;
    uint32_t Flags;
    uint8_t LitBits, PosBits;
} GoUnit;

/* func marshal(cu *C.lzgpu_unit, u *Unit, inOff, outOff uint64) */
static void marshal(lzgpu_unit *cu, const GoUnit *u, uint64_t inOff, uint64_t outOff) {
    cu->in_off = inOff; cu->in_len = u->InLen;
    cu->out_off = outOff; cu->out_cap = u->OutCap;
    cu->kind = u->Kind; cu->flags = u->Flags;
    cu->lc = u->Lc; cu->lp = u->Lp; cu->pb = u->Pb;
    cu->lit_bits = u->LitBits; cu->pos_bits = u->PosBits;
    cu->dict_size = u->DictSize; cu->unpack_size = u->UnpackSize;
}

static uint8_t *slurp(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *p = malloc(*n + 1);
    if (fread(p, 1, *n, f) != *n) { perror("read"); exit(2); }
    fclose(f);
    return p;
}

/* func (e *Engine) DecodeBatch(units []Unit) */
static int decode_batch(lzgpu_ctx *ctx, const GoUnit *units, int64_t n, const uint8_t *want, size_t want_len, const char *what) {
    lzgpu_unit *cu = calloc((size_t)n, sizeof *cu);     /* make([]C.lzgpu_unit, n) */
    uint64_t inSize = 0, outSize = 0;
    for (int64_t i = 0; i < n; i++) {
        marshal(&cu[i], &units[i], inSize, outSize);
        inSize = (inSize + units[i].InLen + 15) & ~(uint64_t)15;
        outSize = (outSize + units[i].OutCap + 15) & ~(uint64_t)15;
    }
    uint8_t *in = calloc(inSize + 16, 1), *out = calloc(outSize + 16, 1);
    for (int64_t i = 0; i < n; i++) memcpy(in + cu[i].in_off, units[i].In, units[i].InLen);
    lzgpu_result *res = calloc((size_t)n, sizeof *res);
    int rc = lzgpu_decode_batch(ctx, cu, n, in, inSize + 16, out, outSize + 16, res, NULL);
    if (rc != LZGPU_E_OK) { fprintf(stderr, "%s: lzgpu_decode_batch: %s\n", what, lzgpu_last_error()); return 1; }
    size_t got = 0;
    int bad = 0;
    for (int64_t i = 0; i < n && !bad; i++) {
        if (res[i].status != LZGPU_OK) {
            fprintf(stderr, "%s: unit %lld: %s site %d\n", what, (long long)i, lzgpu_status_name(res[i].status), res[i].err_site);
            bad = 1;
            break;
        }
        if (got + res[i].bytes_out > want_len || memcmp(out + cu[i].out_off, want + got, res[i].bytes_out) != 0) {
            fprintf(stderr, "%s: unit %lld: bytes differ\n", what, (long long)i);
            bad = 1;
        }
        got += res[i].bytes_out;
    }
    if (!bad && got != want_len) { fprintf(stderr, "%s: %zu bytes, want %zu\n", what, got, want_len); bad = 1; }
    if (!bad) printf("%s: %lld units, %zu bytes ok\n", what, (long long)n, got);
    free(cu); free(in); free(out); free(res);
    return bad;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s stream.lzma2 plain.bin [dict_size]\n", argv[0]); return 2; }
    size_t sl, pl;
    uint8_t *stream = slurp(argv[1], &sl), *plain = slurp(argv[2], &pl);
    const uint32_t dict = argc > 3 ? (uint32_t)strtoul(argv[3], NULL, 0) : (8u << 20);
    lzgpu_ctx *ctx = NULL;
    if (lzgpu_ctx_create(NULL, 0, &ctx) != LZGPU_E_OK) { fprintf(stderr, "ctx: %s\n", lzgpu_last_error()); return 3; }

    /* func ScanLZMA2(data []byte, dictSize uint32) (units []Unit, ...) */
    uint64_t tot = 0;
    int32_t sst = 0;
    int64_t n = lzgpu_scan_lzma2(stream, sl, dict, NULL, 0, &tot, &sst);
    if (n <= 0) { fprintf(stderr, "scan: %lld units\n", (long long)n); return 1; }
    lzgpu_unit *scanned = calloc((size_t)n, sizeof *scanned);
    lzgpu_scan_lzma2(stream, sl, dict, scanned, n, &tot, &sst);
    GoUnit *units = calloc((size_t)n, sizeof *units);
    for (int64_t i = 0; i < n; i++) {
        const lzgpu_unit *u = &scanned[i];
        GoUnit g = {0};
        g.Kind = 2; g.In = stream + u->in_off; g.InLen = u->in_len; g.OutCap = u->out_cap;
        g.Lc = u->lc; g.Lp = u->lp; g.Pb = u->pb; g.DictSize = u->dict_size;
        g.UnpackSize = u->unpack_size; g.Flags = u->flags;
        g.LitBits = u->lit_bits; g.PosBits = u->pos_bits;
        units[i] = g;
    }
    int bad = decode_batch(ctx, units, n, plain, pl, "A (ScanLZMA2 -> DecodeBatch)");
    for (int64_t i = 0; i < n; i++) { units[i].LitBits = units[i].PosBits = 0; units[i].Flags &= ~LZGPU_UF_BITS_KNOWN; }
    bad |= decode_batch(ctx, units, n, plain, pl, "B (table sizes not carried)");
    lzgpu_ctx_destroy(ctx);
    return bad;
}
