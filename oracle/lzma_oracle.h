/* oracle/lzma_oracle.h -- TEST INFRASTRUCTURE ONLY (see lzma_oracle.c). */
#ifndef LZMA_ORACLE_H
#define LZMA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same numbering as the LZGPU_* status codes in include/lzgpu.h. */
enum {
    ORC_OK = 0,                   /* clean end (decompress.go:14-20, 633-641) */
    ORC_OK_INPUT_EXHAUSTED = 1,   /* input ran out == clean EOF in the reference (Q1) */
    ORC_RESULT_ERROR = 2,         /* ErrResultError (errors.go:8) */
    ORC_INCORRECT_PROPERTIES = 3, /* ErrIncorrectProperties (errors.go:7) */
    ORC_UNEXPECTED_EOF = 4,       /* io.EOF / io.ErrUnexpectedEOF from a constructor or chunk header */
    ORC_OUTPUT_OVERFLOW = 5       /* caller's buffer too small (no reference analogue) */
};

typedef struct {
    int32_t status;
    int32_t err_site;     /* decompress.go line of the failing return; 2033 = range_decoder.go:33 */
    uint64_t bytes_out;
    uint64_t bytes_in;    /* bytes consumed from the start of the input */
    uint32_t final_code;  /* rangeDecoder.Code when the run stopped */
    uint32_t pad;
} orc_result;

int orc_decode_prop(uint8_t d, uint8_t *lc, uint8_t *pb, uint8_t *lp);
uint32_t orc_decode_dict_size(const uint8_t p[4]);
uint64_t orc_decode_unpack_size(const uint8_t p[8]);
uint32_t orc_decode_dict_size2(uint8_t b);

int orc_lzma_alone(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, orc_result *res);
int orc_lzma_raw(const uint8_t *in, uint64_t in_len, uint8_t lc, uint8_t lp, uint8_t pb,
                 uint32_t dict_size, uint64_t unpack_size,
                 uint8_t *out, uint64_t out_cap, orc_result *res);
int orc_lzma2(const uint8_t *in, uint64_t in_len, uint32_t dict_size,
              uint8_t *out, uint64_t out_cap, orc_result *res);

/* Threaded driver used by bench.py's cpu_baseline / --impl reference legs:
 * decodes n independent .lzma streams with `threads` pthreads (one stream per
 * task, dynamic queue).  Returns the number of streams whose status != ORC_OK. */
int orc_lzma_alone_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                         orc_result *res, int n, int threads);
#ifdef __cplusplus
}
#endif
#endif
