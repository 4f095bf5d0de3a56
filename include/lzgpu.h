/*
 * lzgpu.h -- C ABI of the B200-native batch LZMA / LZMA2 decoder.
 *
 * This is the drop-in boundary for the in-stream decode path of kulaginds/lzma
 * (reference: pure Go, no FFI of its own).  The reference-side cgo binding a
 * maintainer adds is shown in INTEGRATION.md; every entry point below names the
 * reference interface it replaces (file:line under the reference repo root).
 *
 * Plain pointers and sizes only.  There is NO CPU fallback: without a CUDA
 * device every decode entry point fails with LZGPU_E_NO_DEVICE.
 *
 * Model.  A *unit* is one independently decodable piece of work: a .lzma
 * stream, a headerless LZMA1 stream, or a run of LZMA2 chunks that starts at a
 * dictionary reset.  A batch is an array of units over one input buffer and
 * one output buffer.  Each unit is decoded start to finish by one warp of an
 * sm_100a kernel; units never communicate, so batches shard across GPUs by
 * unit with no collective.
 */
#ifndef LZGPU_H
#define LZGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZGPU_ABI_VERSION 2   /* 2: lzgpu_unit.pos_bits, LZGPU_UF_BITS_KNOWN / SUM_*, lzgpu_decode_batch_sums, CRC-64 */

/* ---- per-unit outcome <-> reference error values (errors.go:5-12) ---- */
enum lzgpu_status {
    LZGPU_OK = 0,                   /* clean end: size reached with Code==0, or EOS marker with Code==0
                                       (decompress.go:14-20, 633-641) */
    LZGPU_OK_INPUT_EXHAUSTED = 1,   /* input ran out inside the symbol loop; the reference treats this as a
                                       clean EOF with short output (decompress.go:35-38, reader1.go:246-249) */
    LZGPU_RESULT_ERROR = 2,         /* ErrResultError (errors.go:8); err_site says where */
    LZGPU_INCORRECT_PROPERTIES = 3, /* ErrIncorrectProperties (errors.go:7; reader1.go:211-213) */
    LZGPU_UNEXPECTED_EOF = 4,       /* io.EOF / io.ErrUnexpectedEOF from a constructor or an LZMA2 chunk
                                       header (reader1.go:78-98,153-156; reader2.go:103-128) */
    LZGPU_OUTPUT_OVERFLOW = 5,      /* out_cap too small (no reference analogue: its reader streams) */
    LZGPU_DICT_OUT_OF_RANGE = 6,    /* ErrDictOutOfRange: unreachable in the reference, never produced */
    LZGPU_UNEXPECTED_LZMA2_CODE = 7,/* ErrUnexpectedLZMA2Code: unreachable in the reference, never produced */
    LZGPU_NOT_RUN = 255             /* result slot not written (infrastructure failure) */
};

/* err_site: the decompress.go line of the failing `return ErrResultError`
 * (46, 636, 643, 648, 652, 662, 687, 691, 941, 1035, 1111), 2033 for
 * range_decoder.go:33 (first range-coder byte != 0), or a 9xxx code for a
 * documented deviation ("Deviations"):                          */
#define LZGPU_SITE_RC_INIT 2033
#define LZGPU_SITE_REP_BEFORE_DICT 9652   /* rep distance reaches before the dictionary start (Q5) */
#define LZGPU_SITE_LZMA2_CHUNK_SIZE 9100  /* LZMA2 chunk did not consume/produce its declared sizes (Q8) */
#define LZGPU_SITE_LZMA2_PROPS 9101       /* LZMA2 chunk needs literal tables larger than the unit declared */
#define LZGPU_SITE_LZMA2_NO_STATE 9102    /* LZMA2 unit begins with a chunk that needs inherited coder state */

/* infrastructure return codes of the entry points (not per-unit outcomes) */
enum lzgpu_error {
    LZGPU_E_OK = 0,
    LZGPU_E_NO_DEVICE = -1,   /* no CUDA device / driver: there is no CPU path */
    LZGPU_E_CUDA = -2,        /* a CUDA call failed; see lzgpu_last_error() */
    LZGPU_E_INVALID = -3,     /* bad argument (range outside buffer, unknown kind, ...) */
    LZGPU_E_NOMEM = -4
};

enum lzgpu_kind {
    LZGPU_KIND_LZMA1_ALONE = 0, /* 13-byte .lzma header in the stream: NewReader1 (reader1.go:18) */
    LZGPU_KIND_LZMA1_RAW = 1,   /* lc/lp/pb/dict/size supplied: NewLZMADecompressorForSevenZip (reader1.go:32) */
    LZGPU_KIND_LZMA2_GROUP = 2  /* LZMA2 chunk run starting at a dictionary reset: NewReader2 (reader2.go:26) */
};

#define LZGPU_UNKNOWN_SIZE UINT64_MAX   /* unpack size of all-ones: EOS marker mandatory (state.go:135-151) */

/* unit flags (LZMA2 groups; filled by lzgpu_scan_lzma2) */
#define LZGPU_UF_LZMA2_LAST 1u    /* last unit of its stream: must end with a terminator, else UNEXPECTED_EOF */
#define LZGPU_UF_LZMA2_FRESH 2u   /* no LZMA chunk precedes this unit in its stream: the first LZMA chunk
                                     builds a new coder whatever its control byte says (reader2.go:146-153) */
#define LZGPU_UF_BITS_KNOWN 4u    /* LZMA2: lit_bits / pos_bits below were computed from the chunk headers
                                     (lzgpu_scan_lzma2 sets it).  Without it lzgpu_decode_batch walks the unit's
                                     chunk headers itself, so a binding that drops the two fields cannot make a
                                     unit fail; lzgpu_plan_create (input already on the device) takes lit_bits
                                     as given and assumes pos_bits = 4 */

#define LZGPU_UF_SUM_CRC32 8u     /* lzgpu_decode_batch_sums: CRC-32 (zlib / .xz CHECK_CRC32) of this unit's decoded bytes */
#define LZGPU_UF_SUM_CRC64 16u    /* lzgpu_decode_batch_sums: CRC-64/XZ (ECMA-182, .xz CHECK_CRC64) of them */

typedef struct lzgpu_unit {
    uint64_t in_off, in_len;    /* compressed bytes: [in_off, in_off+in_len) of the batch input buffer */
    uint64_t out_off, out_cap;  /* where the decoded bytes go; out_cap bounds them */
    uint64_t unpack_size;       /* LZMA1: size from the header or caller, LZGPU_UNKNOWN_SIZE if unknown.
                                   LZMA2: sum of the chunk sizes found by the scanner (informational) */
    uint32_t dict_size;         /* dictionary size in bytes (after the reference's clamping) */
    uint8_t kind;               /* enum lzgpu_kind */
    uint8_t lc, lp, pb;         /* LZMA1: literal/pos bits.  LZMA2: props in force before the unit */
    uint8_t lit_bits;           /* LZMA2: max lc+lp any chunk of the unit uses (sizes the literal tables);
                                   LZMA1: ignored (lc + lp) */
    uint8_t pos_bits;           /* LZMA2: max pb any chunk of the unit uses (sizes the posState-indexed tables);
                                   LZMA1: ignored (pb) */
    uint8_t pad8[2];
    uint32_t flags;             /* LZGPU_UF_* */
    uint64_t user;              /* caller's tag, copied to nothing; keeps the struct at 64 bytes */
} lzgpu_unit;

typedef struct lzgpu_result {
    int32_t status;        /* enum lzgpu_status */
    int32_t err_site;
    uint64_t bytes_out;    /* decoded bytes written at out_off */
    uint64_t bytes_in;     /* compressed bytes consumed from in_off */
    uint32_t final_code;   /* range-coder Code when the unit stopped */
    int32_t device;        /* CUDA device ordinal that ran the unit */
} lzgpu_result;

typedef struct lzgpu_stats {
    double kernel_ms;      /* CUDA-event time around the decode kernels (max over devices) */
    double h2d_ms, d2h_ms; /* host<->device copies (host-buffer entry point only; max over devices) */
    double total_ms;       /* wall time of the call */
    uint64_t bytes_in, bytes_out;
    int32_t launches;      /* kernels launched */
    int32_t devices;
} lzgpu_stats;

/* ---- library / device ---- */
int lzgpu_abi_version(void);
int lzgpu_device_count(void);              /* CUDA devices visible; <= 0 means nothing here can decode */
const char *lzgpu_status_name(int status);
const char *lzgpu_last_error(void);        /* text of the last LZGPU_E_* failure on this thread */

/* ---- header helpers: same arithmetic as the reference's exported helpers ---- */
/* DecodeProp (reader1.go:210-221).  NB the reference returns (lc, pb, lp). */
int lzgpu_decode_prop(uint8_t d, uint8_t *lc, uint8_t *pb, uint8_t *lp);
/* DecodeDictSize (reader1.go:193-208): little-endian, clamped up to 4096. */
uint32_t lzgpu_decode_dict_size(const uint8_t props[4]);
/* DecodeUnpackSize (reader1.go:178-191). */
uint64_t lzgpu_decode_unpack_size(const uint8_t header[8]);
/* DecodeDictSize2 (reader2.go:296-298). */
uint32_t lzgpu_decode_dict_size2(uint8_t encoded);

/* Reader1.initializeFull (reader1.go:77-101) without touching the payload:
 * fills kind/lc/lp/pb/lit_bits/dict_size/unpack_size of *unit from the 13-byte
 * header at in[0..13).  Returns an lzgpu_status (OK, INCORRECT_PROPERTIES or
 * UNEXPECTED_EOF).  in_off/in_len/out_off/out_cap are left to the caller. */
int lzgpu_parse_alone_header(const uint8_t *in, uint64_t in_len, lzgpu_unit *unit);

/* Host-side LZMA2 chunk scanner: Reader2.startChunk's framing rules
 * (reader2.go:100-214) applied to a whole raw LZMA2 stream to find the units
 * (chunk runs that begin at a dictionary reset).  Writes up to max_units
 * descriptors with in_off/out_off relative to the stream start / decoded start
 * (the caller rebases them), returns the number of units the stream has
 * (call again with a larger array if it exceeds max_units), or a negative
 * lzgpu_error.  *total_out = decoded size the chunk headers promise;
 * *stream_status = LZGPU_OK, or LZGPU_UNEXPECTED_EOF when the stream ends
 * without a terminator / inside a chunk (the last unit then carries it). */
int64_t lzgpu_scan_lzma2(const uint8_t *in, uint64_t in_len, uint32_t dict_size,
                         lzgpu_unit *units, int64_t max_units,
                         uint64_t *total_out, int32_t *stream_status);

/* Host-side scheduler: longest-processing-time-first assignment of units to
 * n_shards GPUs / ranks, weighted by compressed size (in_len).  Deterministic,
 * so every rank of a multi-process job computes the same answer.
 * shard_of_unit[i] in [0, n_shards). */
int lzgpu_shard_units(const lzgpu_unit *units, int64_t n_units, int n_shards, int32_t *shard_of_unit);

/* ---- decoding ---- */
typedef struct lzgpu_ctx lzgpu_ctx;

/* One context owns a stream, events and scratch buffers on each listed device
 * (n_devices == 0: all visible devices).  Re-entrant across contexts; one
 * context serves one call at a time. */
int lzgpu_ctx_create(const int *devices, int n_devices, lzgpu_ctx **ctx);
void lzgpu_ctx_destroy(lzgpu_ctx *ctx);
int lzgpu_ctx_device_count(const lzgpu_ctx *ctx);

/* The batch entry point with HOST buffers.  Shards the units over the context's
 * devices by compressed size, decodes, returns output and results.  Synchronous.
 * With PINNED, mapped buffers (lzgpu_alloc_pinned, cudaHostAlloc / cudaHostRegister,
 * torch's pinned tensors) nothing is copied ahead of or after the kernel: the units
 * read their compressed input from the caller's buffer over PCIe while they decode,
 * and write every finished 64 KiB block of output and their tail into the caller's
 * buffer themselves ("push mode"; only the bytes a unit decoded, only to its range).
 * When the host cannot take the bytes as fast as they come (all GPUs of a node
 * writing at once) a device falls back to copy-engine transfers overlapped with the
 * kernel by itself; pageable buffers always go through device slabs and copies.
 * Environment: LZGPU_NO_ZEROCOPY_IN, LZGPU_NO_PUSH_D2H, LZGPU_PUSH_D2H, LZGPU_TRACE.
 * Returns LZGPU_E_OK when the batch executed (per-unit outcome in results[]).
 * Replaces, for many streams at once, NewReader1/NewReader2 + io.Copy
 * (reader1.go:18,223; reader2.go:26,216). */
int lzgpu_decode_batch(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n_units,
                       const uint8_t *in_base, uint64_t in_size,
                       uint8_t *out_base, uint64_t out_size,
                       lzgpu_result *results, lzgpu_stats *stats);

/* lzgpu_decode_batch that also returns, for every unit flagged LZGPU_UF_SUM_CRC32 / _CRC64, that checksum of
 * the bytes it decoded (sums[i]; a CRC-32 in the low 32 bits; 0 for units without a flag), computed on the GPU
 * from the output where it lies while it travels back -- what an .xz reader needs to verify its blocks, or a
 * caller of the reference would do with hash/crc32 / hash/crc64 over what Read returned, without a host pass
 * over the payload.  A block made of several units: fold with lzgpu_crc32_combine / lzgpu_crc64_combine. */
int lzgpu_decode_batch_sums(lzgpu_ctx *ctx, const lzgpu_unit *units, int64_t n_units,
                            const uint8_t *in_base, uint64_t in_size,
                            uint8_t *out_base, uint64_t out_size,
                            lzgpu_result *results, lzgpu_stats *stats, uint64_t *sums);
/* crc(A || B) from crc(A), crc(B) and the length of B (host arithmetic, O(log len_b)). */
uint32_t lzgpu_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);
uint64_t lzgpu_crc64_combine(uint64_t crc_a, uint64_t crc_b, uint64_t len_b);

/* The same with DEVICE-resident buffers on context device `dev_index`, in three
 * steps so that a caller can keep the plan and time the launch alone:
 *   plan_create : validates, orders units longest-first, uploads descriptors
 *   plan_launch : enqueues the decode kernels on `stream` (a cudaStream_t, NULL =
 *                 the context's own stream); asynchronous
 *   plan_results: waits for the launch and returns the per-unit results
 * LZMA1_ALONE units must have been through lzgpu_parse_alone_header. */
typedef struct lzgpu_plan lzgpu_plan;
int lzgpu_plan_create(lzgpu_ctx *ctx, int dev_index, const lzgpu_unit *units, int64_t n_units,
                      uint64_t in_size, uint64_t out_size, lzgpu_plan **plan);
int lzgpu_plan_launch(lzgpu_plan *plan, const uint8_t *d_in, uint8_t *d_out, void *stream);
int lzgpu_plan_results(lzgpu_plan *plan, lzgpu_result *results, lzgpu_stats *stats);
int lzgpu_plan_launch_count(const lzgpu_plan *plan); /* kernels one plan_launch enqueues */
/* Pinned (page-locked, device-mapped) host memory for the caller's input / output buffers, so that a host
 * language without CUDA bindings (the cgo facade) reaches lzgpu_decode_batch's fast path: compressed input read
 * by the kernel straight from host memory, output streamed back while it runs.  NULL on failure (no device,
 * out of memory); lzgpu_last_error() says why.  Buffers from elsewhere (Go slices, malloc) still work: they
 * are staged through device slabs. */
void *lzgpu_alloc_pinned(uint64_t size);
void lzgpu_free_pinned(void *p);

/* On-device verification (no reference analogue): CRC-32 (zlib's crc32, the .xz
 * CHECK_CRC32 polynomial) of every unit's decoded bytes out[out_off, out_off + bytes_out), computed on the
 * GPU after plan_launch on the same stream and returned in crc[n_units] (host memory).  Units that did not
 * run, or produced nothing, give 0.  Lets a batch too large to copy back be checked where it lies. */
int lzgpu_plan_crc32(lzgpu_plan *plan, const uint8_t *d_out, uint32_t *crc);
/* The same for CRC-64/XZ (ECMA-182 polynomial, reflected: the .xz CHECK_CRC64, xz's default check). */
int lzgpu_plan_crc64(lzgpu_plan *plan, const uint8_t *d_out, uint64_t *crc);
void lzgpu_plan_destroy(lzgpu_plan *plan);

#ifdef __cplusplus
}
#endif
#endif /* LZGPU_H */
