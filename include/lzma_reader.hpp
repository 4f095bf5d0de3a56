// lzma_reader.hpp -- C++ host side above the C ABI (include/lzgpu.h): the reader API of kulaginds/lzma
// with the reference's names, argument meaning and error behaviour, over the GPU batch engine.
//
// The reference is a Go package; no Go toolchain exists in the build image, so its host side is
// mirrored here in C++ (the Go/cgo sources a maintainer would add are in go/).
// Go idioms are kept so that tests/cpp/reader_test.cpp reads like reader1_test.go / reader2_test.go:
// functions return (value, error) pairs, errors are compared with errors::Is, io::Reader /
// io::ByteReader / io::ReadCloser are the interfaces the constructors take.
//
//   reference                                              here
//   errors.go:5-12   Err*                                  lzma::Err*            (same texts)
//   reader1.go:18    NewReader1(io.ByteReader)             lzma::NewReader1
//   reader1.go:223   (*Reader1).Read                       lzma::Reader1::Read
//   reader1.go:161   (*Reader1).Reset / :166 Reopen        lzma::Reader1::Reset / Reopen
//   reader1.go:32    NewLZMADecompressorForSevenZip        lzma::NewLZMADecompressorForSevenZip
//   reader1.go:178   DecodeUnpackSize / :193 DecodeDictSize / :210 DecodeProp (returns lc, pb, lp)
//   reader2.go:26    NewReader2(io.Reader, dictSize)       lzma::NewReader2
//   reader2.go:216   (*Reader2).Read                       lzma::Reader2::Read
//   reader2.go:49    NewLZMA2DecompressorForSevenZip       lzma::NewLZMA2DecompressorForSevenZip
//   reader2.go:296   DecodeDictSize2                       lzma::DecodeDictSize2
//   readcloser.go    readCloser                            lzma::readCloser
//   (new)            batch entry point                     lzma::Engine::DecodeBatch
//
// Where the reference decodes symbol by symbol as Read is called, these readers hand the stream to
// liblzgpu.so (one unit per .lzma stream, one unit per dictionary-reset chunk run of an LZMA2 stream,
// all units of a wave in one lzgpu_decode_batch call) and serve Read from the decoded bytes.  There is
// no CPU decode path: without a CUDA device the first Read returns the engine's error.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "lzgpu.h"

namespace lzma {

// ---- Go's error values ------------------------------------------------------------------------
struct ErrorValue {
    std::string msg;
    std::shared_ptr<const ErrorValue> wrapped;   // fmt.Errorf("...: %w", err)
};
using error = std::shared_ptr<const ErrorValue>;   // nullptr == nil

namespace errors {
error New(const std::string &msg);
error Errorf(const std::string &prefix, const error &err);   // fmt.Errorf(prefix + ": %w", err)
bool Is(const error &err, const error &target);
}  // namespace errors

namespace io {
extern const error EOF_;               // io.EOF (EOF is a macro in <cstdio>)
extern const error ErrUnexpectedEOF;   // io.ErrUnexpectedEOF
struct Reader {
    virtual ~Reader() = default;
    virtual std::pair<int, error> Read(uint8_t *p, size_t len) = 0;
};
struct ByteReader {
    virtual ~ByteReader() = default;
    virtual std::pair<uint8_t, error> ReadByte() = 0;
};
struct Closer {
    virtual ~Closer() = default;
    virtual error Close() = 0;
};
struct ReadCloser : Reader, Closer {};
struct Writer {
    virtual ~Writer() = default;
    virtual std::pair<int, error> Write(const uint8_t *p, size_t len) = 0;
};
// io.Copy with a 32 KiB buffer (what the reference's tests drive the readers with)
std::pair<int64_t, error> Copy(Writer &dst, Reader &src);

// bytes.Reader / bufio.Reader stand-in over memory: both interfaces
class BytesReader : public Reader, public ByteReader {
  public:
    BytesReader(const uint8_t *data, size_t len) : d_(data), n_(len) {}
    explicit BytesReader(const std::vector<uint8_t> &v) : d_(v.data()), n_(v.size()) {}
    std::pair<int, error> Read(uint8_t *p, size_t len) override;
    std::pair<uint8_t, error> ReadByte() override;
  private:
    const uint8_t *d_;
    size_t n_, pos_ = 0;
};
}  // namespace io

// errors.go:5-12
extern const error ErrCorrupted, ErrIncorrectProperties, ErrResultError, ErrDictOutOfRange, ErrUnexpectedLZMA2Code,
    ErrNoLZMAReader;
// unexported in the reference (reader1.go:26, reader2.go:43, readcloser.go:14); visible for the tests
extern const error errNeedOneReader, errInsufficientProperties, errAlreadyClosed;
// new: a batch call needs an output capacity; the streaming readers never surface it
extern const error ErrOutputOverflow;

// ---- the engine (new): one per process is enough; readers use Default() unless given another -----
struct Unit : lzgpu_unit {};
struct Result : lzgpu_result {};

class Engine {
  public:
    // devices: CUDA ordinals; empty = every visible device
    static std::pair<std::shared_ptr<Engine>, error> New(const std::vector<int> &devices = {});
    static std::pair<std::shared_ptr<Engine>, error> Default();
    ~Engine();
    // the batch entry point: all units in one lzgpu_decode_batch call (sharded over the engine's GPUs)
    std::pair<std::vector<Result>, error> DecodeBatch(const std::vector<Unit> &units, const uint8_t *in, size_t in_len,
                                                      uint8_t *out, size_t out_len);
    // All folders (coders) of a 7z archive in ONE GPU call -- bodgit/sevenzip hands each folder to a registered
    // decompressor separately (reader1.go:28-61, reader2.go:45-75), which is one reader and, here, one GPU call per
    // folder; an archive reader that knows its folders up front decodes them together.  LZMA folders (props = 5
    // bytes) become headerless LZMA1 units, LZMA2 folders (props = 1 byte) are cut at their dictionary resets.
    // Property errors are the constructors' (ErrIncorrectProperties, errInsufficientProperties); decode errors are
    // the readers' (ErrResultError, io::ErrUnexpectedEOF ...), with the bytes decoded before them.
    struct Folder {
        bool lzma2 = false;
        std::vector<uint8_t> props;
        uint64_t unpackSize = 0;      // LZMA: from the archive header; LZMA2: ignored (the chunk headers say)
        const uint8_t *packed = nullptr;
        size_t packedLen = 0;
    };
    struct FolderResult {
        std::vector<uint8_t> out;
        error err;
    };
    std::pair<std::vector<FolderResult>, error> DecodeFolders(const std::vector<Folder> &folders);
    int Devices() const;
    static error StatusError(int status);   // lzgpu_status -> the error value the reference returns
  private:
    Engine() = default;
    lzgpu_ctx *ctx_ = nullptr;
};

// decoded bytes of a reader: page-locked memory (lzgpu_alloc_pinned) when large, so that the batch call streams
// them back while the kernel runs; ordinary memory otherwise or when pinning fails
class Bytes {
  public:
    Bytes() = default;
    Bytes(const Bytes &) = delete;
    Bytes &operator=(const Bytes &) = delete;
    Bytes(Bytes &&o) noexcept : p_(o.p_), size_(o.size_), cap_(o.cap_), pinned_(o.pinned_) { o.p_ = nullptr; o.size_ = o.cap_ = 0; o.pinned_ = false; }
    Bytes &operator=(Bytes &&o) noexcept {
        if (this != &o) {
            release();
            p_ = o.p_; size_ = o.size_; cap_ = o.cap_; pinned_ = o.pinned_;
            o.p_ = nullptr; o.size_ = o.cap_ = 0; o.pinned_ = false;
        }
        return *this;
    }
    ~Bytes() { release(); }
    void reset(size_t n);                    // contents undefined afterwards
    void truncate(size_t n) { if (n < size_) size_ = n; }
    void clear() { size_ = 0; }
    uint8_t *data() { return p_; }
    const uint8_t *data() const { return p_; }
    size_t size() const { return size_; }
    static void TrimCache();                 // unpin the released buffers kept for reuse (lzma_reader.cpp, PinnedCache)
  private:
    void release();
    uint8_t *p_ = nullptr;
    size_t size_ = 0, cap_ = 0;
    bool pinned_ = false;
};

// a growing byte buffer over Bytes (so that large ones are page-locked and come from the process-wide cache)
class InBuf {
  public:
    size_t size() const { return len_; }
    uint8_t *data() { return mem_.data(); }
    uint8_t &operator[](size_t i) { return mem_.data()[i]; }
    void assign(const uint8_t *p, size_t n) { len_ = 0; append(p, n); }
    void append(const uint8_t *p, size_t n) { reserve(len_ + n); if (n) memcpy(mem_.data() + len_, p, n); len_ += n; }
    void reserve(size_t cap);                 // keeps the contents
    void set_size(size_t n) { len_ = n; }     // n <= capacity
    size_t capacity() const { return cap_; }
    Bytes release() { len_ = cap_ = 0; return std::move(mem_); }
    void adopt(Bytes b, size_t cap) { mem_ = std::move(b); cap_ = cap; len_ = 0; }
  private:
    Bytes mem_;
    size_t len_ = 0, cap_ = 0;
};

// ---- reader1.go ---------------------------------------------------------------------------------
std::tuple<uint8_t, uint8_t, uint8_t, error> DecodeProp(uint8_t d);        // (lc, pb, lp, err)
std::pair<uint32_t, error> DecodeDictSize(const uint8_t properties[4]);
uint64_t DecodeUnpackSize(const uint8_t header[8]);
uint32_t DecodeDictSize2(uint8_t encodedDictSize);

class Reader1 : public io::Reader {
  public:
    std::pair<int, error> Read(uint8_t *p, size_t len) override;
    void Reset();
    error Reopen(io::ByteReader &inStream, uint64_t unpackSize);
    bool isEndOfStream = false;

  private:
    friend std::pair<std::unique_ptr<Reader1>, error> NewReader1(io::ByteReader &, std::shared_ptr<Engine>);
    friend std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMADecompressorForSevenZip(
        const std::vector<uint8_t> &, uint64_t, const std::vector<io::ReadCloser *> &, std::shared_ptr<Engine>);
    error initialize();   // range-coder preamble, reader1.go:149-159
    void decode();
    io::ByteReader *in_ = nullptr;
    std::unique_ptr<io::ByteReader> owned_in_;
    std::shared_ptr<Engine> eng_;
    uint8_t lc_ = 0, lp_ = 0, pb_ = 0;
    uint32_t dict_ = 0;
    uint64_t unpack_ = ~0ull;
    std::vector<uint8_t> payload_;
    Bytes out_;
    size_t pos_ = 0;
    bool decoded_ = false;
    error err_;
};

std::pair<std::unique_ptr<Reader1>, error> NewReader1(io::ByteReader &inStream, std::shared_ptr<Engine> eng = nullptr);
std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMADecompressorForSevenZip(
    const std::vector<uint8_t> &props, uint64_t unpackSize, const std::vector<io::ReadCloser *> &readers,
    std::shared_ptr<Engine> eng = nullptr);

// ---- reader2.go ---------------------------------------------------------------------------------
class Reader2 : public io::Reader {
  public:
    Reader2();
    ~Reader2();
    std::pair<int, error> Read(uint8_t *p, size_t len) override;
    size_t wave_bytes = (size_t)1 << 30;   // decoded bytes per GPU call (at least one unit).  A wave takes as long as its longest unit (~100 ms per MiB of text), however many units it has: large waves are what gives throughput
    bool decode_ahead = true;         // wave k+1 is decoded and wave k+2 read from the input while wave k is being served

  private:
    friend std::pair<std::unique_ptr<Reader2>, error> NewReader2(io::Reader &, int, std::shared_ptr<Engine>);
    friend std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMA2DecompressorForSevenZip(
        const std::vector<uint8_t> &, uint64_t, const std::vector<io::ReadCloser *> &, std::shared_ptr<Engine>);
    error initialize();   // validateDictSize + the first chunk header, reader2.go:77-173
    bool fill(size_t need);
    bool independentFrom(size_t pos);
    bool nextWave(size_t &start, size_t &end, bool &term);
    struct Wave {
        Bytes in;            // holds the wave's compressed bytes, terminated, at [in_off, in_off + in_len) (page-locked when
        size_t in_off = 0;   //  large: the kernel reads them in place)
        size_t in_len = 0;
        Bytes out;           // its decoded bytes
        error err;
        bool last = false;
        double cut_ms = 0;
    };
    std::unique_ptr<Wave> cutWave(std::unique_ptr<Wave> w);                // stage 1: input -> wave bytes
    std::unique_ptr<Wave> decodeWave(std::unique_ptr<Wave> w, Bytes out);  // stage 2: one GPU call
    void advance(std::unique_ptr<Wave> done);                              // stage 3 takes the next decoded wave
    void startPipeline();
    struct Mailbox;
    io::Reader *in_ = nullptr;
    std::shared_ptr<Engine> eng_;
    uint32_t dict_ = 0;
    InBuf buf_;                           // input read ahead (touched only by the cut in flight: never two at a time)
    size_t rd_ = 0, pos_ = 0;
    bool in_eof_ = false;
    std::unique_ptr<Wave> cur_;           // being served
    std::unique_ptr<Mailbox> cut_box_, dec_box_;   // cutter -> decoder -> Read
    std::thread cutter_, decoder_;
    std::mutex free_mu_;
    std::vector<Bytes> free_in_, free_out_;        // buffers of delivered waves
};

std::pair<std::unique_ptr<Reader2>, error> NewReader2(io::Reader &inStream, int dictSize, std::shared_ptr<Engine> eng = nullptr);
std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMA2DecompressorForSevenZip(
    const std::vector<uint8_t> &props, uint64_t unused, const std::vector<io::ReadCloser *> &readers,
    std::shared_ptr<Engine> eng = nullptr);

// ---- readcloser.go ------------------------------------------------------------------------------
class readCloser : public io::ReadCloser {
  public:
    readCloser(io::Closer *c, std::unique_ptr<io::Reader> r) : c_(c), r_(std::move(r)) {}
    error Close() override;
    std::pair<int, error> Read(uint8_t *p, size_t len) override;
  private:
    io::Closer *c_;
    std::unique_ptr<io::Reader> r_;
};

}  // namespace lzma
