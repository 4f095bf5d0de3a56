// lzma_reader.cpp -- the reader API of kulaginds/lzma in C++ over liblzgpu.so (see include/lzma_reader.hpp).
// Host glue only: header parsing, input slurping, LZMA2 wave cutting, status -> error mapping.  Every decoded
// byte comes from lzgpu_decode_batch; there is no CPU decode path here either.
#include "lzma_reader.hpp"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

namespace lzma {

// ---------------------------------------------------------------- errors / io
namespace errors {
error New(const std::string &msg) { return std::make_shared<const ErrorValue>(ErrorValue{msg, nullptr}); }
error Errorf(const std::string &prefix, const error &err) {
    return std::make_shared<const ErrorValue>(ErrorValue{prefix + ": " + (err ? err->msg : "<nil>"), err});
}
bool Is(const error &err, const error &target) {
    for (const ErrorValue *e = err.get(); e; e = e->wrapped.get())
        if (e == target.get()) return true;
    return false;
}
}  // namespace errors

namespace io {
const error EOF_ = errors::New("EOF");
const error ErrUnexpectedEOF = errors::New("unexpected EOF");

std::pair<int64_t, error> Copy(Writer &dst, Reader &src) {
    std::vector<uint8_t> buf(32 * 1024);
    int64_t written = 0;
    for (;;) {
        auto [n, rerr] = src.Read(buf.data(), buf.size());
        if (n > 0) {
            auto [w, werr] = dst.Write(buf.data(), (size_t)n);
            written += w;
            if (werr) return {written, werr};
        }
        if (rerr) return {written, errors::Is(rerr, EOF_) ? nullptr : rerr};
    }
}

std::pair<int, error> BytesReader::Read(uint8_t *p, size_t len) {
    if (pos_ >= n_) return {0, len ? EOF_ : nullptr};
    const size_t n = std::min(len, n_ - pos_);
    memcpy(p, d_ + pos_, n);
    pos_ += n;
    return {(int)n, nullptr};
}
std::pair<uint8_t, error> BytesReader::ReadByte() {
    if (pos_ >= n_) return {0, EOF_};
    return {d_[pos_++], nullptr};
}
}  // namespace io

const error ErrCorrupted = errors::New("corrupted");
const error ErrIncorrectProperties = errors::New("incorrect LZMA properties");
const error ErrResultError = errors::New("result error");
const error ErrDictOutOfRange = errors::New("dictionary capacity is out of range");
const error ErrUnexpectedLZMA2Code = errors::New("unexpected lzma2 code");
const error ErrNoLZMAReader = errors::New("no lzma reader on chunkLZMAResetState");
const error errNeedOneReader = errors::New("lzma: need exactly one reader");
const error errInsufficientProperties = errors::New("lzma2: not enough properties");
const error errAlreadyClosed = errors::New("lzma: already closed");
const error ErrOutputOverflow = errors::New("lzgpu: output capacity too small");

namespace {
// bufio.NewReader(r) as far as the readers need it: a ByteReader over a Reader
class bufReader : public io::ByteReader, public io::Reader {
  public:
    explicit bufReader(io::Reader *r) : r_(r), buf_(64 * 1024) {}
    std::pair<uint8_t, error> ReadByte() override {
        if (lo_ == hi_) {
            if (err_) return {0, err_};
            auto [n, e] = r_->Read(buf_.data(), buf_.size());
            lo_ = 0;
            hi_ = n > 0 ? (size_t)n : 0;
            if (e) err_ = e;
            else if (n == 0) err_ = io::EOF_;
            if (lo_ == hi_) return {0, err_};
        }
        return {buf_[lo_++], nullptr};
    }
    std::pair<int, error> Read(uint8_t *p, size_t len) override {
        if (lo_ < hi_) {
            const size_t n = std::min(len, hi_ - lo_);
            memcpy(p, buf_.data() + lo_, n);
            lo_ += n;
            return {(int)n, nullptr};
        }
        if (err_) return {0, err_};
        return r_->Read(p, len);
    }
  private:
    io::Reader *r_;
    std::vector<uint8_t> buf_;
    size_t lo_ = 0, hi_ = 0;
    error err_;
};

// everything the stream still holds, appended to `to`
void slurp(io::ByteReader *br, std::vector<uint8_t> &to) {
    if (auto *r = dynamic_cast<io::Reader *>(br)) {
        std::vector<uint8_t> chunk(1 << 20);
        for (;;) {
            auto [n, e] = r->Read(chunk.data(), chunk.size());
            if (n > 0) to.insert(to.end(), chunk.begin(), chunk.begin() + n);
            if (e || n == 0) return;
        }
    }
    for (;;) {
        auto [b, e] = br->ReadByte();
        if (e) return;
        to.push_back(b);
    }
}

size_t read_exact(io::Reader *r, uint8_t *p, size_t n) {
    size_t got = 0;
    while (got < n) {
        auto [k, e] = r->Read(p + got, n - got);
        if (k > 0) got += (size_t)k;
        if (e || k == 0) break;
    }
    return got;
}
}  // namespace

// ---------------------------------------------------------------- Bytes
// Page-locking memory costs ~0.4 s per GiB (every page is faulted in and pinned), as much as decoding it: released
// buffers are kept for the next reader of the process (a handful, bounded in bytes) instead of being unpinned.
namespace {
struct PinnedCache {
    std::mutex mu;
    std::vector<std::pair<uint8_t *, size_t>> free_;   // (pointer, capacity)
    size_t bytes = 0;
    static constexpr size_t kMaxBytes = (size_t)8 << 30, kMaxBuffers = 8;
    ~PinnedCache() { /* process exit: the driver is shutting down, leave the pages to the OS */ }
    uint8_t *take(size_t want, size_t *cap) {
        std::lock_guard<std::mutex> lock(mu);
        size_t best = free_.size();
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].second >= want && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
        if (best == free_.size()) return nullptr;
        uint8_t *p = free_[best].first;
        *cap = free_[best].second;
        bytes -= *cap;
        free_.erase(free_.begin() + best);
        return p;
    }
    bool give(uint8_t *p, size_t cap) {
        std::lock_guard<std::mutex> lock(mu);
        if (free_.size() >= kMaxBuffers || bytes + cap > kMaxBytes) return false;
        free_.push_back({p, cap});
        bytes += cap;
        return true;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto &b : free_) lzgpu_free_pinned(b.first);
        free_.clear();
        bytes = 0;
    }
};
PinnedCache &pinned_cache() {
    static PinnedCache *c = new PinnedCache();   // never destroyed: see ~PinnedCache
    return *c;
}
}  // namespace

void Bytes::TrimCache() { pinned_cache().trim(); }

void Bytes::release() {
    if (p_) {
        if (pinned_) { if (!pinned_cache().give(p_, cap_)) lzgpu_free_pinned(p_); }
        else free(p_);
    }
    p_ = nullptr;
    size_ = cap_ = 0;
    pinned_ = false;
}
void Bytes::reset(size_t n) {
    if (n <= cap_) { size_ = n; return; }
    release();
    const size_t want = std::max<size_t>(n, 16);
    // LZMA_READER_NO_PIN=1: ordinary memory only.  Page-locking is what lets the output stream back while the kernel
    // runs and the input be read in place, but the first use of a buffer costs ~1 s per GiB: a one-shot tool that
    // decodes a single stream and exits is better off without, a long-lived process keeps its buffers.
    static const bool no_pin = getenv("LZMA_READER_NO_PIN") != nullptr;
    if (want >= (32u << 20) && !no_pin) {
        size_t cap = 0;
        p_ = pinned_cache().take(want, &cap);
        if (p_) { pinned_ = true; cap_ = cap; size_ = n; return; }
        p_ = static_cast<uint8_t *>(lzgpu_alloc_pinned(want));
        pinned_ = p_ != nullptr;
    }
    if (!p_) p_ = static_cast<uint8_t *>(malloc(want));
    if (!p_) throw std::bad_alloc();
    cap_ = want;
    size_ = n;
}

// ---------------------------------------------------------------- engine
std::pair<std::shared_ptr<Engine>, error> Engine::New(const std::vector<int> &devices) {
    std::shared_ptr<Engine> e(new Engine());
    const int rc = lzgpu_ctx_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &e->ctx_);
    if (rc != LZGPU_E_OK) return {nullptr, errors::New(std::string("lzgpu: ") + lzgpu_last_error())};
    return {e, nullptr};
}
std::pair<std::shared_ptr<Engine>, error> Engine::Default() {
    static std::mutex mu;
    static std::shared_ptr<Engine> def;
    std::lock_guard<std::mutex> lock(mu);
    if (def) return {def, nullptr};
    auto [e, err] = New();
    if (!err) def = e;
    return {e, err};
}
Engine::~Engine() {
    if (ctx_) lzgpu_ctx_destroy(ctx_);
}
int Engine::Devices() const { return lzgpu_ctx_device_count(ctx_); }

std::pair<std::vector<Result>, error> Engine::DecodeBatch(const std::vector<Unit> &units, const uint8_t *in, size_t in_len,
                                                          uint8_t *out, size_t out_len) {
    std::vector<Result> res(units.size());
    static const uint8_t nothing[16] = {0};
    const int rc = lzgpu_decode_batch(ctx_, units.data(), (int64_t)units.size(), in ? in : nothing, in_len, out, out_len,
                                      res.data(), nullptr);
    if (rc != LZGPU_E_OK) return {{}, errors::New(std::string("lzgpu: ") + lzgpu_last_error())};
    return {std::move(res), nullptr};
}

std::pair<std::vector<Engine::FolderResult>, error> Engine::DecodeFolders(const std::vector<Folder> &folders) {
    std::vector<FolderResult> res(folders.size());
    struct Span { size_t first = 0, count = 0; uint64_t out_off = 0; bool lzma1 = false; uint64_t cap = 0; };
    std::vector<Span> spans(folders.size());
    std::vector<size_t> todo;
    for (size_t i = 0; i < folders.size(); i++) {
        const Folder &f = folders[i];
        if (f.lzma2) {   // NewLZMA2DecompressorForSevenZip, reader2.go:49-75
            if (f.props.size() != 1) { res[i].err = errInsufficientProperties; continue; }
        } else {         // NewLZMADecompressorForSevenZip, reader1.go:32-61
            if (f.props.size() < 5) { res[i].err = ErrIncorrectProperties; continue; }
            if (std::get<3>(DecodeProp(f.props[0]))) { res[i].err = ErrIncorrectProperties; continue; }
            // the size field of an archive header is untrusted: first capacity bounded by the packed size (as Reader1)
            spans[i].cap = std::min<uint64_t>(f.unpackSize, std::max<uint64_t>(1 << 16, 8 * (uint64_t)f.packedLen));
        }
        todo.push_back(i);
    }
    static const uint8_t nothing[16] = {0};
    while (!todo.empty()) {
        // one input buffer (the folders' packed bytes, 16-byte aligned), one output buffer, one call
        std::vector<Unit> units;
        uint64_t in_size = 0, out_size = 0;
        for (size_t i : todo) {
            const Folder &f = folders[i];
            Span &sp = spans[i];
            sp.first = units.size();
            sp.out_off = out_size;
            if (f.lzma2) {
                std::vector<Unit> us(16);
                uint64_t total = 0;
                int32_t sst = 0;
                const uint8_t *pk = f.packed ? f.packed : nothing;
                int64_t n = lzgpu_scan_lzma2(pk, f.packedLen, DecodeDictSize2(f.props[0]), us.data(), (int64_t)us.size(), &total, &sst);
                if (n > (int64_t)us.size()) {
                    us.resize((size_t)n);
                    n = lzgpu_scan_lzma2(pk, f.packedLen, DecodeDictSize2(f.props[0]), us.data(), (int64_t)us.size(), &total, &sst);
                }
                for (int64_t k = 0; k < n; k++) {
                    us[(size_t)k].in_off += in_size;
                    us[(size_t)k].out_off += out_size;
                    units.push_back(us[(size_t)k]);
                }
                sp.count = (size_t)std::max<int64_t>(n, 0);
                out_size += (total + 15) & ~(uint64_t)15;
            } else {
                Unit u;
                memset(&u, 0, sizeof u);
                u.kind = LZGPU_KIND_LZMA1_RAW;
                auto [lc, pb, lp, perr] = DecodeProp(f.props[0]);
                (void)perr;
                u.lc = lc; u.lp = lp; u.pb = pb;
                u.dict_size = DecodeDictSize(f.props.data() + 1).first;
                u.unpack_size = f.unpackSize;
                u.in_off = in_size;
                u.in_len = f.packedLen;
                u.out_off = out_size;
                u.out_cap = sp.cap;
                units.push_back(u);
                sp.count = 1;
                sp.lzma1 = true;
                out_size += (sp.cap + 15) & ~(uint64_t)15;
            }
            in_size += ((uint64_t)f.packedLen + 15) & ~(uint64_t)15;
        }
        Bytes in, out;
        in.reset((size_t)in_size + 16);
        out.reset((size_t)out_size + 16);
        uint64_t off = 0;
        for (size_t i : todo) {
            if (folders[i].packedLen) memcpy(in.data() + off, folders[i].packed, folders[i].packedLen);
            off += ((uint64_t)folders[i].packedLen + 15) & ~(uint64_t)15;
        }
        auto [r, err] = DecodeBatch(units, in.data(), in.size(), out.data(), out.size());
        if (err) return {{}, err};
        std::vector<size_t> again;
        for (size_t i : todo) {
            Span &sp = spans[i];
            if (sp.lzma1 && r[sp.first].status == LZGPU_OUTPUT_OVERFLOW && sp.cap < folders[i].unpackSize) {
                sp.cap = std::min<uint64_t>(sp.cap * 8, folders[i].unpackSize);   // the capacity guess was too small: this folder again
                again.push_back(i);
                continue;
            }
            uint64_t n_out = 0;
            error e;
            for (size_t k = sp.first; k < sp.first + sp.count; k++) {   // a folder's units are consecutive, and so are their outputs
                n_out = units[k].out_off - sp.out_off + r[k].bytes_out;
                if (r[k].status != LZGPU_OK) { e = StatusError(r[k].status); break; }
            }
            res[i].out.assign(out.data() + sp.out_off, out.data() + sp.out_off + n_out);
            res[i].err = e;
        }
        todo.swap(again);
    }
    return {std::move(res), nullptr};
}

error Engine::StatusError(int status) {
    switch (status) {
    case LZGPU_OK:
    case LZGPU_OK_INPUT_EXHAUSTED: return nullptr;   // truncated input is a clean EOF in the reference (reader1.go:246-249)
    case LZGPU_RESULT_ERROR: return ErrResultError;
    case LZGPU_INCORRECT_PROPERTIES: return ErrIncorrectProperties;
    case LZGPU_UNEXPECTED_EOF: return io::ErrUnexpectedEOF;
    case LZGPU_OUTPUT_OVERFLOW: return ErrOutputOverflow;
    default: return errors::New("lzgpu: unexpected status " + std::to_string(status));
    }
}

// ---------------------------------------------------------------- reader1.go
std::tuple<uint8_t, uint8_t, uint8_t, error> DecodeProp(uint8_t d) {   // reader1.go:210-221
    uint8_t lc, pb, lp;
    if (lzgpu_decode_prop(d, &lc, &pb, &lp) != LZGPU_OK) return {0, 0, 0, ErrIncorrectProperties};
    return {lc, pb, lp, nullptr};
}
std::pair<uint32_t, error> DecodeDictSize(const uint8_t properties[4]) { return {lzgpu_decode_dict_size(properties), nullptr}; }
uint64_t DecodeUnpackSize(const uint8_t header[8]) { return lzgpu_decode_unpack_size(header); }
uint32_t DecodeDictSize2(uint8_t encodedDictSize) { return lzgpu_decode_dict_size2(encodedDictSize); }

error Reader1::initialize() {   // rangeDecoder.Init: 5 bytes, the first must be 0 (range_decoder.go:27-46)
    payload_.clear();
    for (int i = 0; i < 5; i++) {
        auto [b, e] = in_->ReadByte();
        if (e) {
            if (!payload_.empty() && payload_[0] != 0) return errors::Errorf("rangeDec.Init", ErrResultError);
            return errors::Errorf("rangeDec.Init", e);
        }
        payload_.push_back(b);
    }
    if (payload_[0] != 0) return errors::Errorf("rangeDec.Init", ErrResultError);
    return nullptr;
}

void Reader1::decode() {
    decoded_ = true;
    slurp(in_, payload_);
    if (!eng_) {
        auto [e, err] = Engine::Default();
        if (err) { err_ = err; return; }
        eng_ = e;
    }
    Unit u;
    memset(&u, 0, sizeof u);
    u.kind = LZGPU_KIND_LZMA1_RAW;
    u.lc = lc_; u.lp = lp_; u.pb = pb_;
    u.dict_size = dict_;
    u.unpack_size = unpack_;
    u.in_off = 0;
    u.in_len = payload_.size();
    const bool known = unpack_ != ~0ull;
    uint64_t cap = std::max<uint64_t>(1 << 16, 8 * (uint64_t)payload_.size());
    if (known) cap = std::min<uint64_t>(cap, std::max<uint64_t>(unpack_, 1));
    for (;;) {
        u.out_off = 0;
        u.out_cap = cap;
        out_.reset((size_t)std::max<uint64_t>(cap, 16));
        auto [res, err] = eng_->DecodeBatch({u}, payload_.data(), payload_.size(), out_.data(), out_.size());
        if (err) { out_.clear(); err_ = err; return; }
        const bool can_grow = known ? cap < unpack_ : cap < (1ull << 40);
        if (res[0].status == LZGPU_OUTPUT_OVERFLOW && can_grow) {   // the streaming reader has no capacity: grow, decode again
            cap = known ? std::min<uint64_t>(cap * 8, unpack_) : cap * 8;
            continue;
        }
        out_.truncate((size_t)res[0].bytes_out);
        err_ = Engine::StatusError(res[0].status);
        break;
    }
    payload_.clear();
    payload_.shrink_to_fit();
}

std::pair<int, error> Reader1::Read(uint8_t *p, size_t len) {   // reader1.go:223-254
    if (!decoded_) decode();
    const size_t n = std::min(len, out_.size() - pos_);
    if (n) {
        memcpy(p, out_.data() + pos_, n);
        pos_ += n;
    }
    if (n == len && n > 0) return {(int)n, nullptr};
    isEndOfStream = true;
    if (err_) {
        error e = err_;
        err_ = nullptr;
        return {(int)n, e};
    }
    return {(int)n, io::EOF_};
}

void Reader1::Reset() { isEndOfStream = false; }

error Reader1::Reopen(io::ByteReader &inStream, uint64_t unpackSize) {   // reader1.go:166-176
    isEndOfStream = false;
    in_ = &inStream;
    owned_in_.reset();
    unpack_ = unpackSize;
    out_.clear();
    pos_ = 0;
    decoded_ = false;
    err_ = nullptr;
    return initialize();
}

std::pair<std::unique_ptr<Reader1>, error> NewReader1(io::ByteReader &inStream, std::shared_ptr<Engine> eng) {
    // reader1.go:18-24 + initializeFull (:77-101): header and range-coder preamble read eagerly
    std::unique_ptr<Reader1> r(new Reader1());
    r->in_ = &inStream;
    r->eng_ = std::move(eng);
    uint8_t h[13];
    for (int i = 0; i < 13; i++) {
        auto [b, e] = inStream.ReadByte();
        if (e) {
            if (i == 0) return {std::move(r), e};
            return {std::move(r), errors::Errorf(i < 5 ? "decode dict size" : "decode unpack size", e)};
        }
        h[i] = b;
        if (i == 0) {
            auto [lc, pb, lp, perr] = DecodeProp(b);
            if (perr) return {std::move(r), errors::Errorf("decode prop", perr)};
            r->lc_ = lc; r->pb_ = pb; r->lp_ = lp;
        }
    }
    r->dict_ = DecodeDictSize(h + 1).first;
    r->unpack_ = DecodeUnpackSize(h + 5);
    error e = r->initialize();
    return {std::move(r), e};
}

std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMADecompressorForSevenZip(
    const std::vector<uint8_t> &props, uint64_t unpackSize, const std::vector<io::ReadCloser *> &readers,
    std::shared_ptr<Engine> eng) {   // reader1.go:32-61
    if (readers.size() != 1) return {nullptr, errNeedOneReader};
    if (props.size() < 5) return {nullptr, ErrIncorrectProperties};
    auto [lc, pb, lp, perr] = DecodeProp(props[0]);
    if (perr) return {nullptr, perr};
    auto [dictSize, derr] = DecodeDictSize(props.data() + 1);
    if (derr) return {nullptr, derr};
    std::unique_ptr<Reader1> r(new Reader1());
    if (auto *br = dynamic_cast<io::ByteReader *>(readers[0])) r->in_ = br;
    else {
        r->owned_in_.reset(new bufReader(readers[0]));
        r->in_ = r->owned_in_.get();
    }
    r->eng_ = std::move(eng);
    r->lc_ = lc; r->lp_ = lp; r->pb_ = pb;
    r->dict_ = dictSize;
    r->unpack_ = unpackSize;
    error e = r->initialize();
    return {std::unique_ptr<io::ReadCloser>(new readCloser(readers[0], std::move(r))), e};
}

// ---------------------------------------------------------------- reader2.go
error Reader2::initialize() {
    if (dict_ < (1u << 12)) dict_ = 8u << 20;   // reader2.go:88-91
    uint8_t c;
    if (read_exact(in_, &c, 1) < 1) return io::ErrUnexpectedEOF;   // reader2.go:103-110
    buf_.assign(&c, 1);
    if (c == 0 || (c >= 3 && c < 0x80)) return nullptr;
    const size_t hl = c < 0x80 ? 3 : (c < 0xC0 ? 5 : 6);
    uint8_t rest[6];
    const size_t got = read_exact(in_, rest, hl - 1);
    buf_.append(rest, got);
    if (got < hl - 1) return io::ErrUnexpectedEOF;   // reader2.go:121-128
    if (c >= 0x80) {
        // first LZMA chunk: NewReader1ForReader2 -> DecodeProp + rangeDec.Init (reader2.go:146-153)
        const uint8_t prop = hl == 6 ? buf_[5] : 0;
        if (prop >= 225) return ErrIncorrectProperties;
        uint8_t pre;
        if (read_exact(in_, &pre, 1) < 1) return errors::Errorf("rangeDec.Init", io::EOF_);
        buf_.append(&pre, 1);
        if (pre != 0) return errors::Errorf("rangeDec.Init", ErrResultError);
    }
    return nullptr;
}

bool Reader2::fill(size_t need) {   // at least `need` unread bytes in buf_ (false: the input ended first)
    while (buf_.size() - rd_ < need && !in_eof_) {
        const size_t want = std::max<size_t>(4 << 20, need - (buf_.size() - rd_));
        const size_t old = buf_.size();
        // an input buffer is sized once for a whole wave (text compresses 3-4x: half the wave's output is ample), so that
        // the same few page-locked buffers circulate for the life of the reader -- and, through the cache, of the process
        buf_.reserve(std::max(old + want, std::min<size_t>(wave_bytes / 2, (size_t)1 << 30) + (64u << 10)));
        auto [n, e] = in_->Read(buf_.data() + old, want);
        buf_.set_size(old + (n > 0 ? (size_t)n : 0));
        if (e || n <= 0) in_eof_ = true;
    }
    return buf_.size() - rd_ >= need;
}

// Uncompressed chunk with dictionary reset at `pos`: does the first LZMA chunk after it (if one comes before the
// next reset) reset the state and carry properties?  Otherwise it inherits the previous unit's coder
// (reader2.go:155-165) and must stay in the same wave.
bool Reader2::independentFrom(size_t pos) {
    const size_t p0 = pos;
    for (;;) {
        rd_ = p0;
        if (!fill(pos - p0 + 3)) return true;
        const uint8_t c = buf_[pos];
        if (c == 0 || (c >= 3 && c < 0x80) || c >= 0xE0 || (c == 1 && pos != p0)) return true;
        if (c >= 0x80) return c >= 0xC0;
        pos += 3 + (((size_t)buf_[pos + 1] << 8) | buf_[pos + 2]) + 1;
    }
}

// Walk chunk headers (reader2.go:100-214) from rd_ until wave_bytes of output are covered and the next chunk starts
// a unit that inherits nothing.  The wave is buf_[start, end) -- decoded where it lies, no copy -- and needs a 0x00
// terminator written at buf_[end] when `term` (the byte there is the next wave's first control byte: the caller saves
// and restores it).  Returns true when this wave is the stream's last.
bool Reader2::nextWave(size_t &start, size_t &end, bool &term) {
    start = rd_;
    term = false;
    size_t pos = rd_, out = 0;
    bool first = true;
    for (;;) {
        rd_ = pos;
        if (!fill(1)) { end = buf_.size(); rd_ = buf_.size(); return true; }   // ran off the input: the scanner reports it
        const uint8_t ctrl = buf_[pos];
        if (ctrl == 0 || (ctrl >= 3 && ctrl < 0x80)) {   // end of stream (0x03-0x7F too, reader2.go:185-198)
            rd_ = pos + 1;
            end = pos + 1;
            return true;
        }
        const bool enough = !first && out >= wave_bytes;
        const bool reset = ctrl >= 0xE0 || (ctrl == 1 && enough && independentFrom(pos));
        rd_ = pos;
        if (reset && enough) {   // the wave ends before this chunk; it is terminated there
            end = pos;
            term = true;
            return false;
        }
        const size_t hl = ctrl < 0x80 ? 3 : (ctrl < 0xC0 ? 5 : 6);
        if (!fill(hl)) { end = buf_.size(); rd_ = buf_.size(); return true; }
        size_t usz = (((size_t)buf_[pos + 1] << 8) | buf_[pos + 2]) + 1, payload;
        if (ctrl >= 0x80) {
            usz += (size_t)(ctrl & 0x1F) << 16;
            payload = (((size_t)buf_[pos + 3] << 8) | buf_[pos + 4]) + 1;
        } else payload = usz;
        if (!fill(hl + payload)) { end = buf_.size(); rd_ = buf_.size(); return true; }
        pos += hl + payload;
        out += usz;
        first = false;
    }
}

// Stage 1 of the reader's pipeline: find the next wave in the input (nextWave) and move its bytes, terminated, into
// the wave's own buffer -- page-locked when large, so that the decode kernel reads it straight from host memory.
// `w` brings the buffers of a delivered wave back into circulation.
std::unique_ptr<Reader2::Wave> Reader2::cutWave(std::unique_ptr<Wave> w) {
    if (!w) w.reset(new Wave());
    w->err = nullptr;
    w->out.clear();
    size_t start = 0, end = 0;
    bool term = false;
    const auto t0 = std::chrono::steady_clock::now();
    w->last = nextWave(start, end, term);
    // The wave stays where it was read: the input buffer itself goes to the decoder, and what was read beyond the
    // wave (a few MiB at most) moves to a new input buffer -- w->in, a delivered wave's, when it is large enough.
    const size_t tail = buf_.size() - end;
    InBuf nb;
    const size_t want = std::max<size_t>(tail + (1u << 20), std::min<size_t>(wave_bytes / 2, (size_t)1 << 30) + (64u << 10));
    if (w->in.size() >= want) { const size_t c = w->in.size(); nb.adopt(std::move(w->in), c); }
    nb.reserve(want);
    nb.append(buf_.data() + end, tail);
    buf_.reserve(end + 1);
    buf_[end] = 0;                                    // terminator of a wave that is not the stream's last
    w->in_off = start;
    w->in_len = end - start + (term ? 1 : 0);
    w->in = buf_.release();
    buf_ = std::move(nb);
    rd_ = 0;
    w->cut_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return w;
}

void InBuf::reserve(size_t cap) {
    if (cap <= cap_) return;
    size_t want = std::max<size_t>(cap, cap_ * 2);
    if (want < 4096) want = 4096;
    Bytes nb;
    nb.reset(want);
    if (len_) memcpy(nb.data(), mem_.data(), len_);
    mem_ = std::move(nb);
    cap_ = want;
}

// Stage 2: scan the wave into units and decode them in one GPU call, into `out` (a delivered wave's buffer, or empty).
std::unique_ptr<Reader2::Wave> Reader2::decodeWave(std::unique_ptr<Wave> w, Bytes out) {
    static const bool trace = getenv("LZMA_READER_TRACE") != nullptr;   // phase times of every wave on stderr
    const auto t0 = std::chrono::steady_clock::now();
    auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    w->out = std::move(out);
    if (!eng_) {
        auto [e, err] = Engine::Default();
        if (err) { w->err = err; w->last = true; return w; }
        eng_ = e;
    }
    std::vector<Unit> units(64);
    uint64_t total = 0;
    int32_t sst = 0;
    const uint8_t *wave = w->in.data() + w->in_off;
    int64_t n = lzgpu_scan_lzma2(wave, w->in_len, dict_, units.data(), (int64_t)units.size(), &total, &sst);
    if (n > (int64_t)units.size()) {
        units.resize((size_t)n);
        n = lzgpu_scan_lzma2(wave, w->in_len, dict_, units.data(), (int64_t)units.size(), &total, &sst);
    }
    if (n < 0) { w->err = errors::New(std::string("lzgpu: ") + lzgpu_last_error()); w->last = true; return w; }
    units.resize((size_t)n);
    const double t_scan = ms();
    w->out.reset((size_t)std::max<uint64_t>(total, 16));
    const double t_alloc = ms();
    auto [res, err] = eng_->DecodeBatch(units, wave, w->in_len, w->out.data(), w->out.size());
    if (trace) fprintf(stderr, "[reader2] wave of %zu units, %.0f MiB: cut %.1f ms, scan %.1f, buffer %.1f, decode %.1f\n", units.size(),
                       total / 1048576.0, w->cut_ms, t_scan, t_alloc - t_scan, ms() - t_alloc);
    if (err) { w->out.clear(); w->err = err; w->last = true; return w; }
    // the bytes of the units before the first failing one are delivered with the failure, as the reference's
    // reader would have delivered them
    uint64_t n_out = 0;
    int status = sst == LZGPU_OK || n > 0 ? LZGPU_OK : sst;
    for (size_t k = 0; k < units.size(); k++) {
        n_out = units[k].out_off + res[k].bytes_out;
        if (res[k].status != LZGPU_OK) { status = res[k].status; break; }
    }
    w->out.truncate((size_t)n_out);
    w->err = Engine::StatusError(status);
    w->last = w->last || w->err != nullptr;
    return w;
}

// Three stages run at once (the reference streams chunk by chunk, reader2.go:216-250; here a Read would otherwise
// block for a whole wave): wave k is being served to the caller, wave k+1 is being decoded by the GPU, wave k+2 is
// being read from the input -- a steady reader sees the GPU's throughput.  A cutter thread and a decoder thread hand
// waves on through one-slot mailboxes; a stage starts its next wave only when its mailbox is empty, so at most three
// waves exist.  The buffers of delivered waves go back to the stages through `free_in_` / `free_out_`.
struct Reader2::Mailbox {
    std::mutex mu;
    std::condition_variable cv;
    std::unique_ptr<Wave> slot;
    bool closed = false;   // the producer will not put anything more
    bool stop = false;     // the reader is going away
    bool wait_empty() {    // producer: before starting on the next wave
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !slot || stop; });
        return !stop;
    }
    void put(std::unique_ptr<Wave> w, bool last) {
        std::lock_guard<std::mutex> lk(mu);
        slot = std::move(w);
        closed = last;
        cv.notify_all();
    }
    std::unique_ptr<Wave> take() {   // consumer: null when the producer is done
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return slot || closed || stop; });
        std::unique_ptr<Wave> w = std::move(slot);
        cv.notify_all();
        return w;
    }
    void shutdown() {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
};

void Reader2::startPipeline() {
    cut_box_.reset(new Mailbox());
    dec_box_.reset(new Mailbox());
    cutter_ = std::thread([this]() {
        for (;;) {
            if (!cut_box_->wait_empty()) return;
            std::unique_ptr<Wave> w(new Wave());
            { std::lock_guard<std::mutex> lk(free_mu_); if (!free_in_.empty()) { w->in = std::move(free_in_.back()); free_in_.pop_back(); } }
            w = cutWave(std::move(w));
            const bool last = w->last;
            cut_box_->put(std::move(w), last);
            if (last) return;
        }
    });
    decoder_ = std::thread([this]() {
        for (;;) {
            if (!dec_box_->wait_empty()) return;
            std::unique_ptr<Wave> w = cut_box_->take();
            if (!w) { dec_box_->put(nullptr, true); return; }
            Bytes out;
            { std::lock_guard<std::mutex> lk(free_mu_); if (!free_out_.empty()) { out = std::move(free_out_.back()); free_out_.pop_back(); } }
            w = decodeWave(std::move(w), std::move(out));
            { std::lock_guard<std::mutex> lk(free_mu_); free_in_.push_back(std::move(w->in)); }   // the input is done with
            const bool last = w->last;
            dec_box_->put(std::move(w), last);
            if (last) { cut_box_->shutdown(); return; }   // (an error ends the stream: the cutter need not go on)
        }
    });
}

void Reader2::advance(std::unique_ptr<Wave> done) {
    if (done) { std::lock_guard<std::mutex> lk(free_mu_); free_out_.push_back(std::move(done->out)); }
    pos_ = 0;
    if (!decode_ahead) {   // one wave at a time, on the caller's thread
        std::unique_ptr<Wave> w(new Wave());
        if (!free_in_.empty()) { w->in = std::move(free_in_.back()); free_in_.pop_back(); }
        Bytes out;
        if (!free_out_.empty()) { out = std::move(free_out_.back()); free_out_.pop_back(); }
        cur_ = decodeWave(cutWave(std::move(w)), std::move(out));
        free_in_.push_back(std::move(cur_->in));
        return;
    }
    if (!dec_box_) startPipeline();
    cur_ = dec_box_->take();
    if (!cur_) {   // (cannot happen: the decoder always delivers a last wave)
        cur_.reset(new Wave());
        cur_->last = true;
    }
}

Reader2::Reader2() = default;

Reader2::~Reader2() {
    if (dec_box_) {
        cut_box_->shutdown();
        dec_box_->shutdown();
        if (cutter_.joinable()) cutter_.join();
        if (decoder_.joinable()) decoder_.join();
    }
}

std::pair<int, error> Reader2::Read(uint8_t *p, size_t len) {   // reader2.go:216-250
    if (!cur_) advance(nullptr);
    while (pos_ == cur_->out.size() && !cur_->last && len) advance(std::move(cur_));   // previous wave delivered: the next one
    const size_t n = std::min(len, cur_->out.size() - pos_);
    if (n) {
        memcpy(p, cur_->out.data() + pos_, n);
        pos_ += n;
    }
    if (n == len && n > 0) return {(int)n, nullptr};
    if (!cur_->last) return {(int)n, nullptr};
    if (cur_->err) {
        error e = cur_->err;
        cur_->err = nullptr;
        return {(int)n, e};
    }
    return {(int)n, io::EOF_};
}

std::pair<std::unique_ptr<Reader2>, error> NewReader2(io::Reader &inStream, int dictSize, std::shared_ptr<Engine> eng) {
    std::unique_ptr<Reader2> r(new Reader2());   // reader2.go:26-41
    r->in_ = &inStream;
    r->dict_ = (uint32_t)dictSize;
    r->eng_ = std::move(eng);
    error e = r->initialize();
    return {std::move(r), e};
}

std::pair<std::unique_ptr<io::ReadCloser>, error> NewLZMA2DecompressorForSevenZip(
    const std::vector<uint8_t> &props, uint64_t, const std::vector<io::ReadCloser *> &readers,
    std::shared_ptr<Engine> eng) {   // reader2.go:49-75
    if (readers.size() != 1) return {nullptr, errNeedOneReader};
    if (props.size() != 1) return {nullptr, errInsufficientProperties};
    std::unique_ptr<Reader2> r(new Reader2());
    r->in_ = readers[0];
    r->dict_ = DecodeDictSize2(props[0]);
    r->eng_ = std::move(eng);
    error e = r->initialize();
    return {std::unique_ptr<io::ReadCloser>(new readCloser(readers[0], std::move(r))), e};
}

// ---------------------------------------------------------------- readcloser.go
error readCloser::Close() {
    if (!c_ || !r_) return errAlreadyClosed;   // readcloser.go:17-19
    if (error e = c_->Close()) return errors::Errorf("lzma: error closing", e);
    c_ = nullptr;
    r_.reset();
    return nullptr;
}

std::pair<int, error> readCloser::Read(uint8_t *p, size_t len) {
    if (!r_) return {0, errAlreadyClosed};
    auto [n, err] = r_->Read(p, len);
    if (err && !errors::Is(err, io::EOF_)) err = errors::Errorf("lzma: error reading", err);   // readcloser.go:36-38
    return {n, err};
}

}  // namespace lzma
