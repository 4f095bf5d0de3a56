// The reference's own tests (reader1_test.go:15-107, reader2_test.go:12-29) restated in C++ against
// include/lzma_reader.hpp, plus constructor-error, sevenzip-adapter and LZMA2-wave cases.
//   reader_test <assets dir> [--no-device]     prints one line per check, exits 1 on the first failure.
// --no-device: only what needs no GPU (constructors, header errors, helpers) -- the CPU tier's smoke run.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <string>

#include "lzma_reader.hpp"

using namespace lzma;

// ---- crypto/md5 stand-in (RFC 1321) so the checks read like the Go tests' ----------------------------
struct MD5 : io::Writer {
    uint32_t a = 0x67452301, b = 0xefcdab89, c = 0x98badcfe, d = 0x10325476;
    uint64_t len = 0;
    uint8_t buf[64];
    size_t fill = 0;
    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void block(const uint8_t *p) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
            0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
            0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
            0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
            0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
            0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t M[16];
        for (int i = 0; i < 16; i++) M[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; i++) {
            uint32_t F;
            int g;
            if (i < 16) { F = (B & C) | (~B & D); g = i; }
            else if (i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { F = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { F = C ^ (B | ~D); g = (7 * i) & 15; }
            F += A + K[i] + M[g];
            A = D; D = C; C = B;
            B += rol(F, S[i]);
        }
        a += A; b += B; c += C; d += D;
    }
    std::pair<int, error> Write(const uint8_t *p, size_t n) override {
        len += n;
        for (size_t i = 0; i < n; i++) {
            buf[fill++] = p[i];
            if (fill == 64) { block(buf); fill = 0; }
        }
        return {(int)n, nullptr};
    }
    std::string Sum() {
        const uint64_t bits = len * 8;
        uint8_t pad = 0x80;
        Write(&pad, 1);
        pad = 0;
        while (fill != 56) Write(&pad, 1);
        uint8_t l[8];
        for (int i = 0; i < 8; i++) l[i] = (uint8_t)(bits >> (8 * i));
        Write(l, 8);
        char hex[33];
        const uint32_t v[4] = {a, b, c, d};
        for (int i = 0; i < 16; i++) snprintf(hex + 2 * i, 3, "%02x", (v[i / 4] >> (8 * (i % 4))) & 0xFF);
        return hex;
    }
};
struct Discard : io::Writer {
    int64_t n = 0;
    std::pair<int, error> Write(const uint8_t *, size_t len) override { n += (int64_t)len; return {(int)len, nullptr}; }
};
struct Collect : io::Writer {
    std::vector<uint8_t> v;
    std::pair<int, error> Write(const uint8_t *p, size_t len) override { v.insert(v.end(), p, p + len); return {(int)len, nullptr}; }
};
// an io.ReadCloser over memory that is NOT an io.ByteReader (the sevenzip constructors then wrap it like bufio)
struct MemFile : io::ReadCloser {
    io::BytesReader r;
    int closed = 0;
    error close_err;
    explicit MemFile(const std::vector<uint8_t> &v) : r(v) {}
    std::pair<int, error> Read(uint8_t *p, size_t len) override { return r.Read(p, len); }
    error Close() override { closed++; return close_err; }
};

static std::string dir;
static int checks = 0;
static std::vector<uint8_t> ReadFile(const std::string &name) {
    std::ifstream f(dir + "/" + name, std::ios::binary);
    if (!f) { fprintf(stderr, "cannot open %s/%s\n", dir.c_str(), name.c_str()); exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
#define REQUIRE(cond, what)                                                                     \
    do {                                                                                        \
        checks++;                                                                               \
        if (!(cond)) { printf("FAIL %s:%d %s: %s\n", __FILE__, __LINE__, what, #cond); exit(1); } \
    } while (0)
static const char *text(const error &e) { return e ? e->msg.c_str() : "<nil>"; }

static const char *randomFileMD5 = "b2d18c4275c394a729607ff9fe0caae7";   // reader1_test.go:107
static const char *aTxtMD5 = "57a42eb7f425c13fa644f2618a097ab7";         // plaintext of the a*.lzma assets

static void TestHelpers() {
    auto [lc, pb, lp, err] = DecodeProp(0x5D);
    REQUIRE(!err && lc == 3 && pb == 2 && lp == 0, "DecodeProp(0x5d) = lc3 pb2 lp0 (return order lc, pb, lp: reader1.go:210)");
    REQUIRE(errors::Is(std::get<3>(DecodeProp(225)), ErrIncorrectProperties), "DecodeProp(225)");
    const uint8_t small[4] = {0, 1, 0, 0}, big[4] = {0, 0, 0x80, 0};
    REQUIRE(DecodeDictSize(small).first == 4096, "DecodeDictSize floors at 4 KiB (reader1.go:193-208)");
    REQUIRE(DecodeDictSize(big).first == (8u << 20), "DecodeDictSize 8 MiB");
    const uint8_t us[8] = {1, 2, 0, 0, 0, 0, 0, 0};
    REQUIRE(DecodeUnpackSize(us) == 0x201, "DecodeUnpackSize");
    REQUIRE(DecodeDictSize2(0) == 4096 && DecodeDictSize2(1) == 6144 && DecodeDictSize2(22) == (8u << 20), "DecodeDictSize2 (reader2.go:296)");
    REQUIRE(errors::Is(errors::Errorf("decode prop", ErrIncorrectProperties), ErrIncorrectProperties), "errors.Is sees through %w");
    REQUIRE(!errors::Is(ErrResultError, ErrCorrupted), "distinct error values");
}

static void TestConstructorErrors() {   // reader1.go:77-101,149-159; reader2.go:100-128
    std::vector<uint8_t> a = ReadFile("a.lzma");
    {
        io::BytesReader in(a.data(), 0);
        auto [r, err] = NewReader1(in);
        REQUIRE(err == io::EOF_, "empty input: bare io.EOF");
    }
    {
        io::BytesReader in(a.data(), 3);
        auto [r, err] = NewReader1(in);
        REQUIRE(errors::Is(err, io::EOF_) && err->msg == "decode dict size: EOF", text(err));
    }
    {
        io::BytesReader in(a.data(), 9);
        auto [r, err] = NewReader1(in);
        REQUIRE(errors::Is(err, io::EOF_) && err->msg == "decode unpack size: EOF", text(err));
    }
    {
        io::BytesReader in(a.data(), 15);
        auto [r, err] = NewReader1(in);
        REQUIRE(errors::Is(err, io::EOF_) && err->msg == "rangeDec.Init: EOF", text(err));
    }
    {
        std::vector<uint8_t> b = a;
        b[0] = 225;
        io::BytesReader in(b);
        auto [r, err] = NewReader1(in);
        REQUIRE(errors::Is(err, ErrIncorrectProperties) && err->msg == "decode prop: incorrect LZMA properties", text(err));
    }
    {
        std::vector<uint8_t> b = a;
        b[13] = 1;   // first range-coder byte must be 0 (range_decoder.go:32-34)
        io::BytesReader in(b);
        auto [r, err] = NewReader1(in);
        REQUIRE(errors::Is(err, ErrResultError) && err->msg == "rangeDec.Init: result error", text(err));
    }
    {
        io::BytesReader in(a.data(), 0);
        auto [r, err] = NewReader2(in, 0);
        REQUIRE(err == io::ErrUnexpectedEOF, "LZMA2: no control byte");
    }
    {
        const uint8_t h[3] = {0xE0, 0x00, 0x10};
        io::BytesReader in(h, 3);
        auto [r, err] = NewReader2(in, 0);
        REQUIRE(err == io::ErrUnexpectedEOF, "LZMA2: truncated chunk header");
    }
    {
        MemFile f(a);
        auto [rc, err] = NewLZMADecompressorForSevenZip({0x5D, 0, 0, 0x80, 0}, 10, {});
        REQUIRE(err == errNeedOneReader && !rc, "sevenzip LZMA: needs exactly one reader");
        auto [rc2, err2] = NewLZMA2DecompressorForSevenZip({0x16, 0x00}, 0, {&f});
        REQUIRE(err2 == errInsufficientProperties && !rc2, "sevenzip LZMA2: one property byte");
    }
}

static void TestReader1() {   // reader1_test.go:15-83
    struct { const char *name, *inputFile; bool ok1, ok2; } testCases[] = {
        {"correct_file_with_size", "a.lzma", true, true},
        {"correct_file_with_eos", "a_eos.lzma", true, true},
        {"correct_file_with_eos_and_size", "a_eos_and_size.lzma", true, true},
        {"correct_file_lp1_lc2_pb1", "a_lp1_lc2_pb1.lzma", true, true},
        {"bad_file", "bad_corrupted.lzma", true, false},
        {"bad_file_with_eos_and_incorrect_size", "bad_eos_incorrect_size.lzma", true, false},
        {"bad_file_with_incorrect_size", "bad_incorrect_size.lzma", true, false},
    };
    for (auto &tc : testCases) {
        std::vector<uint8_t> input = ReadFile(tc.inputFile);
        io::BytesReader in(input);
        auto [reader, err] = NewReader1(in);
        REQUIRE((err == nullptr) == tc.ok1, tc.name);
        MD5 sum;
        auto [n, err2] = io::Copy(sum, *reader);
        REQUIRE((err2 == nullptr) == tc.ok2, tc.name);
        if (tc.ok2) REQUIRE(sum.Sum() == aTxtMD5, tc.name);
        else REQUIRE(errors::Is(err2, ErrResultError), text(err2));   // the class the reference returns
        printf("ok   TestReader1/%s  (%lld bytes, err=%s)\n", tc.name, (long long)n, text(err2));
    }
}

static void TestReader1WithFileVerification() {   // reader1_test.go:85-105
    std::vector<uint8_t> compressedData = ReadFile("randomfile.dat.lzma");
    MD5 actualSummator;
    io::BytesReader in(compressedData);
    auto [r, err] = NewReader1(in);
    REQUIRE(!err, text(err));
    auto [n, err2] = io::Copy(actualSummator, *r);
    REQUIRE(!err2, text(err2));
    REQUIRE(actualSummator.Sum() == randomFileMD5, "decompressed data corrupted");
    printf("ok   TestReader1WithFileVerification (%lld bytes)\n", (long long)n);
    // small caller buffers, byte-wise delivery
    io::BytesReader in2(compressedData);
    auto [r2, e2] = NewReader1(in2);
    REQUIRE(!e2, text(e2));
    MD5 s2;
    uint8_t p[7];
    for (;;) {
        auto [k, e] = r2->Read(p, sizeof p);
        s2.Write(p, (size_t)k);
        if (e) { REQUIRE(e == io::EOF_, text(e)); break; }
    }
    REQUIRE(s2.Sum() == randomFileMD5 && r2->isEndOfStream, "7-byte reads");
}

static void TestReader2WithFileVerification() {   // reader2_test.go:12-29
    std::vector<uint8_t> compressedData = ReadFile("randomfile.dat.lzma2");
    MD5 actualSummator;
    io::BytesReader in(compressedData);
    auto [r, err] = NewReader2(in, 0);
    REQUIRE(!err, text(err));
    auto [n, err2] = io::Copy(actualSummator, *r);
    REQUIRE(!err2, text(err2));
    REQUIRE(actualSummator.Sum() == randomFileMD5, "decompressed data corrupted");
    printf("ok   TestReader2WithFileVerification (%lld bytes)\n", (long long)n);
    // the same in waves of one unit each (bounded look-ahead): every wave boundary is exercised
    io::BytesReader in2(compressedData);
    auto [r2, e2] = NewReader2(in2, 0);
    REQUIRE(!e2, text(e2));
    r2->wave_bytes = 1;
    MD5 s2;
    auto [n2, e3] = io::Copy(s2, *r2);
    REQUIRE(!e3 && n2 == n && s2.Sum() == randomFileMD5, "waves of one unit (next wave decoded while this one is served)");
    {
        io::BytesReader in6(compressedData);
        auto [r6, e9] = NewReader2(in6, 0);
        REQUIRE(!e9, text(e9));
        r6->wave_bytes = 1;
        r6->decode_ahead = false;
        MD5 s6;
        auto [n6, e10] = io::Copy(s6, *r6);
        REQUIRE(!e10 && n6 == n && s6.Sum() == randomFileMD5, "waves of one unit, no decode-ahead");
    }
    // a long stream: the asset 40 times over (each copy starts with a dictionary reset -> 40 units in one wave,
    // 40 MiB of output: the reader's buffer is page-locked and the batch call streams into it)
    {
        Collect plain;
        io::BytesReader in4(compressedData);
        auto [r4, e6] = NewReader2(in4, 0);
        REQUIRE(!e6, text(e6));
        REQUIRE(!io::Copy(plain, *r4).second && plain.v.size() == (size_t)n, "plaintext of the asset");
        std::vector<uint8_t> big;
        for (int k = 0; k < 40; k++) big.insert(big.end(), compressedData.begin(), compressedData.end() - 1);
        big.push_back(0);
        io::BytesReader in5(big);
        auto [r5, e7] = NewReader2(in5, 0);
        REQUIRE(!e7, text(e7));
        MD5 got, want;
        auto [n5, e8] = io::Copy(got, *r5);
        for (int k = 0; k < 40; k++) want.Write(plain.v.data(), plain.v.size());
        REQUIRE(!e8 && n5 == 40 * n && got.Sum() == want.Sum(), "40 MiB LZMA2 stream");
        printf("ok   TestReader2 long stream (%lld bytes)\n", (long long)n5);
    }
    // truncated stream: io.ErrUnexpectedEOF after the bytes of the complete chunks
    io::BytesReader in3(compressedData.data(), compressedData.size() - 1);   // terminator cut off
    auto [r3, e4] = NewReader2(in3, 0);
    REQUIRE(!e4, text(e4));
    Discard d3;
    auto [n3, e5] = io::Copy(d3, *r3);
    REQUIRE(errors::Is(e5, io::ErrUnexpectedEOF) && n3 == n, text(e5));
}

static void TestSevenZipAdapters() {   // reader1.go:32-61, reader2.go:49-75, readcloser.go
    std::vector<uint8_t> a = ReadFile("a.lzma");
    {
        // a 7z LZMA folder = the .lzma stream without its 13-byte header; props = its first 5 bytes
        std::vector<uint8_t> body(a.begin() + 13, a.end());
        MemFile f(body);
        std::vector<uint8_t> props(a.begin(), a.begin() + 5);
        auto [rc, err] = NewLZMADecompressorForSevenZip(props, DecodeUnpackSize(a.data() + 5), {&f});
        REQUIRE(!err && rc, text(err));
        MD5 sum;
        auto [n, e2] = io::Copy(sum, *rc);
        REQUIRE(!e2 && sum.Sum() == aTxtMD5, "sevenzip LZMA folder");
        REQUIRE(rc->Close() == nullptr && f.closed == 1, "Close closes the source once");
        REQUIRE(rc->Close() == errAlreadyClosed, "second Close");
        uint8_t p[4];
        REQUIRE(rc->Read(p, 4).second == errAlreadyClosed, "Read after Close");
    }
    {
        std::vector<uint8_t> bad = ReadFile("bad_corrupted.lzma");
        std::vector<uint8_t> body(bad.begin() + 13, bad.end());
        MemFile f(body);
        std::vector<uint8_t> props(bad.begin(), bad.begin() + 5);
        auto [rc, err] = NewLZMADecompressorForSevenZip(props, DecodeUnpackSize(bad.data() + 5), {&f});
        REQUIRE(!err, text(err));
        Discard d;
        auto [n, e2] = io::Copy(d, *rc);
        REQUIRE(errors::Is(e2, ErrResultError) && e2->msg == "lzma: error reading: result error", text(e2));   // readcloser.go:36-38
        f.close_err = errors::New("disk on fire");
        error ce = rc->Close();
        REQUIRE(ce && ce->msg == "lzma: error closing: disk on fire", text(ce));
    }
    {
        std::vector<uint8_t> z = ReadFile("randomfile.dat.lzma2");
        MemFile f(z);
        auto [rc, err] = NewLZMA2DecompressorForSevenZip({0x16}, 0, {&f});   // 0x16 = 8 MiB
        REQUIRE(!err && rc, text(err));
        MD5 sum;
        auto [n, e2] = io::Copy(sum, *rc);
        REQUIRE(!e2 && sum.Sum() == randomFileMD5, "sevenzip LZMA2 folder");
        REQUIRE(rc->Close() == nullptr, "Close");
    }
    printf("ok   TestSevenZipAdapters\n");
}

static void TestDecodeBatch() {   // the new batch entry point: many streams, one call
    auto [eng, err] = Engine::Default();
    REQUIRE(!err, text(err));
    const char *files[] = {"a.lzma", "a_eos.lzma", "a_lp1_lc2_pb1.lzma", "bad_corrupted.lzma", "randomfile.dat.lzma"};
    std::vector<uint8_t> in;
    std::vector<Unit> units;
    uint64_t out_off = 0;
    for (int rep = 0; rep < 8; rep++)
        for (const char *fn : files) {
            std::vector<uint8_t> d = ReadFile(fn);
            Unit u;
            memset(&u, 0, sizeof u);
            u.kind = LZGPU_KIND_LZMA1_ALONE;
            while (in.size() % 16) in.push_back(0);
            u.in_off = in.size();
            u.in_len = d.size();
            u.out_off = out_off;
            u.out_cap = 1 << 20;
            out_off += u.out_cap;
            in.insert(in.end(), d.begin(), d.end());
            units.push_back(u);
        }
    std::vector<uint8_t> out(out_off);
    auto [res, e2] = eng->DecodeBatch(units, in.data(), in.size(), out.data(), out.size());
    REQUIRE(!e2 && res.size() == units.size(), text(e2));
    for (size_t k = 0; k < units.size(); k++) {
        const char *fn = files[k % 5];
        MD5 sum;
        sum.Write(out.data() + units[k].out_off, res[k].bytes_out);
        if (!strcmp(fn, "bad_corrupted.lzma")) REQUIRE(errors::Is(Engine::StatusError(res[k].status), ErrResultError), fn);
        else REQUIRE(res[k].status == LZGPU_OK && sum.Sum() == (strncmp(fn, "random", 6) ? aTxtMD5 : randomFileMD5), fn);
    }
    printf("ok   TestDecodeBatch (%zu units on %d device(s))\n", units.size(), eng->Devices());
}

static void TestDecodeFolders() {   // every folder of an archive in one GPU call (reader1.go:28-61, reader2.go:45-75)
    auto [eng, err] = Engine::Default();
    REQUIRE(!err, text(err));
    const std::vector<uint8_t> a = ReadFile("a.lzma"), a2 = ReadFile("a_lp1_lc2_pb1.lzma"), rnd = ReadFile("randomfile.dat.lzma"),
                               bad = ReadFile("bad_corrupted.lzma"), z = ReadFile("randomfile.dat.lzma2");
    std::vector<Engine::Folder> folders;
    std::vector<const char *> want;   // MD5 of the plaintext, "" = ErrResultError, "props" = constructor error
    auto lzma1 = [&](const std::vector<uint8_t> &f, const char *md5) {
        Engine::Folder fo;
        fo.props.assign(f.begin(), f.begin() + 5);
        fo.unpackSize = DecodeUnpackSize(f.data() + 5);
        fo.packed = f.data() + 13;
        fo.packedLen = f.size() - 13;
        folders.push_back(fo);
        want.push_back(md5);
    };
    for (int rep = 0; rep < 14; rep++) {
        lzma1(a, aTxtMD5);
        lzma1(a2, aTxtMD5);
        lzma1(rnd, randomFileMD5);
        lzma1(bad, "");
        Engine::Folder f2;
        f2.lzma2 = true;
        f2.props = {0x16};
        f2.packed = z.data();
        f2.packedLen = z.size();
        folders.push_back(f2);
        want.push_back(randomFileMD5);
    }
    Engine::Folder p1;                       // errInsufficientProperties / ErrIncorrectProperties, like the constructors
    p1.lzma2 = true;
    folders.push_back(p1);
    want.push_back("props2");
    Engine::Folder p2;
    p2.props = {225, 0, 0, 1, 0};
    folders.push_back(p2);
    want.push_back("props1");
    auto [res, e2] = eng->DecodeFolders(folders);
    REQUIRE(!e2 && res.size() == folders.size() && folders.size() >= 64, text(e2));
    for (size_t i = 0; i < res.size(); i++) {
        if (!strcmp(want[i], "props2")) { REQUIRE(res[i].err == errInsufficientProperties, text(res[i].err)); continue; }
        if (!strcmp(want[i], "props1")) { REQUIRE(res[i].err == ErrIncorrectProperties, text(res[i].err)); continue; }
        if (!want[i][0]) { REQUIRE(errors::Is(res[i].err, ErrResultError), text(res[i].err)); continue; }
        MD5 sum;
        sum.Write(res[i].out.data(), res[i].out.size());
        REQUIRE(!res[i].err && sum.Sum() == want[i], "folder bytes");
    }
    printf("ok   TestDecodeFolders (%zu folders, one call)\n", folders.size());
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: reader_test <assets dir> [--no-device]\n"); return 2; }
    dir = argv[1];
    const bool no_device = argc > 2 && !strcmp(argv[2], "--no-device");
    TestHelpers();
    TestConstructorErrors();
    if (no_device) {
        // the product path must fail loudly without a device: no CPU fallback
        std::vector<uint8_t> a = ReadFile("a.lzma");
        io::BytesReader in(a);
        auto [r, err] = NewReader1(in);
        REQUIRE(!err, text(err));
        if (lzgpu_device_count() <= 0) {
            Discard d;
            auto [n, e2] = io::Copy(d, *r);
            REQUIRE(e2 && n == 0 && e2->msg.find("lzgpu:") == 0, text(e2));
            printf("ok   no device: Read fails with \"%s\"\n", text(e2));
        }
        printf("PASS %d checks (no-device tier)\n", checks);
        return 0;
    }
    TestReader1();
    TestReader1WithFileVerification();
    TestReader2WithFileVerification();
    TestSevenZipAdapters();
    TestDecodeBatch();
    TestDecodeFolders();
    printf("PASS %d checks\n", checks);
    return 0;
}
