// reader2_bench -- throughput of the streaming facade: one raw LZMA2 stream read through
// NewReader2 + Read (the reference's reader2.go:26,216 API) with the next wave decoded while the current one is
// served, against the same stream read with decode-ahead off.  Input: a file holding the stream (scripts/
// bench_reader2.py writes it).  Output: one JSON line.
//   reader2_bench stream.lzma2 [read_buffer_bytes] [wave_bytes]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "lzma_reader.hpp"

using namespace lzma;

static double run(const std::vector<uint8_t> &stream, size_t bufsize, size_t wave, bool ahead, uint64_t *n_out, uint32_t *sum) {
    io::BytesReader in(stream);
    auto [r, err] = NewReader2(in, 8 << 20);
    if (err) { fprintf(stderr, "NewReader2: %s\n", err->msg.c_str()); exit(1); }
    r->wave_bytes = wave;
    r->decode_ahead = ahead;
    std::vector<uint8_t> buf(bufsize);
    uint64_t total = 0;
    uint32_t s = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        auto [n, e] = r->Read(buf.data(), buf.size());
        if (n > 0) { total += (uint64_t)n; s += buf[0] + buf[(size_t)n - 1]; }   // the caller touches what it got
        if (e) {
            if (!errors::Is(e, io::EOF_)) { fprintf(stderr, "Read: %s\n", e->msg.c_str()); exit(1); }
            break;
        }
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *n_out = total;
    *sum = s;
    return sec;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: reader2_bench stream.lzma2 [read_buffer_bytes] [wave_bytes]\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> stream((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const size_t bufsize = argc > 2 ? strtoull(argv[2], nullptr, 0) : (1u << 20);
    const size_t wave = argc > 3 ? strtoull(argv[3], nullptr, 0) : ((size_t)1 << 30);
    uint64_t n = 0;
    uint32_t s = 0;
    const double t_cold = run(stream, bufsize, wave, true, &n, &s);   // first use in the process: context, device slabs, page-locked buffers
    const double t_ahead = run(stream, bufsize, wave, true, &n, &s);
    const double t_plain = run(stream, bufsize, wave, false, &n, &s);
    printf("{\"what\": \"NewReader2 + Read over one raw LZMA2 stream\", \"compressed_bytes\": %zu, \"decoded_bytes\": %llu, "
           "\"read_buffer\": %zu, \"wave_bytes\": %zu, \"first_run_in_process_s\": %.4f, \"decode_ahead_s\": %.4f, \"decode_ahead_GBps\": %.3f, "
           "\"no_decode_ahead_s\": %.4f, \"no_decode_ahead_GBps\": %.3f}\n",
           stream.size(), (unsigned long long)n, bufsize, wave, t_cold, t_ahead, n / t_ahead / 1e9, t_plain, n / t_plain / 1e9);
    return 0;
}
