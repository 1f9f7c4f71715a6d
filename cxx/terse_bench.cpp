// terse_bench.cpp -- end-to-end throughput THROUGH THE DROP-IN CLASS (include/trpx/Terse.hpp), the call sequence a
// user of the reference makes (Terse.hpp:290-302 push_back, :352-389 prolix) with its own pageable std::vectors:
//   jpa::Terse t; t.push_back_frames(pixels.data(), n, F); t.prolix_frames(back.data(), 0, F);
// Input: a raw file of F frames of n uint16 values (bench.py writes its synthetic stack there).  Prints one JSON line.
//   terse_bench <raw u16 file> <values per frame> <frames> <passes>
// Environment: TRPX_PIN_MIN_MB (0: never page-lock the caller's ranges), TRPX_MULTI_GPU=1 (shard over every device).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "trpx/Terse.hpp"

int main(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: terse_bench <raw u16 file> <values per frame> <frames> <passes>\n"); return 2; }
    const std::size_t n = std::strtoull(argv[2], nullptr, 10), F = std::strtoull(argv[3], nullptr, 10);
    const int passes = std::atoi(argv[4]);
    std::vector<std::uint16_t> px(n * F), back(n * F);
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(px.data(), 2, n * F, f) != n * F) { std::fprintf(stderr, "terse_bench: cannot read %s\n", argv[1]); return 2; }
    std::fclose(f);
    std::vector<double> enc, dec;
    std::size_t bytes = 0;
    try {
        for (int k = 0; k <= passes; ++k) {                      // pass 0 warms up (context creation, allocations)
            std::fill(back.begin(), back.end(), std::uint16_t(0xCDCD));
            const auto t0 = std::chrono::steady_clock::now();
            jpa::Terse t;
            t.push_back_frames(px.data(), n, F);
            const auto t1 = std::chrono::steady_clock::now();
            t.prolix_frames(back.data(), 0, F);
            const auto t2 = std::chrono::steady_clock::now();
            if (std::memcmp(px.data(), back.data(), n * F * 2) != 0) { std::fprintf(stderr, "terse_bench: round trip FAILED\n"); return 1; }
            bytes = t.terse_size();
            if (k) {
                enc.push_back(std::chrono::duration<double>(t1 - t0).count());
                dec.push_back(std::chrono::duration<double>(t2 - t1).count());
            }
        }
    } catch (std::exception const& e) {
        std::fprintf(stderr, "terse_bench: %s\n", e.what());
        return 2;
    }
    auto median = [](std::vector<double> v) { std::sort(v.begin(), v.end()); return v.size() % 2 ? v[v.size() / 2] : 0.5 * (v[v.size() / 2 - 1] + v[v.size() / 2]); };
    const double e = median(enc), d = median(dec);
    const char* pin = std::getenv("TRPX_PIN_MIN_MB");
    const char* multi = std::getenv("TRPX_MULTI_GPU");
    std::printf("{\"api\": \"jpa::Terse::push_back_frames + prolix_frames (pageable std::vector)\", \"frames\": %zu, \"payload_bytes\": %zu, "
                "\"encode_s\": %.6f, \"decode_s\": %.6f, \"frames_per_s\": %.1f, \"pin_min_mb\": \"%s\", \"multi_gpu\": \"%s\", \"passes\": %d}\n",
                F, bytes, e, d, double(F) / (e + d), pin ? pin : "default(0: off)", multi ? multi : "0", passes);
    return 0;
}
