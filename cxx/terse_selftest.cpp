// terse_selftest.cpp -- exercises include/trpx/Terse.hpp (the host-side drop-in for jpa::Terse) the way the
// reference's own test intends (test/terse_tests.cpp:15-33: iota(-500..499) -> Terse -> file -> read back ->
// prolix), plus stacks, dims, conversions and error behaviour.
//   terse_selftest --container <in.trpx> <out.trpx>   host only: read a .trpx, write it back (byte identity is
//                                                     checked by the caller); prints the parsed attributes
//   terse_selftest --objects <in> <out>               host only: every <Terse/> object of a stream, read one after another
//   terse_selftest --gpu                              needs a CUDA device: round trips through the kernels
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <vector>

#include <trpx/Grey_tiff_io.hpp>
#include <trpx/Terse.hpp>

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

struct Image {                       // a container with dim(), like the reference's Grey_tif image views
    std::vector<std::uint16_t> px;
    std::vector<std::size_t> d;
    auto begin() const { return px.begin(); }
    auto end() const { return px.end(); }
    auto begin() { return px.begin(); }
    auto end() { return px.end(); }
    std::size_t size() const { return px.size(); }
    std::vector<std::size_t> const& dim() const { return d; }
};

static int container_mode(const char* in, const char* out)
{
    std::ifstream is(in, std::ios::binary);
    jpa::Terse t(is);
    std::printf("prolix_bits=%u signed=%d block=%u memory_size=%zu number_of_values=%zu frames=%zu dims=",
                t.bits_per_val(), int(t.is_signed()), t.block(), t.terse_size(), t.size(), t.number_of_frames());
    for (auto d : t.dim()) std::printf("%zu ", d);
    std::printf("\n");
    std::ofstream os(out, std::ios::binary);
    t.write(os);
    return 0;
}

// host only: a stream may hold several <Terse .../> objects back to back (the reader leaves the stream right after each
// payload, reference Terse.hpp:275-279, XML_element.hpp:216-224): read them all with consecutive constructor calls and
// write them back in order.  Malformed headers must throw, never crash.
static int objects_mode(const char* in, const char* out)
{
    std::ifstream is(in, std::ios::binary);
    std::ofstream os(out, std::ios::binary);
    int n = 0;
    for (;;) {
        try {
            jpa::Terse t(is);
            std::printf("object %d: prolix_bits=%u signed=%d block=%u memory_size=%zu number_of_values=%zu frames=%zu\n", n,
                        t.bits_per_val(), int(t.is_signed()), t.block(), t.terse_size(), t.size(), t.number_of_frames());
            t.write(os);
            ++n;
        } catch (std::exception const& e) {
            std::printf("stopped after %d object(s): %s\n", n, e.what());
            break;
        }
    }
    return n ? 0 : 3;
}

static int tiff_mode(const char* in, const char* out)       // host only: TIFF stack -> TIFF stack through Grey_tiff_io
{
    std::ifstream is(in, std::ios::binary);
    auto imgs = jpa::tiffio::read(is);
    for (auto const& im : imgs)
        std::printf("image %zux%zu bits=%u kind=%d\n", im.width, im.height, im.bits, int(im.kind));
    std::ofstream os(out, std::ios::binary);
    jpa::tiffio::write(os, imgs);
    return 0;
}

static int gpu_mode()
{
    // (1) the reference's documented example: Terse.hpp:127-154
    std::vector<int> numbers(1000);
    std::iota(numbers.begin(), numbers.end(), -500);
    jpa::Terse compressed(numbers);
    CHECK(compressed.size() == 1000 && compressed.is_signed() && compressed.number_of_frames() == 1);
    CHECK(compressed.bits_per_val() == 10 && compressed.terse_size() == 1152);
    std::stringstream file;
    compressed.write(file);
    jpa::Terse from_file(file);
    std::vector<int> back(1000);
    from_file.prolix(back.begin());
    CHECK(back == numbers);
    std::vector<double> as_double(1000);
    from_file.prolix(as_double);                                                       // float outputs: Terse.hpp:379-383
    for (int i = 0; i < 1000; ++i) CHECK(as_double[i] == double(numbers[i]));
    std::vector<std::int8_t> clamped(1000);
    from_file.prolix(clamped.begin());                                                 // narrower type: clamp, Bit_pointer.hpp:747-763
    CHECK(clamped[0] == -128 && clamped[999] == 127 && clamped[500] == 0);

    // (2) a stack built frame by frame and in one batch gives the same bytes; every frame decodes (App. C1/C2)
    const std::size_t N = 512 * 512, F = 5;
    std::vector<std::uint16_t> stack(N * F);
    std::uint64_t x = 88172645463325252ull;
    for (auto& v : stack) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = std::uint16_t((x >> 40) % 7 + ((x & 1023) == 0 ? 3000 : 0)); }
    jpa::Terse a, b;
    for (std::size_t f = 0; f < F; ++f) a.push_back(stack.data() + f * N, N);
    b.push_back_frames(stack.data(), N, F);
    CHECK(a.number_of_frames() == F && b.number_of_frames() == F && a.terse_size() == b.terse_size());
    CHECK(std::memcmp(a.terse_data(), b.terse_data(), a.terse_size()) == 0);
    std::stringstream sf;
    a.write(sf);
    jpa::Terse c(sf);                                                                  // frame sizes unknown now
    for (std::size_t f : {std::size_t(4), std::size_t(0), std::size_t(2)}) {
        std::vector<std::uint16_t> fr(N);
        c.prolix(fr, f);
        CHECK(std::memcmp(fr.data(), stack.data() + f * N, N * 2) == 0);
    }
    std::vector<std::uint32_t> all(N * F);
    c.prolix_frames(all.begin(), 0, F);
    for (std::size_t i = 0; i < N * F; i += 997) CHECK(all[i] == stack[i]);

    // (3) containers with dim(): picked up on construction, written as "dimensions"
    Image img{std::vector<std::uint16_t>(640 * 480, 3), {640, 480}};
    jpa::Terse d(img);
    std::stringstream df;
    d.write(df);
    CHECK(df.str().find("dimensions=\"640 480\"") != std::string::npos);
    Image img2{std::vector<std::uint16_t>(640 * 480), {640, 480}};
    d.prolix(img2);
    CHECK(img2.px == img.px);

    // (4) signed data cannot be unpacked into an unsigned type (Terse.hpp:356-357)
#ifdef NDEBUG
    bool threw = false;
    try { std::vector<unsigned> u(1000); compressed.prolix(u.begin()); } catch (std::invalid_argument const&) { threw = true; }
    CHECK(threw);
#endif
    std::printf(fails ? "selftest FAILED (%d)\n" : "selftest ok\n", fails);
    return fails ? 1 : 0;
}

int main(int argc, char** argv)
{
    try {
        if (argc == 4 && !std::strcmp(argv[1], "--container")) return container_mode(argv[2], argv[3]);
        if (argc == 4 && !std::strcmp(argv[1], "--tiff")) return tiff_mode(argv[2], argv[3]);
        if (argc == 4 && !std::strcmp(argv[1], "--objects")) return objects_mode(argv[2], argv[3]);
        if (argc == 2 && !std::strcmp(argv[1], "--gpu")) return gpu_mode();
    } catch (std::exception const& e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
    std::printf("usage: terse_selftest --container in out | --gpu\n");
    return 64;
}
