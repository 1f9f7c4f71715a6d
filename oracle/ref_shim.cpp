// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" wrapper around the UNMODIFIED reference headers (senikm/trpx include/Terse.hpp,
// Bit_pointer.hpp, XML_element.hpp), compiled from where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libtrpx_ref.so.  No reference source is copied into this repo:
// the headers are only #included through -I/root/reference/include at build time.
//
// Usage rules that follow SURVEY.md App. C: ONE jpa::Terse per frame (multi-frame objects are
// quadratic to build, C3, and mis-decode frames >= 2, C1); payload = write() output minus the XML
// header; decode goes through a temporary file because the only stream constructor takes
// std::ifstream& (Terse.hpp:279).
#include <cmath>      // Terse.hpp uses std::ceil (:503) and std::abs (:514) without including these
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#include "Terse.hpp"

namespace {

enum { U8 = 0, U16, U32, U64, I8, I16, I32, I64 };

template <typename T>
size_t encode_one(const void* px, size_t n, unsigned block, uint8_t* out, size_t cap, unsigned* pb)
{
    jpa::Terse t(static_cast<const T*>(px), n, block);
    std::ostringstream os;
    t.write(os);
    std::string s = os.str();
    size_t h = s.find("/>");
    if (h == std::string::npos) return 0;
    h += 2;
    size_t bytes = s.size() - h;
    if (bytes != t.terse_size() || bytes > cap) return 0;
    std::memcpy(out, s.data() + h, bytes);
    if (pb && t.bits_per_val() > *pb) *pb = t.bits_per_val();
    return bytes;
}

template <typename F>
auto dispatch(int dtype, F&& f)
{
    switch (dtype) {
    case U8:  return f((uint8_t*)nullptr);
    case U16: return f((uint16_t*)nullptr);
    case U32: return f((uint32_t*)nullptr);
    case U64: return f((uint64_t*)nullptr);
    case I8:  return f((int8_t*)nullptr);
    case I16: return f((int16_t*)nullptr);
    case I32: return f((int32_t*)nullptr);
    default:  return f((int64_t*)nullptr);
    }
}

struct RefHandle { jpa::Terse* t; };

} // namespace

extern "C" {

// Encode one frame with the reference; returns payload bytes (0 on failure).
size_t ref_encode_frame(const void* px, int dtype, size_t n, unsigned block, uint8_t* out,
                        size_t cap, unsigned* prolix_bits)
{
    return dispatch(dtype, [&](auto* tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        return encode_one<T>(px, n, block, out, cap, prolix_bits);
    });
}

// The reference's own XML header + payload for a single-frame object (for header parity).
size_t ref_write_file_image(const void* px, int dtype, size_t n, unsigned block, const size_t* dims,
                            size_t n_dims, uint8_t* out, size_t cap)
{
    return dispatch(dtype, [&](auto* tag) -> size_t {
        using T = std::remove_pointer_t<decltype(tag)>;
        jpa::Terse t(static_cast<const T*>(px), n, block);
        if (n_dims) t.dim(std::vector<std::size_t>(dims, dims + n_dims));
        std::ostringstream os;
        t.write(os);
        std::string s = os.str();
        if (s.size() > cap) return 0;
        std::memcpy(out, s.data(), s.size());
        return s.size();
    });
}

// Stack built with push_back (quadratic, App. C3 -- small stacks only): full file image.
size_t ref_write_stack_image(const void* px, int dtype, size_t n, size_t n_frames, uint8_t* out,
                             size_t cap)
{
    return dispatch(dtype, [&](auto* tag) -> size_t {
        using T = std::remove_pointer_t<decltype(tag)>;
        jpa::Terse t;
        for (size_t f = 0; f < n_frames; ++f) t.push_back(static_cast<const T*>(px) + f * n, n);
        std::ostringstream os;
        t.write(os);
        std::string s = os.str();
        if (s.size() > cap) return 0;
        std::memcpy(out, s.data(), s.size());
        return s.size();
    });
}

// Open a single-frame payload as a reference Terse object (through a temp file).
void* ref_open(const uint8_t* payload, size_t bytes, int is_signed, unsigned block,
               unsigned prolix_bits, size_t n)
{
    char path[] = "/tmp/trpx_ref_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) return nullptr;
    std::string hdr = "<Terse prolix_bits=\"" + std::to_string(prolix_bits) + "\" signed=\"" +
                      std::to_string(is_signed ? 1 : 0) + "\" block=\"" + std::to_string(block) +
                      "\" memory_size=\"" + std::to_string(bytes) + "\" number_of_values=\"" +
                      std::to_string(n) + "\" number_of_frames=\"1\"/>";
    bool ok = write(fd, hdr.data(), hdr.size()) == (ssize_t)hdr.size() &&
              write(fd, payload, bytes) == (ssize_t)bytes;
    close(fd);
    RefHandle* h = nullptr;
    if (ok) {
        std::ifstream in(path, std::ios::binary);
        h = new RefHandle{ new jpa::Terse(in) };
    }
    unlink(path);
    return h;
}

// Decode frame 0 with the reference into out_dtype (Terse.hpp:352-389).
int ref_prolix(void* handle, void* out, int out_dtype)
{
    RefHandle* h = static_cast<RefHandle*>(handle);
    if (!h) return -1;
    if (out_dtype == 8) { h->t->prolix(static_cast<float*>(out), 0); return 0; }    // floating-point iterators,
    if (out_dtype == 9) { h->t->prolix(static_cast<double*>(out), 0); return 0; }   // Terse.hpp:379-383
    dispatch(out_dtype, [&](auto* tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        h->t->prolix(static_cast<T*>(out), 0);
        return 0;
    });
    return 0;
}

void ref_close(void* handle)
{
    RefHandle* h = static_cast<RefHandle*>(handle);
    if (h) { delete h->t; delete h; }
}

// CPU baseline: encode every frame with one jpa::Terse per frame, frames statically
// partitioned over `threads` std::threads.  Returns seconds of wall time; payload bytes summed.
double ref_bench_encode(const void* px, int dtype, size_t n, size_t n_frames, unsigned threads,
                        size_t* total_payload_bytes)
{
    if (threads < 1) threads = 1;
    std::vector<size_t> bytes(threads, 0);
    size_t sz = (dtype & 3) == 0 ? 1 : (dtype & 3) == 1 ? 2 : (dtype & 3) == 2 ? 4 : 8;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < threads; ++w)
        pool.emplace_back([&, w] {
            size_t lo = n_frames * w / threads, hi = n_frames * (w + 1) / threads;
            for (size_t f = lo; f < hi; ++f)
                dispatch(dtype, [&](auto* tag) {
                    using T = std::remove_pointer_t<decltype(tag)>;
                    jpa::Terse t(reinterpret_cast<const T*>((const uint8_t*)px + f * n * sz), n);
                    bytes[w] += t.terse_size();
                    return 0;
                });
        });
    for (auto& th : pool) th.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    size_t tot = 0;
    for (size_t b : bytes) tot += b;
    if (total_payload_bytes) *total_payload_bytes = tot;
    return s;
}

// CPU baseline decode: frames are first encoded (untimed) into one Terse each, then prolix() of
// every frame is timed over `threads` threads.  out must hold n_frames*n values of dtype.
double ref_bench_decode(const void* px, int dtype, size_t n, size_t n_frames, unsigned threads,
                        void* out)
{
    if (threads < 1) threads = 1;
    size_t sz = (dtype & 3) == 0 ? 1 : (dtype & 3) == 1 ? 2 : (dtype & 3) == 2 ? 4 : 8;
    std::vector<jpa::Terse*> objs(n_frames, nullptr);
    for (size_t f = 0; f < n_frames; ++f)
        dispatch(dtype, [&](auto* tag) {
            using T = std::remove_pointer_t<decltype(tag)>;
            objs[f] = new jpa::Terse(reinterpret_cast<const T*>((const uint8_t*)px + f * n * sz), n);
            return 0;
        });
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < threads; ++w)
        pool.emplace_back([&, w] {
            size_t lo = n_frames * w / threads, hi = n_frames * (w + 1) / threads;
            for (size_t f = lo; f < hi; ++f)
                dispatch(dtype, [&](auto* tag) {
                    using T = std::remove_pointer_t<decltype(tag)>;
                    objs[f]->prolix(reinterpret_cast<T*>((uint8_t*)out + f * n * sz), 0);
                    return 0;
                });
        });
    for (auto& th : pool) th.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto* o : objs) delete o;
    return s;
}

// Byte identity at full scale: per-frame payload size and FNV-1a-64 of the reference's payload bytes (one jpa::Terse
// per frame, payload = write() minus the XML element), frames partitioned over `threads` threads.  Not timed.
int ref_frame_digests(const void* px, int dtype, size_t n, size_t n_frames, unsigned threads, uint64_t* sizes,
                      uint64_t* fnv)
{
    if (threads < 1) threads = 1;
    size_t sz = (dtype & 3) == 0 ? 1 : (dtype & 3) == 1 ? 2 : (dtype & 3) == 2 ? 4 : 8;
    std::vector<int> bad(threads, 0);
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < threads; ++w)
        pool.emplace_back([&, w] {
            size_t lo = n_frames * w / threads, hi = n_frames * (w + 1) / threads;
            for (size_t f = lo; f < hi; ++f)
                dispatch(dtype, [&](auto* tag) {
                    using T = std::remove_pointer_t<decltype(tag)>;
                    jpa::Terse t(reinterpret_cast<const T*>((const uint8_t*)px + f * n * sz), n);
                    std::ostringstream os;
                    t.write(os);
                    const std::string s = os.str();
                    size_t h = s.find("/>");
                    if (h == std::string::npos || s.size() - (h + 2) != t.terse_size()) { bad[w] = 1; return 0; }
                    h += 2;
                    uint64_t x = 0xcbf29ce484222325ull;
                    for (size_t i = h; i < s.size(); ++i) { x ^= (uint8_t)s[i]; x *= 0x100000001b3ull; }
                    sizes[f] = s.size() - h;
                    fnv[f] = x;
                    return 0;
                });
        });
    for (auto& th : pool) th.join();
    for (int b : bad) if (b) return 1;
    return 0;
}

} // extern "C"
