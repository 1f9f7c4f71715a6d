// trpx/Terse.hpp -- host-side drop-in for the reference's `jpa::Terse` (senikm/trpx include/Terse.hpp:228-474)
// on top of the B200 C ABI (include/trpx_b200.h, libtrpx_b200.so).
//
// Same public surface and meaning as the reference class:
//   Terse()                                      Terse.hpp:237
//   Terse(Container const&)                      :249-253   (picks up dim() when the container has one)
//   Terse(Iterator, size, block = 12)            :263-270
//   Terse(std::istream&)                         :279       (reference: std::ifstream& only, App. C8)
//   push_back(Iterator, size) / (Container)      :290-302, :312-322
//   prolix(Container&, frame) / (Iterator, frame):333-341, :352-389
//   size, number_of_frames, dim, dim(v), is_signed, bits_per_val, terse_size, write   :396-474
// plus two batch entry points that are the natural way to drive a GPU (one launch for a whole stack):
//   push_back_frames(ptr, size, n_frames), prolix_frames(ptr, first_frame, n_frames).
//
// What differs on purpose: every frame of a multi-frame object decodes correctly
// (the reference mis-addresses frames >= 2, C1/C2); building a stack is linear, not quadratic (C3);
// blocks as wide as the type decode correctly (C5); `number_of_frames` may be absent in a header (C8); an UNSIGNED stream
// decoded into a SIGNED type keeps its values -- the reference's get_range sign-extends from bit s-1 whenever the
// OUTPUT type is signed (Bit_pointer.hpp:784-789), so unsigned 5 in a 3-bit block would come back as -3 (C11);
// malformed headers throw instead of reading garbage.
// The arithmetic of the codec is NOT here: construction / push_back call trpx_encode_host, prolix
// calls trpx_decode_host, and both fail (std::runtime_error) when no CUDA device is usable.
#pragma once

#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <iterator>
#include <limits>
#include <memory>
#include <new>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../trpx_b200.h"

namespace jpa {

namespace trpx_detail {

// one context per process and device, created on first use
inline trpx_ctx* context(int device = 0)
{
    struct Holder {
        trpx_ctx* ctx = nullptr;
        int status = TRPX_OK;
        explicit Holder(int dev) { status = trpx_ctx_create(dev, &ctx); }
        ~Holder() { trpx_ctx_destroy(ctx); }
    };
    static Holder h(device);
    if (h.status != TRPX_OK)
        throw std::runtime_error(std::string("trpx_b200: ") + trpx_strerror(h.status));
    return h.ctx;
}

// Several GPUs (TRPX_MULTI_GPU=1 in the environment, or jpa::Terse::use_all_devices()): multi-frame calls are sharded
// by frame over every visible device (trpx_pool_*: one context and one host thread per device, host concatenation).
inline bool& multi_gpu()
{
    static bool on = [] { const char* e = std::getenv("TRPX_MULTI_GPU"); return e && *e && *e != '0'; }();
    return on;
}
inline trpx_pool* pool()
{
    struct Holder {
        trpx_pool* p = nullptr;
        int status = TRPX_OK;
        Holder() { status = trpx_pool_create(nullptr, 0, &p); }
        ~Holder() { trpx_pool_destroy(p); }
    };
    static Holder h;
    if (h.status != TRPX_OK) throw std::runtime_error(std::string("trpx_b200: ") + trpx_strerror(h.status));
    return h.p;
}

// Page-locks a caller's range for the duration of one call.  OFF by default: measured on the B200 hosts, registering a
// 1 GB range costs more than the bounce-buffer copy it saves (bench.py, e2e.dropin: 1.4 k frames/s with, 3.5 k frames/s
// without).  TRPX_PIN_MIN_MB=<n> turns it on for ranges of at least n MB -- for callers that reuse one buffer for many
// calls it is cheaper to pin it once themselves (trpx_host_pin).  A range that is pinned already is left alone.
class ScopedPin {
public:
    ScopedPin(const void* p, std::size_t bytes)
    {
        static const std::size_t min_bytes = [] {
            const char* e = std::getenv("TRPX_PIN_MIN_MB");
            return (e && *e ? std::size_t(std::strtoull(e, nullptr, 10)) : std::size_t(0)) << 20;
        }();
        if (p && min_bytes && bytes >= min_bytes && trpx_host_pin(const_cast<void*>(p), bytes) == TRPX_OK) d_p = const_cast<void*>(p);
    }
    ~ScopedPin() { if (d_p) trpx_host_unpin(d_p); }
    ScopedPin(ScopedPin const&) = delete;
    ScopedPin& operator=(ScopedPin const&) = delete;
private:
    void* d_p = nullptr;
};

// One grow-only pinned buffer per thread for the payload of an encode call: the worst-case capacity (raw size + 6 %) is
// never value-initialised (the reference zero-fills it, Terse.hpp:503) and the payload D2H runs at the link rate; only
// the actual payload is then copied into the object.
inline std::uint8_t* pinned_scratch(std::size_t bytes)
{
    struct Holder {
        void* p = nullptr;
        std::size_t cap = 0;
        ~Holder() { trpx_host_free(p); }
    };
    static thread_local Holder h;
    if (h.cap < bytes) {
        trpx_host_free(h.p);
        h.p = trpx_host_alloc(bytes + bytes / 8);
        h.cap = h.p ? bytes + bytes / 8 : 0;
    }
    return static_cast<std::uint8_t*>(h.p);
}

// The payload lives in a byte vector that is NOT value-initialised when it grows: first-touching fresh pages by zero-
// filling them and then overwriting them costs more than the whole GPU encode (measured: 226 MB appended in 98 ms by
// one thread, page faults included, against 43 ms for encoding the 1 GB of pixels behind them).
template <typename T>
struct default_init_allocator : std::allocator<T> {
    template <typename U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    template <typename U> void construct(U* p) noexcept(std::is_nothrow_default_constructible_v<U>) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using byte_vector = std::vector<std::uint8_t, default_init_allocator<std::uint8_t>>;

// dst += [src, src + n): large appends are copied (and their fresh pages faulted in) by a few threads side by side
inline void append_bytes(byte_vector& dst, const std::uint8_t* src, std::size_t n)
{
    const std::size_t old = dst.size();
    dst.resize(old + n);
    std::uint8_t* d = dst.data() + old;
    constexpr std::size_t PER_THREAD = std::size_t(16) << 20;
    std::size_t threads = n / PER_THREAD;
    if (threads > 6) threads = 6;
    if (threads < 2) { std::memcpy(d, src, n); return; }
    std::vector<std::thread> th;
    const std::size_t step = ((n + threads - 1) / threads + 4095) & ~std::size_t(4095);
    for (std::size_t t = 0; t < threads; ++t) {
        const std::size_t lo = t * step, hi = lo + step < n ? lo + step : n;
        if (lo < hi) th.emplace_back([=] { std::memcpy(d + lo, src + lo, hi - lo); });
    }
    for (auto& x : th) x.join();
}

inline void check(int status, trpx_ctx* ctx)
{
    if (status == TRPX_OK) return;
    std::string msg = std::string("trpx_b200: ") + trpx_strerror(status);
    const char* detail = ctx ? trpx_last_error(ctx) : "";
    if (detail && *detail) msg += std::string(" (") + detail + ")";
    throw std::runtime_error(msg);
}

template <typename T>
constexpr int dtype_of()
{
    static_assert(std::is_integral_v<T>, "TERSE encodes integral pixels only (Terse.hpp:245)");
    constexpr bool s = std::is_signed_v<T>;
    return sizeof(T) == 1 ? (s ? TRPX_I8 : TRPX_U8)
         : sizeof(T) == 2 ? (s ? TRPX_I16 : TRPX_U16)
         : sizeof(T) == 4 ? (s ? TRPX_I32 : TRPX_U32)
                          : (s ? TRPX_I64 : TRPX_U64);
}

template <typename It>
using value_t = typename std::iterator_traits<It>::value_type;

template <typename It>
constexpr bool is_raw_pointer = std::is_pointer_v<It>;

// value of attribute `name` inside the text of one XML start tag, "" when absent
inline std::string attribute(std::string const& tag, std::string const& name)
{
    std::size_t pos = 0;
    while ((pos = tag.find(name, pos)) != std::string::npos) {
        const bool starts = pos == 0 || tag[pos - 1] == ' ' || tag[pos - 1] == '\t' || tag[pos - 1] == '\n';
        std::size_t q = pos + name.size();
        while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) ++q;
        if (starts && q < tag.size() && tag[q] == '=') {
            ++q;
            while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) ++q;
            if (q < tag.size() && (tag[q] == '"' || tag[q] == '\'')) {
                const char quote = tag[q];
                const std::size_t e = tag.find(quote, q + 1);
                if (e != std::string::npos) return tag.substr(q + 1, e - q - 1);
            }
        }
        pos += name.size();
    }
    return std::string();
}

// An unsigned decimal attribute; anything else (absent, empty, a sign, letters, out of range) is a malformed header and
// throws -- std::stoul alone would read "-3" as a huge value.
inline unsigned long long parse_unsigned(std::string const& tag, const char* name)
{
    const std::string v = attribute(tag, name);
    if (v.empty() || v.size() > 19 || v.find_first_not_of("0123456789") != std::string::npos)
        throw std::runtime_error(std::string("trpx: malformed <Terse> header: attribute ") + name);
    return std::stoull(v);
}

// Scan forward to "<Terse", return the text up to (not including) the closing '>', leaving the stream on
// the first payload byte (what XML_element(istream, "Terse") does for the reference, XML_element.hpp:216-224).
inline bool find_terse_tag(std::istream& in, std::string& tag)
{
    static const char key[] = "<Terse";
    std::size_t matched = 0;
    int c;
    while ((c = in.get()) != std::char_traits<char>::eof()) {
        if (static_cast<char>(c) == key[matched]) {
            if (++matched == sizeof(key) - 1) break;
        } else {
            matched = static_cast<char>(c) == key[0] ? 1 : 0;
        }
    }
    if (matched != sizeof(key) - 1) return false;
    tag.clear();
    while ((c = in.get()) != std::char_traits<char>::eof() && static_cast<char>(c) != '>') tag.push_back(static_cast<char>(c));
    return c != std::char_traits<char>::eof();
}

} // namespace trpx_detail

class Terse {
public:
    Terse() = default;

    template <typename Container>
        requires (requires (Container c) { c.begin(); c.size(); })
    Terse(Container const& data) : Terse(data.begin(), data.size())
    {
        if constexpr (requires (Container& c) { c.dim(); })
            for (auto d : data.dim()) d_dim.push_back(d);
    }

    template <typename Iterator>
    Terse(Iterator const data, std::size_t const size, unsigned int const block = 12)
        : d_signed(std::is_signed_v<trpx_detail::value_t<Iterator>>), d_block(block), d_size(size)
    {
        f_encode(data, 1);
    }

    // Reads the next <Terse .../> object of a stream; the stream is left on the byte after its payload.
    explicit Terse(std::istream& istream)
    {
        std::string tag;
        if (!trpx_detail::find_terse_tag(istream, tag)) throw std::runtime_error("trpx: no <Terse .../> element in stream");
        auto attr = [&](const char* n) { return trpx_detail::attribute(tag, n); };
        auto num = [&](const char* n) { return trpx_detail::parse_unsigned(tag, n); };
        if (tag.empty() || tag.back() != '/') throw std::runtime_error("trpx: malformed <Terse> header: element not closed");
        d_prolix_bits = unsigned(num("prolix_bits"));
        d_signed = num("signed") != 0;
        d_block = unsigned(num("block"));
        d_size = num("number_of_values");
        if (d_prolix_bits > 73 || d_block == 0 || d_size == 0) throw std::runtime_error("trpx: malformed <Terse> header");
        std::istringstream dims(attr("dimensions"));
        for (std::size_t v; dims >> v;) d_dim.push_back(v);
        d_terse_data.resize(std::size_t(num("memory_size")));
        istream.read(reinterpret_cast<char*>(d_terse_data.data()), std::streamsize(d_terse_data.size()));
        if (std::size_t(istream.gcount()) != d_terse_data.size()) throw std::runtime_error("trpx: truncated TERSE payload");
        const std::string nf = attr("number_of_frames");
        const std::size_t frames = nf.empty() ? 1 : std::size_t(num("number_of_frames"));
        if (frames == 0 || frames > d_terse_data.size()) throw std::runtime_error("trpx: malformed <Terse> header: number_of_frames");
        d_frame_bytes.assign(frames, 0);                                // 0: size not known yet
        if (d_frame_bytes.size() == 1) d_frame_bytes[0] = d_terse_data.size();
        if (d_block == 0 || d_size == 0) throw std::runtime_error("trpx: malformed <Terse> header");
    }

    template <typename Iterator>
    void push_back(Iterator const data, std::size_t const size) { push_back_frames(data, size, 1); }

    template <typename Container>
        requires requires (Container& c) { c.begin(), c.end(), c.size(); }
    void push_back(Container const& data)
    {
        if constexpr (requires (Container& c) { c.dim(); }) {
            for (std::size_t i = 0; i != data.dim().size(); ++i)
                if (number_of_frames() == 0) d_dim.push_back(data.dim()[i]);
                else assert(d_dim[i] == data.dim()[i]);
        }
        push_back(data.begin(), data.size());
    }

    // Batch form: `n_frames` frames of `size` values each, contiguous from `data`; ONE pass on the GPU.
    template <typename Iterator>
    void push_back_frames(Iterator const data, std::size_t const size, std::size_t const n_frames)
    {
        constexpr bool sgn = std::is_signed_v<trpx_detail::value_t<Iterator>>;
        if (number_of_frames() == 0) {
            d_size = size;
            d_signed = sgn;
        } else {
            assert(this->size() == size);
            assert(d_signed == sgn);
            if (this->size() != size || d_signed != sgn) throw std::invalid_argument("trpx: frame size / signedness mismatch");
        }
        f_encode(data, n_frames);
    }

    template <typename Container>
        requires requires (Container& c) { c.begin(), c.end(), c.size(); }
    void prolix(Container& data, std::size_t frame = 0)
    {
        assert(this->size() == data.size());
        if (this->size() != data.size()) throw std::invalid_argument("trpx: prolix: container size differs from the frame size");
        if constexpr (requires (Container& c) { c.dim(); })
            for (std::size_t i = 0; i != d_dim.size(); ++i) assert(d_dim[i] == data.dim()[i]);
        prolix(data.begin(), frame);
    }

    template <typename Iterator>
        requires requires (Iterator& i) { *i; }
    void prolix(Iterator begin, std::size_t frame = 0) { prolix_frames(begin, frame, 1); }

    // Batch form: frames [first_frame, first_frame + n_frames) into n_frames * size() values from `begin`.
    template <typename Iterator>
    void prolix_frames(Iterator begin, std::size_t first_frame, std::size_t n_frames)
    {
        using V = trpx_detail::value_t<Iterator>;
        assert(first_frame + n_frames <= number_of_frames());
        if (first_frame + n_frames > number_of_frames()) throw std::out_of_range("trpx: prolix: frame index");
        if (d_signed) {
            assert(std::is_signed_v<V>);
            if (!std::is_signed_v<V>) throw std::invalid_argument("trpx: signed data cannot be unpacked into an unsigned type");
        }
        const std::size_t n = n_frames * d_size;
        if constexpr (std::is_integral_v<V> && !std::is_same_v<V, bool>) {
            if constexpr (trpx_detail::is_raw_pointer<Iterator>) {
                f_decode(begin, trpx_detail::dtype_of<V>(), first_frame, n_frames);
            } else {
                std::vector<V> tmp(n);
                f_decode(tmp.data(), trpx_detail::dtype_of<V>(), first_frame, n_frames);
                std::copy(tmp.begin(), tmp.end(), begin);
            }
        } else if constexpr (std::is_same_v<V, float> || std::is_same_v<V, double>) {
            // floating point (Terse.hpp:379-383): the device converts through a 64-bit integer and a double
            constexpr int code = std::is_same_v<V, float> ? TRPX_F32 : TRPX_F64;
            if constexpr (trpx_detail::is_raw_pointer<Iterator>) {
                f_decode(begin, code, first_frame, n_frames);
            } else {
                std::vector<V> tmp(n);
                f_decode(tmp.data(), code, first_frame, n_frames);
                std::copy(tmp.begin(), tmp.end(), begin);
            }
        } else {                                            // other arithmetic types (long double ...): via 64-bit integers on the host
            if (d_signed) {
                std::vector<std::int64_t> tmp(n);
                f_decode(tmp.data(), TRPX_I64, first_frame, n_frames);
                for (std::size_t i = 0; i < n; ++i, ++begin) *begin = V(double(tmp[i]));
            } else {
                std::vector<std::uint64_t> tmp(n);
                f_decode(tmp.data(), TRPX_U64, first_frame, n_frames);
                for (std::size_t i = 0; i < n; ++i, ++begin) *begin = V(double(tmp[i]));
            }
        }
    }

    std::size_t size() const { return d_size; }
    std::size_t number_of_frames() const { return d_frame_bytes.size(); }
    std::vector<std::size_t> const& dim() const { return d_dim; }
    std::vector<std::size_t> const& dim(std::vector<std::size_t> const& dim)
    {
        assert(d_dim.size() == 0);
        return d_dim = dim;
    }
    bool is_signed() const { return d_signed; }
    unsigned bits_per_val() const { return d_prolix_bits; }
    std::size_t terse_size() const { return d_terse_data.size(); }
    unsigned block() const { return d_block; }
    std::uint8_t const* terse_data() const { return d_terse_data.data(); }
    // shard multi-frame push_back_frames / prolix_frames calls over every visible GPU (default: TRPX_MULTI_GPU, else off)
    static void use_all_devices(bool on = true) { trpx_detail::multi_gpu() = on; }

    // XML element + payload, byte-identical to the reference's writer (Terse.hpp:454-474)
    void write(std::ostream& ostream) const
    {
        std::string h = "<Terse prolix_bits=\"" + std::to_string(d_prolix_bits) + "\" signed=\"" + (d_signed ? "1" : "0") +
                        "\" block=\"" + std::to_string(d_block) + "\" memory_size=\"" + std::to_string(d_terse_data.size()) +
                        "\" number_of_values=\"" + std::to_string(d_size) + "\"";
        if (!d_dim.empty()) {
            h += " dimensions=\"";
            for (std::size_t i = 0; i < d_dim.size(); ++i) h += (i ? " " : "") + std::to_string(d_dim[i]);
            h += "\"";
        }
        h += " number_of_frames=\"" + std::to_string(d_frame_bytes.size()) + "\"/>";
        ostream.write(h.data(), std::streamsize(h.size()));
        ostream.write(reinterpret_cast<const char*>(d_terse_data.data()), std::streamsize(d_terse_data.size()));
        ostream.flush();
    }

private:
    bool d_signed = false;
    unsigned d_block = 12;
    std::size_t d_size = 0;
    unsigned d_prolix_bits = 0;
    std::vector<std::size_t> d_dim;
    trpx_detail::byte_vector d_terse_data;
    std::vector<std::size_t> d_frame_bytes;         // payload bytes of each frame; 0 = not known yet (read from a file)

    template <typename Iterator>
    void f_encode(Iterator data, std::size_t n_frames)
    {
        using V = trpx_detail::value_t<Iterator>;
        static_assert(std::is_integral_v<V>, "TERSE encodes integral pixels only");
        if (n_frames == 0 || d_size == 0) return;
        const std::size_t n = d_size * n_frames;
        std::vector<V> staged;
        const V* src;
        if constexpr (trpx_detail::is_raw_pointer<Iterator>) {
            src = data;
        } else {                                            // gather a non-contiguous range first
            staged.reserve(n);
            Iterator it = data;
            for (std::size_t i = 0; i < n; ++i, ++it) staged.push_back(*it);
            src = staged.data();
        }
        const bool multi = trpx_detail::multi_gpu() && n_frames > 1;
        trpx_ctx* ctx = multi ? nullptr : trpx_detail::context();
        const int dt = trpx_detail::dtype_of<V>();
        const std::size_t cap = trpx_max_compressed_bytes(d_size, dt, d_block, n_frames);
        std::vector<std::size_t> fb(n_frames);
        std::size_t total = 0;
        unsigned pb = 0;
        int rc;
        std::uint8_t* scratch = trpx_detail::pinned_scratch(cap);
        std::vector<std::uint8_t> pageable;                  // (only if the host refuses that much pinned memory)
        if (!scratch) { pageable.resize(cap); scratch = pageable.data(); }
        {
            trpx_detail::ScopedPin pin_in(src, n * sizeof(V));
            rc = multi ? trpx_pool_encode_host(trpx_detail::pool(), src, dt, d_size, n_frames, d_block, scratch, cap, fb.data(), &total, &pb)
                       : trpx_encode_host(ctx, src, dt, d_size, n_frames, d_block, scratch, cap, fb.data(), &total, &pb);
        }
        if (multi && rc != TRPX_OK) throw std::runtime_error(std::string("trpx_b200: ") + trpx_strerror(rc) + " (" + trpx_pool_last_error(trpx_detail::pool()) + ")");
        trpx_detail::check(rc, ctx);
        trpx_detail::append_bytes(d_terse_data, scratch, total);
        d_frame_bytes.insert(d_frame_bytes.end(), fb.begin(), fb.end());
        if (pb > d_prolix_bits) d_prolix_bits = pb;
    }

    void f_decode(void* out, int out_dtype, std::size_t first_frame, std::size_t n_frames)
    {
        if (n_frames == 0) return;
        const bool multi = trpx_detail::multi_gpu() && n_frames > 1;
        trpx_ctx* ctx = multi ? nullptr : trpx_detail::context();
        bool known = true;
        for (std::size_t b : d_frame_bytes) known = known && b != 0;
        std::vector<std::size_t> recovered(known ? 0 : d_frame_bytes.size());
        int rc;
        {
            trpx_detail::ScopedPin pin_in(d_terse_data.data(), d_terse_data.size()), pin_out(out, n_frames * d_size * trpx_dtype_size(out_dtype));
            rc = multi ? trpx_pool_decode_host(trpx_detail::pool(), d_terse_data.data(), d_terse_data.size(), d_signed ? 1 : 0, d_block, d_size,
                                               d_frame_bytes.size(), first_frame, n_frames, known ? d_frame_bytes.data() : nullptr,
                                               known ? nullptr : recovered.data(), out, out_dtype)
                       : trpx_decode_host(ctx, d_terse_data.data(), d_terse_data.size(), d_signed ? 1 : 0, d_block, d_size,
                                          d_frame_bytes.size(), first_frame, n_frames, known ? d_frame_bytes.data() : nullptr,
                                          known ? nullptr : recovered.data(), out, out_dtype);
        }
        if (multi && rc != TRPX_OK) throw std::runtime_error(std::string("trpx_b200: ") + trpx_strerror(rc) + " (" + trpx_pool_last_error(trpx_detail::pool()) + ")");
        trpx_detail::check(rc, ctx);
        if (!known) d_frame_bytes = recovered;               // frame boundaries are cached (absolute, cf. App. C1)
    }
};

} // namespace jpa
