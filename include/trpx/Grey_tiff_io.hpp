// trpx/Grey_tiff_io.hpp -- minimal host-side reader / writer for uncompressed greyscale TIFF stacks, as
// much as the `terse` / `prolix` CLIs need (the reference keeps this host-side too: include/Grey_tif.hpp,
// BASELINE north_star "Grey_tif/TIFF I/O ... remain host-side").  Written from the TIFF 6.0 baseline layout,
// not from the reference's class: classic (32-bit offset) TIFF, II or MM byte order on read, II on write,
// one sample per pixel, Compression = 1, strips; 8 / 16 / 32-bit unsigned or signed integers and 32 / 64-bit
// IEEE floats (SampleFormat 1 / 2 / 3); several IFDs = a stack of images.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace jpa::tiffio {

enum class Kind { Uint, Int, Float };

struct Image {
    std::size_t width = 0, height = 0;
    unsigned bits = 0;                       // bits per sample: 8, 16, 32, 64
    Kind kind = Kind::Uint;
    std::vector<std::uint8_t> data;          // width * height samples, host byte order, row-major
    std::size_t pixels() const { return width * height; }
    template <typename T> T const* as() const { return reinterpret_cast<T const*>(data.data()); }
    template <typename T> T* as() { return reinterpret_cast<T*>(data.data()); }
};

namespace detail {

struct Reader {
    std::vector<std::uint8_t> const& f;
    bool big;
    std::uint64_t get(std::size_t off, unsigned n) const
    {
        if (n > f.size() || off > f.size() - n) throw std::runtime_error("TIFF: truncated file");   // (no overflow in off + n)
        std::uint64_t v = 0;
        for (unsigned i = 0; i < n; ++i) v |= std::uint64_t(f[off + (big ? n - 1 - i : i)]) << (8 * i);
        return v;
    }
};

inline unsigned type_size(unsigned t)
{
    switch (t) { case 1: case 2: case 6: case 7: return 1; case 3: case 8: return 2; case 4: case 9: case 11: return 4;
                 case 5: case 10: case 12: return 8; default: return 0; }
}

// values of one IFD entry (integers only)
inline std::vector<std::uint64_t> entry_values(Reader const& r, std::size_t e)
{
    const unsigned type = unsigned(r.get(e + 2, 2));
    const std::uint64_t count = r.get(e + 4, 4);
    const unsigned ts = type_size(type);
    if (!ts || (type != 1 && type != 3 && type != 4)) throw std::runtime_error("TIFF: unsupported tag type");
    if (count == 0) throw std::runtime_error("TIFF: tag without a value");
    if (count > r.f.size()) throw std::runtime_error("TIFF: tag count larger than the file");
    std::size_t off = ts * count <= 4 ? e + 8 : std::size_t(r.get(e + 8, 4));
    std::vector<std::uint64_t> v(count);
    for (std::uint64_t i = 0; i < count; ++i) v[i] = r.get(off + i * ts, ts);
    return v;
}

inline void put(std::vector<std::uint8_t>& o, std::uint64_t v, unsigned n)
{
    for (unsigned i = 0; i < n; ++i) o.push_back(std::uint8_t(v >> (8 * i)));
}

} // namespace detail

// Reads every image of a TIFF stack.
inline std::vector<Image> read(std::istream& in)
{
    std::vector<std::uint8_t> f((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (f.size() < 8) throw std::runtime_error("TIFF: file too short");
    const bool big = f[0] == 'M' && f[1] == 'M';
    if (!big && !(f[0] == 'I' && f[1] == 'I')) throw std::runtime_error("TIFF: bad byte-order mark");
    detail::Reader r{f, big};
    if (r.get(2, 2) != 42) throw std::runtime_error("TIFF: not a classic TIFF (BigTIFF is not supported)");
    std::vector<Image> images;
    std::size_t ifd = std::size_t(r.get(4, 4));
    std::vector<std::size_t> seen;                            // a cyclic chain of IFD offsets must not loop for ever
    while (ifd) {
        for (std::size_t s : seen)
            if (s == ifd) throw std::runtime_error("TIFF: cyclic image directory");
        seen.push_back(ifd);
        if (seen.size() > f.size() / 14 + 1) throw std::runtime_error("TIFF: more image directories than the file can hold");
        const unsigned n = unsigned(r.get(ifd, 2));
        Image img;
        unsigned compression = 1, spp = 1, fmt = 1;
        std::vector<std::uint64_t> offs, counts;
        for (unsigned i = 0; i < n; ++i) {
            const std::size_t e = ifd + 2 + 12 * i;
            const unsigned tag = unsigned(r.get(e, 2));
            switch (tag) {
            case 256: img.width = std::size_t(detail::entry_values(r, e)[0]); break;
            case 257: img.height = std::size_t(detail::entry_values(r, e)[0]); break;
            case 258: img.bits = unsigned(detail::entry_values(r, e)[0]); break;
            case 259: compression = unsigned(detail::entry_values(r, e)[0]); break;
            case 273: offs = detail::entry_values(r, e); break;
            case 277: spp = unsigned(detail::entry_values(r, e)[0]); break;
            case 279: counts = detail::entry_values(r, e); break;
            case 339: fmt = unsigned(detail::entry_values(r, e)[0]); break;
            default: break;
            }
        }
        if (compression != 1) throw std::runtime_error("TIFF: compressed images are not supported");
        if (spp != 1) throw std::runtime_error("TIFF: only greyscale (one sample per pixel) is supported");
        if (img.bits != 8 && img.bits != 16 && img.bits != 32 && img.bits != 64) throw std::runtime_error("TIFF: unsupported bits per sample");
        img.kind = fmt == 2 ? Kind::Int : fmt == 3 ? Kind::Float : Kind::Uint;
        if (img.kind == Kind::Float && img.bits < 32) throw std::runtime_error("TIFF: unsupported float width");
        if (img.kind != Kind::Float && img.bits == 64) throw std::runtime_error("TIFF: 64-bit integer samples are not supported");
        const std::size_t bps = img.bits / 8;
        if (img.width == 0 || img.height == 0 || img.width > f.size() || img.height > f.size() / img.width / bps)
            throw std::runtime_error("TIFF: image larger than the file");   // (also keeps width * height * bps from overflowing)
        const std::size_t want = img.pixels() * bps;
        img.data.resize(want);
        std::size_t got = 0;
        for (std::size_t s = 0; s < offs.size() && got < want; ++s) {
            std::size_t cnt = s < counts.size() ? std::size_t(counts[s]) : want - got;
            if (cnt > want - got) cnt = want - got;
            if (cnt > f.size() || offs[s] > f.size() - cnt) throw std::runtime_error("TIFF: strip outside the file");
            std::memcpy(img.data.data() + got, f.data() + offs[s], cnt);
            got += cnt;
        }
        if (got != want) throw std::runtime_error("TIFF: image data incomplete");
        if (big && bps > 1)                                   // to host (little-endian) order
            for (std::size_t i = 0; i < img.pixels(); ++i)
                for (std::size_t b = 0; b < bps / 2; ++b) std::swap(img.data[i * bps + b], img.data[i * bps + bps - 1 - b]);
        images.push_back(std::move(img));
        ifd = std::size_t(r.get(ifd + 2 + 12 * n, 4));
    }
    return images;
}

// Writes a stack (little-endian classic TIFF, one strip per image).
inline void write(std::ostream& out, std::vector<Image> const& images)
{
    std::vector<std::uint8_t> o;
    o.push_back('I'); o.push_back('I');
    detail::put(o, 42, 2);
    detail::put(o, 0, 4);                                     // patched: offset of the first IFD
    std::size_t link = 4;                                     // where the next IFD's offset goes
    for (Image const& img : images) {
        while (o.size() & 1) o.push_back(0);
        const std::size_t data_off = o.size();
        o.insert(o.end(), img.data.begin(), img.data.end());
        while (o.size() & 1) o.push_back(0);
        const std::size_t ifd = o.size();
        if (ifd + 200 > 0xffffffffull) throw std::runtime_error("TIFF: stack too large for classic TIFF");
        for (unsigned i = 0; i < 4; ++i) o[link + i] = std::uint8_t(ifd >> (8 * i));
        struct E { unsigned tag, type; std::uint64_t v; };
        const E es[] = {{256, 4, img.width}, {257, 4, img.height}, {258, 3, img.bits}, {259, 3, 1}, {262, 3, 1},
                        {273, 4, data_off}, {277, 3, 1}, {278, 4, img.height}, {279, 4, img.data.size()},
                        {339, 3, img.kind == Kind::Int ? 2u : img.kind == Kind::Float ? 3u : 1u}};
        detail::put(o, sizeof(es) / sizeof(es[0]), 2);
        for (E const& e : es) {
            detail::put(o, e.tag, 2);
            detail::put(o, e.type, 2);
            detail::put(o, 1, 4);
            detail::put(o, e.v, 4);
        }
        link = o.size();
        detail::put(o, 0, 4);
    }
    out.write(reinterpret_cast<const char*>(o.data()), std::streamsize(o.size()));
    out.flush();
}

} // namespace jpa::tiffio
