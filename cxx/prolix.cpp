// prolix -- expand .trpx files to TIFF stacks on the GPU.  Same command line and file semantics as the reference CLI
// (src/prolix.cpp:18-128): every *.trpx argument becomes a .tif next to it and the .trpx is deleted; output type
// int16 / uint16 for <= 16 bits per value, int32 / uint32 up to 32 (src/prolix.cpp:69-97; the reference's 17..32-bit
// branches decode through 16-bit views -- SURVEY App. C6 -- here they use the 32-bit type); square images are assumed
// when the header holds no dimensions (:61-65).  All frames of a file are decoded in ONE call (Terse::prolix_frames), and
// the next file is read while the current one is expanded and written (cli_common.hpp: for_each_prefetched).
#include <array>
#include <chrono>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "cli_common.hpp"
#include <trpx/Grey_tiff_io.hpp>
#include <trpx/Terse.hpp>

template <typename T>
static std::vector<jpa::tiffio::Image> expand(jpa::Terse& t, std::size_t w, std::size_t h, jpa::tiffio::Kind kind)
{
    const std::size_t n = t.size(), frames = t.number_of_frames();
    std::vector<T> all(n * frames);
    t.prolix_frames(all.data(), 0, frames);
    std::vector<jpa::tiffio::Image> imgs(frames);
    for (std::size_t f = 0; f < frames; ++f) {
        imgs[f].width = w; imgs[f].height = h; imgs[f].bits = 8 * sizeof(T); imgs[f].kind = kind;
        imgs[f].data.resize(w * h * sizeof(T));                      // zero-filled if the frame is smaller (sqrt rounding)
        std::memcpy(imgs[f].data.data(), all.data() + f * n, std::min(n, w * h) * sizeof(T));
    }
    return imgs;
}

int main(int argc, char const* argv[])
{
    using namespace trpx_cli;
    const Args args(argc, argv);
    if (args.help) {
        std::cout << "prolix [-help] [-verbose] [file ...]\n"
                     "  expands all files with the .trpx extension to tiff files with the .tif extension.\n"
                     "\nkeywords:\n  -help     print help\n  -verbose  print expanded filenames and compute times\n";
        return 0;
    }
    Report rep;
    std::vector<fs::path> todo;
    for (fs::path const& source : args.files)
        if (fs::is_regular_file(source) && has_extension(source, {".trpx"})) todo.push_back(source);
    struct Loaded {
        bool opened = false;
        std::unique_ptr<jpa::Terse> packed;
        std::string error;
        Seconds io{0};
    };
    auto load = [](fs::path const& source) {                           // (helper thread: reads the container, never prints)
        Loaded l;
        const auto t_open = Clock::now();
        try {
            std::ifstream in(source, std::ios::binary);
            l.opened = in.is_open();
            if (l.opened) l.packed = std::make_unique<jpa::Terse>(in);
        } catch (std::exception const& e) {
            l.error = e.what();
        }
        l.io = Clock::now() - t_open;
        return l;
    };
    bool give_up = false;
    for_each_prefetched<Loaded>(todo, load, [&](fs::path const& source, Loaded& got) {
        if (give_up) return;
        try {
            if (!got.opened) { std::cerr << "Failed to open input file " << source << std::endl; return; }
            if (!got.error.empty()) throw std::runtime_error(got.error);
            jpa::Terse& packed = *got.packed;
            const auto t_gpu = Clock::now();
            rep.io += got.io;
            std::size_t w, h;                                          // no dimensions in the header: a square image
            if (packed.dim().size() < 2) w = h = std::size_t(std::sqrt(double(packed.size())));
            else { w = packed.dim()[0]; h = packed.dim()[1]; }
            // (the reference truncates or zero-fills silently when w * h differs from the frame size, src/prolix.cpp:61-65,
            // and deletes the .trpx all the same: here the source is kept whenever the image cannot hold every value)
            const bool lossless = w * h == packed.size();
            using jpa::tiffio::Kind;
            const unsigned bits = packed.bits_per_val();
            std::vector<jpa::tiffio::Image> stack;
            if (bits > 32) {
                std::cerr << "Terse file " << source << " encodes data that requires 64 bits per pixel." << std::endl;
                std::cerr << "Prolix cannot process such trpx-stacks." << std::endl;
                give_up = true;                                        // (the reference returns here, src/prolix.cpp:93-97)
                return;
            }
            if (packed.is_signed()) stack = bits <= 16 ? expand<std::int16_t>(packed, w, h, Kind::Int) : expand<std::int32_t>(packed, w, h, Kind::Int);
            else stack = bits <= 16 ? expand<std::uint16_t>(packed, w, h, Kind::Uint) : expand<std::uint32_t>(packed, w, h, Kind::Uint);
            const auto t_write = Clock::now();
            rep.user += t_write - t_gpu;
            fs::path target = source;
            target.replace_extension(".tif");
            fs::path tmp = target;
            tmp += ".part";
            std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
            if (!out.is_open()) {
                std::cerr << "Failed to open tif file " << target << std::endl;
            } else {
                jpa::tiffio::write(out, stack);
                const bool wrote = out.good();
                const std::uintmax_t expect = std::uintmax_t(out.tellp());
                out.close();
                std::error_code ec;
                if (!wrote || out.fail() || fs::file_size(tmp, ec) != expect || ec) {     // full disk, quota, I/O error: keep the source
                    fs::remove(tmp, ec);
                    throw std::runtime_error("Failed to write the tif file (the trpx file is kept).");
                }
                fs::rename(tmp, target);
                if (lossless) fs::remove(source);
                else std::cerr << "Frame size " << packed.size() << " is not " << w << " x " << h << ": keeping " << source << std::endl;
                ++rep.done;
            }
            rep.io += Clock::now() - t_write;
        } catch (std::exception const& e) {
            std::cerr << "Error processing " << source << ": " << e.what() << std::endl;
        }
    });
    if (give_up) return 0;
    if (args.verbose) rep.print("Expanded", "Prolix expanded : ", args);
    return 0;
}
