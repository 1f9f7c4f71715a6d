// terse -- compress TIFF stacks to .trpx on the GPU.  Same command line and file semantics as the reference CLI
// (src/terse.cpp:20-104): every *.tif / *.tiff argument becomes a .trpx next to it and the TIFF is deleted;
// -help, -verbose ("Terse compressed", "User time", "IO time", "Compression rate").  The whole stack of a file goes
// to the GPU in ONE call (Terse::push_back_frames) instead of one push_back per image, and the next file is read while
// the current one is compressed and written (cli_common.hpp: for_each_prefetched).
#include <chrono>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "cli_common.hpp"
#include <trpx/Grey_tiff_io.hpp>
#include <trpx/Terse.hpp>

template <typename T>
static void push_stack(jpa::Terse& t, std::vector<jpa::tiffio::Image> const& imgs)
{
    const std::size_t n = imgs[0].pixels();
    std::vector<T> all(n * imgs.size());
    for (std::size_t i = 0; i < imgs.size(); ++i) std::copy_n(imgs[i].as<T>(), n, all.begin() + i * n);
    t.push_back_frames(all.data(), n, imgs.size());
}

template <typename F>
static void push_stack_float(jpa::Terse& t, std::vector<jpa::tiffio::Image> const& imgs)   // as src/terse.cpp:119-124: via int64
{
    const std::size_t n = imgs[0].pixels();
    std::vector<std::int64_t> all(n * imgs.size());
    for (std::size_t i = 0; i < imgs.size(); ++i)
        for (std::size_t k = 0; k < n; ++k) all[i * n + k] = std::int64_t(imgs[i].as<F>()[k]);
    t.push_back_frames(all.data(), n, imgs.size());
}

int main(int argc, char const* argv[])
{
    using namespace trpx_cli;
    const Args args(argc, argv);
    if (args.help) {
        std::cout << "terse [-help] [-verbose] [file ...]\n"
                     "  compresses all files with .tiff or .tif extensions to terse files with .trpx extensions.\n"
                     "Examples:\n"
                     "   terse *                   // all tiff files in this directory are compressed to trpx files.\n"
                     "   terse ~/dir/my_img*       // compresses all tiff files in the directory ~/dir that start with my_img\n"
                     "\nkeywords:\n  -help     print help\n  -verbose  print compressed filenames, compute times and compression rate\n";
        return 0;
    }
    Report rep;
    double trpx_bytes = 0, tiff_bytes = 0;
    std::vector<fs::path> todo;
    for (fs::path const& tif : args.files)
        if (fs::is_regular_file(tif) && has_extension(tif, {".tiff", ".tif", ".TIFF", ".TIF"})) todo.push_back(tif);
    struct Loaded {
        bool opened = false;
        std::vector<jpa::tiffio::Image> stack;
        std::string error;
        Seconds io{0};
    };
    auto load = [](fs::path const& tif) {                              // (helper thread: reads and parses, never prints)
        Loaded l;
        const auto t_open = Clock::now();
        try {
            std::ifstream in(tif, std::ios::binary);
            l.opened = in.is_open();
            if (l.opened) l.stack = jpa::tiffio::read(in);
        } catch (std::exception const& e) {
            l.error = e.what();
        }
        l.io = Clock::now() - t_open;
        return l;
    };
    for_each_prefetched<Loaded>(todo, load, [&](fs::path const& tif, Loaded& got) {
        try {
            if (!got.opened) { std::cerr << "Failed to open input file " << tif << std::endl; return; }
            if (!got.error.empty()) throw std::runtime_error(got.error);
            std::vector<jpa::tiffio::Image> const& stack = got.stack;
            if (stack.empty()) throw std::runtime_error("TIFF file contains no image.");
            jpa::tiffio::Image const& first = stack.front();
            for (auto const& im : stack) {
                tiff_bytes += double(im.data.size());
                if (im.width != first.width || im.height != first.height || im.bits != first.bits || im.kind != first.kind)
                    throw std::runtime_error("TIFF file contains a stack of images with varying sizes.");
            }
            const auto t_gpu = Clock::now();
            jpa::Terse packed;
            packed.dim({first.width, first.height});                  // "width height", the order Grey_tif reports
            using jpa::tiffio::Kind;
            switch (first.kind == Kind::Float ? 100 + first.bits : (first.kind == Kind::Int ? 1000 : 0) + first.bits) {
            case 8: push_stack<std::uint8_t>(packed, stack); break;
            case 16: push_stack<std::uint16_t>(packed, stack); break;
            case 32: push_stack<std::uint32_t>(packed, stack); break;
            case 1008: push_stack<std::int8_t>(packed, stack); break;
            case 1016: push_stack<std::int16_t>(packed, stack); break;
            case 1032: push_stack<std::int32_t>(packed, stack); break;
            case 132: push_stack_float<float>(packed, stack); break;
            case 164: push_stack_float<double>(packed, stack); break;
            default: throw std::runtime_error("unsupported TIFF pixel type.");
            }
            trpx_bytes += double(packed.terse_size());
            fs::path target = tif;
            target.replace_extension(".trpx");
            // The source is the only copy of the data: write to a temporary name, check the stream after the write AND
            // after the close (a full disk or a quota shows up only there), check the size on disk, rename into place, and
            // only then delete the TIFF (the reference removes it unconditionally, src/terse.cpp:79-82).
            fs::path tmp = target;
            tmp += ".part";
            std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
            if (!out.is_open()) throw std::runtime_error("Failed to open trpx file for output.");
            packed.write(out);
            const bool wrote = out.good();
            const std::uintmax_t expect = std::uintmax_t(out.tellp());
            out.close();
            std::error_code ec;
            if (!wrote || out.fail() || fs::file_size(tmp, ec) != expect || ec || expect < packed.terse_size()) {
                fs::remove(tmp, ec);
                throw std::runtime_error("Failed to write the trpx file (the TIFF file is kept).");
            }
            fs::rename(tmp, target);
            std::cout << "Deleting original TIFF file: " << tif << std::endl;
            fs::remove(tif);
            ++rep.done;
            rep.user += Clock::now() - t_gpu;
            rep.io += got.io;
        } catch (std::exception const& e) {
            std::cerr << "Error processing " << tif << ": " << e.what() << std::endl;
        }
    });
    if (args.verbose) {
        rep.print("Compressed", "Terse compressed: ", args);
        if (tiff_bytes > 0) std::cout << "Compression rate: " << std::round(1000 * (1 - trpx_bytes / tiff_bytes)) / 10 << "%\n";
    }
    return 0;
}
