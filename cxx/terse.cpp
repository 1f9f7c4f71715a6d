// terse -- compress TIFF stacks to .trpx on the GPU.  Same command line and file semantics as the reference CLI
// (src/terse.cpp:20-104): every *.tif / *.tiff argument becomes a .trpx next to it and the TIFF is deleted;
// -help, -verbose ("Terse compressed", "User time", "IO time", "Compression rate").  The whole stack of a file goes
// to the GPU in ONE call (Terse::push_back_frames) instead of one push_back per image.
#include <chrono>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <trpx/Grey_tiff_io.hpp>
#include <trpx/Terse.hpp>

namespace fs = std::filesystem;

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
    bool help = false, verbose = false;
    std::vector<fs::path> params;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-help") help = true;
        else if (a == "-verbose") verbose = true;
        else params.emplace_back(a);
    }
    if (help) {
        std::cout << "terse [-help] [-verbose] [file ...]\n"
                     "  compresses all files with .tiff or .tif extensions to terse files with .trpx extensions.\n"
                     "Examples:\n"
                     "   terse *                   // all tiff files in this directory are compressed to trpx files.\n"
                     "   terse ~/dir/my_img*       // compresses all tiff files in the directory ~/dir that start with my_img\n"
                     "\nkeywords:\n  -help     print help\n  -verbose  print compressed filenames, compute times and compression rate\n";
        return 0;
    }
    std::chrono::duration<double> user_time(0), io_time(0);
    double total_trpx = 0, total_tiff = 0;
    std::size_t compressed_files = 0;
    for (fs::path const& tif : params) {
        const std::string ext = tif.extension().string();
        if (!fs::is_regular_file(tif) || !(ext == ".tiff" || ext == ".tif" || ext == ".TIFF" || ext == ".TIF")) continue;
        try {
            auto t0 = std::chrono::high_resolution_clock::now();
            std::ifstream in(tif, std::ios::binary);
            if (!in.is_open()) { std::cerr << "Failed to open input file " << tif << std::endl; continue; }
            std::vector<jpa::tiffio::Image> imgs = jpa::tiffio::read(in);
            in.close();
            if (imgs.empty()) throw std::runtime_error("TIFF file contains no image.");
            for (auto const& im : imgs) {
                total_tiff += double(im.data.size());
                if (im.width != imgs[0].width || im.height != imgs[0].height || im.bits != imgs[0].bits || im.kind != imgs[0].kind)
                    throw std::runtime_error("TIFF file contains a stack of images with varying sizes.");
            }
            auto t1 = std::chrono::high_resolution_clock::now();
            jpa::Terse compressed;
            compressed.dim({imgs[0].width, imgs[0].height});          // "width height", as Grey_tif reports it
            using jpa::tiffio::Kind;
            const unsigned b = imgs[0].bits;
            const Kind k = imgs[0].kind;
            if (k == Kind::Uint && b == 8) push_stack<std::uint8_t>(compressed, imgs);
            else if (k == Kind::Uint && b == 16) push_stack<std::uint16_t>(compressed, imgs);
            else if (k == Kind::Uint && b == 32) push_stack<std::uint32_t>(compressed, imgs);
            else if (k == Kind::Int && b == 8) push_stack<std::int8_t>(compressed, imgs);
            else if (k == Kind::Int && b == 16) push_stack<std::int16_t>(compressed, imgs);
            else if (k == Kind::Int && b == 32) push_stack<std::int32_t>(compressed, imgs);
            else if (k == Kind::Float && b == 32) push_stack_float<float>(compressed, imgs);
            else if (k == Kind::Float && b == 64) push_stack_float<double>(compressed, imgs);
            else throw std::runtime_error("unsupported TIFF pixel type.");
            total_trpx += double(compressed.terse_size());
            fs::path trpx = tif;
            trpx.replace_extension(".trpx");
            std::ofstream out(trpx, std::ios::binary);
            if (!out.is_open()) throw std::runtime_error("Failed to open trpx file for output.");
            compressed.write(out);
            out.close();
            std::cout << "Deleting original TIFF file: " << tif << std::endl;
            fs::remove(tif);
            ++compressed_files;
            auto t2 = std::chrono::high_resolution_clock::now();
            user_time += t2 - t1;
            io_time += t1 - t0;
        } catch (std::exception const& e) {
            std::cerr << "Error processing " << tif << ": " << e.what() << std::endl;
        }
    }
    if (verbose) {
        for (fs::path const& f : params) std::cout << "Compressed: " << f << std::endl;
        std::cout << "Terse compressed: " << compressed_files << " files\n";
        std::cout << "User time       : " << user_time.count() << " seconds\n";
        std::cout << "IO time         : " << io_time.count() << " seconds\n";
        if (total_tiff > 0) std::cout << "Compression rate: " << std::round(1000 * (1 - total_trpx / total_tiff)) / 10 << "%\n";
    }
    return 0;
}
