// prolix -- expand .trpx files to TIFF stacks on the GPU.  Same command line and file semantics as the reference CLI
// (src/prolix.cpp:18-128): every *.trpx argument becomes a .tif next to it and the .trpx is deleted; output type
// int16 / uint16 for <= 16 bits per value, int32 / uint32 up to 32 (src/prolix.cpp:69-97; the reference's 17..32-bit
// branches decode through 16-bit views -- SURVEY App. C6 -- here they use the 32-bit type); square images are assumed
// when the header holds no dimensions (:61-65).  All frames of a file are decoded in ONE call (Terse::prolix_frames).
#include <array>
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
    bool help = false, verbose = false;
    std::vector<fs::path> params;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-help") help = true;
        else if (a == "-verbose") verbose = true;
        else params.emplace_back(a);
    }
    if (help) {
        std::cout << "prolix [-help] [-verbose] [file ...]\n"
                     "  expands all files with the .trpx extension to tiff files with the .tif extension.\n"
                     "\nkeywords:\n  -help     print help\n  -verbose  print expanded filenames and compute times\n";
        return 0;
    }
    std::chrono::duration<double> user_time(0), io_time(0);
    std::size_t expanded_files = 0;
    for (fs::path filename : params) {
        if (!fs::is_regular_file(filename) || filename.extension() != ".trpx") continue;
        auto t0 = std::chrono::high_resolution_clock::now();
        try {
            std::ifstream in(filename, std::ios::binary);
            if (!in.is_open()) { std::cerr << "Failed to open input file " << filename << std::endl; continue; }
            jpa::Terse trpx(in);
            in.close();
            auto t1 = std::chrono::high_resolution_clock::now();
            io_time += t1 - t0;
            std::size_t w, h;
            if (trpx.dim().size() < 2) w = h = std::size_t(std::sqrt(double(trpx.size())));
            else { w = trpx.dim()[0]; h = trpx.dim()[1]; }
            using jpa::tiffio::Kind;
            std::vector<jpa::tiffio::Image> imgs;
            if (trpx.bits_per_val() <= 16 && trpx.is_signed()) imgs = expand<std::int16_t>(trpx, w, h, Kind::Int);
            else if (trpx.bits_per_val() <= 16) imgs = expand<std::uint16_t>(trpx, w, h, Kind::Uint);
            else if (trpx.bits_per_val() <= 32 && trpx.is_signed()) imgs = expand<std::int32_t>(trpx, w, h, Kind::Int);
            else if (trpx.bits_per_val() <= 32) imgs = expand<std::uint32_t>(trpx, w, h, Kind::Uint);
            else {
                std::cerr << "Terse file " << filename << " encodes data that requires 64 bits per pixel." << std::endl;
                std::cerr << "Prolix cannot process such trpx-stacks." << std::endl;
                return 0;
            }
            auto t2 = std::chrono::high_resolution_clock::now();
            user_time += t2 - t1;
            fs::path tif = filename;
            tif.replace_extension(".tif");
            std::ofstream out(tif, std::ios::binary);
            if (!out.is_open()) {
                std::cerr << "Failed to open tif file " << tif << std::endl;
            } else {
                jpa::tiffio::write(out, imgs);
                out.close();
                fs::remove(filename);
                ++expanded_files;
            }
            io_time += std::chrono::high_resolution_clock::now() - t2;
        } catch (std::exception const& e) {
            std::cerr << "Error processing " << filename << ": " << e.what() << std::endl;
        }
    }
    if (verbose) {
        for (fs::path const& f : params) std::cout << "Expanded: " << f << std::endl;
        std::cout << "Prolix expanded : " << expanded_files << " files\n";
        std::cout << "User time       : " << user_time.count() << " seconds\n";
        std::cout << "IO time         : " << io_time.count() << " seconds\n";
    }
    return 0;
}
