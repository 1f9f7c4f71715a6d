// cli_common.hpp -- what the two command-line tools share: argument scan, wall-clock buckets, the -verbose report.
#pragma once

#include <chrono>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>

namespace trpx_cli {

namespace fs = std::filesystem;
using Clock = std::chrono::high_resolution_clock;
using Seconds = std::chrono::duration<double>;

struct Args {
    bool help = false, verbose = false;
    std::vector<fs::path> files;
    Args(int argc, char const* argv[])
    {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "-help") help = true;
            else if (a == "-verbose") verbose = true;
            else files.emplace_back(a);
        }
    }
};

inline bool has_extension(fs::path const& p, std::initializer_list<const char*> exts)
{
    const std::string e = p.extension().string();
    for (const char* x : exts)
        if (e == x) return true;
    return false;
}

struct Report {
    Seconds user{0}, io{0};
    std::size_t done = 0;
    void print(const char* verb, const char* summary, Args const& a) const
    {
        for (fs::path const& f : a.files) std::cout << verb << ": " << f << std::endl;
        std::cout << summary << done << " files\n";
        std::cout << "User time       : " << user.count() << " seconds\n";
        std::cout << "IO time         : " << io.count() << " seconds\n";
    }
};

} // namespace trpx_cli
