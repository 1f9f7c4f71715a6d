// cli_common.hpp -- what the two command-line tools share: argument scan, wall-clock buckets, the -verbose report.
#pragma once

#include <chrono>
#include <filesystem>
#include <future>
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

// File I/O overlapped with the GPU: the tools spend their time reading and writing files, so
// file k + 1 is read and parsed by a helper thread while file k is on the GPU and being written.  `load` runs on the
// helper and must not print (messages stay in file order: it returns them); `process` runs on the calling thread, in
// order.  At most two files are in memory at a time.
template <typename Loaded, typename Load, typename Process>
void for_each_prefetched(std::vector<fs::path> const& files, Load load, Process process)
{
    std::future<Loaded> next;
    if (!files.empty()) next = std::async(std::launch::async, load, files[0]);
    for (std::size_t i = 0; i < files.size(); ++i) {
        Loaded cur = next.get();
        if (i + 1 < files.size()) next = std::async(std::launch::async, load, files[i + 1]);
        process(files[i], cur);
    }
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
