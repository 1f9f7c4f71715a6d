"""Builds trpx_b200/libtrpx_b200.so (the C-ABI library of include/trpx_b200.h) IN-TREE with nvcc for
sm_100a.  nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtrpx_b200.so")
SOURCES = [os.path.join(CSRC, "trpx_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("simt.cuh", "terse_encode.cuh", "prolix_decode.cuh",
                                                   "codec_launch.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "trpx_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-cudart", "static"]


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", OUT]
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libtrpx_b200.so")
    return OUT


ROOT = os.path.dirname(HERE)
HOST_TARGETS = {"terse_selftest": [os.path.join(ROOT, "cxx", "terse_selftest.cpp")],
                "terse": [os.path.join(ROOT, "cxx", "terse.cpp")],
                "prolix": [os.path.join(ROOT, "cxx", "prolix.cpp")]}


def build_host(force=False):
    """Host-side C++ on top of the C ABI (include/trpx/Terse.hpp): g++ -std=c++20, linked against the in-tree .so."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    outs = []
    for name, srcs in HOST_TARGETS.items():
        srcs = [s for s in srcs if os.path.exists(s)]
        if not srcs:
            continue
        out = os.path.join(ROOT, "cxx", name)
        deps = srcs + [os.path.join(ROOT, "include", "trpx", "Terse.hpp"), os.path.join(ROOT, "include", "trpx", "Grey_tiff_io.hpp"), os.path.join(ROOT, "cxx", "cli_common.hpp"),
                       os.path.join(ROOT, "include", "trpx_b200.h"), OUT]
        if force or not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps if os.path.exists(d)):
            cmd = [cxx, "-std=c++20", "-O2", "-DNDEBUG", "-Wall", "-I", os.path.join(ROOT, "include")] + srcs + [
                "-L", HERE, "-ltrpx_b200", "-Wl,-rpath," + HERE, "-o", out]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                raise RuntimeError("g++ failed building " + name)
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
